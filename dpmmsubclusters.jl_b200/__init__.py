"""dpmmsubclusters.jl_b200 -- B200-native data-parallel sweep for the DPMM sub-cluster sampler.

The directory name contains a dot, so it is not importable with a plain `import`; use
`dpmm_pkg.load()` at the repository root (registers the package as `dpmmsubclusters_jl_b200`).
"""
from . import _lib  # noqa: F401
from .sweep import GpuSweep, NIW, MULTINOMIAL  # noqa: F401

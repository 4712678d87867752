"""Import helper: the package directory is named `dpmmsubclusters.jl_b200` (not a valid Python
identifier), so it is registered in sys.modules as `dpmmsubclusters_jl_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "dpmmsubclusters.jl_b200")
NAME = "dpmmsubclusters_jl_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod

#!/bin/bash
# development aid: time the C2 sweep kernels with the default library and every build/var_*.so variant
echo "default: $(python tools/tc2_probe.py c2 20)"
for v in dpmmsubclusters.jl_b200/build/var_*.so; do
  echo "$(basename $v): $(DPMM_LIB_PATH=$PWD/$v python tools/tc2_probe.py c2 20)"
done

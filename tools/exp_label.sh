#!/bin/bash
# sweep label-kernel launch variants (development aid)
for cfg in "2 128 20" "2 128 4" "1 128 20" "1 128 4" "4 64 4" "4 128 4"; do
  set -- $cfg
  echo "== P=$1 T=$2 KC=$3"
  DPMM_LABEL_P=$1 DPMM_LABEL_T=$2 DPMM_LABEL_KC=$3 python tools/quick_timing.py niw 1e6 32 20 10 2>&1 | grep -E "label|error|Error" | head -2
done

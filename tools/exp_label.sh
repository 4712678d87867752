#!/bin/bash
for cfg in "1 12" "1 10" "1 8" "0 0"; do
  set -- $cfg
  echo "== FORM=$1 W=$2"
  if [ "$2" = "0" ]; then
    DPMM_LABEL_FORM=$1 python tools/quick_timing.py niw 1e6 32 20 10 2>&1 | grep -E "label|rror" | head -2
  else
    DPMM_LABEL_FORM=$1 DPMM_LABEL_W=$2 python tools/quick_timing.py niw 1e6 32 20 10 2>&1 | grep -E "label|rror" | head -2
  fi
done

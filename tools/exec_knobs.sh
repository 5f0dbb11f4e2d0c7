#!/bin/bash
# GPU-box helper: exec kernel variants (device-resident 1 GiB)
for cfg in "5 5" "6 6" "6 5" "8 8" "8 6" "8 7"; do
  set -- $cfg
  echo "variant=$1 warps=$2: $(ZK_EXEC_W5=$1 ZK_EXEC_WARPS=$2 ZK_PROF_REPS=3 python tools/prof_codec.py 2>&1 | tail -1)"
done

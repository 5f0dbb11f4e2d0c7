#!/bin/bash
# GPU-box helper: exec kernel warps per frame (device-resident 1 GiB; 0 = the heuristic, which picks the 96-register
# build with 5 warps from four frames per SM on)
for w in 0 2 4 5 8 16; do
  echo "warps=$w: $(ZK_EXEC_WARPS=$w ZK_PROF_REPS=3 python tools/prof_codec.py 2>&1 | tail -1)"
done

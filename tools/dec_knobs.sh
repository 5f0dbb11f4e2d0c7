#!/bin/bash
# GPU-box helper: entropy-stage co-scheduling knobs of the decoder (device-resident 1 GiB)
for cfg in "4 5 1" "4 5 0" "2 5 1" "3 5 1" "3 4 1" "2 3 1" "4 3 1" "3 3 1" "4 2 1"; do
  set -- $cfg
  echo "seq_ctas=$1 huf_ctas=$2 side=$3: $(ZK_SEQ_CTAS=$1 ZK_HUF_CTAS=$2 ZK_DEV_SIDE=$3 ZK_PROF_REPS=3 python tools/prof_codec.py 2>&1 | tail -1)"
done

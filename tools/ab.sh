#!/bin/bash
# GPU-box A/B helper: device-resident configs[1] step (1 GiB, L1) under different environment settings, same box, same run.
# usage: tools/ab.sh "VAR=val VAR2=val" "VAR=val" ...   (each argument is one variant; "-" = defaults)
for v in "$@"; do
  if [ "$v" = "-" ]; then v=""; fi
  echo "== variant: [$v]"
  env $v ZK_DEV_SUB_BYTES=${SUB:-1073741824} LVL=${LVL:-1} CK=${CK:-0} python tools/c4_probe.py ${GIB:-1} x 2>&1 | grep -v Warn
done

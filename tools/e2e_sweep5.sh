#!/bin/bash
# GPU-box: host-pointer decompress/compress throughput over pipeline shapes (sub-batch bytes x slots x share)
for sub in 67108864 134217728 268435456; do for slots in 4 8; do for share in 2 3 5; do
  echo "== sub $((sub>>20)) slots $slots share $share"
  ZK_HOST_SUB_BYTES=$sub ZK_HOST_SLOTS=$slots ZK_HOST_SHARE=$share REPS=2 timeout 120 python tools/host_probe.py 1 2>&1 | grep e2e
done; done; done

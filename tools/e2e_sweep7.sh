#!/bin/bash
for conn in 8 32; do for ramp in 0 2 3; do for side in 1 0; do
  echo "== conn $conn ramp $ramp side $side"
  CUDA_DEVICE_MAX_CONNECTIONS=$conn ZK_HOST_SIDE=$side ZK_HOST_RAMP=$ramp REPS=3 timeout 120 python tools/host_probe.py 1 2>&1 | grep e2e
done; done; done

#!/bin/bash
for ramp in 0 1 2 3 4; do for sub in 134217728 268435456; do
  echo "== ramp $ramp sub $((sub>>20))"
  ZK_HOST_RAMP=$ramp ZK_HOST_SUB_BYTES=$sub REPS=3 timeout 120 python tools/host_probe.py 1 2>&1 | grep e2e
done; done

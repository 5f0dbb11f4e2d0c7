#!/bin/bash
# GPU-box helper: host-pointer pipelines vs hardware queues (side streams on/off, CUDA_DEVICE_MAX_CONNECTIONS)
sed -n '/^cat > \/tmp\/e2e_one.py/,/^PY$/p' tools/e2e_sweep2.sh > /tmp/mk.sh; bash /tmp/mk.sh
CFG="4,128,128,2 6,128,128,3 8,128,128,4 8,64,64,4 6,256,128,3"
for side in 0 1; do for conn in 8 32; do
  echo "side=$side connections=$conn"
  ZK_HOST_SIDE=$side CUDA_DEVICE_MAX_CONNECTIONS=$conn ZK_HOST_PRIO=1 python /tmp/e2e_one.py $CFG
done; done
echo "prio=0 side=0 conn=32"; ZK_HOST_SIDE=0 CUDA_DEVICE_MAX_CONNECTIONS=32 ZK_HOST_PRIO=0 python /tmp/e2e_one.py $CFG

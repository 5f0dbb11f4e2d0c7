#!/bin/bash
# GPU-box helper: share hint / slots of the host-pointer pipelines after the exec register-budget change
sed -n '/^cat > \/tmp\/e2e_one.py/,/^PY$/p' tools/e2e_sweep2.sh > /tmp/mk.sh; bash /tmp/mk.sh
CFG="8,128,128,1 8,128,128,2 8,128,128,3 8,256,256,2 8,256,64,1 4,128,128,2 4,128,256,2 3,128,128,2 5,128,128,2 4,128,64,2 7,128,128,3"
python /tmp/e2e_one.py $CFG

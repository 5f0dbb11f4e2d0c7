#!/bin/bash
# GPU-box helper: one round of knob checks (entropy-stage CTAs, host pipeline shapes)
echo "base: $(ZK_PROF_REPS=3 python tools/prof_codec.py | tail -1)"
for cfg in "4 6" "3 8" "2 8" "3 6" "4 9"; do set -- $cfg; echo "seq_ctas=$1 huf_ctas=$2: $(ZK_SEQ_CTAS=$1 ZK_HUF_CTAS=$2 ZK_PROF_REPS=3 python tools/prof_codec.py | tail -1)"; done
sed -n '/^cat > \/tmp\/e2e_one.py/,/^PY$/p' tools/e2e_sweep2.sh > /tmp/mk.sh; bash /tmp/mk.sh
ZK_HOST_SLOTS_ENC=4 python /tmp/e2e_one.py 8,128,128,3 8,96,128,3 8,192,128,3 8,128,128,2 7,128,96,3 8,160,160,3

#!/bin/bash
# GPU-box helper (round 2): everything committed under profiles/ from one box.
#   usage: tools/gpu_final_r2.sh [tests|bench|ncu|seek ...]   (default: all)
mkdir -p gpurun_out
what="${@:-tests bench ncu seek}"
for w in $what; do case $w in
tests) (time timeout 600 python -m pytest tests -m gpu -q --timeout=150 2>&1 | tail -4) > gpurun_out/r2_final_tests.log 2>&1; cat gpurun_out/r2_final_tests.log;;
bench)
  python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r2_reference_arm.json
  python bench.py 2> gpurun_out/bench_r2_n1.err | tail -1 > gpurun_out/bench_r2_n1.json
  python -c "
import json;d=json.load(open('gpurun_out/bench_r2_n1.json'));r=json.load(open('gpurun_out/bench_r2_reference_arm.json'))
print('ours', d['value'], d['compress_GiBps'], d['decompress_GiBps'], 'e2e', d['e2e']['value'], d['e2e']['compress_GiBps'], d['e2e']['decompress_GiBps'], 'c4', d['config4_one_gpu']['value'])
print('ref', r['value'], r['compress_GiBps'], r['decompress_GiBps']); print(d['roofline']['kernel_ms'], d['roofline']['frac'])";;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ncu_launches_bench_r2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2_ncu_bench.log 2>&1
  CK=1 LVL=1 ZK_PROF_REPS=2 timeout 1000 ncu --set full --import-source on --clock-control none -k regex:"zk_(exec|match|seq_enc|lit_enc|huf|seq2|xxh64|frame_hash|scan)_kernel" -s 9 -c 9 -o gpurun_out/ncu_r2 --force-overwrite python tools/prof_codec.py 2>&1 | tail -2;;
c3) timeout 500 python tools/config3_bench.py > gpurun_out/config3_r2.json 2> gpurun_out/config3_r2.log || tail -5 gpurun_out/config3_r2.log; cat gpurun_out/config3_r2.json;;
smoke) python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -3;;
seek) ZK_SEEK_SINGLE=150 python tools/seek_bench.py > gpurun_out/seek_r2.json 2> gpurun_out/seek_r2.log || tail -3 gpurun_out/seek_r2.log; cat gpurun_out/seek_r2.json;;
esac; done
ls -la gpurun_out | tail -6

#!/bin/bash
# GPU-box helper: the evidence committed under profiles/ (bench line, reference arm, ncu launch list, ncu full capture)
mkdir -p gpurun_out
python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log > gpurun_out/bench_final_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_final_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_final.csv python bench.py --steps 1 --warmup 1 > gpurun_out/final_ncu_bench.log 2>&1
ZK_PROF_REPS=1 timeout 1000 ncu --set full --import-source on --clock-control none -k regex:"zk_(exec|match|seq_enc|lit_enc|huf|seq)_kernel" -c 14 -o gpurun_out/ncu_final --force-overwrite python tools/prof_codec.py 2>&1 | tail -2
python tools/seek_bench.py > gpurun_out/seek_final.json 2> gpurun_out/seek_final.log || true
ls -la gpurun_out | tail -8

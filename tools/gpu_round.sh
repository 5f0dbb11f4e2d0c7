#!/bin/bash
# GPU-box helper: bench + parity tests + one ncu capture of the hot kernels.  usage: tools/gpu_round.sh <tag> [ncu]
tag=${1:-x}
mkdir -p gpurun_out
python bench.py 2>&1 | tail -1 > gpurun_out/bench_$tag.json
python - <<PY
import json;d=json.load(open("gpurun_out/bench_$tag.json"));print(d["value"],d["compress_GiBps"],d["decompress_GiBps"],d["e2e"]["compress_GiBps"],d["e2e"]["decompress_GiBps"],d["roofline"]["kernel_ms"])
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
if [ "$2" = "ncu" ]; then
  ZK_PROF_REPS=1 timeout 1000 ncu --set full --import-source on --clock-control none -k regex:"zk_(exec|match|seq_enc|lit_enc|huf|seq)_kernel" -c 12 -o gpurun_out/ncu_$tag --force-overwrite python tools/prof_codec.py 2>&1 | tail -2
fi

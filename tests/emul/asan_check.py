"""TEST INFRASTRUCTURE: the CUDA sources compiled for the CPU emulation (tests/emul/cuda_emul.h) with
-fsanitize=address,undefined, driven over the shared test bodies.  Stands in for compute-sanitizer memcheck in this
GPU-less container (out-of-bounds shared/global accesses, misaligned loads, signed overflow, bad shifts).

  ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \\
  LD_PRELOAD=/usr/lib/x86_64-linux-gnu/libasan.so.8:/usr/lib/x86_64-linux-gnu/libubsan.so.1 \\
  CXX=/usr/bin/g++ python tests/emul/asan_check.py [seconds of random cases]
"""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import zeekstd_b200 as zk
from zeekstd_b200 import _native as N, corpus
from zeekstd_b200.build import build_emul
import cases

lib = N.load(build_emul(sanitize=True)); ctx = zk.Context(0, lib)
cases.check_golden_archives(ctx)
x = corpus.make_class("text", 70_000, 9).numpy()
for n, fs in ((0, 100), (1, 100), (15, 100), (16, 16), (31, 100), (100, 100), (101, 100), (32_767, 1 << 20), (32_768, 32_768), (32_769, 65_536),
              (65_537, 1 << 20), (3000, 1), (3000, 7)):
    cases.check_compress_roundtrip(ctx, x[:n], fs, 3, n % 2 == 0)
rng = np.random.default_rng(3)
cases.check_compress_roundtrip(ctx, rng.integers(0, 64, 80_000, dtype=np.uint8), 1 << 20, 1, False)      # tiled Huffman packer
cases.check_range_reads_stop_early(ctx, n=400_000, frame_size=200_000, reads=8)
cases.check_corruption_is_detected(ctx, trials=8)          # includes the crafted frames of the round-1 advisory
cases.check_special_entries(ctx); cases.check_patch_cycle(ctx); cases.check_prefix_batches(ctx, n=120_000)
cases.check_cycle_tiny_buffers(ctx); cases.check_decoder_state_machine(ctx); cases.check_libzstd_archive_through_decoder(ctx)
kinds = ["text", "structured", "lowent", "random", "runs"]
t0 = time.time(); it = 0
while time.time() - t0 < (float(sys.argv[1]) if len(sys.argv) > 1 else 60):
    parts = []; left = int(rng.integers(1, 200_000))
    while left > 0:
        m = min(left, int(rng.integers(1, 80_000))); parts.append(corpus.make_class(kinds[rng.integers(5)], m, int(rng.integers(1 << 30))).numpy()); left -= m
    d = np.concatenate(parts); fs = int(rng.choice([100, 4096, 32768, 40000, 100000, 1 << 20])); lv = int(rng.choice([1, 3])); ck = bool(rng.integers(2))
    cases.check_compress_roundtrip(ctx, d, fs, lv, ck); cases.check_decode_matches_libzstd(ctx, d, fs, lv, ck); it += 1
print("asan/ubsan clean;", it, "random cases")

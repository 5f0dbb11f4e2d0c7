"""TEST INFRASTRUCTURE: random inputs through the ENCODE kernels (every level tier, checksum on/off, ragged frame sizes, optional raw-content
prefix): libzstd must restore every frame byte for byte, and so must the decode kernels.  Runs on the sanitized emulation build (default) or,
with ZK_FUZZ_LIB=product, on a real GPU.    usage: python tests/emul/fuzz_encode.py [seconds=120] [seed=1]"""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import zeekstd_b200 as zk
from zeekstd_b200 import _native as N, corpus
from zeekstd_b200.build import build_emul
from oracle import oracle as O
from util import offsets

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
gpu = os.environ.get("ZK_FUZZ_LIB") == "product"
lib = N.load() if gpu else N.load(build_emul(sanitize=os.environ.get("ZK_FUZZ_SANITIZE", "1") == "1"))
ctx = zk.Context(0, lib)
kinds = ["text", "structured", "lowent", "random", "runs"]
big = 3_000_000 if gpu else 150_000
t0 = time.time(); n_cases = n_bytes = 0
while time.time() - t0 < secs:
    parts = []; left = int(rng.integers(0, big)) if rng.integers(6) else int(rng.integers(0, 40))
    while left > 0:
        m = min(left, int(rng.integers(1, 200_000))); parts.append(corpus.make_class(kinds[rng.integers(5)], m, int(rng.integers(1 << 30))).numpy()); left -= m
    d = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    fs = int(rng.choice([1, 7, 100, 4096, 32767, 32768, 32769, 65536, 100_000, 1 << 20, 1 << 21, int(rng.integers(1, 300_000))]))
    if d.size // fs > 3000:
        fs = max(fs, d.size // 3000 + 1)
    lvl = int(rng.choice([1, 2, 3, 4, 7, 10, 13, 19])); ck = bool(rng.integers(2))
    pfx = None
    if rng.integers(4) == 0:
        pfx = corpus.make_class(kinds[rng.integers(3)], int(rng.integers(1, 100_000)), int(rng.integers(1 << 30))).numpy()
        if d.size > 2000 and pfx.size > 1000:
            d = d.copy(); k = min(d.size // 2, pfx.size - 500); d[:k] = pfx[-k - 300: -300]          # something worth finding in the prefix
    comp, cs, ds = ctx.compress_frames(d, fs, lvl, ck, prefix=pfx)
    nf = max(1, -(-d.size // fs))
    assert len(cs) == nf and int(cs.sum()) == comp.size and int(ds.sum()) == d.size, (d.size, fs, lvl)
    out, sizes = O.ref_decompress_frames(comp, offsets(cs), offsets(ds), threads=8, prefix=pfx)
    assert list(sizes) == [int(x) for x in ds] and np.array_equal(out[: d.size], d), ("libzstd does not restore", d.size, fs, lvl, ck, None if pfx is None else pfx.size)
    back, st, rc = ctx.decompress_frames(comp, offsets(cs), offsets(ds), True, prefix=pfx)
    assert rc == 0 and np.array_equal(back[: d.size], d), ("own decode", rc, d.size, fs, lvl, ck)
    n_cases += 1; n_bytes += d.size
print(f"encode fuzz clean: {n_cases} inputs, {n_bytes} bytes, levels 1/2/3/4/7/10/13/19, prefix on a quarter")

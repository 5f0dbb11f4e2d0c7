"""GPU-box: find which frames of large single-class inputs fail to round-trip, and whether the archive or the decode is wrong."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zeekstd_b200 as zk
from zeekstd_b200 import corpus
from oracle import oracle as O
ctx = zk.Context(0)
kind = sys.argv[1] if len(sys.argv) > 1 else "lowent"
FS = 2 << 20
saved = 0
for seed in [int(a) for a in sys.argv[2:]] or [7, 263, 519, 775]:
    x = corpus.make_class(kind, 256 << 20, seed=seed, device="cuda").cpu().numpy()
    for rep in range(2):
        comp, cs, ds = ctx.compress_frames(x, FS, 1, False)
        co = np.concatenate([[0], np.cumsum(cs)]).astype(np.uint64); do = np.concatenate([[0], np.cumsum(ds)]).astype(np.uint64)
        out, sizes = O.ref_decompress_frames(comp, co, do, threads=64)
        bad_ref = [i for i, s in enumerate(sizes) if s != int(ds[i]) or out[int(do[i]):int(do[i + 1])].tobytes() != x[int(do[i]):int(do[i + 1])].tobytes()]
        back, st, rc = ctx.decompress_frames(np.concatenate([comp, np.zeros(64, np.uint8)]), co, do, True)
        bad_gpu = [i for i in range(len(cs)) if st[i] != 0 or back[int(do[i]):int(do[i + 1])].tobytes() != x[int(do[i]):int(do[i + 1])].tobytes()]
        print(json.dumps({"kind": kind, "seed": seed, "rep": rep, "alphabet": int(len(np.unique(x[:100000]))), "ratio": round(x.size / comp.size, 3), "rc": int(rc),
                          "bad_libzstd": bad_ref[:8], "bad_gpu": bad_gpu[:8], "status": [int(st[i]) for i in bad_gpu[:8]]}), flush=True)
        for i in (bad_ref + bad_gpu)[:1]:
            if saved < 2:
                x[int(do[i]):int(do[i + 1])].tofile(f"gpurun_out/bad_{kind}_{seed}_{i}.bin"); comp[int(co[i]):int(co[i + 1])].tofile(f"gpurun_out/bad_{kind}_{seed}_{i}.zst"); saved += 1

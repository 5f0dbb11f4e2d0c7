"""GPU-box: device-resident round trip of a large single-class input (zk_*_frames_dev, one 1 GiB sub-batch): which side fails?"""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zeekstd_b200 as zk
from zeekstd_b200 import corpus, parallel, _native as N
from oracle import oracle as O
ctx = zk.Context(0); codec = parallel.DeviceCodec(ctx); lib = ctx.lib
kind = sys.argv[1] if len(sys.argv) > 1 else "lowent"
g = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
FS = 2 << 20
nb = int(g * 2**30)
x = torch.cat([corpus.make_class(kind, min(256 << 20, nb - o), seed=7 + (o >> 20), device="cuda") for o in range(0, nb, 256 << 20)])
for rep in range(3):
    comp, cs, ds = codec.compress(x, FS, 1, False)
    co = np.concatenate([[0], np.cumsum(cs)]).astype(np.uint64); do = np.concatenate([[0], np.cumsum(ds)]).astype(np.uint64)
    hc = comp.cpu().numpy(); hx = x.cpu().numpy()
    out, sizes = O.ref_decompress_frames(hc, co, do, threads=64)
    bad_ref = [i for i, s in enumerate(sizes) if s != int(ds[i]) or out[int(do[i]):int(do[i + 1])].tobytes() != hx[int(do[i]):int(do[i + 1])].tobytes()]
    src = torch.cat([comp, torch.zeros(64, dtype=torch.uint8, device="cuda")]); dst = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
    st = np.zeros(len(cs), dtype=np.int32)
    rc = lib.zk_decompress_frames_dev(ctx._h, src.data_ptr(), co.ctypes.data_as(N.u64p), do.ctypes.data_as(N.u64p), len(cs), dst.data_ptr(), 1, st.ctypes.data_as(N.i32p), None)
    hb = dst[:nb].cpu().numpy()
    bad_gpu = [i for i in range(len(cs)) if st[i] != 0 or hb[int(do[i]):int(do[i + 1])].tobytes() != hx[int(do[i]):int(do[i + 1])].tobytes()]
    print(json.dumps({"kind": kind, "rep": rep, "frames": len(cs), "ratio": round(nb / int(co[-1]), 3), "rc": int(rc), "n_bad_libzstd": len(bad_ref), "bad_libzstd": bad_ref[:6],
                      "n_bad_gpu": len(bad_gpu), "bad_gpu": bad_gpu[:6], "status": [int(st[i]) for i in bad_gpu[:6]]}), flush=True)

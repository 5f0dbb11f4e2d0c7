#!/bin/bash
# GPU-box helper: host-pointer pipeline shapes (slots x sub-batch bytes x share hint x stream priorities)
cat > /tmp/e2e_one.py <<'PY'
import ctypes, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import zeekstd_b200 as zk
from zeekstd_b200 import corpus, _native as N
lib = N.load(); ctx = zk.Context(0, lib)
n = 1 << 30; FRAME = 2 << 20
x = corpus.make_mix(n, seed=20260924, device="cuda")
h_src = torch.empty(n, dtype=torch.uint8).pin_memory(); h_src.copy_(x.cpu())
cap = lib.zk_compress_bound(n, FRAME)
h_comp = torch.empty(cap + 64, dtype=torch.uint8).pin_memory(); h_back = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
nfmax = n // FRAME + 2; cs = np.zeros(nfmax, dtype=np.uint32); ds = np.zeros(nfmax, dtype=np.uint32); nf = ctypes.c_uint32(); dl = ctypes.c_size_t()
def comp():
    rc = lib.zk_compress_frames(ctx._h, h_src.data_ptr(), n, FRAME, 1, 0, h_comp.data_ptr(), cap, cs.ctypes.data_as(N.u32p), ds.ctypes.data_as(N.u32p), nfmax, ctypes.byref(nf), ctypes.byref(dl)); assert rc == 0
def dec():
    k = nf.value; co = np.zeros(k + 1, dtype=np.uint64); co[1:] = np.cumsum(cs[:k]); do = np.zeros(k + 1, dtype=np.uint64); do[1:] = np.cumsum(ds[:k])
    rc = lib.zk_decompress_frames(ctx._h, h_comp.data_ptr(), co.ctypes.data_as(N.u64p), do.ctypes.data_as(N.u64p), k, h_back.data_ptr(), 0, None); assert rc == 0
for cfg in sys.argv[1:]:
    slots, sub, sube, share = cfg.split(",")
    os.environ["ZK_HOST_SLOTS"] = slots; os.environ["ZK_HOST_SUB_BYTES"] = str(int(sub) << 20); os.environ["ZK_HOST_SUB_BYTES_ENC"] = str(int(sube) << 20)
    os.environ["ZK_HOST_SHARE"] = share
    comp(); dec(); comp(); dec()
    tc = td = 1e9
    for _ in range(4):
        t = time.perf_counter(); comp(); tc = min(tc, time.perf_counter() - t)
        t = time.perf_counter(); dec(); td = min(td, time.perf_counter() - t)
    print(json.dumps(dict(prio=os.environ.get("ZK_HOST_PRIO", "0"), slots=slots, sub_dec=sub, sub_enc=sube, share=share, comp_ms=round(tc * 1e3, 1), dec_ms=round(td * 1e3, 1))), flush=True)
assert bytes(h_back[:n].numpy()[:1 << 20]) == bytes(h_src.numpy()[:1 << 20])
PY
CFG="4,256,64,4 2,256,64,1 3,256,128,1 3,256,128,2 4,128,64,2 4,128,32,4 6,128,64,3 6,64,32,6 8,64,32,4 8,128,64,4 8,64,64,8 6,128,128,2"
for p in 0 1; do ZK_HOST_PRIO=$p python /tmp/e2e_one.py $CFG; done

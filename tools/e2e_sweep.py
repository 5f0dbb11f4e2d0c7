"""GPU-box tuning helper: PCIe bandwidth and the host-pointer (e2e) paths for several pipeline shapes."""
import ctypes, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zeekstd_b200 as zk
from zeekstd_b200 import corpus, _native as N
lib = N.load(); ctx = zk.Context(0, lib)
n = 1 << 30; FRAME = 2 << 20
x = corpus.make_mix(n, seed=20260924, device="cuda")
h_src = torch.empty(n, dtype=torch.uint8).pin_memory(); h_src.copy_(x.cpu())
# PCIe
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h_src, non_blocking=True)), ("d2h", lambda: h_src.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, "GB/s", round(3 * n / (time.perf_counter() - t) / 1e9, 1), flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(h_src, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("bidir GB/s each", round(3 * n / (time.perf_counter() - t) / 1e9, 1), flush=True)
h_src.copy_(x.cpu())
cap = lib.zk_compress_bound(n, FRAME)
h_comp = torch.empty(cap + 64, dtype=torch.uint8).pin_memory(); h_back = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
nfmax = n // FRAME + 2; cs = np.zeros(nfmax, dtype=np.uint32); ds = np.zeros(nfmax, dtype=np.uint32); nf = ctypes.c_uint32(); dl = ctypes.c_size_t()
def comp():
    rc = lib.zk_compress_frames(ctx._h, h_src.data_ptr(), n, FRAME, 1, 0, h_comp.data_ptr(), cap, cs.ctypes.data_as(N.u32p), ds.ctypes.data_as(N.u32p), nfmax, ctypes.byref(nf), ctypes.byref(dl)); assert rc == 0
def dec():
    k = nf.value; co = np.zeros(k + 1, dtype=np.uint64); co[1:] = np.cumsum(cs[:k]); do = np.zeros(k + 1, dtype=np.uint64); do[1:] = np.cumsum(ds[:k])
    rc = lib.zk_decompress_frames(ctx._h, h_comp.data_ptr(), co.ctypes.data_as(N.u64p), do.ctypes.data_as(N.u64p), k, h_back.data_ptr(), 0, None); assert rc == 0
for slots in (3, 4, 6):
    for sub in (32, 64, 128, 256):
        os.environ["ZK_HOST_SLOTS"] = str(slots); os.environ["ZK_HOST_SUB_BYTES"] = str(sub << 20)
        comp(); dec(); comp(); dec()
        tc = td = 1e9
        for _ in range(3):
            t = time.perf_counter(); comp(); tc = min(tc, time.perf_counter() - t)
            t = time.perf_counter(); dec(); td = min(td, time.perf_counter() - t)
        print(json.dumps(dict(slots=slots, sub_mib=sub, comp_GiBps=round(1 / tc, 2), dec_GiBps=round(1 / td, 2))), flush=True)

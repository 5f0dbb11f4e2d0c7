"""GPU-box helper: ZK_E2E_TRACE timeline of one host-pointer compress and decompress call (per sub-batch: H2D / kernels / D2H)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for _ in range(3): comp(); dec()
t = time.perf_counter(); comp(); tc = time.perf_counter() - t
t = time.perf_counter(); dec(); td = time.perf_counter() - t
print(f"untraced: comp {tc*1e3:.2f} ms  dec {td*1e3:.2f} ms", flush=True)
os.environ["ZK_E2E_TRACE"] = "1"
t = time.perf_counter(); comp(); tc = time.perf_counter() - t
t = time.perf_counter(); dec(); td = time.perf_counter() - t
print(f"traced: comp {tc*1e3:.2f} ms  dec {td*1e3:.2f} ms", flush=True)

"""GPU-box probe: device-resident decode/encode kernel times per corpus class (which data the exec kernel is slow on).
usage: python tools/class_probe.py [GiB per class] [classes...]   (env: ZK_* knobs, LVL, CK)"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from zeekstd_b200 import corpus
g = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
kinds = sys.argv[2:] or ["text", "lowent", "structured", "runs", "random", "mix"]
rig = bench.Rig(0)
lvl, ck = int(os.environ.get("LVL", "1")), os.environ.get("CK", "0") == "1"
for kind in kinds:
    nb = int(g * 2**30)
    if kind == "mix":
        x = bench.gen_mix(nb, bench.SEED, device=rig.dev)
    else:
        x = torch.cat([corpus.make_class(kind, min(256 << 20, nb - o), seed=7 + (o >> 20), device=rig.dev) for o in range(0, nb, 256 << 20)])
    rig.lib.zk_ctx_profile(rig.ctx._h, 1)
    c, d, clen = rig.timed_device(x, lvl, ck, 1, 2)
    kms = (ctypes.c_float * 8)(); kcnt = (ctypes.c_uint32 * 8)()
    rig.lib.zk_ctx_profile_read(rig.ctx._h, kms, kcnt)
    rig.lib.zk_ctx_profile(rig.ctx._h, 0)
    print(json.dumps({"class": kind, "GiB": g, "ratio": round(nb / clen, 3), "compress_GiBps": round(g / (c / 1e3), 2), "decompress_GiBps": round(g / (d / 1e3), 2),
                      "ms_per_GiB": {bench.KERNEL_NAMES[i].replace("zk_", "").replace("_kernel", ""): round(float(kms[i]) / max(1, int(kcnt[i])) / g, 2) for i in range(8) if kcnt[i]}}), flush=True)
    del x
    torch.cuda.empty_cache()

"""GPU-box probe: device-resident level-3 + checksum round trip of the configs[3] mix at several device sub-batch sizes
(how much codec efficiency a pipeline chunk of the multi-GPU path gives up).  usage: python tools/c4_probe.py [GiB]"""
import ctypes, json, os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2:
    import numpy as np, torch
    import bench
    from zeekstd_b200 import corpus
    rig = bench.Rig(0)
    nb = int(float(sys.argv[1]) * 2**30)
    x = bench.gen_mix(nb, bench.C4_SEED, corpus.CLASS_MIX_MIXED, device=rig.dev)
    lvl, ck = int(os.environ.get("LVL", "3")), os.environ.get("CK", "1") == "1"
    rig.lib.zk_ctx_profile(rig.ctx._h, 1)
    c, d, clen = rig.timed_device(x, lvl, ck, 2, 3)
    kms = (ctypes.c_float * 8)(); kcnt = (ctypes.c_uint32 * 8)()
    rig.lib.zk_ctx_profile_read(rig.ctx._h, kms, kcnt)
    print(json.dumps({"sub": os.environ.get("ZK_DEV_SUB_BYTES"), "level": lvl, "ck": ck, "compress_GiBps": round(nb / 2**30 / (c / 1e3), 2), "decompress_GiBps": round(nb / 2**30 / (d / 1e3), 2),
                      "ms_per_GiB": {bench.KERNEL_NAMES[i]: round(float(kms[i]) / 3 / (nb / 2**30), 2) for i in range(8) if kcnt[i]}}), flush=True)
else:
    g = sys.argv[1] if len(sys.argv) > 1 else "2"
    for lvl, ck in ((3, "1"), (3, "0"), (1, "0")):
        for sub in (256 << 20, 512 << 20, 1 << 30, 2 << 30):
            subprocess.run([sys.executable, __file__, g, "x"], env=dict(os.environ, ZK_DEV_SUB_BYTES=str(sub), LVL=str(lvl), CK=ck))

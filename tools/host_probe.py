"""GPU-box: host-pointer round trip (zk_compress_frames / zk_decompress_frames) with timings.  usage: python tools/host_probe.py [GiB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
g = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rig = bench.Rig(0)
x = bench.gen_mix(int(g * 2**30), bench.SEED, device=rig.dev)
print("generated", flush=True)
ec, ed, clen = rig.timed_host(x, int(os.environ.get("LVL", "1")), os.environ.get("CK", "0") == "1", int(os.environ.get("REPS", "3")))
print({"e2e_compress_GiBps": round(g / ec, 2), "e2e_decompress_GiBps": round(g / ed, 2)}, flush=True)

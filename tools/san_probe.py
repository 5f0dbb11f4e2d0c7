"""GPU-box: small device-path round trip for compute-sanitizer runs.  usage: python tools/san_probe.py [MiB] [class]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from zeekstd_b200 import corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "mix"
rig = bench.Rig(0)
x = bench.gen_mix(mib << 20, bench.SEED, device=rig.dev) if kind == "mix" else corpus.make_class(kind, mib << 20, 7, device=rig.dev)
step = rig.device_step_fn(x, int(os.environ.get("LVL", "1")), os.environ.get("CK", "0") == "1")
for i in range(int(os.environ.get("REPS", "2"))):
    a, b, c = step()
    torch.cuda.synchronize()
    print("step", i, round(a, 2), round(b, 2), bool(torch.equal(step.back[: x.numel()], x)), flush=True)

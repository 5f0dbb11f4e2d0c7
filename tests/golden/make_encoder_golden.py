"""Writes tests/golden/encoder_golden.json: size + SHA-256 of what OUR encoder emits for tests/golden/dickens_96k.txt (40 000-byte frames, checksum on)
at one level of every tier, and for prefix mode -- from the CPU emulation build of the sources (tests/emul), whose output does not depend on the
warp-scheduling seed (ZK_EMUL_SEED).  The GPU suite asserts that the nvcc build reproduces these bytes (tests/cases.py check_encoder_golden).
Re-run after any intended change of the encoder's output:   python tests/golden/make_encoder_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from zeekstd_b200 import _native
from zeekstd_b200.build import build_emul
from util import make_ctx

lib = _native.load(build_emul()); ctx = make_ctx(lib)
d = np.frombuffer(open(os.path.join(HERE, "dickens_96k.txt"), "rb").read(), dtype=np.uint8)
out = {}
for lvl in (1, 3, 4, 7, 10, 13):
    comp, cs, ds = ctx.compress_frames(d, 40_000, lvl, True)
    out[str(lvl)] = {"size": int(comp.size), "sha256": hashlib.sha256(comp.tobytes()).hexdigest()}
comp, cs, ds = ctx.compress_frames(d[20_000:], 40_000, 3, True, prefix=d[:30_000])
out["3+prefix"] = {"size": int(comp.size), "sha256": hashlib.sha256(comp.tobytes()).hexdigest()}
json.dump(out, open(os.path.join(HERE, "encoder_golden.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

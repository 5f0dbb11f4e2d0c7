"""torchrun target for the GPU box: NCCL scatter / all-gather / gather around the device codec (SURVEY.md 8e, config 4 shape)."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zeekstd_b200 as zk
from zeekstd_b200 import corpus, parallel
from oracle import oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = zk.Context(local); codec = parallel.DeviceCodec(ctx)
n = int(os.environ.get("ZK_MG_BYTES", str(2 << 30))); fs = 2 << 20
x = corpus.make_mix(n, seed=20260925, mix=corpus.CLASS_MIX_MIXED, device=dev) if rank == 0 else None
for it in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    frames, cs, ds = parallel.sharded_compress(codec, x, n, fs, 3, True, device=dev)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    back = parallel.sharded_decompress(codec, frames, cs, ds, True, device=dev)
    torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
if rank == 0:
    ok = torch.equal(back, x)
    c_off = np.concatenate([[0], np.cumsum(cs)]); d_off = np.concatenate([[0], np.cumsum(ds)])
    sub = slice(0, 64)
    out, sizes = O.ref_decompress_frames(frames[: int(c_off[64])].cpu().numpy(), c_off[:65], d_off[:65], threads=32)
    ok_ref = out.tobytes() == x[: int(d_off[64])].cpu().numpy().tobytes()
    print(json.dumps(dict(world=world, bytes=n, frames=len(cs), level=3, checksum=True, roundtrip_ok=bool(ok), libzstd_ok=bool(ok_ref), ratio=round(n / int(c_off[-1]), 3),
                          scatter_compress_gather_GiBps=round(n / 2**30 / (t1 - t0), 2), scatter_decompress_gather_GiBps=round(n / 2**30 / (t2 - t1), 2))), flush=True)
dist.destroy_process_group()

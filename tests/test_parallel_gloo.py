"""N > 1 host logic on CPU: world_size-2 gloo run of the frame sharding (scatter -> local codec -> all-gather of sizes ->
variable gather), with the emulated codec build on tiny inputs; the root's archive must be what libzstd decodes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, frame_size, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zeekstd_b200 as zk
    from zeekstd_b200 import _native, corpus, parallel
    from zeekstd_b200.build import build_emul
    lib = _native.load(build_emul()); _native.set_default_lib(lib)
    ctx = zk.Context(0, lib)
    codec = parallel.HostCodec(ctx)
    x = corpus.make_class("text", n_total, 4) if rank == 0 else None
    frames, cs, ds = parallel.sharded_compress(codec, x, n_total, frame_size, 3, True)
    back = parallel.sharded_decompress(codec, frames, cs, ds, True)
    if rank == 0:
        from oracle import oracle as O
        c_off = np.concatenate([[0], np.cumsum(cs)]); d_off = np.concatenate([[0], np.cumsum(ds)])
        out, sizes = O.ref_decompress_frames(frames.numpy(), c_off, d_off)
        ok = out.tobytes() == x.numpy().tobytes() and back.numpy().tobytes() == x.numpy().tobytes() and int(ds.sum()) == n_total
        q.put((ok, len(cs), [int(v) for v in ds]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total,frame_size,expect_frames", [(45_000, 10_000, 5), (30_000, 30_000, 1), (0, 1000, 1)])
def test_sharded_roundtrip_world2(n_total, frame_size, expect_frames):
    from zeekstd_b200.build import build_emul
    build_emul()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, n_total, frame_size, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, nf, ds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok and nf == expect_frames
    assert ds == [min(frame_size, n_total - i * frame_size) for i in range(expect_frames)] if n_total else ds == [0]


def test_sharded_roundtrip_world3_with_idle_rank():
    """2 frames over 3 ranks: the last rank owns no frame and must still take part in every exchange"""
    from zeekstd_b200.build import build_emul
    build_emul()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_worker, args=(r, 3, port, 15_000, 10_000, q)) for r in range(3)]
    for p in procs:
        p.start()
    ok, nf, ds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok and nf == 2 and ds == [10_000, 5_000]


def test_frame_ranges():
    from zeekstd_b200.parallel import frame_ranges
    assert frame_ranges(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert frame_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert frame_ranges(0, 2) == [(0, 0), (0, 0)]
    assert frame_ranges(8192, 8)[-1] == (7168, 8192)


def _leg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import zeekstd_b200 as zk
    from zeekstd_b200 import _native, corpus, parallel
    from zeekstd_b200.build import build_emul
    lib = _native.load(build_emul()); _native.set_default_lib(lib)
    codec = parallel.HostCodec(zk.Context(0, lib))
    nb, frame = 52_345, 8_000
    xr = None
    if rank == 0:
        xr = codec.empty(nb)
        xr[:nb] = corpus.make_class("text", nb, 6); xr[nb:] = 0
    leg = bench.c4_host_leg(parallel, codec, xr, nb, frame, 3, "cpu", rank, 1, lambda: None, False)
    if rank == 0:
        q.put(leg)
    else:
        assert leg is None
    dist.barrier()
    dist.destroy_process_group()


def test_bench_configs3_host_leg_world2():
    """bench.py's N > 1 e2e leg (root-held host buffers <-> sharded passes) on CPU tensors over gloo: same code, tiny input"""
    from zeekstd_b200.build import build_emul
    build_emul()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + (os.getpid() + 777) % 2000
    procs = [ctxm.Process(target=_leg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    leg = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert leg is not None and leg["value"] >= 0 and leg["h2d_bytes_per_step"] == leg["d2h_bytes_per_step"] > 52_345 and leg["steps"] == 1

#!/usr/bin/env python
"""bench.py -- the hot path of rorosen/zeekstd (per-frame compress + decompress of a seekable archive) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a 1 GiB synthetic Silesia-like mix per GPU, 2 MiB frames, level 1.
One STEP = compress the buffer into a seekable archive AND decompress that archive again (one pass of the hot path
in both directions).  `value` = uncompressed GiB moved per second over the step (2 x buffer / step time), inputs
resident in HBM; `compress_GiBps` / `decompress_GiBps` split it.  `e2e` is the same step through the host-pointer
C ABI (zk_compress_frames / zk_decompress_frames: pinned host buffers, H2D + D2H inside the timed region).
N > 1: every rank owns one GPU and its own shard of frames (frames are independent, seekable_format.md:23-29);
the only exchange is an all-gather of the per-frame sizes that make up the global seek table ("scaling": "weak").

--impl reference times the reference's own CPU path: libzstd driven through zeekstd's call sequence
(oracle/libzstd_driver.c; the Rust crate itself cannot be built in this image -- no cargo, no network), the frames of
the same workload spread over every host thread; what one thread reaches (zeekstd itself is single-threaded) is
reported beside it as `single_thread`.  `roofline.traffic` comes from the committed ncu capture of this workload
(profiles/traffic_r1.json): a bench value is never taken under a profiler.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = 2 << 20
LEVEL = 1
WORKLOAD_BYTES = int(os.environ.get("ZK_BENCH_BYTES", str(1 << 30)))
METRIC = "GiB/s compress + decompress (2 MiB frames)"
KERNEL_NAMES = ["zk_scan_kernel", "zk_seq_kernel", "zk_huf_kernel", "zk_exec_kernel", "zk_xxh64_kernel", "zk_match_kernel",
                "zk_entropy_enc_kernel", "zk_frame_*_kernels"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU during the timed region (NVML; nvidia-smi semantics)"""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference(data: np.ndarray, threads: int, reps: int = 1):
    """the reference path on the host cores: libzstd through zeekstd's call sequence -> (GiB/s step, compress, decompress)"""
    from oracle import oracle as O
    best_c = best_d = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        frames, cs, ds = O.ref_compress_frames(data, FRAME, LEVEL, False, threads=threads)
        t1 = time.perf_counter()
        comp = np.frombuffer(b"".join(frames), dtype=np.uint8)
        co = np.zeros(len(cs) + 1, dtype=np.uint64); co[1:] = np.cumsum(cs)
        do = np.zeros(len(ds) + 1, dtype=np.uint64); do[1:] = np.cumsum(ds)
        t2 = time.perf_counter()
        out, sizes = O.ref_decompress_frames(comp, co, do, threads=threads)
        t3 = time.perf_counter()
        assert all(s == d for s, d in zip(sizes, ds))
        best_c, best_d = min(best_c, t1 - t0), min(best_d, t3 - t2)
    gib = data.size / 2**30
    return 2 * gib / (best_c + best_d), gib / best_c, gib / best_d, data.size / sum(cs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    import torch
    from zeekstd_b200 import corpus

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the same workload, generated on the CPU so that this arm needs no GPU.  Headline: every host thread (one
        # CCtx/DCtx per thread over disjoint frame ranges -- the most the path can use); zeekstd itself drives one thread,
        # which is reported beside it on a bounded sample.
        from oracle import oracle as O
        data = corpus.make_mix(WORKLOAD_BYTES, seed=20260924).numpy()
        W = max(args.warmup, 0); K = max(args.steps, 1)
        for _ in range(min(W, 2)):
            cpu_reference(data[: 256 << 20], ncores)
        t0 = time.perf_counter()
        vals = [cpu_reference(data, ncores) for _ in range(K)]
        dt = (time.perf_counter() - t0) / K
        v = max(x[0] for x in vals); vc = max(x[1] for x in vals); vd = max(x[2] for x in vals)
        one_bytes = min(WORKLOAD_BYTES, 128 << 20)
        one = cpu_reference(data[:one_bytes], 1)
        line = {"metric": METRIC, "value": round(v, 4), "unit": "GiB/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": round(dt * 1e3, 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference",
                "config": {"workload": "silesia-mix", "frame_size": FRAME, "level": LEVEL, "bytes_per_gpu": WORKLOAD_BYTES, "step": "compress+decompress"},
                "compress_GiBps": round(vc, 4), "decompress_GiBps": round(vd, 4), "ratio": round(vals[0][3], 4),
                "cpu_baseline": {"value": round(v, 4), "unit": "GiB/s", "cores": ncores, "kind": "reference",
                                 "sample": f"the whole {WORKLOAD_BYTES >> 20} MiB workload per step; libzstd {O.libzstd_version()} through the reference's call sequence "
                                           f"(oracle/libzstd_driver.c), frames spread over {ncores} host threads",
                                 "single_thread": {"cores": 1, "value": round(one[0], 4), "compress_GiBps": round(one[1], 4), "decompress_GiBps": round(one[2], 4),
                                                   "sample": f"{one_bytes >> 20} MiB", "note": "what zeekstd's own single-threaded Encoder/Decoder reaches"}},
                "e2e": {"value": round(v, 4), "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: zeekstd_b200 has no CPU fallback"}), flush=True)
        return 2
    import zeekstd_b200 as zk
    from zeekstd_b200 import _native as N
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = N.load()
    ctx = zk.Context(local, lib)
    dev = torch.device("cuda", local)

    n = WORKLOAD_BYTES
    x = corpus.make_mix(n, seed=20260924 + rank, device=dev)                     # synthetic Silesia-like mix, unique per rank
    src = torch.cat([x, torch.zeros(64, dtype=torch.uint8, device=dev)])
    cap = lib.zk_compress_bound(n, FRAME)
    comp = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
    back = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
    nfmax = n // FRAME + 2
    cs = np.zeros(nfmax, dtype=np.uint32); ds = np.zeros(nfmax, dtype=np.uint32)
    nf = ctypes.c_uint32(); dl = ctypes.c_size_t()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                # > 126 MB L2: written between steps

    def step_device():
        rc = lib.zk_compress_frames_dev(ctx._h, src.data_ptr(), n, FRAME, LEVEL, 0, comp.data_ptr(), cap, cs.ctypes.data_as(N.u32p),
                                        ds.ctypes.data_as(N.u32p), nfmax, ctypes.byref(nf), ctypes.byref(dl), None)
        assert rc == 0, rc
        t_c = ctx.last_device_ms
        k = nf.value
        co = np.zeros(k + 1, dtype=np.uint64); co[1:] = np.cumsum(cs[:k]); do = np.zeros(k + 1, dtype=np.uint64); do[1:] = np.cumsum(ds[:k])
        rc = lib.zk_decompress_frames_dev(ctx._h, comp.data_ptr(), co.ctypes.data_as(N.u64p), do.ctypes.data_as(N.u64p), k, back.data_ptr(), 0, None, None)
        assert rc == 0, rc
        return t_c, ctx.last_device_ms, int(co[-1])

    # warm-up (also grows the workspaces), then a correctness check of the whole step
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step_device()
    assert torch.equal(back[:n], x), "round trip mismatch"
    launches0 = ctx.kernel_launches
    lib.zk_ctx_profile(ctx._h, 1)
    sampler = ClockSampler(local); sampler.start()
    tc = td = 0.0
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(args.steps):
        flush.fill_(1); torch.cuda.synchronize()                                   # L2 flush between timed iterations
        a, b, clen = step_device()
        tc += a; td += b
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = sampler.result()
    launches = ctx.kernel_launches - launches0
    kms = (ctypes.c_float * 8)(); kcnt = (ctypes.c_uint32 * 8)()
    lib.zk_ctx_profile_read(ctx._h, kms, kcnt)
    lib.zk_ctx_profile(ctx._h, 0)
    t = torch.tensor([tc, td], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                                  # device time, max over ranks
        # the one real exchange of the path: every rank learns all frame sizes -> global seek table
        sizes = torch.from_numpy(cs[: nf.value].astype(np.int64)).to(dev)
        gathered = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(gathered, sizes)
    tc_ms, td_ms = float(t[0]) / args.steps, float(t[1]) / args.steps
    gib = n / 2**30
    step_ms = tc_ms + td_ms
    value = world * 2 * gib / (step_ms / 1e3)

    # ---- e2e through the host-pointer C ABI with pinned buffers (H2D + D2H inside the timed region)
    h_src = torch.empty(n, dtype=torch.uint8).pin_memory(); h_src.copy_(x.cpu())
    h_comp = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(n + 64, dtype=torch.uint8).pin_memory()

    def step_host():
        t0 = time.perf_counter()
        rc = lib.zk_compress_frames(ctx._h, h_src.data_ptr(), n, FRAME, LEVEL, 0, h_comp.data_ptr(), cap, cs.ctypes.data_as(N.u32p), ds.ctypes.data_as(N.u32p),
                                    nfmax, ctypes.byref(nf), ctypes.byref(dl))
        assert rc == 0, rc
        t1 = time.perf_counter()
        k = nf.value
        co = np.zeros(k + 1, dtype=np.uint64); co[1:] = np.cumsum(cs[:k]); do = np.zeros(k + 1, dtype=np.uint64); do[1:] = np.cumsum(ds[:k])
        t2 = time.perf_counter()
        rc = lib.zk_decompress_frames(ctx._h, h_comp.data_ptr(), co.ctypes.data_as(N.u64p), do.ctypes.data_as(N.u64p), k, h_back.data_ptr(), 0, None)
        assert rc == 0, rc
        t3 = time.perf_counter()
        return t1 - t0, t3 - t2, int(co[-1])

    for _ in range(2):
        step_host()
    assert bytes(h_back[:4096].numpy()) == bytes(h_src[:4096].numpy())
    if dist:
        dist.barrier()
    ec = ed = 0.0
    for _ in range(args.steps):
        a, b, clen_h = step_host()
        ec += a; ed += b
    te = torch.tensor([ec, ed], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    ec_s, ed_s = float(te[0]) / args.steps, float(te[1]) / args.steps
    assert torch.equal(h_back[:n], h_src), "host round trip mismatch"
    e2e_value = world * 2 * gib / (ec_s + ed_s)

    if rank != 0:
        return 0
    # ---- roofline of the dominant kernel: algorithmic bytes / its own CUDA-event duration
    peak, peak_src = peaks()
    per = [float(kms[i]) / max(1, int(kcnt[i])) for i in range(8)]
    dom = int(np.argmax(per))
    # DRAM traffic of the dominant kernel per launch: taken from the committed ncu capture of this exact workload (profiling
    # inside a timed run is not allowed); null for any other workload size
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r1.json")))
        if n == (1 << 30) and KERNEL_NAMES[dom] in tj["bytes_per_launch"]:
            traffic, traffic_src = tj["bytes_per_launch"][KERNEL_NAMES[dom]], tj["source"]
    except (OSError, ValueError, KeyError):
        pass
    alg_bytes = n + clen            # decode: read C + write D ; compress: read D + write C -- the same sum (seek-table sizes)
    achieved = alg_bytes / (per[dom] / 1e3) / 1e9 if per[dom] > 0 else 0.0
    # ---- CPU baseline on a bounded sample (rank 0, same run)
    sample = x[: min(n, 128 << 20)].cpu().numpy()
    cb = cpu_reference(sample, 1)
    cb_all = cpu_reference(x[: min(n, 1 << 30)].cpu().numpy(), ncores)
    from oracle import oracle as O
    line = {"metric": METRIC, "value": round(value, 3), "unit": "GiB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "silesia-mix (configs[1])", "frame_size": FRAME, "level": LEVEL, "bytes_per_gpu": n, "step": "compress+decompress",
                       "l2": "256 MiB buffer written between timed iterations; inputs (1 GiB) exceed L2", "exchange": "all_gather of frame sizes" if world > 1 else "none"},
            "compress_GiBps": round(world * gib / (tc_ms / 1e3), 3), "decompress_GiBps": round(world * gib / (td_ms / 1e3), 3),
            "ratio": round(n / clen, 4), "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 3), "unit": "GiB/s", "h2d_bytes_per_step": int(n + clen_h), "d2h_bytes_per_step": int(clen_h + n),
                    "compress_GiBps": round(world * gib / ec_s, 3), "decompress_GiBps": round(world * gib / ed_s, 3), "api": "zk_compress_frames + zk_decompress_frames (pinned host buffers)"},
            "roofline": {"bound": "hbm", "kernel": KERNEL_NAMES[dom], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel_ms": {KERNEL_NAMES[i]: round(per[i], 3) for i in range(8) if kcnt[i]},
                         "note": "the path is bound by serial entropy / match dependencies, not by HBM (SURVEY.md 8d)"},
            "cpu_baseline": {"value": round(cb_all[0], 4), "unit": "GiB/s", "cores": ncores, "kind": "reference", "compress_GiBps": round(cb_all[1], 4),
                             "decompress_GiBps": round(cb_all[2], 4), "ratio": round(cb_all[3], 4),
                             "sample": f"the {min(n, 1 << 30) >> 20} MiB workload, one pass; libzstd {O.libzstd_version()} through the reference's call sequence "
                                       f"(oracle/libzstd_driver.c), frames spread over {ncores} host threads",
                             "single_thread": {"cores": 1, "value": round(cb[0], 4), "compress_GiBps": round(cb[1], 4), "decompress_GiBps": round(cb[2], 4),
                                               "sample": f"{sample.size >> 20} MiB", "note": "what zeekstd's own single-threaded Encoder/Decoder reaches"}}}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

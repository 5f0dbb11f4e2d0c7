"""Frame sharding across GPUs: frames of the seekable format are independent
(seekable_format.md:23-29), so rank r of W owns a contiguous range of frames and no collective touches the codec itself.
The exchanges around it are plain torch.distributed calls (NCCL over NVLink on the GPU box, gloo in the CPU tests):

  compress:   root scatters contiguous input byte ranges -> every rank compresses its frames -> all-gather of the per-frame
              sizes (4 B/frame; this is what makes the global seek table) -> ranks send their compressed frames to the
              root, which places them at the scanned offsets and appends the seek table
  decompress: root scatters compressed ranges (by seek-table offsets) -> every rank decodes -> gather of the outputs

One process per GPU; `codec` is any object with compress(t, frame_size, level, checksum) -> (comp, c_sizes, d_sizes) and
decompress(comp, c_off, d_off, verify) -> out operating on uint8 tensors of the rank's device.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _native


def frame_ranges(n_frames: int, world: int):
    """contiguous frame ranges, ceil(N/W) per rank (rank order == frame order)"""
    per = -(-n_frames // world) if n_frames else 0
    return [(min(r * per, n_frames), min((r + 1) * per, n_frames)) for r in range(world)]


def torch_stream_handle(device) -> ctypes.c_void_p:
    """cudaStream_t of torch's current stream; torch's default stream is the legacy NULL stream, which the C ABI
    spells cudaStreamLegacy (0x1) because NULL there means "the context's own stream"."""
    h = torch.cuda.current_stream(device).cuda_stream
    return ctypes.c_void_p(h if h else 1)


class DeviceCodec:
    """zero-copy codec over CUDA tensors (zk_*_frames_dev).  The kernels are enqueued on torch's CURRENT stream so they
    are ordered after the tensor ops / NCCL receives that produced their inputs."""

    def __init__(self, ctx):
        self.ctx, self.lib = ctx, ctx.lib

    def compress(self, x: torch.Tensor, frame_size: int, level: int, checksum: bool):
        n = x.numel()
        src = torch.cat([x, torch.zeros(64, dtype=torch.uint8, device=x.device)])
        cap = self.lib.zk_compress_bound(n, frame_size)
        dst = torch.empty(cap + 64, dtype=torch.uint8, device=x.device)
        nfmax = n // frame_size + 2
        cs = np.zeros(nfmax, dtype=np.uint32); ds = np.zeros(nfmax, dtype=np.uint32)
        nf = ctypes.c_uint32(); dl = ctypes.c_size_t()
        rc = self.lib.zk_compress_frames_dev(self.ctx._h, src.data_ptr(), n, frame_size, level, int(checksum), dst.data_ptr(), cap,
                                             cs.ctypes.data_as(_native.u32p), ds.ctypes.data_as(_native.u32p), nfmax, ctypes.byref(nf),
                                             ctypes.byref(dl), torch_stream_handle(x.device))
        if rc:
            raise RuntimeError(f"zk_compress_frames_dev: {rc}")
        return dst[: dl.value], cs[: nf.value].astype(np.int64), ds[: nf.value].astype(np.int64)

    def decompress(self, comp: torch.Tensor, c_off, d_off, verify: bool):
        co = np.ascontiguousarray(c_off, dtype=np.uint64); do = np.ascontiguousarray(d_off, dtype=np.uint64)
        n = len(co) - 1
        src = torch.cat([comp, torch.zeros(64, dtype=torch.uint8, device=comp.device)])
        out = torch.empty(int(do[-1]) + 64, dtype=torch.uint8, device=comp.device)
        rc = self.lib.zk_decompress_frames_dev(self.ctx._h, src.data_ptr(), co.ctypes.data_as(_native.u64p), do.ctypes.data_as(_native.u64p), n,
                                               out.data_ptr(), int(verify), None, torch_stream_handle(comp.device))
        if rc:
            raise RuntimeError(f"zk_decompress_frames_dev: {rc}")
        return out[: int(do[-1])]


class HostCodec:
    """same interface over CPU tensors through the host-pointer entry points (used by the gloo tests)"""

    def __init__(self, ctx):
        self.ctx = ctx

    def compress(self, x, frame_size, level, checksum):
        comp, cs, ds = self.ctx.compress_frames(x.numpy(), frame_size, level, checksum)
        return torch.from_numpy(comp.copy()), cs.astype(np.int64), ds.astype(np.int64)

    def decompress(self, comp, c_off, d_off, verify):
        out, st, rc = self.ctx.decompress_frames(np.concatenate([comp.numpy(), np.zeros(64, dtype=np.uint8)]), c_off, d_off, verify)
        if rc:
            raise RuntimeError(f"zk_decompress_frames: {rc}")
        return torch.from_numpy(out.copy())


def _dev(t_like_device):
    return t_like_device


def sharded_compress(codec, x_root, n_total: int, frame_size: int, level: int = 1, checksum: bool = False, root: int = 0, device="cpu"):
    """-> on root: (frames tensor, c_sizes, d_sizes) for the WHOLE input; on other ranks: (None, c_sizes, d_sizes)"""
    rank, world = dist.get_rank(), dist.get_world_size()
    n_frames = max(1, -(-n_total // frame_size))
    ranges = frame_ranges(n_frames, world)
    lo, hi = ranges[rank]
    b0, b1 = min(lo * frame_size, n_total), min(hi * frame_size, n_total)
    # 1. scatter contiguous byte ranges (fixed size except the tail)
    if rank == root:
        reqs = []
        for r, (flo, fhi) in enumerate(ranges):
            if r == root:
                continue
            s0, s1 = min(flo * frame_size, n_total), min(fhi * frame_size, n_total)
            if s1 > s0:
                reqs.append(dist.isend(x_root[s0:s1].contiguous(), dst=r))
        mine = x_root[b0:b1]
        for q in reqs:
            q.wait()
    else:
        mine = torch.empty(b1 - b0, dtype=torch.uint8, device=device)
        if b1 > b0:
            dist.recv(mine, src=root)
    # 2. local compress (ranks past the end of a short input have no frames; an empty input is one empty frame on rank 0)
    if hi > lo or (n_total == 0 and rank == 0):
        comp, cs, ds = codec.compress(mine, frame_size, level, checksum)
    else:
        comp, cs, ds = torch.empty(0, dtype=torch.uint8, device=device), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    # 3. all-gather the per-frame sizes (padded to the per-rank maximum) -> global seek table on every rank
    per = max(1, max(h - l for l, h in ranges))
    pad = torch.full((2, per), -1, dtype=torch.int64, device=device)
    pad[0, : len(cs)] = torch.from_numpy(cs).to(device); pad[1, : len(ds)] = torch.from_numpy(ds).to(device)
    allp = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(allp, pad)
    c_all, d_all, counts = [], [], []
    for p in allp:
        p = p.cpu().numpy()
        k = int((p[0] >= 0).sum())
        counts.append(k); c_all.extend(p[0, :k].tolist()); d_all.extend(p[1, :k].tolist())
    c_all, d_all = np.array(c_all, dtype=np.int64), np.array(d_all, dtype=np.int64)
    # 4. variable-size gather of the compressed frames to the root at the scanned offsets
    starts = np.concatenate([[0], np.cumsum(c_all)])
    first = np.concatenate([[0], np.cumsum(counts)])
    if rank == root:
        out = torch.empty(int(starts[-1]), dtype=torch.uint8, device=device)
        for r in range(world):
            o0, o1 = int(starts[first[r]]), int(starts[first[r + 1]])
            if o1 == o0:
                continue
            if r == root:
                out[o0:o1] = comp
            else:
                dist.recv(out[o0:o1], src=r)
        return out, c_all, d_all
    if comp.numel():
        dist.send(comp.contiguous(), dst=root)
    return None, c_all, d_all


def sharded_decompress(codec, comp_root, c_sizes, d_sizes, verify: bool = True, root: int = 0, device="cpu"):
    """c_sizes / d_sizes known on every rank (the seek table). -> on root: the decompressed tensor"""
    rank, world = dist.get_rank(), dist.get_world_size()
    n_frames = len(c_sizes)
    c_off = np.concatenate([[0], np.cumsum(c_sizes)]).astype(np.int64); d_off = np.concatenate([[0], np.cumsum(d_sizes)]).astype(np.int64)
    ranges = frame_ranges(n_frames, world)
    lo, hi = ranges[rank]
    if rank == root:
        reqs = []
        for r, (flo, fhi) in enumerate(ranges):
            if r != root and c_off[fhi] > c_off[flo]:
                reqs.append(dist.isend(comp_root[int(c_off[flo]): int(c_off[fhi])].contiguous(), dst=r))
        mine = comp_root[int(c_off[lo]): int(c_off[hi])]
        for q in reqs:
            q.wait()
    else:
        mine = torch.empty(int(c_off[hi] - c_off[lo]), dtype=torch.uint8, device=device)
        if mine.numel():
            dist.recv(mine, src=root)
    if hi > lo:
        out_local = codec.decompress(mine, c_off[lo: hi + 1] - c_off[lo], d_off[lo: hi + 1] - d_off[lo], verify)
    else:
        out_local = torch.empty(0, dtype=torch.uint8, device=device)
    if rank == root:
        out = torch.empty(int(d_off[-1]), dtype=torch.uint8, device=device)
        for r, (flo, fhi) in enumerate(ranges):
            o0, o1 = int(d_off[flo]), int(d_off[fhi])
            if o1 == o0:
                continue
            if r == root:
                out[o0:o1] = out_local
            else:
                dist.recv(out[o0:o1], src=r)
        return out
    if out_local.numel():
        dist.send(out_local.contiguous(), dst=root)
    return None

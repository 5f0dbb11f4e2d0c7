"""Batched random-access reads (BASELINE config 5): many `set_offset(o); set_offset_limit(o+len); read`
requests served by decoding every touched frame ONCE, in one batch per `max_batch_bytes` of frames, instead of frame by
frame.  Semantics per request are those of Decoder (decode.rs:402-437): bytes [o, o+len) of the decompressed stream."""
from __future__ import annotations

import numpy as np


def read_ranges(ctx, archive: np.ndarray, c_off, d_off, offsets, length: int, max_batch_bytes: int = 1 << 30, verify: bool = False,
                scratch: np.ndarray | None = None):
    """archive: np.uint8 (frames; padded by >= 64 bytes), c_off/d_off: N+1 cumulative offsets (the seek table).
    scratch: optional (pinned) np.uint8 buffer of >= max_batch_bytes + 64 that receives each decoded batch.
    -> list of bytes objects, one per offset, plus the number of frames decoded"""
    c_off = np.asarray(c_off, dtype=np.uint64); d_off = np.asarray(d_off, dtype=np.uint64)
    offsets = np.asarray(offsets, dtype=np.uint64)
    f_lo = np.searchsorted(d_off, offsets, side="right") - 1
    f_hi = np.searchsorted(d_off, offsets + np.uint64(length - 1), side="right") - 1
    need = np.zeros(len(d_off) - 1, dtype=bool)
    for a, b in zip(f_lo, f_hi):
        need[a: b + 1] = True
    frames = np.nonzero(need)[0]
    out = [None] * len(offsets)
    # contiguous runs of needed frames, cut into batches
    i = 0
    while i < len(frames):
        j = i
        size = 0
        while j < len(frames) and (j == i or (frames[j] == frames[j - 1] + 1 and size + int(d_off[frames[j] + 1] - d_off[frames[j]]) <= max_batch_bytes)):
            size += int(d_off[frames[j] + 1] - d_off[frames[j]]); j += 1
        lo, hi = int(frames[i]), int(frames[j - 1]) + 1
        sel = np.nonzero((f_lo >= lo) & (f_lo < hi))[0]
        if len(sel):
            hi = max(hi, int(f_hi[sel].max()) + 1)          # a read that starts in this batch may end in the next frame
        co = c_off[lo: hi + 1] - c_off[lo]; do = d_off[lo: hi + 1] - d_off[lo]
        buf, st, rc = ctx.decompress_frames(archive[int(c_off[lo]): int(c_off[hi]) + 64], co, do, verify, out=scratch)
        if rc:
            raise RuntimeError(f"decode failed: {rc}")
        # hand out the slices of this batch right away (the scratch buffer is reused by the next batch)
        base = int(d_off[lo])
        for k in sel:
            o = int(offsets[k]) - base
            out[k] = buf[o: o + length].tobytes()
        i = j
    return out, int(need.sum())

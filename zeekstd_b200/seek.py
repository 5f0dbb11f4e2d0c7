"""Batched random-access reads (BASELINE config 5): many `set_offset(o); set_offset_limit(o+len); read`
requests served by decoding every touched frame ONCE, in one batch per `max_batch_bytes` of frames, instead of frame by
frame.  Semantics per request are those of Decoder (decode.rs:402-437): bytes [o, o+len) of the decompressed stream."""
from __future__ import annotations

import numpy as np


def read_ranges(ctx, archive: np.ndarray, c_off, d_off, offsets, length: int, max_batch_bytes: int = 1 << 30, verify: bool = False,
                scratch: np.ndarray | None = None, partial: bool = True):
    """archive: np.uint8 (frames; padded by >= 64 bytes), c_off/d_off: N+1 cumulative offsets (the seek table).
    scratch: optional (pinned) np.uint8 buffer that receives each decoded batch (a batch that does not fit gets a fresh
    buffer).  Reads are clipped at the end of the stream, like Decoder with offset_limit = min(o + length, size).
    partial: decode every touched frame only as far as the last byte some read wants of it (zk_decompress_frames_upto).
    -> list of bytes objects, one per offset, plus the number of frames decoded"""
    c_off = np.asarray(c_off, dtype=np.uint64); d_off = np.asarray(d_off, dtype=np.uint64)
    offsets = np.asarray(offsets, dtype=np.uint64)
    nfr = len(d_off) - 1
    total = int(d_off[-1]) if nfr > 0 else 0
    out = [b""] * len(offsets)
    if nfr <= 0 or length <= 0 or len(offsets) == 0:
        return out, 0
    ends = np.minimum(offsets + np.uint64(length), np.uint64(total))
    live = np.nonzero(offsets < np.uint64(total))[0]                       # a read at or past the end returns nothing
    f_lo = np.clip(np.searchsorted(d_off, offsets, side="right") - 1, 0, nfr - 1)
    f_hi = np.clip(np.searchsorted(d_off, np.maximum(ends, np.uint64(1)) - np.uint64(1), side="right") - 1, 0, nfr - 1)
    f_hi = np.maximum(f_hi, f_lo)
    # per frame: how many leading bytes some read needs
    want = np.zeros(nfr, dtype=np.uint64)
    for k in live:
        a, b = int(f_lo[k]), int(f_hi[k])
        if b > a:
            want[a:b] = d_off[a + 1: b + 1] - d_off[a:b]
        want[b] = max(int(want[b]), int(ends[k]) - int(d_off[b]))
    glue = np.zeros(nfr + 1, dtype=bool)                                   # glue[f]: some read starts before frame f and reaches into it
    for k in live:
        glue[int(f_lo[k]) + 1: int(f_hi[k]) + 1] = True
    frames = np.nonzero(want)[0]
    order = live[np.argsort(f_lo[live], kind="stable")]
    pos = 0                                                                # reads are handed out in frame order
    i = 0
    while i < len(frames):
        j = i
        size = 0
        # contiguous runs of needed frames, cut into batches of <= max_batch_bytes; a read never straddles two batches
        while j < len(frames) and (j == i or (frames[j] == frames[j - 1] + 1 and
                                              (glue[frames[j]] or size + int(d_off[frames[j] + 1] - d_off[frames[j]]) <= max_batch_bytes))):
            size += int(d_off[frames[j] + 1] - d_off[frames[j]]); j += 1
        lo, hi = int(frames[i]), int(frames[j - 1]) + 1
        co = c_off[lo: hi + 1] - c_off[lo]; do = d_off[lo: hi + 1] - d_off[lo]
        need = want[lo:hi].astype(np.uint32) if partial else None
        buf_out = scratch if (scratch is not None and scratch.size >= int(do[-1])) else None
        buf, st, rc = ctx.decompress_frames(archive[int(c_off[lo]): int(c_off[hi]) + 64], co, do, verify, out=buf_out, need=need)
        if rc:
            raise RuntimeError(f"decode failed: {rc}")
        base = int(d_off[lo])
        while pos < len(order) and int(f_lo[order[pos]]) < hi:
            k = order[pos]; pos += 1
            out[k] = buf[int(offsets[k]) - base: int(ends[k]) - base].tobytes()
        i = j
    return out, int(len(frames))

"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the zeekstd_b200 package).

ctypes access to the two CPU checkers built by oracle/Makefile into oracle/_ref/:

  libzk_ref.so     the REAL libzstd of the image driven through the reference's call sequence
                   (oracle/libzstd_driver.c; encode.rs:340-346,442-464; decode.rs:221-256)
  libzk_oracle.so  the plain-C restatement of the zstd frame decoder (oracle/zstd_oracle.c)

plus a pure-Python restatement of the seekable-format seek table
(lib/src/seek_table.rs:144-225, 513-525, 916-934, 967-1005; seekable_format.md:45-157).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
SYSTEM_LIBZSTD = "/usr/lib/x86_64-linux-gnu/libzstd.so.1"

SEEKABLE_MAGIC_NUMBER = 0x8F92EAB1          # lib.rs:52
SKIPPABLE_MAGIC_NUMBER = 0x184D2A5E         # seek_table.rs:89
SEEKABLE_MAX_FRAMES = 0x08000000            # lib.rs:54
SEEKABLE_MAX_FRAME_SIZE = 0x40000000        # lib.rs:58


def build(force: bool = False) -> None:
    """compile the checkers (gcc only; a few seconds)"""
    pairs = (("libzk_ref.so", "libzstd_driver.c"), ("libzk_oracle.so", "zstd_oracle.c"))
    missing = not all(os.path.exists(os.path.join(_REF, so)) for so, _ in pairs)
    stale = not missing and any(os.path.getmtime(os.path.join(_HERE, c)) > os.path.getmtime(os.path.join(_REF, so)) for so, c in pairs)
    if force or missing or stale:          # (a prebuilt pair that is merely older than a fresh checkout's sources still loads if make is absent)
        r = subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), capture_output=True)
        if r.returncode != 0 and (force or missing):
            raise RuntimeError("oracle build failed:\n" + r.stderr.decode(errors="replace"))


_ref = None
_orc = None

_u8p = ctypes.POINTER(ctypes.c_uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def ref_lib():
    global _ref
    if _ref is None:
        build()
        lib = ctypes.CDLL(os.path.join(_REF, "libzk_ref.so"))
        lib.zkr_open.argtypes = [ctypes.c_char_p]
        lib.zkr_version.restype = ctypes.c_char_p
        lib.zkr_compress_bound.restype = ctypes.c_size_t
        lib.zkr_compress_bound.argtypes = [ctypes.c_size_t]
        lib.zkr_compress_frames.restype = ctypes.c_int64
        lib.zkr_compress_frames.argtypes = [_u8p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                            _u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int64), ctypes.c_uint32,
                                            ctypes.c_int]
        lib.zkr_decompress_frames.restype = ctypes.c_int64
        lib.zkr_decompress_frames.argtypes = [_u8p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                              ctypes.c_uint32, _u8p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
        lib.zkr_decompress_any.restype = ctypes.c_int64
        lib.zkr_decompress_any.argtypes = [_u8p, ctypes.c_size_t, _u8p, ctypes.c_size_t]
        lib.zkr_compress_frames_prefix.restype = ctypes.c_int64
        lib.zkr_compress_frames_prefix.argtypes = lib.zkr_compress_frames.argtypes + [_u8p, ctypes.c_size_t]
        lib.zkr_decompress_frames_prefix.restype = ctypes.c_int64
        lib.zkr_decompress_frames_prefix.argtypes = lib.zkr_decompress_frames.argtypes + [_u8p, ctypes.c_size_t]
        lib.zkr_compress_simple.restype = ctypes.c_int64
        lib.zkr_compress_simple.argtypes = [_u8p, ctypes.c_size_t, _u8p, ctypes.c_size_t, ctypes.c_int]
        path = os.environ.get("ZK_LIBZSTD", SYSTEM_LIBZSTD)
        rc = lib.zkr_open(path.encode())
        if rc != 0:
            raise RuntimeError(f"cannot dlopen libzstd at {path} (rc={rc})")
        _ref = lib
    return _ref


def orc_lib():
    global _orc
    if _orc is None:
        build()
        lib = ctypes.CDLL(os.path.join(_REF, "libzk_oracle.so"))
        lib.zko_decompress.restype = ctypes.c_int64
        lib.zko_decompress.argtypes = [_u8p, ctypes.c_size_t, _u8p, ctypes.c_size_t, ctypes.c_int]
        lib.zko_xxh64.restype = ctypes.c_uint64
        lib.zko_xxh64.argtypes = [_u8p, ctypes.c_size_t, ctypes.c_uint64]
        lib.zko_frame_stats.restype = ctypes.c_int64
        lib.zko_frame_stats.argtypes = [_u8p, ctypes.c_size_t, ctypes.c_void_p]
        lib.zko_decompress_ex.restype = ctypes.c_int64
        lib.zko_decompress_ex.argtypes = [_u8p, ctypes.c_size_t, _u8p, ctypes.c_size_t, ctypes.c_int, _u8p, ctypes.c_size_t, ctypes.c_void_p]
        _orc = lib
    return _orc


def libzstd_version() -> str:
    return ref_lib().zkr_version().decode()


# ----------------------------------------------------------------------------- reference codec (libzstd)
def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf.view(np.uint8).reshape(-1))
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def ref_compress_frames(data, frame_size: int = 0x200000, level: int = 0, checksum: bool = False, threads: int = 1, prefix=None):
    """-> (list[bytes] frames, c_sizes, d_sizes) exactly as RawEncoder+FrameSizePolicy::Uncompressed would emit
    (encode.rs:311-354, 438-472, 528-544).  Empty input -> zero frames... except that the reference's
    finish() always ends one (possibly empty) frame (encode.rs:755-756); callers add that themselves."""
    lib = ref_lib()
    src = _as_u8(data)
    n = src.size
    frame_size = min(frame_size, SEEKABLE_MAX_FRAME_SIZE)
    n_frames = (n + frame_size - 1) // frame_size
    if n_frames == 0:
        return [], [], []
    slot = lib.zkr_compress_bound(min(frame_size, n)) + 64
    dst = np.empty(n_frames * slot, dtype=np.uint8)
    sizes = (ctypes.c_int64 * n_frames)()
    if prefix is not None:                       # compress_with_prefix: the same prefix for every frame (encode.rs:332-338)
        pf = _as_u8(prefix)
        rc = lib.zkr_compress_frames_prefix(_ptr(src), n, frame_size, level, int(checksum), _ptr(dst), slot, sizes, n_frames,
                                            threads, _ptr(pf), pf.size)
    else:
        rc = lib.zkr_compress_frames(_ptr(src), n, frame_size, level, int(checksum), _ptr(dst), slot, sizes, n_frames,
                                     threads)
    if rc != 0:
        raise RuntimeError(f"libzstd compress error {rc}")
    frames = [dst[i * slot: i * slot + sizes[i]].tobytes() for i in range(n_frames)]
    d_sizes = [min(frame_size, n - i * frame_size) for i in range(n_frames)]
    return frames, [int(s) for s in sizes], d_sizes


def ref_decompress_frames(comp, c_off, d_off, threads: int = 1, prefix=None):
    """decompress frames given cumulative offsets -> (np.uint8 output, per-frame sizes or -code)"""
    lib = ref_lib()
    src = _as_u8(comp)
    nf = len(c_off) - 1
    co = (ctypes.c_uint64 * (nf + 1))(*[int(x) for x in c_off])
    do = (ctypes.c_uint64 * (nf + 1))(*[int(x) for x in d_off])
    out = np.empty(max(1, int(d_off[-1])), dtype=np.uint8)
    sizes = (ctypes.c_int64 * max(nf, 1))()
    if prefix is not None:                       # decompress_with_prefix (decode.rs:211-214, 246-255)
        pf = _as_u8(prefix)
        lib.zkr_decompress_frames_prefix(_ptr(src), co, do, nf, _ptr(out), sizes, threads, _ptr(pf), pf.size)
    else:
        lib.zkr_decompress_frames(_ptr(src), co, do, nf, _ptr(out), sizes, threads)
    return out[: int(d_off[-1])], [int(sizes[i]) for i in range(nf)]


def ref_decompress_any(comp, cap: int):
    """decompress concatenated frames -> bytes, or raises ZstdError(code)"""
    lib = ref_lib()
    src = _as_u8(comp)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    r = lib.zkr_decompress_any(_ptr(src), src.size, _ptr(out), cap)
    if r < 0:
        raise ZstdError(int(-r))
    return out[:r].tobytes()


def ref_compress_simple(data, level: int = 3) -> bytes:
    """ZSTD_compress(): a one-shot frame (Frame_Content_Size present; Single_Segment when the content fits the window)"""
    lib = ref_lib()
    src = _as_u8(data)
    dst = np.empty(lib.zkr_compress_bound(src.size) + 64, dtype=np.uint8)
    r = lib.zkr_compress_simple(_ptr(src), src.size, _ptr(dst), dst.size, level)
    if r < 0:
        raise ZstdError(int(-r))
    return dst[:r].tobytes()


def skippable_frame(payload: bytes, nibble: int = 0) -> bytes:
    """a skippable frame (RFC 8878 3.1.2): magic 0x184D2A5? + size + payload"""
    return struct.pack("<II", 0x184D2A50 | (nibble & 15), len(payload)) + payload


class ZstdError(Exception):
    def __init__(self, code: int):
        super().__init__(f"zstd error code {code}")
        self.code = code


# ----------------------------------------------------------------------------- restated decoder
def oracle_decompress(comp, cap: int, verify_checksum: bool = True) -> bytes:
    lib = orc_lib()
    src = _as_u8(comp)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    r = lib.zko_decompress(_ptr(src), src.size, _ptr(out), cap, int(verify_checksum))
    if r < 0:
        raise ZstdError(int(-r))
    return out[:r].tobytes()


class SeqStats(ctypes.Structure):
    _fields_ = [("rep_idx", ctypes.c_uint64 * 4)] + [(n, ctypes.c_uint64) for n in ("rep_ll0", "overlap", "explicit_offsets", "max_offset", "prefix_matches")]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n in ("rep_ll0", "overlap", "explicit_offsets", "max_offset", "prefix_matches")}
        d.update({"rep1": int(self.rep_idx[0]), "rep2": int(self.rep_idx[1]), "rep3": int(self.rep_idx[2]), "rep1_minus_1": int(self.rep_idx[3])})
        return d


def oracle_decompress_ex(comp, cap: int, prefix=None, verify_checksum: bool = True):
    """restated decoder with an optional raw-content prefix -> (bytes, sequence statistics dict)"""
    lib = orc_lib()
    src = _as_u8(comp)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    ss = SeqStats()
    pf = _as_u8(prefix) if prefix is not None else None
    r = lib.zko_decompress_ex(_ptr(src), src.size, _ptr(out), cap, int(verify_checksum), _ptr(pf) if pf is not None else None,
                              pf.size if pf is not None else 0, ctypes.byref(ss))
    if r < 0:
        raise ZstdError(int(-r))
    return out[:r].tobytes(), ss.as_dict()


def oracle_xxh64(data, seed: int = 0) -> int:
    src = _as_u8(data)
    return int(orc_lib().zko_xxh64(_ptr(src), src.size, seed))


class FrameStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "n_raw", "n_rle", "n_comp", "lit_raw", "lit_rle", "lit_huf", "lit_treeless", "lit_1stream", "lit_4stream",
        "huf_direct", "huf_fse", "mode_predef", "mode_rle", "mode_fse", "mode_repeat", "nseq0_blocks",
        "checksum_frames", "single_segment_frames", "skippable_frames", "zstd_frames")] + [("n_seq", ctypes.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def frame_stats(comp) -> dict:
    src = _as_u8(comp)
    st = FrameStats()
    r = orc_lib().zko_frame_stats(_ptr(src), src.size, ctypes.byref(st))
    if r < 0:
        raise ZstdError(int(-r))
    return st.as_dict()


# ----------------------------------------------------------------------------- seek table restatement
class OracleSeekTable:
    """Pure-Python restatement of SeekTable / Parser / Serializer (seek_table.rs)."""

    def __init__(self):
        self.c = [0]      # cumulative compressed offsets, N+1 entries (seek_table.rs:97-101, 314-321)
        self.d = [0]

    def log_frame(self, c_size: int, d_size: int):          # seek_table.rs:513-525
        if self.num_frames() >= SEEKABLE_MAX_FRAMES:
            raise IndexError("frame index too large")
        self.c.append(self.c[-1] + c_size)
        self.d.append(self.d[-1] + d_size)

    def num_frames(self) -> int:
        return len(self.c) - 1

    def frame_index_at(self, offset: int, arr) -> int:      # seek_table.rs:916-934
        n = self.num_frames()
        if offset >= arr[n]:
            return n - 1
        low, high = 0, n
        while low + 1 < high:
            mid = (low + high) // 2
            if arr[mid] <= offset:
                low = mid
            else:
                high = mid
        return low

    def frame_index_comp(self, off):
        return self.frame_index_at(off, self.c)

    def frame_index_decomp(self, off):
        return self.frame_index_at(off, self.d)

    def serialize(self, fmt: str = "foot") -> bytes:        # seek_table.rs:967-1005
        n = self.num_frames()
        frames = b"".join(struct.pack("<II", self.c[i + 1] - self.c[i], self.d[i + 1] - self.d[i]) for i in range(n))
        integrity = struct.pack("<IBI", n, 0, SEEKABLE_MAGIC_NUMBER)
        head = struct.pack("<II", SKIPPABLE_MAGIC_NUMBER, 9 + 8 * n)
        return head + (integrity + frames if fmt == "head" else frames + integrity)

    @classmethod
    def parse(cls, buf: bytes, fmt: str = "foot") -> "OracleSeekTable":   # seek_table.rs:144-225, 379-436
        if fmt == "foot":
            if len(buf) < 9:
                raise ValueError("offset out of range")
            integ = buf[-9:]
        else:
            if len(buf) < 17:
                raise ValueError("offset out of range")
            integ = buf[8:17]
        n, desc, magic = struct.unpack("<IBI", integ)
        if magic != SEEKABLE_MAGIC_NUMBER:
            raise ZstdError(10)
        if (desc >> 2) & 0x1F:
            raise ZstdError(20)
        if n > SEEKABLE_MAX_FRAMES:
            raise IndexError("frame index too large")
        per = 12 if desc & 0x80 else 8
        size = n * per + 17
        if size > len(buf):
            raise ValueError("offset out of range")
        tbl = buf[-size:] if fmt == "foot" else buf[:size]
        m, fsz = struct.unpack("<II", tbl[:8])
        if m != SKIPPABLE_MAGIC_NUMBER:
            raise ZstdError(10)
        if fsz + 8 != size:
            raise ZstdError(20)
        body = tbl[8:-9] if fmt == "foot" else tbl[17:]
        st = cls()
        for i in range(n):
            c, d = struct.unpack_from("<II", body, i * per)
            st.log_frame(c, d)
        return st


def ref_seekable_archive(data, frame_size=0x200000, level=0, checksum=False, threads=1):
    """Full seekable archive as Encoder::finish() would write it (encode.rs:743-775):
    frames + Foot seek table.  Mirrors the reference's behaviour that finish() always closes one frame, so an
    empty input yields a single empty frame. -> (bytes, OracleSeekTable)"""
    frames, cs, ds = ref_compress_frames(data, frame_size, level, checksum, threads)
    n = _as_u8(data).size
    if n == 0:
        # finish() -> end_frame() on a frame with no input yet: libzstd emits an empty frame.  (A frame that
        # filled exactly is only closed lazily at the next compress() call or by finish() itself,
        # encode.rs:317-327, so exact multiples of the frame size do NOT get an extra empty frame.)
        f2, c2, d2 = _empty_frame(level, checksum)
        frames.append(f2); cs.append(c2); ds.append(d2)
    st = OracleSeekTable()
    for c, d in zip(cs, ds):
        st.log_frame(c, d)
    return b"".join(frames) + st.serialize("foot"), st


def _empty_frame(level, checksum):
    lib = ref_lib()
    src = np.zeros(1, dtype=np.uint8)
    dst = np.empty(64, dtype=np.uint8)
    sizes = (ctypes.c_int64 * 1)()
    # frame_size=1, n=0 is not expressible through zkr_compress_frames(n_frames=ceil(0/1)=0); call with n_frames=1
    rc = lib.zkr_compress_frames(_ptr(src), 0, 1, level, int(checksum), _ptr(dst), 64, sizes, 1, 1)
    if rc != 0:
        raise RuntimeError("libzstd empty frame failed")
    return dst[: sizes[0]].tobytes(), int(sizes[0]), 0

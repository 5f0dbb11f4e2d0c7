"""Deterministic synthetic corpora for tests and bench.py (BASELINE.json configs).

A "Silesia-like" mix of classes (text / structured binary / low-entropy / incompressible / runs) drawn per 256 KiB
segment.  The text class is cut from the reference's own corpus, assets/dickens.txt, committed as the fixture
tests/golden/dickens.txt (the reference tree does not exist on the GPU box): every segment is a slice at a random
offset, as SURVEY.md 8d config 2 prescribes; without the fixture a Zipf pseudo-vocabulary stands in (`text_source()`).
Generation is written with torch ops over a COUNTER-BASED generator (splitmix64 of the element index, integer ops
only), so the bytes are identical on the CPU and on cuda:0 -- bench.py's two arms and the tests see the same input.
Nothing here is on the timed path.
"""
from __future__ import annotations

import os

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _i64(v: int) -> int:
    """python int -> the int64 with the same 64-bit pattern"""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """logical shift right of int64 lanes"""
    return (x >> k) & ((1 << (64 - k)) - 1)


class Gen:
    """counter-based uniform generator: draw #d, element #i = splitmix64(seed, d, i) -- the same on every device"""

    def __init__(self, seed: int, device="cpu"):
        self.seed, self.device, self.draw = int(seed), device, 0

    def bits(self, n: int) -> torch.Tensor:
        self.draw += 1
        key = (self.seed * 0x9E3779B97F4A7C15 + self.draw * 0xD1B54A32D192ED03) & _M64
        x = torch.arange(n, device=self.device, dtype=torch.int64) * _i64(0x9E3779B97F4A7C15) + _i64(key)
        x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
        x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
        return x ^ _lsr(x, 31)

    def rand(self, n: int) -> torch.Tensor:
        """float64 in [0, 1)"""
        return _lsr(self.bits(n), 11).to(torch.float64) * (1.0 / (1 << 53))

    def randint(self, lo: int, hi: int, n: int) -> torch.Tensor:
        """int64 in [lo, hi)  (hi - lo < 2^31)"""
        return lo + (_lsr(self.bits(n), 33) * (hi - lo) >> 31)

SEGMENT = 256 * 1024

CLASS_MIX_SILESIA = {"text": 0.40, "structured": 0.30, "lowent": 0.15, "random": 0.10, "runs": 0.05}
CLASS_MIX_MIXED = {"text": 0.25, "structured": 0.25, "lowent": 0.20, "random": 0.20, "runs": 0.10}

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2,
                      2.0, 2.0, 1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()


class _Vocab:
    """Zipf-weighted pseudo-English vocabulary (word bytes include the trailing separator)."""

    def __init__(self, n_words: int = 6000, seed: int = 1234):
        rng = np.random.default_rng(seed)
        lens = np.clip(rng.poisson(4.2, n_words) + 1, 1, 14)
        # frequent words are short
        order = np.argsort(lens + rng.normal(0, 2.0, n_words))
        lens = lens[order]
        words = []
        for i, ln in enumerate(lens):
            w = rng.choice(_LETTERS, size=int(ln), p=_LETTER_P).tobytes()
            r = rng.random()
            sep = b" " if r < 0.86 else (b", " if r < 0.93 else (b". " if r < 0.98 else b"\n"))
            if i > 50 and rng.random() < 0.04:
                w = w.capitalize()
            words.append(w + sep)
        self.lens = np.array([len(w) for w in words], dtype=np.int64)
        self.starts = np.concatenate([[0], np.cumsum(self.lens)[:-1]]).astype(np.int64)
        self.flat = np.frombuffer(b"".join(words), dtype=np.uint8).copy()
        p = 1.0 / np.arange(1, n_words + 1) ** 1.05
        self.cdf = np.cumsum(p / p.sum())
        self.mean_len = float((self.lens * (p / p.sum())).sum())


_VOCAB = None


def _vocab() -> _Vocab:
    global _VOCAB
    if _VOCAB is None:
        _VOCAB = _Vocab()
    return _VOCAB


_DICKENS = None
DICKENS_PATH = os.environ.get("ZK_DICKENS", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dickens.txt"))


def dickens() -> np.ndarray | None:
    """the reference's corpus (assets/dickens.txt, 10 192 446 bytes) as committed under tests/golden/, or None"""
    global _DICKENS
    if _DICKENS is None:
        _DICKENS = np.fromfile(DICKENS_PATH, dtype=np.uint8) if os.path.exists(DICKENS_PATH) else False
    return None if _DICKENS is False else _DICKENS


def text_source() -> str:
    return "dickens.txt slices" if dickens() is not None else "zipf pseudo-vocabulary (fixture tests/golden/dickens.txt missing)"


def _gen_text(n: int, gen: "Gen", device) -> torch.Tensor:
    d = dickens()
    if d is None:
        return _gen_pseudotext(n, gen, device)
    # every SEGMENT-sized piece is a slice of the corpus at a random offset
    n_seg = (n + SEGMENT - 1) // SEGMENT
    offs = gen.randint(0, d.size - SEGMENT, n_seg)
    src = torch.from_numpy(d).to(device)
    idx = (offs[:, None] + torch.arange(SEGMENT, device=device, dtype=torch.int64)[None, :]).reshape(-1)[:n]
    return src[idx].contiguous()


def _gen_pseudotext(n: int, gen: "Gen", device) -> torch.Tensor:
    v = _vocab()
    n_words = int(n / v.mean_len * 1.15) + 64
    cdf = torch.from_numpy(v.cdf).to(device=device, dtype=torch.float64)
    u = gen.rand(n_words)
    ids = torch.searchsorted(cdf, u).clamp_(max=len(v.lens) - 1)
    # phrase reuse: ~8 % of positions start a 2..5-word phrase copied from 5..400 words back, which
    # creates the medium-length repeats natural language has ("of the", "it was the")
    idx = torch.arange(n_words, device=device)
    start = (gen.rand(n_words) < 0.08) & (idx >= 400)
    plen = gen.randint(2, 6, n_words)
    back = gen.randint(5, 400, n_words)
    last_start = torch.cummax(torch.where(start, idx, torch.full_like(idx, -1)), 0).values
    ls = last_start.clamp(min=0)
    in_phrase = (last_start >= 0) & (idx - ls < plen[ls])
    src = torch.where(in_phrase, idx - back[ls], idx)
    ids = ids[src]
    lens = torch.from_numpy(v.lens).to(device)[ids]
    starts = torch.from_numpy(v.starts).to(device)[ids]
    ends = torch.cumsum(lens, 0)
    total = int(ends[-1].item())
    word_of_byte = torch.repeat_interleave(torch.arange(n_words, device=device), lens, output_size=total)
    within = torch.arange(total, device=device) - (ends - lens)[word_of_byte]
    flat = torch.from_numpy(v.flat).to(device)
    out = flat[starts[word_of_byte] + within]
    if total < n:  # extremely unlikely; pad by wrapping
        out = out.repeat((n + total - 1) // total)
    return out[:n].contiguous()


def _gen_structured(n: int, gen: Gen, device) -> torch.Tensor:
    """little-endian u32 arithmetic ramps, random stride per 4 KiB run, 10 % low-byte noise"""
    n32 = (n + 3) // 4
    run = 1024
    n_runs = (n32 + run - 1) // run
    base = gen.randint(0, 1 << 24, n_runs).reshape(-1, 1)
    stride = gen.randint(1, 64, n_runs).reshape(-1, 1)
    vals = (base + stride * torch.arange(run, device=device, dtype=torch.int64)).reshape(-1)[:n32]
    noise = gen.rand(n32) < 0.10
    lowb = gen.randint(0, 256, n32)
    vals = torch.where(noise, (vals & ~0xFF) | lowb, vals) & 0xFFFFFFFF
    b = torch.stack([(vals >> s) & 0xFF for s in (0, 8, 16, 24)], dim=1).to(torch.uint8).reshape(-1)
    return b[:n].contiguous()


def _gen_lowent(n: int, gen: Gen, device) -> torch.Tensor:
    k = int(gen.randint(4, 17, 1).item())
    p = 1.0 / torch.arange(1, k + 1, device=device, dtype=torch.float64) ** 1.3
    cdf = torch.cumsum(p / p.sum(), 0)
    u = gen.rand(n)
    sym = torch.searchsorted(cdf, u).clamp_(max=k - 1)
    alphabet = gen.randint(0, 256, k)
    return alphabet[sym].to(torch.uint8)


def _gen_random(n: int, gen: Gen, device) -> torch.Tensor:
    return (gen.bits(n) & 0xFF).to(torch.uint8)


def _gen_runs(n: int, gen: Gen, device) -> torch.Tensor:
    """zeros with occasional runs of another byte"""
    n_runs = max(1, n // 2048)
    lens = gen.randint(1, 4096, n_runs)
    vals = torch.where(gen.rand(n_runs) < 0.7,
                       torch.zeros(n_runs, device=device, dtype=torch.int64),
                       gen.randint(0, 256, n_runs))
    out = torch.repeat_interleave(vals, lens)
    if out.numel() < n:
        out = torch.cat([out, torch.zeros(n - out.numel(), device=device, dtype=torch.int64)])
    return out[:n].to(torch.uint8).contiguous()


_GEN = {"text": _gen_text, "pseudotext": _gen_pseudotext, "structured": _gen_structured, "lowent": _gen_lowent, "random": _gen_random,
        "runs": _gen_runs}


def make_class(kind: str, n: int, seed: int = 0, device="cpu") -> torch.Tensor:
    return _GEN[kind](n, Gen(seed, device), device)


def make_mix(n: int, seed: int = 20260924, mix=None, device="cpu", segment: int = SEGMENT) -> torch.Tensor:
    """n bytes: consecutive `segment`-byte pieces, each drawn from `mix`."""
    mix = mix or CLASS_MIX_SILESIA
    names = list(mix)
    p = np.array([mix[k] for k in names], dtype=np.float64)
    rng = np.random.default_rng(seed)
    n_seg = (n + segment - 1) // segment
    kinds = rng.choice(len(names), size=n_seg, p=p / p.sum())
    out = torch.empty(n, dtype=torch.uint8, device=device)
    gen = Gen(seed, device)
    # generate each class in one go (fast on GPU), then scatter segments
    for ci, name in enumerate(names):
        segs = np.nonzero(kinds == ci)[0]
        if len(segs) == 0:
            continue
        blob = _GEN[name](len(segs) * segment, gen, device)
        for j, s in enumerate(segs):
            lo = int(s) * segment
            hi = min(n, lo + segment)
            out[lo:hi] = blob[j * segment: j * segment + (hi - lo)]
    return out


def make_text(n: int, seed: int = 7, device="cpu") -> torch.Tensor:
    """enwik-like text (config 3)"""
    return make_class("text", n, seed, device)


def as_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()

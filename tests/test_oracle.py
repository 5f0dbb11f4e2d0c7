"""The oracle itself is pinned here: the plain-C restatement (oracle/zstd_oracle.c) against the REAL libzstd of the
image driven through the reference's call sequence (oracle/libzstd_driver.c), on the reference's own corpus
(committed golden fixtures made from assets/dickens.txt) and on generated data; the seek-table restatement against
the known-answer vectors derived from seek_table.rs:967-1005 / seekable_format.md:59-102."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from zeekstd_b200 import corpus
from util import golden_bytes, golden_meta


def test_libzstd_present():
    assert O.libzstd_version().startswith("1.")


def test_golden_archives_decode_identically():
    meta = golden_meta()
    src = golden_bytes("dickens_96k.txt")
    assert hashlib.sha256(src).hexdigest() == meta["slice"]["sha256"]
    for name, info in meta["archives"].items():
        a = golden_bytes(name + ".zst")
        assert hashlib.sha256(a).hexdigest() == info["sha256"]
        st = O.OracleSeekTable.parse(a, "foot")
        assert st.num_frames() == info["num_frames"]
        want = src[: info["src_bytes"]]
        body = a[: st.c[-1]]
        assert O.ref_decompress_any(body, len(want) + 1) == want          # real libzstd
        assert O.oracle_decompress(body, len(want) + 1) == want           # restatement
        # frame by frame through the seek table
        for i in range(st.num_frames()):
            fr = a[st.c[i]: st.c[i + 1]]
            assert O.oracle_decompress(fr, st.d[i + 1] - st.d[i] + 1) == want[st.d[i]: st.d[i + 1]]


def test_golden_format_coverage():
    """the fixtures exercise the decoder coverage matrix of SURVEY.md 8a"""
    tot = {}
    for info in golden_meta()["archives"].values():
        for k, v in info["stats"].items():
            tot[k] = tot.get(k, 0) + v
    for k in ("n_comp", "lit_huf", "lit_1stream", "lit_4stream", "huf_fse", "mode_fse", "mode_predef", "checksum_frames"):
        assert tot[k] > 0, k


@pytest.mark.parametrize("kind", ["text", "structured", "lowent", "random", "runs"])
@pytest.mark.parametrize("level", [1, 3, 7, 19])
def test_restatement_matches_libzstd_on_generated(kind, level):
    x = corpus.make_class(kind, 200_000, seed=level).numpy()
    frames, cs, ds = O.ref_compress_frames(x, 70_000, level, True)
    assert O.oracle_decompress(b"".join(frames), x.size + 1) == x.tobytes()


def test_restatement_rejects_corruption_like_libzstd():
    x = corpus.make_class("text", 50_000, 1).numpy()
    frames, _, _ = O.ref_compress_frames(x, 50_000, 3, True)
    good = bytearray(frames[0])
    rng = np.random.default_rng(5)
    agree = 0
    for _ in range(60):
        bad = bytearray(good)
        pos = int(rng.integers(0, len(bad)))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        def run(fn):
            try:
                return fn(bytes(bad), x.size + 1) == x.tobytes()
            except O.ZstdError:
                return "err"
        a, b = run(O.ref_decompress_any), run(O.oracle_decompress)
        assert b != True or a == True          # never accept what libzstd rejects as wrong data
        agree += (a == "err") == (b == "err")
    assert agree >= 55                          # both detect (nearly) the same corruptions; checksum catches the rest


def test_xxh64_known_answers():
    assert O.oracle_xxh64(b"") == 0xEF46DB3751D8E999
    assert O.oracle_xxh64(b"a") == 0xD24EC4F1A98C6E5B
    meta = golden_meta()["xxh64"]
    src = golden_bytes("dickens_96k.txt")
    assert O.oracle_xxh64(src) == meta["slice"] and O.oracle_xxh64(src[:1000]) == meta["slice_1000"]


def test_seek_table_known_answer_vectors():
    st = O.OracleSeekTable()
    assert st.serialize("foot").hex() == "5e2a4d1809000000" + "00000000" + "00" + "b1ea928f"
    assert st.serialize("head").hex() == "5e2a4d1809000000" + "00000000" + "00" + "b1ea928f"
    st.log_frame(123, 456)
    assert st.serialize("foot").hex() == "5e2a4d18110000007b000000c80100000100000000b1ea928f"
    assert st.serialize("head").hex() == "5e2a4d18110000000100000000b1ea928f7b000000c8010000"
    st.log_frame(333, 444)
    foot = st.serialize("foot")
    assert foot.hex() == "5e2a4d18190000007b000000c80100004d010000bc0100000200000000b1ea928f" and len(foot) == 33
    back = O.OracleSeekTable.parse(foot, "foot")
    assert back.c == [0, 123, 456] and back.d == [0, 456, 900]


def test_full_dickens_known_answers_recorded():
    """sizes libzstd produces on the reference's benchmark input, recorded when the fixtures were made"""
    m = golden_meta()
    assert m["dickens_bytes"] == 10192446
    assert m["l1_2m"]["c_sizes"] == [870829, 874683, 884095, 870740, 767099]
    assert m["l1_2m"]["d_sizes"] == [2097152] * 4 + [1803838]

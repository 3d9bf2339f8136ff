"""Pins the CPU oracle (oracle/lz4_oracle.c) against everything the reference's own tests hold for the block
path, against the committed known-answers, and against an independent implementation (system liblz4 1.9.4 /
pyarrow), mirroring the reference's test strategy.  CPU only."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from lz4_flex_b200 import corpus
from vectors import DECODE_KATS, NO_PANIC, ROUNDTRIP

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _liblz4():
    try:
        L = ctypes.CDLL("liblz4.so.1")
    except OSError:
        return None
    L.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    L.LZ4_decompress_safe.restype = ctypes.c_int
    L.LZ4_compress_default.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    L.LZ4_compress_default.restype = ctypes.c_int
    L.LZ4_compressBound.argtypes = [ctypes.c_int]
    L.LZ4_compressBound.restype = ctypes.c_int
    return L


LIBLZ4 = _liblz4()


def c_decompress(comp: bytes, n: int) -> bytes:
    buf = ctypes.create_string_buffer(max(n, 1))
    r = LIBLZ4.LZ4_decompress_safe(comp, buf, len(comp), n)
    assert r >= 0, r
    return buf.raw[:r]


def c_compress(data: bytes) -> bytes:
    cap = LIBLZ4.LZ4_compressBound(len(data))
    buf = ctypes.create_string_buffer(max(cap, 1))
    r = LIBLZ4.LZ4_compress_default(data, buf, len(data), cap)
    assert r > 0 or not data
    return buf.raw[:r]


# ---- decoder KATs from the reference's unit tests -----------------------------------------------------

@pytest.mark.parametrize("name,stream,cap,status,out,expected", DECODE_KATS, ids=[k[0].split()[0] + str(i) for i, k in enumerate(DECODE_KATS)])
def test_decode_kats(name, stream, cap, status, out, expected):
    st_, o, e1, e2 = oracle.decompress_block(bytes(stream), cap)
    assert st_ == status, name
    if out is not None:
        assert o == out
    if expected is not None:
        assert (e1, e2) == (expected, cap)


def test_decode_single_zero_token():
    # a lone token with no literals ends the stream: Ok(0)
    assert oracle.decompress_block(b"\x00", 10)[:2] == (0, b"")


@pytest.mark.parametrize("data", NO_PANIC)
def test_no_panic_corpus(data):
    size = int.from_bytes(bytes(data[:4]), "little")
    if size > 20_000_000:          # tests/tests.rs:497-501
        return
    oracle.decompress_size_prepended(bytes(data))


def test_no_output_leak():
    # fuzz_decomp_no_output_leak.rs:38-43 — the result may not depend on the buffer's previous content
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(1, 80))
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        a = oracle.decompress_block(data, 300)
        b = oracle.decompress_block(data, 300)
        assert a == b


# ---- encoder: the indirect pins the reference offers -------------------------------------------------

def test_ratio_bounds_block():
    # tests/tests.rs:158-171
    for f, bound in [("compression_34k.txt", 0.585), ("compression_65k.txt", 0.574), ("compression_66k_JSON.txt", 0.229)]:
        d = corpus.load(f)
        assert len(oracle.compress_block(d)) / len(d) < bound


def test_ratio_bounds_frame():
    # tests/tests.rs:173-192 (default FrameInfo: Auto block size, independent)
    for f, bound in [("compression_34k.txt", 0.585), ("compression_65k.txt", 0.574), ("compression_66k_JSON.txt", 0.235)]:
        d = corpus.load(f)
        assert len(oracle.frame_compress(d)) / len(d) < bound


def test_conformant_last_block():
    # src/block/compress.rs:952-988: 12 equal bytes are not compressible, 13/14/15 are
    assert len(oracle.compress_block(b"a" * 12)) == 13
    assert oracle.compress_block(b"a" * 12) == b"\xc0" + b"a" * 12
    for n in (13, 14, 15):
        c = oracle.compress_block(b"a" * n)
        assert len(c) < n
        assert oracle.decompress_block(c, n)[:2] == (0, b"a" * n)
    assert oracle.compress_block(b"a" * 13) == bytes([0x12, 0x61, 1, 0, 0x60]) + b"a" * 6
    assert oracle.compress_block(b"") == b"\x00"
    assert oracle.compress_block(b"a") == b"\x10a"
    assert oracle.compress_block(b"Hello people, what's up?") == b"\xf0\x09Hello people, what's up?"


def test_zero_block_layout():
    z = oracle.compress_block(bytes(65536))
    assert z == bytes([0x1F, 0, 1, 0]) + b"\xff" * 256 + bytes([0xE6, 0x60, 0, 0, 0, 0, 0, 0])


def test_max_output_size():
    # compress.rs:588-590
    assert oracle.max_output_size(0) == 20
    assert oracle.max_output_size(65536) == 72109
    assert oracle.max_output_size(66675) == 73362
    assert oracle.max_output_size(4 << 20) == 4613754
    d = corpus.load("compression_1k.txt")
    assert oracle.compress_into(d, oracle.max_output_size(len(d)) - 1) is None      # compress.rs:338-340
    assert oracle.compress_into(d, oracle.max_output_size(len(d))) is not None


@pytest.mark.parametrize("entry", GOLDEN["block"], ids=[e["name"] for e in GOLDEN["block"]])
def test_known_answers_block(entry):
    inputs = {"json_tiled_block0": lambda: corpus.tiled("compression_66k_JSON.txt", 131072).tobytes()[:65536],
              "json_tiled_block1": lambda: corpus.tiled("compression_66k_JSON.txt", 131072).tobytes()[65536:],
              "zeros_65536": lambda: bytes(65536),
              "hdfs_first_4MiB": lambda: corpus.load("hdfs.json")[: 4 << 20],
              "xorshift_65536": lambda: corpus.xorshift64star_bytes(65536).tobytes()}
    data = inputs[entry["name"]]() if entry["name"] in inputs else corpus.load(entry["name"])
    assert sha(data) == entry["input_sha256"]
    for mode, fn in (("block_api", oracle.compress_block), ("frame_fresh", oracle.compress_block_fresh_h5),
                     ("frame_cont", oracle.compress_block_cont)):
        c = fn(data)
        assert (len(c), sha(c)) == (entry[mode]["len"], entry[mode]["sha256"]), mode
        assert oracle.decompress_block(c, len(data))[:2] == (0, data)


def test_survey_known_answers():
    # sizes/hashes from an independent restatement written during the survey
    exp = {"compression_1k.txt": (558, "ada55e5c"), "compression_34k.txt": (19888, "26566a37"),
           "compression_65k.txt": (37150, "ec7b7436"), "compression_66k_JSON.txt": (15268, "46ef8571"),
           "dickens.txt": (6367512, "afba52f0")}
    for f, (n, h) in exp.items():
        c = oracle.compress_block(corpus.load(f))
        assert len(c) == n and sha(c).startswith(h)
    t = corpus.tiled("compression_66k_JSON.txt", 131072).tobytes()
    assert sha(oracle.compress_block(t[:65536])).startswith("af682043")
    assert sha(oracle.compress_block(t[65536:])).startswith("59dcc127")
    assert sha(oracle.compress_block_cont(t[65536:])).startswith("4107a535")


# ---- interop with an independent implementation (the reference's lz4_cpp_compatibility) ---------------

@pytest.mark.skipif(LIBLZ4 is None, reason="liblz4 not present")
def test_liblz4_cross_decode():
    # tests/tests.rs:109-147: flex-compress -> C-decompress, C-compress -> flex-decompress
    items = [corpus.load(f) for f in ("compression_1k.txt", "compression_34k.txt", "compression_65k.txt",
                                      "compression_66k_JSON.txt")] + [corpus.load("dickens.txt")[:1 << 20]] + ROUNDTRIP
    for d in items:
        assert c_decompress(oracle.compress_block(d), len(d)) == d
        assert c_decompress(oracle.compress_block_cont(d), len(d)) == d
        cc = c_compress(d)
        if d:
            assert oracle.decompress_block(cc, len(d))[:2] == (0, d)


@pytest.mark.parametrize("i", range(len(ROUNDTRIP)))
def test_roundtrip_vectors(i):
    d = ROUNDTRIP[i]
    c = oracle.compress_block(d)
    assert oracle.decompress_block(c, len(d))[:2] == (0, d)
    p = oracle.compress_prepend_size(d)
    assert p[:4] == len(d).to_bytes(4, "little") and p[4:] == c
    assert oracle.decompress_size_prepended(p)[:2] == (0, d)
    for bsid in (0, 4, 5, 6, 7):
        for flags in (0, 7):
            f = oracle.frame_compress(d, bsid, flags)
            assert oracle.frame_decompress(f, len(d) + 16)[:2] == (0, d)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.lists(st.integers(0, 5), max_size=40), max_size=60))
def test_roundtrip_property(vv):
    # tests/tests.rs:593-623 (low-entropy nested vectors)
    d = bytes(b for v in vv for b in v)
    c = oracle.compress_block(d)
    assert oracle.decompress_block(c, len(d))[:2] == (0, d)
    if LIBLZ4 is not None and d:
        assert c_decompress(c, len(d)) == d


@settings(max_examples=150, deadline=None)
@given(st.binary(max_size=200), st.integers(0, 400))
def test_garbage_never_crashes(data, cap):
    # fuzz_decomp_corrupt_block.rs: random input, arbitrary capacity
    s, o, e1, e2 = oracle.decompress_block(data, cap)
    assert 0 <= s <= 6 and len(o) <= cap


# ---- frame container -------------------------------------------------------------------------------

def test_frame_header_kats():
    # fuzz_decomp_corrupt_frame.rs:26-27: 04 22 4d 18 60 40 82 = independent, 64 KiB, no checksums
    f = oracle.frame_compress(b"x" * 100, 4)
    assert f[:7] == bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x82])
    assert oracle.frame_compress(b"x" * 100, 5)[:7] == bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x50, 0xFB])
    assert oracle.frame_compress(b"x" * 100, 6)[:7] == bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x60, 0x51])
    assert oracle.frame_compress(b"x" * 100, 7)[:7] == bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x70, 0x73])
    assert oracle.xxh32(b"") == 0x02CC5D05
    assert oracle.xxh32(b"a") == 0x550D7456
    assert oracle.xxh32(b"abc") == 0x32D153FF
    assert oracle.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F


def test_frame_content_size_header_len():
    # tests/tests.rs:726-734: with content_size the header is 15 bytes
    f = oracle.frame_compress(corpus.load("compression_1k.txt"), 4, oracle.F_CONTENT_SIZE)
    assert int.from_bytes(f[6:14], "little") == 725
    assert f[4] & 0x08


def test_legacy_frame_fixture():
    # tests/tests.rs:741-745: benches/dickens.lz4 (legacy frame made by the C lz4 CLI) -> dickens.txt
    leg = corpus.load("dickens.lz4")
    s, o, be = oracle.frame_decompress(leg, 10_000_000)
    assert s == 0 and o == corpus.load("dickens.txt")


def test_frame_concatenated_and_checksums():
    # tests/tests.rs:633-700
    a, b = corpus.load("compression_1k.txt"), corpus.load("compression_34k.txt")
    f = oracle.frame_compress(a, 4, 3) + oracle.frame_compress(b, 4, 3)
    assert oracle.frame_decompress(f, len(a) + len(b))[:2] == (0, a + b)
    one = bytearray(oracle.frame_compress(a, 4, oracle.F_BLOCK_CHECKSUMS))
    one[20] ^= 1
    assert oracle.frame_decompress(bytes(one), 2000)[0] == oracle.FERR_BLOCK_CHECKSUM
    two = bytearray(oracle.frame_compress(a, 4, oracle.F_CONTENT_CHECKSUM))
    two[-1] ^= 1
    assert oracle.frame_decompress(bytes(two), 2000)[0] == oracle.FERR_CONTENT_CHECKSUM


def test_frame_block_size_improves_ratio():
    # tests/tests.rs:703-718
    d = corpus.load("dickens.txt")[: 6 << 20]
    sizes = [len(oracle.frame_compress(d, b)) for b in (4, 5, 6, 7)]
    assert sizes[0] > sizes[1] > sizes[2] > sizes[3]


def test_frame_vs_pyarrow():
    pa = pytest.importorskip("pyarrow")
    d = corpus.load("compression_66k_JSON.txt") * 3
    f = oracle.frame_compress(d, 4, 3)
    assert pa.decompress(f, decompressed_size=len(d), codec="lz4").to_pybytes() == d
    # pyarrow's own frames use linked blocks (LZ4F default), which is outside the GPU path's scope
    g = pa.compress(d, codec="lz4").to_pybytes()
    assert oracle.frame_decompress(g, len(d))[:2] == (0, d)          # linked blocks are decoded too


def test_frame_fresh_cont_epochs():
    """Block k of a frame is parsed FRESH iff it starts a table epoch: the frame equals
    per-block compression with the mode chosen by index."""
    d = corpus.tiled("compression_66k_JSON.txt", 10 * 65536 + 1234).tobytes()
    f = oracle.frame_compress(d, 4)
    pos, k = 7, 0
    while True:
        w = int.from_bytes(f[pos:pos + 4], "little"); pos += 4
        if w == 0:
            break
        blk = d[k * 65536:(k + 1) * 65536]
        exp = oracle.compress_block_fresh_h5(blk) if k == 0 else oracle.compress_block_cont(blk)
        assert f[pos:pos + w] == exp, k
        pos += w; k += 1
    assert k == 11 and pos == len(f)


def test_frame_flush_shifts_phase():
    d = corpus.tiled("compression_66k_JSON.txt", 300000).tobytes()
    f = oracle.frame_compress(d, 4, 0, 100000)
    assert oracle.frame_decompress(f, len(d))[:2] == (0, d)
    assert f != oracle.frame_compress(d, 4)


# ---- external dictionary ------------------------------------------------------------------------

def _liblz4_dict():
    if LIBLZ4 is None:
        return None
    LIBLZ4.LZ4_decompress_safe_usingDict.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_char_p, ctypes.c_int]
    LIBLZ4.LZ4_decompress_safe_usingDict.restype = ctypes.c_int
    return LIBLZ4


def test_dict_reference_unit_tests():
    from dict_cases import REF_INPUT
    # compress.rs:892-911 test_dict
    c = oracle.compress_with_dict(REF_INPUT, REF_INPUT)
    assert len(c) < len(oracle.compress_block(REF_INPUT))
    assert oracle.decompress_with_dict(c, len(REF_INPUT), REF_INPUT)[:2] == (0, REF_INPUT)
    # compress.rs:913-919 test_dict_no_panic: a 3-byte dictionary is ignored
    assert oracle.compress_with_dict(REF_INPUT, bytes([10, 12, 14])) == oracle.compress_block(REF_INPUT)
    # decompress.rs:593-601: offset beyond dictionary + output
    st = oracle.decompress_with_dict(bytes([0x0E, 255, 0, 0x70, 0, 0, 0, 0, 0, 0, 0]), 256, bytes(250))[0]
    assert st == oracle.ERR_OFFSET_OOB
    # compress.rs:921-950 test_dict_match_crossing, restated: decoding with the whole dictionary equals decoding
    # with its first half as dictionary and its second half already in the output (virtual concatenation)
    dct = b"0123456789ABCDE"
    stream = bytes([0x0B, 5, 0, 0x50]) + b"vwxyz"
    assert oracle.decompress_with_dict(stream, 64, dct)[:2] == (0, b"ABCDE" * 3 + b"vwxyz")


def test_dict_roundtrip_and_foreign_decoder():
    """Every dictionary case round-trips through the oracle and decodes identically with liblz4's
    LZ4_decompress_safe_usingDict (the reference's interop strategy, tests/tests.rs:109-147)."""
    from dict_cases import dict_cases
    L = _liblz4_dict()
    for name, data, dct in dict_cases():
        c = oracle.compress_with_dict(data, dct)
        st, o, _, _ = oracle.decompress_with_dict(c, len(data), dct)
        assert (st, o) == (0, data), name
        assert len(c) <= oracle.max_output_size(len(data)), name
        if L is not None and len(data) > 0:
            eff = dct[-65536:] if len(dct) > 3 else b""
            buf = ctypes.create_string_buffer(len(data))
            r = L.LZ4_decompress_safe_usingDict(c, buf, len(c), len(data), eff if eff else None, len(eff))
            assert r == len(data) and buf.raw[: len(data)] == data, name


def test_dict_helps_and_window_is_trimmed():
    j = corpus.load("compression_66k_JSON.txt")
    plain = oracle.compress_block(j[30000:50000])
    with_dict = oracle.compress_with_dict(j[30000:50000], j[:30000])
    assert len(with_dict) < len(plain)
    # only the last 64 KiB of a dictionary are used (init_dict, compress.rs:571-575)
    assert oracle.compress_with_dict(j[:5000], j) == oracle.compress_with_dict(j[:5000], j[-65536:])


@settings(max_examples=150, deadline=None)
@given(st.binary(max_size=600), st.binary(max_size=600), st.integers(min_value=0, max_value=5))
def test_dict_hypothesis_roundtrip(data, dct, k):
    data = data * k
    c = oracle.compress_with_dict(data, dct)
    assert oracle.decompress_with_dict(c, len(data), dct)[:2] == (0, data)


# ---- BlockMode::Linked frames, decode side -----------------------------------------------------------

def _linked_fixtures():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_linked_frames import sources
    src = sources()
    d = os.path.join(os.path.dirname(__file__), "golden", "linked")
    for m in json.load(open(os.path.join(d, "manifest.json"))):
        f = open(os.path.join(d, m["file"]), "rb").read()
        assert sha(f) == m["frame_sha256"] and sha(src[m["source"]]) == m["source_sha256"]
        yield m["file"], f, src[m["source"]]


def test_linked_frames_from_liblz4_decode():
    """Frames written by liblz4's LZ4F with linked blocks (incl. stored blocks, checksums, an LZ4HC parse) decode to
    their sources; a block may reach up to 64 KiB into the frame's earlier output (frame/decompress.rs:196-305)."""
    for name, f, want in _linked_fixtures():
        st, got, berr = oracle.frame_decompress(f, len(want) + 64)
        assert (st, got) == (0, want), name


def test_linked_frame_corruption_is_detected():
    name, f, want = next(_linked_fixtures())                   # has block + content checksums
    bad = bytearray(f); bad[len(f) // 2] ^= 0x55
    assert oracle.frame_decompress(bytes(bad), len(want) + 64)[0] == oracle.FERR_BLOCK_CHECKSUM


# ---- the performance-oriented CPU baseline (oracle/lz4_cpu_baseline.c) must be the same function as the oracle ------

def test_fast_baseline_equals_oracle():
    rng = np.random.default_rng(9)
    inputs = [corpus.load(f) for f in ("compression_1k.txt", "compression_34k.txt", "compression_65k.txt", "compression_66k_JSON.txt")]
    inputs += [corpus.load("dickens.txt")[: 1 << 20], corpus.load("hdfs.json")[: 300000], bytes(65536), bytes(70000)]
    inputs += [corpus.load("compression_66k_JSON.txt")[:n] for n in list(range(0, 40)) + [65534, 65535, 65536]]
    inputs += [bytes(rng.integers(0, a, n, dtype=np.uint8)) for a in (2, 4, 256) for n in (13, 100, 5000, 66000)]
    for d in inputs:
        c = oracle.compress_block(d)
        assert oracle.fast_compress_block(d) == c, len(d)
        for cap in (len(d), len(d) + 5, len(d) + 100):
            assert oracle.fast_decompress_block(c, cap)[:2] == (0, d)
        if len(d) > 20:
            assert oracle.fast_decompress_block(c, len(d) - 1) == oracle.decompress_block(c, len(d) - 1)
    for name, stream, cap, status, out, expected in DECODE_KATS:
        st_, o, e1, e2 = oracle.fast_decompress_block(bytes(stream), cap)
        assert (st_, o if status == 0 else b"") == (status, bytes(out) if status == 0 else b"")
    good = oracle.compress_block(corpus.load("compression_66k_JSON.txt")[:30000])
    for i in range(2000):
        if i % 2:
            b = bytearray(good)
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            s, cap = bytes(b), 30000
        else:
            s, cap = bytes(rng.integers(0, 256, int(rng.integers(1, 120)), dtype=np.uint8)), int(rng.integers(0, 300))
        a, b_ = oracle.fast_decompress_block(s, cap), oracle.decompress_block(s, cap)
        assert a[0] == b_[0] and (a[0] != 0 or a[1] == b_[1]) and (a[0] != 2 or a[2:] == b_[2:]), i


def test_pool_batches_equal_oracle():
    nb, B, slot = 300, 65536, 72112
    data = corpus.tiled("compression_66k_JSON.txt", nb * B)
    offs = np.arange(nb, dtype=np.uint64) * B
    lens = np.full(nb, B, dtype=np.uint32)
    soff = np.arange(nb, dtype=np.uint64) * slot
    scap = np.full(nb, slot, dtype=np.uint32)
    c1, c2 = np.zeros(nb * slot, dtype=np.uint8), np.zeros(nb * slot, dtype=np.uint8)
    l1, s1 = oracle.compress_batch(data, offs, lens, c1, soff, scap, 4)
    pool = oracle.Pool(5)
    for _ in range(3):                                        # the pool is reusable
        l2, s2 = pool.compress(data, offs, lens, c2, soff, scap)
        assert np.array_equal(l1, l2) and not s2.any()
        assert all(c1[b * slot: b * slot + int(l1[b])].tobytes() == c2[b * slot: b * slot + int(l1[b])].tobytes() for b in range(nb))
        back = np.zeros(nb * B, dtype=np.uint8)
        ol, st = pool.decompress(c2, soff, l2, back, offs, lens)
        assert not st.any() and (ol == B).all() and np.array_equal(back, data)
    pool.close()

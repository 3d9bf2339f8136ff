"""CPU parity of the per-thread codec (lz4_flex_b200/csrc/lz4b200_thread_codec.cuh): the exact functions the
K1-T / K2-T kernels run per lane, compiled for the host by g++ (tests/cpp/thread_codec_host.cpp) and compared with
the oracle byte for byte — all three parse modes, both hash variants, every small length, unaligned buffers,
overlapping copies of every period, and garbage streams with identical status / expected fields.
No GPU needed; the GPU tests then only have to show that the kernels run this same code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from lz4_flex_b200 import corpus
from vectors import DECODE_KATS

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "thread_codec_host.cpp")
HDR = os.path.join(os.path.dirname(HERE), "lz4_flex_b200", "csrc", "lz4b200_thread_codec.cuh")
SO = os.path.join(HERE, "cpp", "_build", "libtchost.so")

CONT, H5 = 1, 2


@pytest.fixture(scope="module")
def tch():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", SO, SRC], check=True)
    L = C.CDLL(SO)
    L.tc_host_compress.restype = C.c_uint32
    L.tc_host_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.tc_host_compress_solo.restype = C.c_uint32
    L.tc_host_compress_solo.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.tc_host_decompress.restype = C.c_int
    L.tc_host_decompress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    return L


def _place(data: bytes, mis: int):
    """data inside a padded array at byte misalignment `mis` (the codec reads whole aligned words that overlap the
    buffer, so 8 bytes of slack either side keep a host build inside one allocation)."""
    a = np.full(len(data) + 64, 0xA5, dtype=np.uint8)
    base = a.ctypes.data
    off = (16 - base % 16) % 16 + 16 + mis
    a[off:off + len(data)] = np.frombuffer(data, dtype=np.uint8)
    return a, off


def t_compress(tch, data: bytes, flags=0, mis_in=0, mis_out=0, solo=False) -> bytes:
    a, ao = _place(data, mis_in)
    cap = oracle.max_output_size(len(data))
    o, oo = _place(bytes(cap), mis_out)
    fn = tch.tc_host_compress_solo if solo else tch.tc_host_compress
    n = fn(a.ctypes.data + ao, len(data), o.ctypes.data + oo, flags)
    assert n <= cap
    assert (o[:oo] == 0xA5).all(), "bytes before the output were modified"
    assert (o[oo + cap:] == 0xA5).all(), "bytes after the output slot were modified"
    return o[oo:oo + n].tobytes()


def t_decompress(tch, comp: bytes, cap: int, mis_in=0, mis_out=0):
    a, ao = _place(comp, mis_in)
    o, oo = _place(bytes(cap), mis_out)
    w, e = C.c_uint32(0), C.c_uint64(0)
    st = tch.tc_host_decompress(a.ctypes.data + ao, len(comp), o.ctypes.data + oo, cap, C.byref(w), C.byref(e))
    assert (o[:oo] == 0xA5).all(), "bytes before the output were modified"
    assert (o[oo + cap:] == 0xA5).all(), "bytes after the output were modified"
    return st, o[oo:oo + w.value].tobytes(), e.value


def oracle_modes(data: bytes):
    return {0: oracle.compress_block(data), H5: oracle.compress_block_fresh_h5(data), CONT | H5: oracle.compress_block_cont(data)}


CASES = [
    ("compression_1k.txt", 0, None), ("compression_34k.txt", 0, None), ("compression_65k.txt", 0, None),
    ("compression_66k_JSON.txt", 0, None), ("compression_66k_JSON.txt", 0, 65536), ("compression_66k_JSON.txt", 1139, 65536),
    ("dickens.txt", 0, 65536), ("dickens.txt", 65536 * 7 + 3, 65535), ("dickens.txt", 100, 65534),
    ("hdfs.json", 0, 65536), ("hdfs.json", 12345, 300000), ("dickens.txt", 0, 1 << 20),
]


@pytest.mark.parametrize("name,start,length", CASES)
def test_compress_matches_oracle_all_modes(tch, name, start, length):
    data = corpus.load(name)[start:start + length if length else None]
    for flags, want in oracle_modes(data).items():
        got = t_compress(tch, data, flags)
        assert got == want, (name, start, length, flags)
        st, out, _ = t_decompress(tch, got, len(data))
        assert st == 0 and out == data


def test_compress_small_and_boundary_lengths(tch):
    j = corpus.load("compression_66k_JSON.txt")
    rng = np.random.default_rng(5)
    lengths = list(range(0, 80)) + [15 + 255 * k + d for k in (0, 1, 2) for d in (-1, 0, 1)] + [4095, 4096, 4097, 65533, 65534, 65535, 65536, 65537]
    for n in lengths:
        for data in (j[:n], bytes(n), bytes(rng.integers(0, 3, n, dtype=np.uint8)), b"ab" * (n // 2) + b"a" * (n % 2)):
            for flags, want in oracle_modes(data).items():
                assert t_compress(tch, data, flags) == want, (n, flags)


def test_compress_random_and_lowentropy(tch):
    rng = np.random.default_rng(11)
    for i in range(60):
        n = int(rng.integers(13, 70000))
        alpha = int(rng.choice([2, 4, 16, 256]))
        data = bytes(rng.integers(0, alpha, n, dtype=np.uint8))
        if i % 3 == 0:                                         # long repeats with literal runs between them
            data = (data[:997] * (n // 997 + 1))[:n]
        for flags, want in oracle_modes(data).items():
            assert t_compress(tch, data, flags) == want, (i, n, alpha, flags)


def test_unaligned_buffers(tch):
    data = corpus.load("compression_66k_JSON.txt")[:20000]
    want = oracle.compress_block(data)
    for mi in range(8):
        for mo in range(8):
            got = t_compress(tch, data, 0, mi, mo)
            assert got == want, (mi, mo)
            st, out, _ = t_decompress(tch, got, len(data), mo, mi)
            assert st == 0 and out == data, (mi, mo)


def test_long_literal_and_match_length_encodings(tch):
    rng = np.random.default_rng(3)
    rnd = bytes(rng.integers(0, 256, 3000, dtype=np.uint8))
    for lit in (14, 15, 16, 269, 270, 271, 524, 525, 526, 2000):
        for m in (4, 18, 19, 20, 273, 274, 275, 529, 5000):
            data = rnd[:lit] + b"Z" * 8 + rnd[100:100 + lit][::-1] + (b"0123456789abcdefXYZ" * (m // 19 + 2))[:m + 19] + rnd[:13]
            for flags, want in oracle_modes(data).items():
                got = t_compress(tch, data, flags)
                assert got == want, (lit, m, flags)
                st, out, _ = t_decompress(tch, got, len(data) + 7)
                assert st == 0 and out == data


def test_decode_reference_kats(tch):
    for name, stream, cap, status, out, expected in DECODE_KATS:
        for mi in (0, 3):
            st, o, e = t_decompress(tch, bytes(stream), cap, mi, (mi * 5) % 8)
            assert st == status, name
            if status == 0:
                assert o == bytes(out), name
            if expected is not None:
                assert e == expected, name


def test_decode_every_period_and_alignment(tch):
    # overlapping copies (duplicate_overlapping, decompress.rs:57-82): period 1..70, match lengths around the word
    # and chunk boundaries, every output alignment
    for dist in list(range(1, 48)) + [63, 64, 65, 70]:
        seed = bytes((i * 37 + dist) & 0xff for i in range(dist))
        for mlen in (4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 39, 40, 41, 64, 65, 100, 300, 1000):
            lit_tok = min(dist, 15)
            s = bytearray([(lit_tok << 4) | min(mlen - 4, 15)])
            if dist >= 15:
                s += _ext(dist - 15)
            s += seed + bytes([dist & 0xff, dist >> 8])
            if mlen - 4 >= 15:
                s += _ext(mlen - 4 - 15)
            s += bytes([0x50]) + b"tail!"
            want_st, want, _, _ = oracle.decompress_block(bytes(s), dist + mlen + 5)
            assert want_st == 0
            for mo in (0, 1, 5, 7):
                st, o, _ = t_decompress(tch, bytes(s), dist + mlen + 5, 0, mo)
                assert st == 0 and o == want, (dist, mlen, mo)


def _ext(v):
    b = bytearray()
    while v >= 255:
        b.append(255); v -= 255
    b.append(v)
    return bytes(b)


def test_decode_garbage_matches_oracle(tch):
    rng = np.random.default_rng(77)
    j = corpus.load("compression_66k_JSON.txt")
    good = oracle.compress_block(j[:30000])
    n_err = 0
    for i in range(3000):
        kind = i % 4
        if kind == 0:
            s = bytes(rng.integers(0, 256, int(rng.integers(1, 200)), dtype=np.uint8))
            cap = int(rng.integers(0, 400))
        elif kind == 1:                                        # mutated valid stream
            b = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            s, cap = bytes(b), 30000
        elif kind == 2:                                        # truncated valid stream
            s, cap = good[:int(rng.integers(1, len(good)))], 30000
        else:                                                  # valid stream, output too small
            s, cap = good, int(rng.integers(0, 30000))
        want_st, want, e1, _ = oracle.decompress_block(s, cap)
        st, o, e = t_decompress(tch, s, cap, i % 8, (i // 8) % 8)
        assert st == want_st, (i, kind)
        if st == 0:
            assert o == want
        else:
            n_err += 1
            if st == 2:
                assert e == e1, (i, kind)
    assert n_err > 1000


def test_decode_foreign_and_big(tch):
    d = corpus.load("dickens.txt")[:1 << 20]
    comp = oracle.compress_block(d)
    st, o, _ = t_decompress(tch, comp, len(d))
    assert st == 0 and o == d
    z = bytes(65536)
    cz = oracle.compress_block(z)
    assert len(cz) == 268
    for mo in range(8):
        st, o, _ = t_decompress(tch, cz, 65536, 0, mo)
        assert st == 0 and o == z


# ---- K1-S: the same parse over the shared-memory ring views (host ring, memcpy for the TMA) -----------------------------

def test_solo_ring_parse_matches_oracle(tch):
    """Slot reuse, look-ahead and the 64 KiB history bound of lz4b200_solo_ring.cuh: a slot overwritten too early would
    change the parse.  4 MiB blocks (1 024 ring laps), far matches (distance close to 65 535), long matches that run the
    cursor far ahead, long backward extensions that reach behind the ring, every input misalignment mod 16."""
    rng = np.random.default_rng(123)
    h, d = corpus.load("hdfs.json"), corpus.load("dickens.txt")
    cases = [h[: 4 << 20], d[: 4 << 20], d[5: 5 + (1 << 20)], h[3: 3 + 70000], bytes(300000), corpus.load("compression_66k_JSON.txt")]
    # far matches: a random page repeated at distances around the 65 535 limit
    page = bytes(rng.integers(0, 256, 4096, dtype=np.uint8))
    for gap in (65535 - 4096 - 8, 65535 - 4096, 65536 - 4096, 65540 - 4096, 60000):
        cases.append((page + bytes(rng.integers(0, 256, gap, dtype=np.uint8))) * 4)
    # long backward extension: literal run that matches backwards for several KiB once a 4-byte match is found late
    blob = bytes(rng.integers(0, 256, 70000, dtype=np.uint8))
    cases.append(blob + b"#" + blob[1:] )                      # second copy found late => backtrack
    cases.append(blob[:30000] + bytes(200000) + blob[:30000] + bytes(1000))
    for i, data in enumerate(cases):
        for flags, want in oracle_modes(data).items():
            assert t_compress(tch, data, flags, solo=True) == want, (i, len(data), flags)
    data = h[: 200000]
    want = oracle.compress_block_cont(data)
    for mi in range(16):
        assert t_compress(tch, data, CONT | H5, mi, 0, solo=True) == want, mi
    for n in list(range(0, 40)) + [2047, 2048, 2049, 2048 * 40 - 1, 2048 * 40, 2048 * 40 + 17]:
        data = (d[:n])
        assert t_compress(tch, data, 0, 5, 0, solo=True) == oracle.compress_block(data), n

"""(Run by tests/test_gpu_kernel_variants.py in a child process against the A/B library, liblz4b200_ab.so.)
Every K1 / K2 kernel family against the oracle on the same ragged batch, each forced through the launcher's
environment switches (a fresh context reads them at creation):
  K1-T / K2-T  one block per thread          (lz4b200_thread_kernels.cuh)   LZ4B200_THREAD_MIN=1
  K1-S / K1-S2 one chain per CTA, smem ring  (lz4b200_solo_kernel.cuh)      LZ4B200_ENC_SOLO=1|2, _SOLO_SMALL_MAX=1000000
  warp kernels matcher/emitter warps, lane groups (default)                 LZ4B200_THREAD_MIN=4e9, LZ4B200_ENC_SOLO=0
               (LZ4B200_ENC_GTAG=71: tagged table entries; LZ4B200_ENC_G16=62: two chains per matcher warp;
                LZ4B200_ENC_NIB=1: shared-memory nibble tags, first-2 verification)
Bit-exact compressed bytes in all three parse modes, exact round trips, identical error codes / expected fields."""
import os

import numpy as np
import pytest

import oracle
from lz4_flex_b200 import block, corpus
from vectors import DECODE_KATS

pytestmark = pytest.mark.gpu

VARIANTS = {
    "thread": {"LZ4B200_THREAD_MIN": "1", "LZ4B200_ENC_SOLO": "0"},
    "thread8": {"LZ4B200_THREAD_MIN": "1", "LZ4B200_ENC_SOLO": "0", "LZ4B200_ENC_THREAD_LANES": "8", "LZ4B200_DEC_THREAD_LANES": "8"},
    "thread_few": {"LZ4B200_THREAD_MIN": "1", "LZ4B200_ENC_SOLO": "0", "LZ4B200_ENC_THREADS": "64", "LZ4B200_DEC_THREADS": "64"},
    "solo": {"LZ4B200_ENC_SOLO": "1", "LZ4B200_ENC_SOLO_SMALL_MAX": "1000000", "LZ4B200_THREAD_MIN": "4000000000"},
    "solo2": {"LZ4B200_ENC_SOLO": "2", "LZ4B200_ENC_SOLO_SMALL_MAX": "1000000", "LZ4B200_THREAD_MIN": "4000000000"},
    "warp": {"LZ4B200_THREAD_MIN": "4000000000", "LZ4B200_ENC_SOLO": "0"},
    "warp_tagged": {"LZ4B200_ENC_GTAG": "71"},
    "warp_half": {"LZ4B200_ENC_G16": "62"},
    "warp_half71": {"LZ4B200_ENC_G16": "71"},
    "warp_quarter": {"LZ4B200_ENC_G16": "871"},
    "warp_quarter62": {"LZ4B200_ENC_G16": "862", "LZ4B200_ENC_G16_CTAS": "6"},
    "warp_nib": {"LZ4B200_ENC_NIB": "1"},
    "warp_nib6": {"LZ4B200_ENC_NIB": "3"},
    "warp_tag8": {"LZ4B200_ENC_NIB": "2"},
    "warp_tagg8": {"LZ4B200_ENC_NIB": "4"},
    "warp_tagg16": {"LZ4B200_ENC_NIB": "5"},
}


@pytest.fixture(params=list(VARIANTS))
def vctx(request):
    saved = {k: os.environ.get(k) for k in list(os.environ) if k.startswith("LZ4B200_")}
    for k in saved:
        del os.environ[k]
    os.environ.update(VARIANTS[request.param])
    c = block.Context(0)
    for k in VARIANTS[request.param]:
        del os.environ[k]
    os.environ.update({k: v for k, v in saved.items() if v is not None})
    yield c
    c.close()


def _cases():
    rng = np.random.default_rng(2024)
    j, d, h = corpus.load("compression_66k_JSON.txt"), corpus.load("dickens.txt"), corpus.load("hdfs.json")
    cases = [j[:n] for n in list(range(0, 48)) + [255, 256, 270, 271, 272, 4095, 4096, 4097, 65533, 65534, 65535, 65536]]
    cases += [bytes(n) for n in (13, 14, 300, 4096, 65535, 65536)]
    cases += [bytes(rng.integers(0, a, n, dtype=np.uint8)) for a in (2, 3, 16, 256) for n in (100, 3000, 65536)]
    cases += [d[i * 65536:(i + 1) * 65536] for i in range(12)] + [h[i * 40000: i * 40000 + 65536] for i in range(12)]
    cases += [(b"abcdefgh" * 9000)[:n] for n in (64, 1000, 65536)] + [b"ab" * 20000, b"a" * 50000 + b"b" * 10000]
    rnd = bytes(rng.integers(0, 256, 3000, dtype=np.uint8))
    for lit in (15, 269, 270, 525, 2000):
        for m in (19, 20, 274, 529, 5000):
            cases.append(rnd[:lit] + b"Z" * 8 + rnd[100:100 + lit][::-1] + (b"0123456789abcdefXYZ" * (m // 19 + 2))[:m + 19] + rnd[:13])
    return cases


CASES = _cases()


def test_compress_all_modes_and_roundtrip(vctx):
    for fl, ref in ((None, oracle.compress_block), (block.BLOCK_HASH5_ALWAYS, oracle.compress_block_fresh_h5),
                    (block.BLOCK_HASH5_ALWAYS | block.BLOCK_CONT, oracle.compress_block_cont)):
        comp = block.compress_blocks(CASES, None if fl is None else [fl] * len(CASES), vctx)
        for i, (x, c) in enumerate(zip(CASES, comp)):
            assert c == ref(x), (i, len(x), fl)
    outs, st, _ = block.decompress_blocks(comp, [max(len(x), 1) for x in CASES], vctx)
    assert not st.any() and outs == CASES


def test_big_blocks_and_unaligned(vctx):
    h, d = corpus.load("hdfs.json"), corpus.load("dickens.txt")
    cases = [h[: 4 << 20], d[1: (1 << 20) + 7], h[3: 3 + (1 << 20)], bytes(1 << 20)]
    comp = block.compress_blocks(cases, [block.BLOCK_HASH5_ALWAYS | block.BLOCK_CONT] * len(cases), vctx)
    for x, c in zip(cases, comp):
        assert c == oracle.compress_block_cont(x), len(x)
    outs, st, _ = block.decompress_blocks(comp, [len(x) for x in cases], vctx)
    assert outs == cases
    # unaligned sub-buffers: every input / output misalignment mod 8 through the batch descriptors
    src = corpus.tiled("compression_66k_JSON.txt", 100000)
    lens = np.array([9000 + 7 * k for k in range(64)], dtype=np.uint32)
    offs = np.array([k * 1001 + (k % 8) + 8 * (k % 3) for k in range(64)], dtype=np.uint64)
    out, ooff, olen = block.compress_batch(src, offs, lens, ctx=vctx)
    for k in range(64):
        want = oracle.compress_block(src[int(offs[k]): int(offs[k]) + int(lens[k])].tobytes())
        assert out[int(ooff[k]): int(ooff[k]) + int(olen[k])].tobytes() == want, k
    back = np.zeros(int(lens.sum()) + 64 * 8, dtype=np.uint8)
    assert int(offs[-1]) + int(lens[-1]) <= src.size
    boff = np.cumsum(np.concatenate([[0], lens[:-1].astype(np.uint64) + np.arange(1, 64, dtype=np.uint64) % 8])).astype(np.uint64)
    ol, st, _ = block.decompress_batch(out, ooff, olen, back, boff, lens, ctx=vctx)
    assert not st.any()
    for k in range(64):
        assert back[int(boff[k]): int(boff[k]) + int(lens[k])].tobytes() == src[int(offs[k]): int(offs[k]) + int(lens[k])].tobytes()


def test_decode_errors_match_oracle(vctx):
    rng = np.random.default_rng(31)
    good = oracle.compress_block(corpus.load("compression_66k_JSON.txt")[:30000])
    streams, caps = [bytes(k[1]) for k in DECODE_KATS], [k[2] for k in DECODE_KATS]
    for i in range(1500):
        kind = i % 4
        if kind == 0:
            streams.append(bytes(rng.integers(0, 256, int(rng.integers(1, 200)), dtype=np.uint8))); caps.append(int(rng.integers(0, 400)))
        elif kind == 1:
            b = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            streams.append(bytes(b)); caps.append(30000)
        elif kind == 2:
            streams.append(good[:int(rng.integers(1, len(good)))]); caps.append(30000)
        else:
            streams.append(good); caps.append(int(rng.integers(0, 30000)))
    outs, st, exp = block.decompress_blocks(streams, caps, vctx, raise_on_error=False)
    nerr = 0
    for i, (s, cap) in enumerate(zip(streams, caps)):
        w_st, w_out, e1, _ = oracle.decompress_block(s, cap)
        assert int(st[i]) == w_st, i
        if w_st == 0:
            assert outs[i] == w_out, i
        else:
            nerr += 1
            if w_st == 2:
                assert int(exp[i]) == e1, i
    assert nerr > 500


def test_overlapping_periods(vctx):
    streams, caps, want = [], [], []
    for dist in list(range(1, 48)) + [63, 64, 65, 70]:
        seed = bytes((i * 37 + dist) & 0xff for i in range(dist))
        for mlen in (4, 7, 8, 9, 31, 32, 33, 39, 40, 41, 65, 300, 70000):
            s = bytearray([(min(dist, 15) << 4) | min(mlen - 4, 15)])
            if dist >= 15:
                s += _ext(dist - 15)
            s += seed + bytes([dist & 0xff, dist >> 8])
            if mlen - 4 >= 15:
                s += _ext(mlen - 4 - 15)
            s += bytes([0x50]) + b"tail!"
            st, o, _, _ = oracle.decompress_block(bytes(s), dist + mlen + 5)
            assert st == 0
            streams.append(bytes(s)); caps.append(dist + mlen + 5); want.append(o)
    outs, st, _ = block.decompress_blocks(streams, caps, vctx)
    assert not st.any() and outs == want


def _ext(v):
    b = bytearray()
    while v >= 255:
        b.append(255); v -= 255
    b.append(v)
    return bytes(b)


def test_many_blocks_global_table_kernels(vctx):
    """More blocks than the shared-memory-table kernel keeps in flight (24 per SM): the launcher's global-table kernel
    (tagged entries by default) — ragged block lengths, all three parse modes, every block compared with the oracle."""
    src = corpus.tiled("compression_66k_JSON.txt", 40 << 20)
    d = np.frombuffer(corpus.load("dickens.txt"), dtype=np.uint8)
    src[20 << 20: (20 << 20) + d.size] = d
    rng = np.random.default_rng(77)
    # low-entropy noise and word soup: slot clashes inside a probe batch with real matches behind them — the in-batch
    # forwarding paths (and gnib's continued batch, tests/test_warp_emulation.py::test_nibble_partial_batches_happen)
    at = 31 << 20
    for a in (2, 3, 5, 8, 13, 21, 34):
        src[at: at + (1 << 19)] = rng.integers(0, a, 1 << 19, dtype=np.uint8)
        at += 1 << 19
        words = [rng.integers(0, 256, int(rng.integers(3, 9)), dtype=np.uint8) for _ in range(a * 8)]
        soup = np.concatenate([words[int(i)] for i in rng.integers(0, len(words), 120000)])[: 1 << 19]
        src[at: at + soup.size] = soup
        at += 1 << 19
    lens = rng.integers(1500, 12000, 6000).astype(np.uint32)
    lens[:64] = np.arange(64, dtype=np.uint32)
    lens[64:80] = 65536
    offs = np.zeros(lens.size, dtype=np.uint64)
    offs[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    assert int(offs[-1]) + int(lens[-1]) <= src.size
    nb = lens.size
    slot = 72112
    soff = np.arange(nb, dtype=np.uint64) * slot
    scap = np.full(nb, slot, dtype=np.uint32)
    for fl, tabfn in ((0, None), (block.BLOCK_HASH5_ALWAYS, oracle.compress_block_fresh_h5),
                      (block.BLOCK_HASH5_ALWAYS | block.BLOCK_CONT, oracle.compress_block_cont)):
        out, ooff, olen = block.compress_batch(src, offs, lens, None if fl == 0 else np.full(nb, fl, dtype=np.uint8), ctx=vctx)
        if tabfn is None:
            want = np.zeros(nb * slot, dtype=np.uint8)
            wlen, wst = oracle.compress_batch(src, offs, lens, want, soff, scap, os.cpu_count())
            assert np.array_equal(wlen, olen)
            packed = np.concatenate([want[b * slot: b * slot + int(wlen[b])] for b in range(nb)])
            assert np.array_equal(out[: packed.size], packed)
        else:
            for b in list(range(0, 90)) + list(range(90, nb, 61)):
                a, n = int(offs[b]), int(lens[b])
                assert out[int(ooff[b]): int(ooff[b]) + int(olen[b])].tobytes() == tabfn(src[a:a + n].tobytes()), (fl, b)
    back = np.zeros(int(offs[-1]) + int(lens[-1]), dtype=np.uint8)
    ol, st, _ = block.decompress_batch(out, ooff, olen, back, offs, lens, ctx=vctx)
    assert not st.any() and np.array_equal(back, src[: back.size])

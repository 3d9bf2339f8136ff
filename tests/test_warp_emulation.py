"""CPU model of K1's matcher (lz4_flex_b200/csrc/lz4b200_enc_split.cuh::match_block) — 32 probes per batch, speculative
pre-batch candidates, shuffle / match.any collision handling, commit of the probes the reference really executed — run
lane by lane in plain Python and compared with the oracle.  It pins the *scheme* (that evaluating the sequential probe loop
of compress_internal, compress.rs:373-439, 32 positions at a time is exact) independently of CUDA, in all three parse modes;
the kernel itself is checked against the oracle on the GPU (test_gpu_block.py)."""
import numpy as np
import pytest

import oracle
from lz4_flex_b200 import corpus

INVALID = 0xFFFFFFFF
MASK64 = (1 << 64) - 1


def _h4(b, p):
    v = int.from_bytes(b[p:p + 4], "little")
    return ((v * 2654435761) & 0xFFFFFFFF) >> 20


def _h5(b, p):
    v = int.from_bytes(b[p:p + 8], "little")                     # the reference reads 8 bytes, hashes the low 5
    return ((((v << 24) & MASK64) * 889523592379) & MASK64) >> 52


def _ext(v):
    out = bytearray()
    while v >= 255:
        out.append(255); v -= 255
    out.append(v)
    return bytes(out)


def _last_literals(b, start):
    n = len(b) - start
    return bytes([min(n, 15) << 4]) + (_ext(n - 15) if n >= 15 else b"") + b[start:]


def warp_encode(b: bytes, cont: bool = False, h5=None, stats=None) -> bytes:
    n = len(b)
    if n < 13:
        return _last_literals(b, 0)
    if h5 is None:
        h5 = n >= 65535
    H = _h5 if h5 else _h4
    tab = [INVALID if cont else 0] * 4096
    last_probe, lim = n - 12, n - 6
    out = bytearray()
    anchor = cur = 0
    ri = False
    if not cont:
        tab[H(b, 0)] = 0
        cur = 1
    while True:
        base, stride = cur, 1
        while True:                                              # one probe batch = 32 lanes
            p = [base + i * stride for i in range(32)]
            term = [x > last_probe for x in p]
            if ri:                                               # re-insert of the previous sequence, before the table reads
                tab[H(b, cur - 2)] = cur - 2
                ri = False
            key = [H(b, p[i]) if not term[i] else (0x10000 | i) for i in range(32)]
            cnd = [tab[key[i]] if not term[i] else INVALID for i in range(32)]

            def check(i, c):
                return (not term[i]) and c != INVALID and p[i] - c <= 65535 and b[c:c + 4] == b[p[i]:p[i] + 4]

            hit = [check(i, cnd[i]) for i in range(32)]
            w0 = hit.index(True) if True in hit else 32
            exact = w0 == 0
            if 1 <= w0 <= 3:                                     # three shuffles: keys of lanes 0..2 against lanes <= w0
                exact = not any(key[j] == key[i] for j in range(1, w0 + 1) for i in range(j))
            same = [[i] for i in range(32)]
            if not exact:                                        # match.any + forwarding of in-batch writes
                same = [[j for j in range(32) if key[j] == key[i]] for i in range(32)]
                prior = [[j for j in same[i] if j < i] for i in range(32)]
                upto0 = 31 if w0 >= 31 else w0
                if any(prior[i] for i in range(upto0 + 1)):
                    for i in range(32):
                        if prior[i]:
                            cnd[i] = p[prior[i][-1]]
                            hit[i] = check(i, cnd[i])
                if stats is not None:
                    stats["slow"] = stats.get("slow", 0) + 1
            win = hit.index(True) if True in hit else 32
            tfirst = term.index(True) if True in term else 32
            if tfirst < win:                                     # compress.rs:381-384
                return bytes(out) + _last_literals(b, anchor)
            upto = win if win < 32 else 31
            for i in range(upto + 1):                            # commit: last writer per slot among the executed probes
                if max(j for j in same[i] if j <= upto) == i:
                    tab[key[i]] = p[i]
            if win < 32:
                mpos, cand = p[win], cnd[win]
                break
            base += 32 * stride
            stride += 1
        dist = mpos - cand
        while cand > 0 and mpos > anchor and b[mpos - 1] == b[cand - 1]:
            mpos -= 1; cand -= 1
        end, c = mpos + 4, cand + 4
        while end < lim and b[end] == b[c]:
            end += 1; c += 1
        lit, extra = mpos - anchor, end - mpos - 4
        out.append((min(lit, 15) << 4) | min(extra, 15))
        if lit >= 15:
            out += _ext(lit - 15)
        out += b[anchor:mpos]
        out += dist.to_bytes(2, "little")
        if extra >= 15:
            out += _ext(extra - 15)
        anchor = cur = end
        ri = True
        if stats is not None:
            stats["seqs"] = stats.get("seqs", 0) + 1


def _cases():
    rng = np.random.default_rng(9)
    cases = [(f, corpus.load(f)) for f in ("compression_1k.txt", "compression_34k.txt", "compression_65k.txt")]
    cases.append(("json_block", corpus.tiled("compression_66k_JSON.txt", 65536).tobytes()))
    cases.append(("zeros", bytes(20000)))
    cases.append(("period7", (b"abcdefg" * 3000)[:20000]))
    for k in range(12):
        a = int(rng.integers(2, 6))
        cases.append((f"low_entropy_{k}", rng.integers(0, a, int(rng.integers(13, 6000)), dtype=np.uint8).tobytes()))
    cases.append(("random", rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()))
    for n in (0, 1, 12, 13, 14, 20, 64):
        cases.append((f"tiny_{n}", (b"xyxyzxyxyz" * 10)[:n]))
    return cases


CASES = _cases()


@pytest.mark.parametrize("name,data", CASES, ids=[c[0] for c in CASES])
def test_batched_probe_scheme_equals_sequential_parse(name, data):
    assert warp_encode(data) == oracle.compress_block(data)                              # block API (FRESH)
    assert warp_encode(data, cont=True, h5=True) == oracle.compress_block_cont(data)     # frame block, table carried over
    assert warp_encode(data, cont=False, h5=True) == oracle.compress_block_fresh_h5(data)  # frame block 0 / Large table


def test_collision_path_is_exercised():
    stats = {}
    data = bytes(20000)                                           # every probe of a batch lands in one slot
    assert warp_encode(data, stats=stats) == oracle.compress_block(data)
    s2 = {}
    low = np.random.default_rng(1).integers(0, 2, 30000, dtype=np.uint8).tobytes()
    assert warp_encode(low, stats=s2) == oracle.compress_block(low)
    assert s2.get("slow", 0) > 0 and s2["seqs"] > 100

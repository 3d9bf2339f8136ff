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


def _tag16(b, p):
    return ((int.from_bytes(b[p:p + 4], "little") * 2246822519) & 0xFFFFFFFF) >> 16


def warp_encode(b: bytes, cont: bool = False, h5=None, stats=None, tagged: bool = False, G: int = 32) -> bytes:
    """tagged=True models lz4_compress_blocks_gtag: table entries are (tag << 16) | position, a probe only looks at its
    candidate's bytes when the tags agree, and a candidate forwarded inside the batch is compared register to register."""
    n = len(b)
    if n < 13:
        return _last_literals(b, 0)
    if h5 is None:
        h5 = n >= 65535
    H = _h5 if h5 else _h4
    tab = [INVALID if cont else 0] * 4096
    fetched = [0]                                                # candidate fetches from memory (the traffic tags remove)
    if tagged:
        assert n <= 65536
        tab = [INVALID if cont else (_tag16(b, 0) << 16)] * 4096
    last_probe, lim = n - 12, n - 6
    out = bytearray()
    anchor = cur = 0
    ri = False
    if not cont:
        tab[H(b, 0)] = (_tag16(b, 0) << 16) if tagged else 0
        cur = 1
    R = range(G)
    while True:
        base, nbatch = cur, 0
        while True:                                              # one probe batch = G lanes (32, or 16 for the half-warp kernel)
            stride = nbatch // (32 // G) + 1                     # compress.rs:374-378: 32 probes per step value
            p = [base + i * stride for i in R]
            term = [x > last_probe for x in p]
            if ri:                                               # re-insert of the previous sequence, before the table reads
                tab[H(b, cur - 2)] = ((cur - 2) | (_tag16(b, cur - 2) << 16)) if tagged else cur - 2
                ri = False
            key = [H(b, p[i]) if not term[i] else (0x10000 | i) for i in R]
            cnd = [tab[key[i]] if not term[i] else INVALID for i in R]
            tag_ok = [True] * G
            if tagged:
                for i in R:
                    if cnd[i] != INVALID:
                        tag_ok[i] = (cnd[i] >> 16) == _tag16(b, p[i])
                        cnd[i] &= 0xFFFF

            def check(i, c, use_tag=True):
                if term[i] or c == INVALID or p[i] - c > 65535 or (use_tag and not tag_ok[i]):
                    return False
                fetched[0] += 1
                return b[c:c + 4] == b[p[i]:p[i] + 4]

            hit = [check(i, cnd[i]) for i in R]
            w0 = hit.index(True) if True in hit else G
            exact = w0 == 0
            if 1 <= w0 <= 3:                                     # three shuffles: keys of lanes 0..2 against lanes <= w0
                exact = not any(key[j] == key[i] for j in range(1, w0 + 1) for i in range(j))
            same = [[i] for i in R]
            if not exact:                                        # match.any + forwarding of in-batch writes
                same = [[j for j in R if key[j] == key[i]] for i in R]
                prior = [[j for j in same[i] if j < i] for i in R]
                upto0 = G - 1 if w0 >= G - 1 else w0
                if any(prior[i] for i in range(upto0 + 1)):
                    for i in R:
                        if prior[i]:
                            cnd[i] = p[prior[i][-1]]
                            hit[i] = check(i, cnd[i], use_tag=False)      # forwarded: compared lane to lane, no tag
                if stats is not None:
                    stats["slow"] = stats.get("slow", 0) + 1
            win = hit.index(True) if True in hit else G
            tfirst = term.index(True) if True in term else G
            if tfirst < win:                                     # compress.rs:381-384
                return bytes(out) + _last_literals(b, anchor)
            upto = win if win < G else G - 1
            for i in range(upto + 1):                            # commit: last writer per slot among the executed probes
                if max(j for j in same[i] if j <= upto) == i:
                    tab[key[i]] = (p[i] | (_tag16(b, p[i]) << 16)) if tagged else p[i]
            if win < G:
                mpos, cand = p[win], cnd[win]
                break
            base += G * stride
            nbatch += 1
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
            stats["fetched"] = fetched[0]


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


@pytest.mark.parametrize("name,data", [c for c in CASES if len(c[1]) <= 65536], ids=[c[0] for c in CASES if len(c[1]) <= 65536])
def test_tagged_entries_keep_the_parse(name, data):
    """lz4_compress_blocks_gtag: (tag, position) entries.  A tag mismatch proves the 4-byte comparison fails, so skipping
    the candidate fetch cannot change the parse — in any mode, including the FRESH table whose empty slots stand for
    position 0 and must carry position 0's tag."""
    assert warp_encode(data, tagged=True) == oracle.compress_block(data)
    assert warp_encode(data, cont=True, h5=True, tagged=True) == oracle.compress_block_cont(data)
    assert warp_encode(data, cont=False, h5=True, tagged=True) == oracle.compress_block_fresh_h5(data)


def test_tags_remove_the_false_candidate_fetches():
    """What is left after the tag filter are (almost only) true 4-byte matches: on JSON ~18 of a batch's 32 probes really
    match their candidate, and those candidates are consecutive addresses (one or two sectors).  The ~45 % that the tags
    remove are the scattered ones — stale or colliding slots — each of which cost a sector of its own."""
    data = corpus.tiled("compression_66k_JSON.txt", 65536).tobytes()
    a, t = {}, {}
    assert warp_encode(data, h5=True, stats=a) == warp_encode(data, h5=True, stats=t, tagged=True)
    assert t["fetched"] < 0.65 * a["fetched"], (a["fetched"], t["fetched"])


@pytest.mark.parametrize("name,data", [c for c in CASES if len(c[1]) <= 65536], ids=[c[0] for c in CASES if len(c[1]) <= 65536])
def test_lane_group_batches_keep_the_parse(name, data):
    """lz4_compress_blocks_gtabg<G>: G probes per batch, 32/G batches per step value of compress.rs:374-378."""
    for G in (16, 8):
        assert warp_encode(data, G=G) == oracle.compress_block(data)
        assert warp_encode(data, cont=True, h5=True, G=G) == oracle.compress_block_cont(data)
        assert warp_encode(data, cont=False, h5=True, G=G) == oracle.compress_block_fresh_h5(data)


# ---------------------------------------------------------------------------------------------------------------------
# lz4_compress_blocks_gnib: 4-bit tags in SHARED memory beside the global position table, candidates verified first-K.
# ---------------------------------------------------------------------------------------------------------------------
_TAG_BITS = [4]                                                  # 4: nibble tags, 8: byte tags (gnib<.., 8, ..>)


def _tag4(b, p):
    return ((int.from_bytes(b[p:p + 4], "little") * 2246822519) & 0xFFFFFFFF) >> (32 - _TAG_BITS[0])


def warp_encode_nib(b: bytes, cont: bool = False, h5=None, stats=None, K: int = 2, G: int = 32) -> bytes:
    """Model of match_block_view<kNib>: every slot has a 4-bit tag of the 4 bytes at its position (shared memory); a probe
    whose tag disagrees cannot pass the 4-byte comparison (compress.rs:432-438), so it touches neither the position table
    nor the candidate.  The tag-matching probes of a batch are verified K at a time in probe order and the search stops at
    the first real hit — the sequential loop never looks past it either.  If in-batch forwarding (an earlier probe of the
    batch owns the slot) takes the hit away and nothing at or before it matches, the probes up to there are committed and
    the batch continues behind them inside the same 32-probe step group (`gi`)."""
    n = len(b)
    if n < 13:
        return _last_literals(b, 0)
    if h5 is None:
        h5 = n >= 65535
    H = _h5 if h5 else _h4
    tab = [INVALID if cont else 0] * 4096
    ntag = [_tag4(b, 0) if n >= 4 else 0] * 4096                 # FRESH: empty slot = position 0 (cont: value irrelevant)
    last_probe, lim = n - 12, n - 6
    out = bytearray()
    anchor = cur = 0
    ri = False
    st = stats if stats is not None else {}
    for k in ("seqs", "batches", "rounds", "tab_loads", "cand_loads", "partial", "slow"):
        st.setdefault(k, 0)
    if not cont:
        tab[H(b, 0)] = 0
        ntag[H(b, 0)] = _tag4(b, 0)
        cur = 1
    R = range(G)                                                 # G < 32: a lane group of gtagg (32/G chains per warp)
    while True:
        gbase, stride, gi = cur, 1, 0
        while True:
            width = min(G, 32 - gi)                              # a batch never crosses the 32-probe step group
            p = [gbase + (gi + i) * stride for i in R]
            act = [i < width for i in R]
            term = [act[i] and p[i] > last_probe for i in R]
            live = [act[i] and not term[i] for i in R]
            if ri:
                tab[H(b, cur - 2)] = cur - 2
                ntag[H(b, cur - 2)] = _tag4(b, cur - 2)
                ri = False
            key = [H(b, p[i]) if live[i] else (0x10000 | i) for i in R]
            mytag = [_tag4(b, p[i]) if live[i] else 256 for i in R]
            pend = [i for i in R if live[i] and ntag[key[i]] == mytag[i]]
            st["batches"] += 1
            hit = [False] * G
            cnd = [INVALID] * G

            def verify(i):
                c = tab[key[i]]
                st["tab_loads"] += 1
                cnd[i] = c
                if c == INVALID or p[i] - c > 65535:
                    return False
                st["cand_loads"] += 1
                return b[c:c + 4] == b[p[i]:p[i] + 4]

            while pend:
                sel, pend = pend[:K], pend[K:]
                st["rounds"] += 1
                for i in sel:
                    hit[i] = verify(i)
                if any(hit[i] for i in sel):
                    break
            w0 = hit.index(True) if True in hit else G
            upto0 = min(w0, width - 1)
            exact = w0 == 0
            if 1 <= w0 <= 3:
                exact = not any(key[j] == key[i] for j in range(1, w0 + 1) for i in range(j))
            same = [[i] for i in R]
            win = w0
            if not exact:
                same = [[j for j in R if key[j] == key[i]] for i in R]
                prior = [[j for j in same[i] if j < i] for i in R]
                if any(prior[i] for i in range(upto0 + 1)):
                    st["slow"] += 1
                    for i in range(upto0 + 1):
                        if prior[i]:                             # forwarded candidate: lane-to-lane compare, no memory
                            cnd[i] = p[prior[i][-1]]
                            hit[i] = b[cnd[i]:cnd[i] + 4] == b[p[i]:p[i] + 4]
                    hh = [hit[i] for i in range(upto0 + 1)]
                    win = hh.index(True) if True in hh else G
            tfirst = term.index(True) if True in term else G
            if win == G and w0 < G and tfirst > w0:
                # forwarding took the hit at w0 away and nothing before it matches: lanes 0..w0 were executed probes
                upto = w0
                st["partial"] += 1
            else:
                if tfirst < win:                                 # compress.rs:381-384
                    return bytes(out) + _last_literals(b, anchor)
                upto = win if win < G else width - 1
            for i in range(upto + 1):
                if max(j for j in same[i] if j <= upto) == i:
                    tab[key[i]] = p[i]
                    ntag[key[i]] = mytag[i]
            if win < G:
                mpos, cand = p[win], cnd[win]
                break
            gi += upto + 1
            if gi == 32:
                gbase += 32 * stride
                stride += 1
                gi = 0
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
        st["seqs"] += 1


@pytest.mark.parametrize("name,data", CASES, ids=[c[0] for c in CASES])
def test_nibble_tags_first_k_keep_the_parse(name, data):
    for bits, ks, gs in ((4, (1, 2, 4), (32,)), (8, (2,), (32,)), (2, (2,), (8, 16))):
        _TAG_BITS[0] = bits
        try:
            for K in ks:
                for G in gs:
                    assert warp_encode_nib(data, K=K, G=G) == oracle.compress_block(data)
                    assert warp_encode_nib(data, cont=True, h5=True, K=K, G=G) == oracle.compress_block_cont(data)
                    assert warp_encode_nib(data, cont=False, h5=True, K=K, G=G) == oracle.compress_block_fresh_h5(data)
        finally:
            _TAG_BITS[0] = 4


def test_nibble_tags_cut_the_table_and_candidate_traffic():
    """On a JSON block the untagged matcher loads 32 table entries and 32 candidates per batch; first-2 verification behind
    4-bit tags needs ~1 table load and ~1 candidate load per batch."""
    data = corpus.tiled("compression_66k_JSON.txt", 65536).tobytes()
    s = {}
    assert warp_encode_nib(data, h5=True, stats=s) == oracle.compress_block_fresh_h5(data)
    assert s["tab_loads"] < 3 * s["batches"] and s["cand_loads"] < 3 * s["batches"], s


def test_nibble_partial_batches_happen():
    """The continued batch (forwarding removes the speculative hit, nothing before it matches) is a real path: low-entropy
    noise and word soup reach it a few times per block, still byte-identical to the oracle."""
    rng = np.random.default_rng(5)
    partial = 0
    for a in (2, 13, 15, 21):
        data = rng.integers(0, a, 20000, dtype=np.uint8).tobytes()
        words = [rng.integers(0, 256, int(rng.integers(3, 9)), dtype=np.uint8).tobytes() for _ in range(a * 8)]
        soup = b"".join(words[int(i)] for i in rng.integers(0, len(words), 4000))
        for x in (data, soup):
            s = {}
            assert warp_encode_nib(x, stats=s) == oracle.compress_block(x)
            partial += s["partial"]
    assert partial >= 5, partial


def test_tag_scheme_on_random_structured_inputs():
    """Property check of the shared-memory-tag scheme (all tag widths, warp and lane-group batches) on seeded random inputs
    built to collide: tiny alphabets, word soup, long runs with single-byte defects."""
    rng = np.random.default_rng(20240923)
    for trial in range(24):
        kind = trial % 3
        n = int(rng.integers(13, 9000))
        if kind == 0:
            data = rng.integers(0, int(rng.integers(2, 30)), n, dtype=np.uint8).tobytes()
        elif kind == 1:
            words = [rng.integers(0, 256, int(rng.integers(3, 9)), dtype=np.uint8).tobytes() for _ in range(int(rng.integers(4, 200)))]
            data = b"".join(words[int(i)] for i in rng.integers(0, len(words), n // 4 + 1))[:n]
        else:
            base = bytearray(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes() * (n // 2 + 1))[:n]
            for _ in range(int(rng.integers(0, 12))):
                base[int(rng.integers(0, n))] ^= 0x55
            data = bytes(base)
        want = (oracle.compress_block(data), oracle.compress_block_cont(data), oracle.compress_block_fresh_h5(data))
        for bits, G in ((2, 8), (2, 16), (4, 32), (8, 32)):
            _TAG_BITS[0] = bits
            try:
                assert warp_encode_nib(data, G=G) == want[0], (trial, bits, G)
                assert warp_encode_nib(data, cont=True, h5=True, G=G) == want[1], (trial, bits, G)
                assert warp_encode_nib(data, cont=False, h5=True, G=G) == want[2], (trial, bits, G)
            finally:
                _TAG_BITS[0] = 4

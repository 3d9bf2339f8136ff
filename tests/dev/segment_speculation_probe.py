"""Round-2 groundwork (CPU only): can a 4 MiB block be parsed in parallel segments and still be byte-identical?

A segment starting at S is parsed speculatively from S - W with an empty table (every slot "invalid", like a CONT block).
If the speculative parse re-synchronises with the true (sequential) parse early enough — same sequence boundaries from some
point P <= S - 65 536 on — then from S onwards both see the same table (entries older than P are out of the 16-bit offset
range for every cursor >= S) and produce the same sequences.  This script measures how long a warm-up W that takes on the
BASELINE config-4 data (hdfs.json) and on dickens.txt, by comparing the (match start, match end, offset) triples of the two
parses after S."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lz4_flex_b200 import corpus

MASK64 = (1 << 64) - 1


def h5(b, p):
    v = int.from_bytes(b[p:p + 8], "little")
    return ((((v << 24) & MASK64) * 889523592379) & MASK64) >> 52


def parse(b, start, stop, fresh):
    """Greedy parse of b[start:] (positions absolute), table empty; returns [(anchor, mpos, end, dist)] until cur >= stop."""
    n = len(b)
    tab = {}
    last_probe, lim = n - 12, n - 6
    seqs = []
    anchor = cur = start
    if fresh:
        tab[h5(b, 0)] = 0
        cur = 1
    while True:
        misses, nxt = 32, cur
        while True:
            step = misses >> 5; misses += 1
            cur = nxt; nxt += step
            if cur > last_probe or cur >= stop:
                return seqs
            s = h5(b, cur)
            cand = tab.get(s, -1)
            tab[s] = cur
            if cand < 0 or cur - cand > 65535:
                continue
            if b[cand:cand + 4] == b[cur:cur + 4]:
                break
        while cand > start and cur > anchor and b[cur - 1] == b[cand - 1]:
            cur -= 1; cand -= 1
        mpos = cur
        dist = cur - cand
        cur += 4; cand += 4
        while cur < lim and b[cur] == b[cand]:
            cur += 1; cand += 1
        tab[h5(b, cur - 2)] = cur - 2
        seqs.append((anchor, mpos, cur, dist))
        anchor = cur


def study(name, data, seg=256 << 10):
    true = parse(data, 0, len(data), True)
    by_anchor = {s[0]: i for i, s in enumerate(true)}
    print(f"{name}: {len(data)} bytes, {len(true)} sequences")
    for W in (4 << 10, 16 << 10, 64 << 10, 66 << 10, 72 << 10, 96 << 10, 128 << 10):
        ok = tot = 0
        worst = 0
        for S in range(seg, len(data) - seg + 1, seg):
            if S - W < 0:
                continue
            tot += 1
            spec = parse(data, S - W, S + seg, False)
            # first speculative sequence whose anchor is a true sequence anchor >= S ... must then agree to the end
            tail = [q for q in spec if q[0] >= S]
            if not tail or tail[0][0] not in by_anchor:
                continue
            i = by_anchor[tail[0][0]]
            same = all(true[i + k] == q for k, q in enumerate(tail) if i + k < len(true))
            # and the sequence that crosses S must be the same one in both parses
            cross_t = [q for q in true if q[0] < S <= q[2]]
            cross_s = [q for q in spec if q[0] < S <= q[2]]
            if same and cross_t == cross_s:
                ok += 1
            # resync distance: first position after S - W from which the anchors coincide for good
            sa = [q[0] for q in spec]
            k = len(sa) - 1
            while k >= 0 and sa[k] in by_anchor and true[by_anchor[sa[k]]] == spec[k]:
                k -= 1
            worst = max(worst, (sa[k + 1] - (S - W)) if k + 1 < len(sa) else W)
        print(f"  warm-up {W >> 10:4d} KiB: {ok}/{tot} segments identical after their start; worst resync distance {worst} B")


if __name__ == "__main__":
    study("hdfs.json[:4 MiB]", corpus.load("hdfs.json")[: 4 << 20])
    study("dickens.txt[:4 MiB]", corpus.load("dickens.txt")[: 4 << 20])

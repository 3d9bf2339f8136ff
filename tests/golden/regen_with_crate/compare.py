"""Compares the real crate's outputs (written by `cargo run --release -- <benches> <out>`) with
tests/golden/known_answers.json.  A full match turns this repo's "compress parity unpinned" into pinned:
    python tests/golden/regen_with_crate/compare.py <out>
Outputs the crate cannot produce through its public API (a block the frame encoder stores raw, inputs larger than
the biggest frame block) are reported as skipped."""
import hashlib
import json
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
ka = json.load(open(os.path.join(os.path.dirname(here), "known_answers.json")))
out = sys.argv[1]
bad = skipped = ok = 0
for e in ka["block"]:
    for mode in ("block_api", "frame_fresh", "frame_cont"):
        p = os.path.join(out, f"{e['name']}.{mode}.lz4")
        if not os.path.exists(p):
            skipped += 1
            print("skipped", e["name"], mode)
            continue
        b = open(p, "rb").read()
        good = len(b) == e[mode]["len"] and hashlib.sha256(b).hexdigest() == e[mode]["sha256"]
        ok += good
        bad += not good
        if not good:
            print("MISMATCH", e["name"], mode, len(b), e[mode]["len"])
for e in ka["frame"]:
    b = open(os.path.join(out, f"{e['name']}.frame.lz4"), "rb").read()
    good = len(b) == e["len"] and hashlib.sha256(b).hexdigest() == e["sha256"]
    ok += good
    bad += not good
    if not good:
        print("MISMATCH frame", e["name"], len(b), e["len"])
print(f"{ok} match, {bad} mismatch, {skipped} skipped")
sys.exit(1 if bad else 0)

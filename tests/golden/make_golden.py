"""Regenerates tests/golden/known_answers.json from the CPU oracle.

IMPORTANT provenance note: lz4_flex is Rust and cannot be built in this image, and its own tests pin no
compressed bytes.  These known-answers are therefore outputs of the C oracle that were
cross-checked against an independent restatement written during the survey (the sha256 values listed in
SURVEY.md §8c "Provisional known-answers": all of them reproduce).  They freeze today's behaviour so that
any later change to the oracle or the kernels that alters a single byte is caught."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from lz4_flex_b200 import corpus  # noqa: E402


def h(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out = {"provenance": __doc__, "block": [], "frame": []}
    inputs = {f: corpus.load(f) for f in ["compression_1k.txt", "compression_34k.txt", "compression_65k.txt",
                                           "compression_66k_JSON.txt", "dickens.txt"]}
    t = corpus.tiled("compression_66k_JSON.txt", 131072).tobytes()
    inputs["json_tiled_block0"] = t[:65536]
    inputs["json_tiled_block1"] = t[65536:]
    inputs["zeros_65536"] = bytes(65536)
    inputs["hdfs_first_4MiB"] = corpus.load("hdfs.json")[: 4 << 20]
    inputs["xorshift_65536"] = corpus.xorshift64star_bytes(65536).tobytes()
    for name, data in inputs.items():
        e = {"name": name, "input_len": len(data), "input_sha256": h(data)}
        for mode, fn in (("block_api", oracle.compress_block), ("frame_fresh", oracle.compress_block_fresh_h5),
                         ("frame_cont", oracle.compress_block_cont)):
            c = fn(data)
            e[mode] = {"len": len(c), "sha256": h(c)}
        out["block"].append(e)
    j = corpus.load("compression_66k_JSON.txt")
    d1m = corpus.load("dickens.txt")[: 1 << 20]
    for name, data, bsid, flags in [("json66k_auto", j, 0, 0), ("json66k_64k", j, 4, 0),
                                    ("json66k_64k_all_flags", j, 4, 7), ("dickens1M_64k", d1m, 4, 0),
                                    ("dickens1M_256k_checksums", d1m, 5, 3), ("empty_auto", b"", 0, 0)]:
        f = oracle.frame_compress(data, bsid, flags)
        out["frame"].append({"name": name, "block_size_id": bsid, "flags": flags, "input_sha256": h(data),
                             "input_len": len(data), "len": len(f), "sha256": h(f), "head": f[:16].hex()})
    json.dump(out, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    print("wrote", len(out["block"]), "block and", len(out["frame"]), "frame known-answers")


if __name__ == "__main__":
    main()

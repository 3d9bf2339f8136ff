"""Shared inputs for the dictionary tests (CPU oracle pins and GPU parity)."""
import numpy as np

from lz4_flex_b200 import corpus

# the reference's own dictionary test input (src/block/compress.rs:892-950)
REF_INPUT = bytes([10, 12, 14, 16, 18] * 4)


def dict_cases():
    """(name, input, dictionary) triples: reference unit-test inputs, corpus files, the table-layout switch at
    dict + input = 65 535, dictionaries beyond the 64 KiB window, tiny dictionaries, matches crossing the
    dictionary/input seam."""
    rng = np.random.default_rng(7)
    j = corpus.load("compression_66k_JSON.txt")
    d34 = corpus.load("compression_34k.txt")
    cases = [
        ("ref_test_dict", REF_INPUT, REF_INPUT),
        ("ref_no_panic_3byte_dict", REF_INPUT, bytes([10, 12, 14])),
        ("empty_dict", d34[:5000], b""),
        ("empty_input", b"", d34[:100]),
        ("tiny_input", b"abc", d34[:100]),
        ("dict4", d34[:3000], d34[:4]),
        ("dict7", d34[:3000], d34[:7]),
        ("dict8", d34[:3000], d34[:8]),
        ("text_self", d34[:20000], d34[:20000]),
        ("text_other_half", d34[17000:34000], d34[:17000]),
        ("json_small_u16", j[30000:50000], j[:30000]),                 # 50 000 < 65 535: u16 table, hash4
        ("switch_65534", j[:35534], j[35534:65534]),                   # dict + input = 65 534: last u16/hash4 size
        ("switch_65535", j[:35535], j[35535:65535]),                   # 65 535: hash5, positions still fit u16
        ("switch_65536", j[:35536], j[35536:65536]),
        ("switch_65537", j[:35537], j[35537:65537]),                   # u32 table
        ("json_full_dict_trim", j, j),                                 # dictionary > 64 KiB: last 65 536 bytes used
        ("big_input_small_dict", corpus.tiled("compression_66k_JSON.txt", 300000).tobytes(), j[:1000]),
        ("zeros", bytes(70000), bytes(70000)),
        ("low_entropy", rng.integers(0, 3, 40000, dtype=np.uint8).tobytes(), rng.integers(0, 3, 30000, dtype=np.uint8).tobytes()),
        ("random", rng.integers(0, 256, 20000, dtype=np.uint8).tobytes(), rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()),
    ]
    # the input continues the dictionary: matches start in the dictionary's tail and stop at its end
    base = rng.integers(0, 4, 9000, dtype=np.uint8).tobytes()
    cases.append(("seam", base[3000:9000] + base[2500:3500], base[:3000] + base[2000:3000]))
    for n in (13, 14, 20, 64, 300, 4097):
        cases.append((f"small_{n}", (b"abcdefghij" * 500)[:n], b"abcdefghij" * 7))
    return cases

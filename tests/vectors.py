"""Vectors shared by the oracle tests (CPU) and the GPU parity tests.

DECODE_KATS are the reference's own hand-written decoder tests with their expected results
(src/block/decompress.rs:530-623, identical set in src/block/decompress_safe.rs:392-485).
NO_PANIC and ROUNDTRIP inputs come from tests/tests.rs (line numbers beside each)."""
OK, C_SMALL, SMALL, LIT_OOB, EAB, OFF0, OFF_OOB = 0, 1, 2, 3, 4, 5, 6

# (name, stream, capacity, status, output or None, expected(OutputTooSmall) or None)
DECODE_KATS = [
    ("all_literal              decompress.rs:535-537", [0x30, ord("a"), ord("4"), ord("9")], 3, OK, b"a49", None),
    ("incomplete_input empty   decompress.rs:542-545", [], 255, EAB, None, None),
    ("incomplete_input 0xF0    decompress.rs:546-549", [0xF0], 255, EAB, None, None),
    ("incomplete_input 0x0F,0  decompress.rs:552-555", [0x0F, 0], 255, EAB, None, None),
    ("incomplete_input 0F,1,0  decompress.rs:556-559", [0x0F, 1, 0], 255, EAB, None, None),
    ("literal oob              decompress.rs:566-569", [0x40, ord("a"), 1, 0], 4, LIT_OOB, None, None),
    ("output too small lit     decompress.rs:571-577", [0x20, ord("a"), ord("a"), 1, 0], 1, SMALL, None, 2),
    ("output too small match   decompress.rs:579-585", [0x10, ord("a"), 1, 0], 4, SMALL, None, 5),
    ("offset oob fast path     decompress.rs:588-594", [0x0E, 255] + [0] * 18, 256, OFF_OOB, None, None),
    ("offset oob slow path     decompress.rs:605-608", [0x0F, 1, 0, 1, 0x70] + [0] * 20, 256, OFF_OOB, None, None),
    ("offset oob after lits    decompress.rs:610-613", [0x40, 0, 0, 0, 0, 255, 0, 0x70] + [0] * 20, 256, OFF_OOB, None, None),
    ("offset zero              decompress.rs:618-621", [0x0E, 0, 0, 0x70] + [0] * 20, 256, OFF0, None, None),
]

# decompress_size_prepended inputs that must not crash (tests/tests.rs:326-350, :507-526)
NO_PANIC = [
    [122, 1, 0, 1, 0, 10, 1, 0],
    [44, 251, 49, 0, 0, 0, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 16, 0, 0, 0, 0, 0, 0, 0, 0],
    [7, 0, 0, 0, 0, 0, 0, 11, 0, 0, 7, 16, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 1, 0, 0],
    [0, 61, 0, 0, 0, 7, 0],
    [8, 0, 0, 0, 4, 0, 0, 0],
    [39, 0, 0, 0, 0, 0, 0, 237, 0, 0, 0, 0, 0, 0, 16, 0, 0, 4, 0, 0, 0, 39, 32, 0, 2, 0, 162, 5, 36, 0, 0, 0, 0, 7, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 10, 0, 0, 10],
]

# round-trip inputs (tests/tests.rs:353-495, :529)
ROUNDTRIP = [
    b"AAAAAAAAAAAAAAAAAAAAAAAAaAAAAAAAAAAAAAAAAAAAAAAAA",
    b"AAAAAAAAAAAAAAAAAAAAAAAABBBBBBBBBaAAAAAAAAAAAAAAAAAAAAAAAA",
    b"AAAAAAAAAAAAAAAAAAAAAAAABBBBBBBBBaAAAAAAAAAAAAAAAAAAAAAAAABBBBBBBBBa",
    b"AAAAAAAAAAAZZZZZZZZAAAAAAAA",
    b"to live or not to live", b"Love is a wonderful terrible thing",
    b"There is nothing either good or bad, but thinking makes it so.", b"I burn, I pine, I perish.",
    b"Save water, it doesn't grow on trees.", b"The panda bear has an amazing black-and-white fur.",
    b"The average panda eats as much as 9 to 14 kg of bamboo shoots a day.",
    b"You are 60% water. Save 60% of yourself!", b"To cute to die! Save the red panda!",
    b"as6yhol.;jrew5tyuikbfewedfyjltre22459ba", b"jhflkdjshaf9p8u89ybkvjsdbfkhvg4ut08yfrr",
    b"ahhd", b"ahd", b"x-29", b"x", b"k", b".", b"ajsdh", b"aaaaaa", b"aaaaaabcbcbcbc", b"", b"\0" * 13,
    bytes([8, 6] + [0] * 288 + [46, 0, 0, 8, 0, 138]),                                           # bug_fuzz   :432
    bytes([122] + [0] * 15 + [8] + [0] * 81 + [65, 0, 0, 128, 10, 1, 10, 1, 0, 122]),              # bug_fuzz_2 :447
    bytes([36, 16, 0, 0, 79, 177, 176, 176, 171, 1, 0, 255, 207, 79, 79, 79, 79, 79, 1, 1, 49, 0, 16, 0, 79, 79, 79,
           79, 79, 1, 0, 255, 36, 79, 79, 79, 79, 79, 1, 0, 255, 207, 79, 79, 79, 79, 79, 1, 0, 255, 255, 255, 255,
           255, 255, 255, 255, 255, 255, 8, 207, 1, 207, 207, 79, 199, 79, 79, 40, 79, 1, 1, 1, 1, 1, 1] + [15] * 27 +
          [79, 15, 15, 14] + [15] * 16 + [61] + [15] * 10 + [0, 48, 45, 0, 1, 0, 0, 1, 0]),         # bug_fuzz_3 :457
    bytes([147]),                                                                                # bug_fuzz_4 :471
    bytes([255, 255, 255, 255, 253, 235, 156, 140, 8, 0, 140, 45, 169, 0, 27, 128, 48, 0, 140, 0, 0, 255, 255, 255,
           253, 235, 156, 140, 8, 61, 255, 255, 255, 255, 65, 239, 254]),                          # buf_fuzz_5 :476
    bytes([181, 181, 181, 181, 181, 147, 147, 147, 0, 0, 255, 218, 44, 0, 177, 44, 0, 233, 177, 74, 85, 47, 95, 146,
           189, 177, 1, 0, 255, 2, 109, 180, 255, 255, 0, 0, 0, 181, 181, 181, 147, 147, 147, 0, 0, 255, 218, 146,
           146, 181, 0, 0, 181]),                                                                # bug_fuzz_6 :486
    bytes(30000),                                                                                # so_many_zeros :529
]

"""GPU parity of the external-dictionary block API: compress_into_with_dict /
compress_with_dict / compress_prepend_size_with_dict / decompress_into_with_dict / decompress_with_dict /
decompress_size_prepended_with_dict (reference src/block/compress.rs:554-583, 610-616, 685-694;
src/block/decompress.rs:85-109, 462-528), through the C ABI, byte-identical to the oracle."""
import numpy as np
import pytest

import oracle
from dict_cases import REF_INPUT, dict_cases
from lz4_flex_b200 import block, corpus, errors

pytestmark = pytest.mark.gpu
CASES = dict_cases()


@pytest.mark.parametrize("name,data,dct", CASES, ids=[c[0] for c in CASES])
def test_dict_bytes_match_oracle(ctx, name, data, dct):
    c = block.compress_with_dict(data, dct, ctx)
    assert c == oracle.compress_with_dict(data, dct)
    assert block.decompress_with_dict(c, len(data), dct, ctx) == data
    p = block.compress_prepend_size_with_dict(data, dct, ctx)
    assert p == len(data).to_bytes(4, "little") + c
    assert block.decompress_size_prepended_with_dict(p, dct, ctx) == data


def test_reference_dict_tests(ctx):
    # compress.rs:892-911 test_dict: a dictionary equal to the input shrinks the output and round-trips
    c = block.compress_with_dict(REF_INPUT, REF_INPUT, ctx)
    assert len(c) < len(block.compress(REF_INPUT, ctx))
    out = np.zeros(len(REF_INPUT), dtype=np.uint8)
    n = block.decompress_into_with_dict(c, out, REF_INPUT, ctx)
    assert out[:n].tobytes() == REF_INPUT
    # compress.rs:913-919 test_dict_no_panic
    block.compress_with_dict(REF_INPUT, bytes([10, 12, 14]), ctx)
    # decompress.rs:593-601: offset beyond dictionary + output
    with pytest.raises(errors.OffsetOutOfBounds):
        block.decompress_with_dict(bytes([0x0E, 255, 0, 0x70, 0, 0, 0, 0, 0, 0, 0]), 256, bytes(250), ctx)
    # tests.rs:504: garbage with itself as dictionary must not crash
    for v in (b"", b"\x00", bytes([0x40, 1, 0, 0, 0, 2, 0, 0x70] + [0] * 20)):
        st, o, e1, _ = oracle.decompress_with_dict(v, 64, v)
        outs, status, exp = block.decompress_blocks_with_dict([v], [64], v, ctx) if v else ([b""], [4], [0])
        assert int(status[0]) == st and (st != 0 or outs[0] == o)


def test_dict_match_crossing(ctx):
    # compress.rs:921-950: a match that starts in the dictionary and runs on into the output.  Hand-made stream:
    # token 0x0B -> no literals, match length 15; offset 5 with an empty output reaches 5 bytes into the dictionary,
    # the remaining 10 bytes repeat the output's first 5 bytes (LZ77 overlap); then 5 literals.
    dct = b"0123456789ABCDE"
    stream = bytes([0x0B, 5, 0, 0x50]) + b"vwxyz"
    want = b"ABCDE" + b"ABCDE" + b"ABCDE" + b"vwxyz"
    st, o, _, _ = oracle.decompress_with_dict(stream, 64, dct)
    assert st == 0 and o == want
    assert block.decompress_with_dict(stream, 64, dct, ctx) == want
    # offset larger than the dictionary + output: OffsetOutOfBounds, same as the oracle
    bad = bytes([0x0B, 16, 0, 0x50]) + b"vwxyz"
    assert oracle.decompress_with_dict(bad, 64, dct)[0] == oracle.ERR_OFFSET_OOB
    with pytest.raises(errors.OffsetOutOfBounds):
        block.decompress_with_dict(bad, 64, dct, ctx)


def test_dict_decode_errors_and_foreign_streams(ctx):
    rng = np.random.default_rng(3)
    d34 = corpus.load("compression_34k.txt")
    dct = d34[:6000]
    good = oracle.compress_with_dict(d34[6000:12000], dct)
    streams, caps = [], []
    for t in range(300):
        s = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, len(s)))] = int(rng.integers(0, 256))
        cut = int(rng.integers(1, len(s) + 1)) if t % 3 == 0 else len(s)
        streams.append(bytes(s[:cut]))
        caps.append(int(rng.integers(0, 7000)))
    outs, status, exp = block.decompress_blocks_with_dict(streams, caps, dct, ctx)
    for i, (s, cap) in enumerate(zip(streams, caps)):
        st, o, e1, _ = oracle.decompress_with_dict(s, cap, dct)
        assert int(status[i]) == st, i
        if st == 0:
            assert outs[i] == o, i
        if st == oracle.ERR_OUTPUT_TOO_SMALL:
            assert int(exp[i]) == e1, i


def test_dict_batch_shared_dictionary(ctx):
    j = corpus.tiled("compression_66k_JSON.txt", 400000).tobytes()
    dct = j[:20000]
    rng = np.random.default_rng(11)
    blocks = []
    for _ in range(200):
        a = int(rng.integers(0, len(j) - 70000))
        blocks.append(j[a: a + int(rng.integers(0, 70000))])
    comp = block.compress_blocks_with_dict(blocks, dct, ctx)
    for b, c in zip(blocks, comp):
        assert c == oracle.compress_with_dict(b, dct)
    outs, status, _ = block.decompress_blocks_with_dict(comp, [max(len(b), 1) for b in blocks], dct, ctx)
    assert not status.any()
    assert all(o == b for o, b in zip(outs, blocks))


def test_compress_into_with_table(ctx):
    """block::compress_into_with_table (compress.rs:744-766): the table variant selects the hash; a Small table handed an
    input of >= 65 535 bytes becomes Large and stays Large."""
    small_in = corpus.load("compression_34k.txt")
    big_in = corpus.load("compression_66k_JSON.txt")
    out = np.zeros(block.get_maximum_output_size(len(big_in)), dtype=np.uint8)
    t = block.CompressTable.small()
    n = block.compress_into_with_table(small_in, out, t, ctx)
    assert out[:n].tobytes() == oracle.compress_block(small_in) and t.kind == block.CompressTable.SMALL
    n = block.compress_into_with_table(big_in, out, t, ctx)                 # upgrade
    assert out[:n].tobytes() == oracle.compress_block(big_in) and t.kind == block.CompressTable.LARGE
    n = block.compress_into_with_table(small_in, out, t, ctx)               # stays Large: 5-byte hash on a small input
    assert out[:n].tobytes() == oracle.compress_block_fresh_h5(small_in)
    assert out[:n].tobytes() != oracle.compress_block(small_in)
    t2 = block.CompressTable.large()
    n = block.compress_into_with_table(b"", out, t2, ctx)
    assert out[:n].tobytes() == oracle.compress_block(b"")
    with pytest.raises(errors.CompressOutputTooSmall):
        block.compress_into_with_table(small_in, np.zeros(10, dtype=np.uint8), t2, ctx)

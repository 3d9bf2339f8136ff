"""GPU parity tests of the frame path (independent blocks): FrameEncoder/FrameDecoder mirrors and the one-shot
C-ABI calls against the oracle's frame encoder — whole frames byte-identical."""
import io

import numpy as np
import pytest

import oracle
from lz4_flex_b200 import corpus, errors, frame
from lz4_flex_b200.frame import BlockMode, BlockSize, FrameDecoder, FrameEncoder, FrameInfo

pytestmark = pytest.mark.gpu


def _flags(info: FrameInfo) -> int:
    return (oracle.F_BLOCK_CHECKSUMS if info.block_checksums else 0) | \
           (oracle.F_CONTENT_CHECKSUM if info.content_checksum else 0) | \
           (oracle.F_CONTENT_SIZE if info.content_size is not None else 0)


@pytest.mark.parametrize("bs", [BlockSize.Auto, BlockSize.Max64KB, BlockSize.Max256KB, BlockSize.Max1MB, BlockSize.Max4MB])
@pytest.mark.parametrize("checks", [False, True])
def test_one_shot_frames_byte_identical(ctx, bs, checks):
    for data in (b"", b"a", corpus.load("compression_1k.txt"), corpus.load("compression_66k_JSON.txt"),
                 corpus.tiled("compression_66k_JSON.txt", 5 * 65536 + 777).tobytes(),
                 corpus.load("dickens.txt")[: (9 << 20) + 5], bytes(300000),
                 corpus.adversarial_blocks(6, 0.5).tobytes()):
        info = FrameInfo(block_size=bs, block_checksums=checks, content_checksum=checks,
                         content_size=len(data) if checks else None)
        f = frame.compress_frame(data, info, ctx)
        assert f == oracle.frame_compress(data, int(bs), _flags(info)), (bs, checks, len(data))
        assert frame.decompress_frame(f, ctx) == data
        assert oracle.frame_decompress(f, len(data) + 1)[:2] == (0, data)


def test_frame_encoder_write_patterns(ctx):
    """io::Write behaviour: chunked writes, flush() producing short blocks (stream-offset phase shift),
    multi-frame reuse, empty input."""
    data = corpus.tiled("compression_66k_JSON.txt", 700000).tobytes()
    # one write_all
    w = io.BytesIO()
    FrameEncoder(w, FrameInfo(block_size=BlockSize.Max64KB), ctx).write(data)
    # not finished: last block still buffered
    enc = FrameEncoder(io.BytesIO(), FrameInfo(block_size=BlockSize.Max64KB), ctx)
    for i in range(0, len(data), 1000):
        enc.write(data[i:i + 1000])
    assert enc.finish().getvalue() == oracle.frame_compress(data, 4)
    # Auto block size from the first write (header.rs:57-67)
    enc = FrameEncoder(io.BytesIO(), None, ctx)
    enc.write(data[:70000]); enc.write(data[70000:])
    assert enc.frame_info().block_size == BlockSize.Max256KB
    got = enc.finish().getvalue()
    assert got[:7] == bytes([4, 0x22, 0x4D, 0x18, 0x60, 0x50, 0xFB])
    assert frame.decompress_frame(got, ctx) == data
    # flush every 100 000 bytes
    enc = FrameEncoder(io.BytesIO(), FrameInfo(block_size=BlockSize.Max64KB), ctx)
    for i in range(0, len(data), 100000):
        enc.write(data[i:i + 100000]); enc.flush()
    assert enc.finish().getvalue() == oracle.frame_compress(data, 4, 0, 100000)
    # empty input (compress.rs:176-181)
    assert FrameEncoder(io.BytesIO(), None, ctx).finish().getvalue() == oracle.frame_compress(b"")
    # auto_finish + second frame from the same encoder
    sink = io.BytesIO()
    enc = FrameEncoder(sink, FrameInfo(block_size=BlockSize.Max64KB), ctx)
    enc.write(data[:200000]); enc.try_finish()
    enc.write(data[200000:]); enc.try_finish()
    assert sink.getvalue() == oracle.frame_compress(data[:200000], 4) + oracle.frame_compress(data[200000:], 4)
    dec = FrameDecoder(io.BytesIO(sink.getvalue()), ctx)                      # concatenated frames: tests.rs:633-647 —
    assert dec.read_to_end() == data[:200000]                                 # read_to_end stops at each EndMark
    assert dec.read_to_end() == data[200000:]
    assert dec.read_to_end() == b""
    assert frame.decompress_frame(sink.getvalue(), ctx) == data               # the join-all convenience
    out1, used1, err1 = frame.decompress_next_frame(sink.getvalue(), ctx)     # C ABI: one frame per call + consumed bytes
    assert (out1, err1) == (data[:200000], None) and used1 == len(oracle.frame_compress(data[:200000], 4))
    with FrameEncoder(io.BytesIO(), FrameInfo(block_size=BlockSize.Max64KB), ctx).auto_finish() as af:
        af.write(data[:5000])
        inner = af.encoder.get_ref()
    assert inner.getvalue() == oracle.frame_compress(data[:5000], 4)


def test_frame_decoder_reads_and_errors(ctx):
    a = corpus.load("compression_34k.txt")
    f = oracle.frame_compress(a, 4, 7)
    dec = FrameDecoder(io.BytesIO(f), ctx)
    parts = []
    while True:
        p = dec.read(5000)
        if not p:
            break
        parts.append(p)
    assert b"".join(parts) == a
    # checksum corruption (tests.rs:650-700)
    bad = bytearray(oracle.frame_compress(a, 4, oracle.F_BLOCK_CHECKSUMS)); bad[30] ^= 1
    with pytest.raises(errors.BlockChecksumError):
        FrameDecoder(io.BytesIO(bytes(bad)), ctx).read()
    bad = bytearray(oracle.frame_compress(a, 4, oracle.F_CONTENT_CHECKSUM)); bad[-1] ^= 1
    with pytest.raises(errors.ContentChecksumError):
        FrameDecoder(io.BytesIO(bytes(bad)), ctx).read()
    # content size mismatch (tests.rs:721-734)
    k = corpus.load("compression_1k.txt")
    with pytest.raises(errors.ContentLengthError) as e:
        enc = FrameEncoder(io.BytesIO(), FrameInfo(content_size=3), ctx); enc.write(k); enc.finish()
    assert (e.value.expected, e.value.actual) == (3, 725)
    # header problems
    with pytest.raises(errors.WrongMagicNumber):
        frame.decompress_frame(b"\x00\x01\x02\x03\x04\x05\x06\x07", ctx)
    with pytest.raises(errors.HeaderChecksumError):
        frame.decompress_frame(bytes([4, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x83]) + b"\0\0\0\0", ctx)
    with pytest.raises(errors.SkippableFrame):
        frame.decompress_frame(bytes([0x50, 0x2A, 0x4D, 0x18, 4, 0, 0, 0, 1, 2, 3, 4]), ctx)
    with pytest.raises(errors.LinkedBlocksUnsupported):
        frame.compress_frame(b"abc", FrameInfo(block_mode=BlockMode.Linked), ctx)
    # corrupt block inside a frame: data before it is still delivered, then DecompressionError
    d = corpus.tiled("compression_66k_JSON.txt", 3 * 65536).tobytes()
    g = bytearray(oracle.frame_compress(d, 4))
    first = int.from_bytes(g[7:11], "little")
    second_payload = 7 + 4 + first + 4
    g[second_payload + 3] = 0; g[second_payload + 4] = 0      # poke an offset to zero somewhere early
    out, err = frame.decompress_frame(bytes(g), ctx, partial=True)
    es, eo, eb = oracle.frame_decompress(bytes(g), len(d))
    if es == 0:
        assert err is None and out == eo
    else:
        assert isinstance(err, errors.DecompressionError) and out[:65536] == d[:65536]


def test_frame_decoder_content_length_and_partial_delivery(ctx):
    """A header whose content_size is SMALLER than the real content: every byte is still delivered, then
    ContentLengthError{expected, actual} (decompress.rs:312-321) — the bound must not trust the header."""
    a = corpus.load("compression_34k.txt")
    f = bytearray(oracle.frame_compress(a, 4, oracle.F_CONTENT_SIZE))
    assert int.from_bytes(f[6:14], "little") == len(a)
    f[6:14] = (100).to_bytes(8, "little")
    f[14] = (frame.xxh32(bytes(f[4:14])) >> 8) & 0xFF                        # re-seal the header checksum
    out, used, err = frame.decompress_next_frame(bytes(f), ctx)
    assert out == a and used == len(f)
    assert isinstance(err, errors.ContentLengthError) and (err.expected, err.actual) == (100, len(a))
    dec = FrameDecoder(io.BytesIO(bytes(f)), ctx)
    got = bytearray()
    with pytest.raises(errors.ContentLengthError) as e:
        while True:
            p = dec.read(7000)
            if not p:
                break
            got += p
    assert bytes(got) == a and (e.value.expected, e.value.actual) == (100, len(a))
    # larger than the real content
    f[6:14] = (10**9).to_bytes(8, "little")
    f[14] = (frame.xxh32(bytes(f[4:14])) >> 8) & 0xFF
    out, used, err = frame.decompress_next_frame(bytes(f), ctx)
    assert out == a and (err.expected, err.actual) == (10**9, len(a))
    # a block that needs more than its frame's block size: OutputTooSmall{expected, actual} inside DecompressionError
    big = oracle.compress_block(bytes(70000))
    g = bytes([4, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x82]) + len(big).to_bytes(4, "little") + big + b"\0\0\0\0"
    out, used, err = frame.decompress_next_frame(g, ctx)
    assert isinstance(err, errors.DecompressionError) and isinstance(err.inner, errors.OutputTooSmall)
    assert err.inner.actual == 65536 and err.inner.expected > 65536
    # end-of-data conventions of read_frame_info (decompress.rs:113-128)
    assert frame.decompress_next_frame(b"", ctx) == (b"", 0, None)
    assert frame.decompress_next_frame(b"\x04\x22\x4d\x18", ctx) == (b"", 4, None)
    assert isinstance(frame.decompress_next_frame(b"\x04\x22", ctx)[2], errors.IoError)


def test_frame_decoder_bounded_memory(ctx):
    """20 000 tiny blocks in a 4 MiB-block frame (a FrameEncoder stream with frequent flush()): decoded with a device
    budget far below #blocks x block size, in several groups, byte-exact; and streamed through FrameDecoder in small
    groups from a reader that returns short reads."""
    from lz4_flex_b200 import _native
    data = corpus.tiled("compression_66k_JSON.txt", 20000 * 150).tobytes()
    f = oracle.frame_compress(data, 7, oracle.F_CONTENT_CHECKSUM, 150)      # flush every 150 bytes
    bound = ctypes_size()
    _native.lib().lz4b200_frame_decoded_bound(f, len(f), bound)
    assert bound.value < 20000 * (4 << 20) // 100                            # not 20 000 x 4 MiB
    _native.lib().lz4b200_ctx_set_frame_budget(ctx.handle, 1 << 20)
    try:
        assert frame.decompress_frame(f, ctx) == data
    finally:
        _native.lib().lz4b200_ctx_set_frame_budget(ctx.handle, 256 << 20)

    class Dribble(io.RawIOBase):
        def __init__(self, b):
            self.b, self.p, self.k = b, 0, 0
        def readable(self):
            return True
        def read(self, n=-1):
            self.k += 1
            n = len(self.b) - self.p if n is None or n < 0 else min(n, 1 + self.k % 977)
            out = self.b[self.p:self.p + n]
            self.p += len(out)
            return out

    big = corpus.tiled("dickens.txt", 5 << 20).tobytes()
    ff = oracle.frame_compress(big, 4, 7) + f
    dec = FrameDecoder(Dribble(ff), ctx, group_bytes=1 << 20)
    assert dec.read_to_end() == big
    got = bytearray()
    while True:
        p = dec.read(100000)
        if not p:
            break
        got += p
    assert bytes(got) == data and dec.read() == b""


def ctypes_size():
    import ctypes
    return ctypes.c_size_t(0)


def test_legacy_frame_fixture(ctx):
    # tests/tests.rs:741-745 — two 8 MiB-class blocks produced by the C lz4 CLI
    assert frame.decompress_frame(corpus.load("dickens.lz4"), ctx) == corpus.load("dickens.txt")


def test_stored_blocks(ctx):
    """Incompressible blocks are stored raw (frame/compress.rs:301-306) and decoded from the raw copy."""
    rnd = corpus.xorshift64star_bytes(3 * 65536 + 100).tobytes()
    f = frame.compress_frame(rnd, FrameInfo(block_size=BlockSize.Max64KB), ctx)
    assert f == oracle.frame_compress(rnd, 4)
    assert int.from_bytes(f[7:11], "little") == (65536 | 0x80000000)
    assert frame.decompress_frame(f, ctx) == rnd


def test_config4_reduced_epochs(ctx):
    """BASELINE config 4 at reduced size: hdfs log data, 4 MiB blocks, frame format — 24 blocks (96 MiB) via
    the device block-range entry point, split into two 'ranks' whose outputs concatenate to the oracle frame."""
    import ctypes as C
    import torch
    from lz4_flex_b200 import _native
    bs, nblk = 4 << 20, 24
    data = corpus.tiled("hdfs.json", nblk * bs)
    L = _native.lib()
    dev = torch.device("cuda", 0)
    parts = []
    for r in range(2):
        lo, hi = r * nblk // 2, (r + 1) * nblk // 2
        d_in = torch.from_numpy(data[lo * bs: hi * bs]).to(dev)
        bound = L.lz4b200_frame_blocks_bound(d_in.numel(), bs)
        d_out = torch.zeros(bound, dtype=torch.uint8, device=dev)
        d_total = torch.zeros(1, dtype=torch.int64, device=dev)
        st = L.lz4b200_frame_compress_blocks_device(ctx.handle, d_in.data_ptr(), d_in.numel(), bs, lo, d_out.data_ptr(),
                                                    bound, d_total.data_ptr(), None,
                                                    torch.cuda.current_stream().cuda_stream)
        assert st == 0
        torch.cuda.synchronize()
        parts.append(d_out[: int(d_total.item())].cpu().numpy().tobytes())
    got = FrameInfo(block_size=BlockSize.Max4MB).header_bytes() + b"".join(parts) + b"\0\0\0\0"
    assert got == oracle.frame_compress(data, 7)
    assert frame.decompress_frame(got, ctx) == data.tobytes()


def test_config4_fresh_block_at_511(ctx):
    """The table reposition of FrameEncoder (frame/compress.rs:266-271) on hardware: with 4 MiB blocks the block at
    absolute index 511 starts a fresh table epoch (FRESH parse), its neighbours continue one (CONT).  Blocks 509..513
    of the config-4 stream through the device block-range entry point vs the oracle's persistent table driven from
    the same stream offset."""
    import torch
    from lz4_flex_b200 import _native
    bs, first, nblk = 4 << 20, 509, 5
    h = np.frombuffer(corpus.load("hdfs.json"), dtype=np.uint8)
    idx = (np.arange(nblk * bs, dtype=np.int64) + first * bs) % h.size          # config 4: tiled by absolute offset
    data = h[idx]
    table = oracle.FrameTable()
    table.offset = first * bs                                                    # an empty table at this offset == CONT
    want = b""
    fresh_seen = []
    for k in range(nblk):
        blk = data[k * bs:(k + 1) * bs].tobytes()
        before = table.offset
        c = table.compress(blk, bs)
        fresh_seen.append(table.offset - len(blk) != before)                     # offset was reset before this block
        assert len(c) < len(blk)
        want += len(c).to_bytes(4, "little") + c
    assert fresh_seen == [False, False, True, False, False]
    assert oracle.compress_block_fresh_h5(data[2 * bs:3 * bs].tobytes()) in want      # block 511 is a FRESH parse
    L = _native.lib()
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(data).to(dev)
    bound = L.lz4b200_frame_blocks_bound(d_in.numel(), bs)
    d_out = torch.zeros(bound, dtype=torch.uint8, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    st = L.lz4b200_frame_compress_blocks_device(ctx.handle, d_in.data_ptr(), d_in.numel(), bs, first, d_out.data_ptr(),
                                                bound, d_total.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    assert st == 0
    torch.cuda.synchronize()
    got = d_out[: int(d_total.item())].cpu().numpy().tobytes()
    assert got == want


# ---- BlockMode::Linked frames, decode side -----------------------------------------------------------

def _linked_fixtures():
    import hashlib
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_linked_frames import sources
    src = sources()
    d = os.path.join(os.path.dirname(__file__), "golden", "linked")
    for m in json.load(open(os.path.join(d, "manifest.json"))):
        f = open(os.path.join(d, m["file"]), "rb").read()
        assert hashlib.sha256(f).hexdigest() == m["frame_sha256"]
        yield m["file"], f, src[m["source"]]


def test_linked_frames_decode_on_gpu(ctx):
    """liblz4-written frames with linked blocks: the chain of blocks is resolved on the device
    (lz4_decompress_blocks_linked), stored blocks included; content / block checksums verified."""
    for name, f, want in _linked_fixtures():
        assert frame.decompress_frame(f, ctx=ctx) == want, name
    # a linked frame followed by an independent frame and another linked one
    fx = list(_linked_fixtures())
    indep = oracle.frame_compress(fx[0][2][:150000], 4)
    cat = fx[3][1] + indep + fx[0][1]
    assert frame.decompress_frame(cat, ctx=ctx) == fx[3][2] + fx[0][2][:150000] + fx[0][2]


def test_linked_frame_errors_match_oracle(ctx):
    import numpy as np
    name, f, want = list(_linked_fixtures())[4]               # no checksums: corruption reaches the block decoder
    rng = np.random.default_rng(5)
    for t in range(40):
        bad = bytearray(f)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(16, len(bad)))] = int(rng.integers(0, 256))
        cut = len(bad) if t % 4 else int(rng.integers(8, len(bad)))
        blob = bytes(bad[:cut])
        st, got, berr = oracle.frame_decompress(blob, len(want) + 70000)
        out, err = frame.decompress_frame(blob, ctx=ctx, partial=True)
        if st == 0:
            assert err is None and out == got, t
        else:
            want_err = errors.error_from_status(st, berr, "")
            assert type(err) is type(want_err), (t, st, berr, err)
            if isinstance(err, errors.DecompressionError):
                assert type(err.inner) is type(want_err.inner), (t, berr, err.inner)


def test_large_64k_block_frame_uses_global_table_encoder(ctx):
    """A frame of 3 800 x 64 KiB blocks: more blocks in one launch than the shared-memory-table encoder keeps in flight,
    so the frame path runs the global-table kernel — here with CONT-mode blocks (every block but the first continues the
    frame's table epoch).  Byte-identical to the oracle frame; round trip exact."""
    import hashlib
    data = corpus.tiled("compression_66k_JSON.txt", 3800 * 65536)
    f = frame.compress_frame(data, frame.FrameInfo(block_size=frame.BlockSize.Max64KB, content_checksum=True), ctx=ctx)
    want = oracle.frame_compress(data.tobytes(), 4, oracle.F_CONTENT_CHECKSUM)
    assert len(f) == len(want) and hashlib.sha256(f).digest() == hashlib.sha256(want).digest()
    assert frame.decompress_frame(f, ctx=ctx) == data.tobytes()

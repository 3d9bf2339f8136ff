"""Host-side logic on CPU: FrameEncoder block cutting / mode flags / container assembly, and the multi-rank
sharding + gather (world_size 2 and 3, gloo).  The GPU batch compressor is replaced by an oracle-backed stand-in
that honours the per-block flags, so these tests pin the HOST code; the kernels are pinned by test_gpu_*."""
import io
import os
import socket

import numpy as np
import pytest

import oracle
from lz4_flex_b200 import block, corpus, frame, sharded
from lz4_flex_b200.frame import BlockSize, FrameEncoder, FrameInfo


def _oracle_compress_blocks(blocks, flags=None, ctx=None):
    out = []
    for i, b in enumerate(blocks):
        fl = 0 if flags is None else flags[i]
        if fl & block.BLOCK_CONT:
            out.append(oracle.compress_block_cont(b))
        elif fl & block.BLOCK_HASH5_ALWAYS:
            out.append(oracle.compress_block_fresh_h5(b))
        else:
            out.append(oracle.compress_block(b))
    return out


@pytest.fixture
def oracle_backend(monkeypatch):
    monkeypatch.setattr(block, "compress_blocks", _oracle_compress_blocks)
    monkeypatch.setattr(block, "default_context", lambda device=None: None)


def test_frame_encoder_matches_oracle_frames(oracle_backend):
    data = corpus.tiled("compression_66k_JSON.txt", 450000).tobytes()
    for bsid in (4, 5, 7):
        for flags in (0, 7):
            info = FrameInfo(block_size=BlockSize(bsid), block_checksums=bool(flags & 1),
                             content_checksum=bool(flags & 2), content_size=len(data) if flags & 4 else None)
            enc = FrameEncoder(io.BytesIO(), info)
            for i in range(0, len(data), 33333):
                enc.write(data[i:i + 33333])
            assert enc.finish().getvalue() == oracle.frame_compress(data, bsid, flags)
    # Auto block size, flush pattern, tiny batches (queue drains mid-stream)
    enc = FrameEncoder(io.BytesIO(), None, batch_bytes=100000)
    enc.write(data)
    assert enc.finish().getvalue() == oracle.frame_compress(data)
    enc = FrameEncoder(io.BytesIO(), FrameInfo(block_size=BlockSize.Max64KB))
    for i in range(0, len(data), 100000):
        enc.write(data[i:i + 100000]); enc.flush()
    assert enc.finish().getvalue() == oracle.frame_compress(data, 4, 0, 100000)
    assert FrameEncoder(io.BytesIO()).finish().getvalue() == oracle.frame_compress(b"")


def test_fresh_epochs_follow_the_reposition_rule(oracle_backend):
    """Block k is FRESH iff the stream offset was reset before it (frame/compress.rs:266-271): with 4 MiB blocks the
    epoch is 511 blocks; exercised here without compressing 2 GiB by driving _write_block's bookkeeping."""
    enc = FrameEncoder(io.BytesIO(), FrameInfo(block_size=BlockSize.Max4MB))
    enc._is_frame_open = True
    seen = []
    enc._drain = lambda: None
    enc._batch_bytes = 1 << 62
    blk = b"\0" * 16                                  # content does not matter for the flag logic
    for k in range(1100):
        enc._src += blk
        # pretend the block is full-size for offset accounting
        before = len(enc._queue)
        enc._write_block()
        enc._stream_offset += (4 << 20) - len(blk)
        seen.append(enc._queue[before][1])
    fresh = [k for k, f in enumerate(seen) if not (f & block.BLOCK_CONT)]
    assert fresh == [0, 511, 1022]
    assert all(f & block.BLOCK_HASH5_ALWAYS for f in seen)


def test_partition_helpers():
    assert [sharded.block_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert sharded.byte_range(10 * 65536 + 5, 65536, 3, 4) == (8 * 65536, 10 * 65536 + 5)
    assert sharded.byte_range(100, 65536, 1, 2) == (100, 100) or sharded.byte_range(100, 65536, 1, 2) == (0, 100)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, bsid, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = corpus.tiled("compression_66k_JSON.txt", total)
    bs = BlockSize(bsid).get_size()
    lo, hi = sharded.byte_range(total, bs, rank, world)
    local = torch.from_numpy(data[lo:hi].copy())

    def stand_in(local, block_size, first_block):
        raw = local.numpy().tobytes()
        out = bytearray()
        nb = -(-len(raw) // block_size)
        for i in range(nb):
            blk = raw[i * block_size:(i + 1) * block_size]
            k = first_block + i
            c = oracle.compress_block_fresh_h5(blk) if k == 0 else oracle.compress_block_cont(blk)
            if len(c) < len(blk):
                out += len(c).to_bytes(4, "little") + c
            else:
                out += (len(blk) | 0x80000000).to_bytes(4, "little") + blk
        t = torch.frombuffer(bytearray(out) or bytearray(1), dtype=torch.uint8)
        return t, len(out)

    fr = sharded.frame_compress_sharded(local, total, bsid, rank, world, compress_range=stand_in)
    if rank == 0:
        q.put(fr.numpy().tobytes())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total,bsid", [(2, 9 * 65536 + 321, 4), (3, 5 * 65536, 4), (2, 700000, 5)])
def test_sharded_frame_gloo(world, total, bsid):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, bsid, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = corpus.tiled("compression_66k_JSON.txt", total)
    assert got == oracle.frame_compress(data, bsid)


def test_numa_helpers_degrade_gracefully():
    """lz4_flex_b200/numa.py: cpulist parsing, and binding is a no-op (not an error) where the GPU's node is unknown."""
    from lz4_flex_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("") == []
    info = numa.bind_to_gpu_node(0)
    assert info["bound"] in (True, False)
    numa.restore_affinity(info)

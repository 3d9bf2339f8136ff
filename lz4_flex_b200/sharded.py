"""Multi-GPU frame path: the independent blocks of one LZ4 frame sharded over the ranks of one node
(SURVEY.md §8e).  One process per GPU; torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in
the CPU tests).  Rank r owns the contiguous block range [r*B/G, (r+1)*B/G): block modes depend only on the
absolute block index, so every rank compresses its range with no communication; the only exchange step is
the gather of the packed per-rank chunks to rank 0, which prepends the frame header and appends the end mark.
"""
from __future__ import annotations

from typing import Callable, Optional

from .frame import BlockSize, FrameInfo


def block_range(nblocks: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, order-preserving partition of block indices."""
    return rank * nblocks // world, (rank + 1) * nblocks // world


def byte_range(total_len: int, block_size: int, rank: int, world: int) -> tuple[int, int]:
    nblocks = -(-total_len // block_size)
    lo, hi = block_range(nblocks, rank, world)
    return min(lo * block_size, total_len), min(hi * block_size, total_len)


def compress_range_device(d_in, block_size: int, first_block: int, ctx):
    """This rank's blocks -> packed [BlockInfo | payload]* on the device (one C-ABI call, stream-ordered).
    Returns (uint8 CUDA tensor, number of valid bytes)."""
    import torch
    from . import _native
    L = _native.lib()
    n = d_in.numel()
    bound = L.lz4b200_frame_blocks_bound(n, block_size)
    d_out = torch.empty(max(bound, 1), dtype=torch.uint8, device=d_in.device)
    d_total = torch.zeros(1, dtype=torch.int64, device=d_in.device)
    st = L.lz4b200_frame_compress_blocks_device(ctx.handle, d_in.data_ptr() if n else None, n, block_size, first_block,
                                                d_out.data_ptr(), bound, d_total.data_ptr(), None,
                                                torch.cuda.current_stream().cuda_stream)
    if st != 0:
        from .errors import error_from_status
        raise error_from_status(st, detail=ctx.last_cuda_error())
    return d_out, d_total


def gather_frame(part, part_len: int, info: FrameInfo, rank: int, world: int, group=None):
    """Exchange step: all ranks learn the chunk sizes (all_gather of one int64), then every rank sends its chunk
    to rank 0, which lays out header | chunk_0 | ... | chunk_{G-1} | EndMark.  `part` is a uint8 tensor (CUDA
    with NCCL, CPU with gloo).  Returns the frame tensor on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    sizes = [torch.zeros(1, dtype=torch.int64, device=part.device) for _ in range(world)]
    mine = torch.tensor([part_len], dtype=torch.int64, device=part.device)
    if world > 1:
        dist.all_gather(sizes, mine, group=group)
    else:
        sizes = [mine]
    sizes = [int(s.item()) for s in sizes]
    if rank != 0:
        if part_len:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, part[:part_len].contiguous(), 0, group)]):
                req.wait()
        return None
    header = info.header_bytes()
    total = len(header) + sum(sizes) + 4
    frame = torch.empty(total, dtype=torch.uint8, device=part.device)
    frame[: len(header)] = torch.frombuffer(bytearray(header), dtype=torch.uint8).to(part.device)
    pos = len(header)
    frame[pos: pos + sizes[0]] = part[: sizes[0]]
    pos += sizes[0]
    ops = []
    for r in range(1, world):
        if sizes[r]:
            ops.append(dist.P2POp(dist.irecv, frame[pos: pos + sizes[r]], r, group))
        pos += sizes[r]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    frame[pos: pos + 4] = 0
    return frame


def frame_compress_sharded(local, total_len: int, block_size_id: int, rank: int, world: int, ctx=None, group=None,
                           compress_range: Optional[Callable] = None):
    """Compress this rank's byte range `local` (uint8 tensor holding bytes byte_range(...)) of a `total_len`-byte
    stream as part of ONE frame with `block_size_id` blocks; returns the whole frame (tensor) on rank 0.

    `compress_range(local, block_size, first_block)` -> (tensor, nbytes) may be injected (the CPU tests inject
    an oracle-backed stand-in; production uses the CUDA C-ABI call)."""
    info = FrameInfo(block_size=BlockSize(block_size_id))
    bs = info.block_size.get_size()
    nblocks = -(-total_len // bs)
    lo, _ = block_range(nblocks, rank, world)
    if compress_range is None:
        from . import block as _block
        ctx = ctx or _block.default_context()
        part, d_total = compress_range_device(local, bs, lo, ctx)
        part_len = int(d_total.item())
    else:
        part, part_len = compress_range(local, bs, lo)
    return gather_frame(part, part_len, info, rank, world, group)

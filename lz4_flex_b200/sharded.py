"""Multi-GPU frame path: the independent blocks of one LZ4 frame sharded over the ranks of one node
(SURVEY.md §8e).  One process per GPU; torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in
the CPU tests).  Rank r owns the contiguous block range [r*B/G, (r+1)*B/G): block modes depend only on the
absolute block index, so every rank compresses its range with no communication; the only exchange step is
the gather of the packed per-rank chunks to rank 0, which prepends the frame header and appends the end mark.

Two implementations of the exchange step:
  * PeerFrameGather (GPUs of one node, the default on CUDA): rank 0 owns the frame buffer and every rank maps it
    over NVLink (CUDA IPC).  Per step: an 8-byte-per-rank all_gather of the packed sizes, a device-side prefix sum,
    and then each rank's PACK KERNEL writes its [BlockInfo | payload] segments straight into rank 0's buffer at its
    offset — the stores of the pack kernel are the transfer; there is no staging copy, no host synchronisation and no
    payload collective.  One small all_reduce tells rank 0 that every rank's stores have landed.
  * gather_frame (any backend; used by the gloo CPU tests and as a fallback): all_gather of sizes + send/recv of the
    packed chunks.
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


class PeerMappingUnavailable(RuntimeError):
    """Raised on EVERY rank when the frame buffer cannot be shared over CUDA IPC; callers fall back to StagedFrameGather."""


class _DevPtr:
    """A raw device allocation as something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerFrameGather:
    """The sharded frame encoder of one node with the gather fused into the pack kernel (module docstring).

    Every rank calls step(d_in) with its byte range resident on its GPU; after the step rank 0 holds the frame in
    `self.frame` (uint8 CUDA tensor view of the shared buffer) and its length in `self.frame_len` (0-d int64 CUDA
    tensor).  Nothing in step() synchronises with the host."""

    transport = "pack kernel stores into rank 0's buffer over NVLink (CUDA IPC mapping); no host sync"

    def __init__(self, total_len: int, block_size_id: int, rank: int, world: int, ctx, group=None, base_block: int = 0):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _native
        self.L = _native.lib()
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.info = FrameInfo(block_size=BlockSize(block_size_id))
        self.bs = self.info.block_size.get_size()
        self.total_len = total_len
        nblocks = -(-total_len // self.bs)
        self.first_block = base_block + block_range(nblocks, rank, world)[0]   # base_block: the stream continues an earlier one
        header = self.info.header_bytes()
        self.header_len = len(header)
        cap = self.header_len + 4
        for r in range(world):
            lo, hi = byte_range(total_len, self.bs, r, world)
            cap += self.L.lz4b200_frame_blocks_bound(hi - lo, self.bs)
        self.cap = cap
        dev = torch.device("cuda", ctx.device)
        self.dev = dev
        handle = (C.c_uint8 * 64)()
        p = C.c_void_p()
        self._owned = self._mapped = None
        # Every rank must reach the same verdict about the mapping, whatever fails where: rank 0 broadcasts its handle
        # (or None), the others try to map it, and a MIN all_reduce of the outcome decides for all.
        box = [None]
        if rank == 0:
            if self.L.lz4b200_peer_alloc(ctx.handle, cap, C.byref(p), handle) == 0:
                self._owned = p.value
                box = [bytes(handle)]
        ok = 1
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
            if box[0] is None:
                ok = 0
            elif rank != 0:
                h = (C.c_uint8 * 64).from_buffer_copy(box[0])
                if self.L.lz4b200_peer_open(ctx.handle, h, C.byref(p)) == 0:
                    self._mapped = p.value
                else:
                    ok = 0
            t = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", ctx.device))
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            ok = int(t.item())
        elif box[0] is None:
            ok = 0
        if not ok:
            if self._mapped:
                self.L.lz4b200_peer_close(ctx.handle, self._mapped)
            if self._owned:
                self.L.lz4b200_peer_free(ctx.handle, self._owned)
            self._owned = self._mapped = None
            raise PeerMappingUnavailable("CUDA IPC mapping of rank 0's frame buffer failed: " + ctx.last_cuda_error())
        self.frame_ptr = p.value
        self.d_total = torch.zeros(1, dtype=torch.int64, device=dev)
        self.d_all = torch.zeros(world, dtype=torch.int64, device=dev)
        self.d_off = torch.zeros(1, dtype=torch.int64, device=dev)
        self.d_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.frame_len = torch.zeros((), dtype=torch.int64, device=dev)
        self.frame = None
        if rank == 0:
            self.frame = torch.as_tensor(_DevPtr(self.frame_ptr, cap), device=dev)
            self.frame[: self.header_len] = torch.frombuffer(bytearray(header), dtype=torch.uint8).to(dev)
            self._four = torch.arange(4, device=dev)
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step(self, d_in, timed: bool = False):
        import torch
        import torch.distributed as dist
        s = torch.cuda.current_stream().cuda_stream
        n = d_in.numel()
        if timed:
            self.ev[0].record()
        st = self.L.lz4b200_frame_range_compress(self.ctx.handle, d_in.data_ptr() if n else None, n, self.bs, self.first_block,
                                                 self.d_total.data_ptr(), None, s)
        if st != 0:
            from .errors import error_from_status
            raise error_from_status(st, detail=self.ctx.last_cuda_error())
        if timed:
            self.ev[1].record()
        if self.world > 1:
            dist.all_gather_into_tensor(self.d_all, self.d_total, group=self.group)
        else:
            self.d_all.copy_(self.d_total)
        ends = torch.cumsum(self.d_all, 0)
        self.d_off.copy_((ends[self.rank] - self.d_all[self.rank] + self.header_len).reshape(1))
        st = self.L.lz4b200_frame_range_pack(self.ctx.handle, self.frame_ptr, self.d_off.data_ptr(), s)
        if st != 0:
            from .errors import error_from_status
            raise error_from_status(st, detail=self.ctx.last_cuda_error())
        if self.world > 1:
            dist.all_reduce(self.d_flag, group=self.group)          # every rank's pack kernel has completed
        if self.rank == 0:
            end = ends[-1] + self.header_len
            self.frame[end + self._four] = 0                        # EndMark
            self.frame_len = end + 4
        if timed:
            self.ev[2].record()

    def timings_ms(self):
        """(compress kernel ms, exchange + pack ms) of the last timed step (after a synchronize)."""
        return self.ev[0].elapsed_time(self.ev[1]), self.ev[1].elapsed_time(self.ev[2])

    def result(self):
        """Rank 0: the frame bytes as a CUDA tensor slice (synchronises to read the length)."""
        if self.rank != 0:
            return None
        return self.frame[: int(self.frame_len.item())]

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self._mapped:
            self.L.lz4b200_peer_close(self.ctx.handle, self._mapped)
            self._mapped = None
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)                          # nobody still maps the buffer
        if self._owned:
            self.frame = None
            self.L.lz4b200_peer_free(self.ctx.handle, self._owned)
            self._owned = None


class StagedFrameGather:
    """Same interface as PeerFrameGather with the backend-neutral exchange (all_gather of sizes + send/recv of the packed
    chunks, gather_frame): the fallback when the peer mapping is unavailable.  Synchronises with the host once per step."""

    transport = "nccl send/recv of packed chunks (host-synchronised sizes)"

    def __init__(self, total_len: int, block_size_id: int, rank: int, world: int, ctx, group=None, base_block: int = 0):
        import torch
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.info = FrameInfo(block_size=BlockSize(block_size_id))
        self.bs = self.info.block_size.get_size()
        nblocks = -(-total_len // self.bs)
        self.first_block = base_block + block_range(nblocks, rank, world)[0]
        self.frame = None
        self.frame_len = torch.zeros((), dtype=torch.int64)
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step(self, d_in, timed: bool = False):
        import torch
        if timed:
            self.ev[0].record()
        part, d_total = compress_range_device(d_in, self.bs, self.first_block, self.ctx)
        if timed:
            self.ev[1].record()
        self.frame = gather_frame(part, int(d_total.item()), self.info, self.rank, self.world, self.group)
        if self.rank == 0:
            self.frame_len = torch.tensor(self.frame.numel(), dtype=torch.int64)
        if timed:
            self.ev[2].record()

    def timings_ms(self):
        return self.ev[0].elapsed_time(self.ev[1]), self.ev[1].elapsed_time(self.ev[2])

    def result(self):
        return self.frame if self.rank == 0 else None

    def close(self):
        self.frame = None


def make_frame_gather(total_len: int, block_size_id: int, rank: int, world: int, ctx, group=None, base_block: int = 0):
    """PeerFrameGather when the ranks can share rank 0's frame buffer, StagedFrameGather otherwise (same verdict on every rank)."""
    import os
    if os.environ.get("LZ4B200_SHARDED_STAGED") != "1":
        try:
            return PeerFrameGather(total_len, block_size_id, rank, world, ctx, group, base_block)
        except PeerMappingUnavailable:
            pass
    return StagedFrameGather(total_len, block_size_id, rank, world, ctx, group, base_block)

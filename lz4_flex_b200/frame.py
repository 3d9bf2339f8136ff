"""lz4_flex::frame, B200 edition (encoder: independent-block mode; decoder: independent and linked blocks).

FrameInfo / BlockSize / BlockMode / FrameEncoder / FrameDecoder with the reference's names and behaviour
(src/frame/header.rs:39-192, src/frame/compress.rs:62-438, src/frame/decompress.rs:48-422).  The container
(header, BlockInfo words, checksums, end mark) is host bookkeeping; the blocks themselves are compressed
and decompressed in GPU batches through the C ABI.  Block boundaries, the stored-raw rule and — crucially —
each block's parse mode (FRESH for the first block of a table epoch, CONT afterwards) follow
FrameEncoder::write/write_block exactly, so the produced frame is byte-identical to lz4_flex's.

BlockMode::Linked chains every block to its frame's earlier output.  The DECODER handles it on the device
(lz4_decompress_blocks_linked: a block waits for the predecessors an offset reaches into), so frames written by
`lz4` / LZ4F / pyarrow with their default linked blocks decode; the ENCODER refuses it (LinkedBlocksUnsupported):
a linked encode is one serial chain per frame, which is not what a GPU batch path is for.
"""
from __future__ import annotations

import ctypes as C
import enum
import io
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native, block as _block
from .errors import (ContentLengthError, LinkedBlocksUnsupported, error_from_status)

WINDOW_SIZE = 64 * 1024
_REPOSITION_LIMIT = 0xFFFFFFFF // 2          # u32::MAX as usize / 2  (frame/compress.rs:266)


class BlockSize(enum.IntEnum):
    """frame::BlockSize (header.rs:39-53)."""
    Auto = 0
    Max64KB = 4
    Max256KB = 5
    Max1MB = 6
    Max4MB = 7
    Max8MB = 8

    @staticmethod
    def from_buf_length(buf_len: int) -> "BlockSize":
        """header.rs:57-67"""
        if buf_len > 256 * 1024:
            return BlockSize.Max4MB
        if buf_len > 64 * 1024:
            return BlockSize.Max256KB
        return BlockSize.Max64KB

    def get_size(self) -> int:
        return {4: 64 << 10, 5: 256 << 10, 6: 1 << 20, 7: 4 << 20, 8: 8 << 20}[int(self)]


class BlockMode(enum.Enum):
    """frame::BlockMode (header.rs:83-91)."""
    Independent = 0
    Linked = 1


@dataclass
class FrameInfo:
    """frame::FrameInfo (header.rs:130-192)."""
    content_size: Optional[int] = None
    dict_id: Optional[int] = None
    block_size: BlockSize = BlockSize.Auto
    block_mode: BlockMode = BlockMode.Independent
    block_checksums: bool = False
    content_checksum: bool = False
    legacy_frame: bool = False

    def to_c(self) -> _native.FrameInfoC:
        return _native.FrameInfoC(int(self.block_size), int(self.block_checksums), int(self.content_checksum),
                                  int(self.content_size is not None), int(self.content_size or 0),
                                  int(self.block_mode == BlockMode.Linked), 0)

    def header_bytes(self) -> bytes:
        """FrameInfo::write (header.rs:232-275)."""
        buf = (C.c_uint8 * 19)()
        ci = self.to_c()
        n = _native.lib().lz4b200_frame_write_header(C.byref(ci), buf, 19)
        return bytes(buf[:n])


class _Xxh32:
    def __init__(self, seed: int = 0):
        self._st = (C.c_uint8 * 48)()
        _native.lib().lz4b200_xxh32_reset(self._st, seed)

    def update(self, data: bytes):
        a = np.frombuffer(data, dtype=np.uint8)
        _native.lib().lz4b200_xxh32_update(self._st, a.ctypes.data if a.size else None, a.size)

    def digest(self) -> int:
        return _native.lib().lz4b200_xxh32_digest(self._st)


def xxh32(data, seed: int = 0) -> int:
    a = np.frombuffer(data, dtype=np.uint8)
    return _native.lib().lz4b200_xxh32(a.ctypes.data if a.size else None, a.size, seed)


# -----------------------------------------------------------------------------------------------------
# one-shot helpers over the C ABI
# -----------------------------------------------------------------------------------------------------

def compress_frame(data, frame_info: FrameInfo | None = None, ctx: _block.Context | None = None) -> bytes:
    """FrameEncoder::with_frame_info(info, Vec::new()); write_all(data); finish() — one C-ABI call."""
    ctx = ctx or _block.default_context()
    info = frame_info or FrameInfo()
    if info.block_mode == BlockMode.Linked:
        raise LinkedBlocksUnsupported()
    src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    ci = info.to_c()
    cap = _native.lib().lz4b200_frame_bound(src.size, C.byref(ci))
    out = np.empty(cap, dtype=np.uint8)
    w = C.c_size_t(0)
    st = _native.lib().lz4b200_frame_compress(ctx.handle, src.ctypes.data if src.size else None, src.size,
                                              C.byref(ci), src.size, out.ctypes.data, cap, C.byref(w))
    if st != 0:
        raise error_from_status(st, detail=ctx.last_cuda_error())
    return out[: w.value].tobytes()


def decompress_frame(data, ctx: _block.Context | None = None, partial: bool = False):
    """FrameDecoder::new(data).read_to_end() — one C-ABI call over all concatenated frames.
    With partial=True returns (bytes_before_error, exception_or_None) instead of raising."""
    ctx = ctx or _block.default_context()
    src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    bound = C.c_size_t(0)
    _native.lib().lz4b200_frame_decoded_bound(src.ctypes.data if src.size else None, src.size, C.byref(bound))
    out = np.empty(max(bound.value, 1), dtype=np.uint8)
    w, bs = C.c_size_t(0), C.c_int(0)
    st = _native.lib().lz4b200_frame_decompress(ctx.handle, src.ctypes.data if src.size else None, src.size,
                                                out.ctypes.data, bound.value, C.byref(w), C.byref(bs))
    err = None if st == 0 else error_from_status(st, bs.value, ctx.last_cuda_error())
    if partial:
        return out[: w.value].tobytes(), err
    if err is not None:
        raise err
    return out[: w.value].tobytes()


# -----------------------------------------------------------------------------------------------------
# FrameEncoder
# -----------------------------------------------------------------------------------------------------

class FrameEncoder:
    """frame::FrameEncoder<W> (compress.rs:62-404): a writer that LZ4-frame-compresses into `w`.

    Blocks are cut exactly where the reference cuts them; they are queued and compressed on the GPU in
    batches of up to `batch_bytes` of input (and at flush()/finish()), so the bytes arriving in `w` are
    identical to lz4_flex's, only later."""

    def __init__(self, w, frame_info: FrameInfo | None = None, ctx: _block.Context | None = None,
                 batch_bytes: int = 256 << 20):
        self.w = w
        self._info = frame_info or FrameInfo()
        if self._info.block_mode == BlockMode.Linked:
            raise LinkedBlocksUnsupported()
        self._ctx = ctx
        self._batch_bytes = batch_bytes
        self._src = bytearray()            # the reference's `src` buffer (one block being filled)
        self._queue: list[tuple[bytes, int]] = []   # (block, flags) not yet compressed
        self._queued_bytes = 0
        self._stream_offset = 0            # src_stream_offset (compress.rs:80)
        self._content_len = 0
        self._hasher = _Xxh32(0)
        self._is_frame_open = False
        self._data_to_frame_written = False

    # -- constructors with the reference's names -----------------------------------------------------
    @classmethod
    def new(cls, w, **kw) -> "FrameEncoder":
        return cls(w, FrameInfo(), **kw)

    @classmethod
    def with_frame_info(cls, frame_info: FrameInfo, w, **kw) -> "FrameEncoder":
        return cls(w, frame_info, **kw)

    def frame_info(self) -> FrameInfo:
        return self._info

    def get_ref(self):
        return self.w

    def get_mut(self):
        return self.w

    def into_inner(self):
        return self.w

    # -- io::Write -------------------------------------------------------------------------------------
    def write(self, buf) -> int:
        buf = bytes(buf)
        if not self._is_frame_open and buf:
            self._begin_frame(len(buf))
        bs = self._info.block_size.get_size() if self._info.block_size != BlockSize.Auto else 0
        view = memoryview(buf)
        while len(view):
            room = bs - len(self._src)
            if room == 0:
                self._write_block()
                continue
            take = min(room, len(view))
            self._src += view[:take]
            view = view[take:]
        return len(buf)

    def write_all(self, buf):
        self.write(buf)

    def flush(self):
        if self._src:
            self._write_block()
        self._drain()
        if hasattr(self.w, "flush"):
            try:
                self.w.flush()
            except Exception:
                pass

    def try_finish(self):
        """compress.rs:166-181"""
        if self._src:
            self._write_block()
        self._drain()
        if not self._is_frame_open and not self._data_to_frame_written:
            self._begin_frame(0)
        self._end_frame()
        self._data_to_frame_written = True

    def finish(self):
        self.try_finish()
        return self.w

    def auto_finish(self) -> "AutoFinishEncoder":
        return AutoFinishEncoder(self)

    # -- internals -----------------------------------------------------------------------------------
    def _begin_frame(self, buf_len: int):
        """compress.rs:234-258"""
        self._is_frame_open = True
        if self._info.block_size == BlockSize.Auto:
            self._info.block_size = BlockSize.from_buf_length(buf_len)
        self.w.write(self._info.header_bytes())
        if self._content_len != 0:
            self._content_len = 0
            self._stream_offset = 0
            self._src.clear()
            self._hasher = _Xxh32(0)

    def _end_frame(self):
        """compress.rs:209-230"""
        self._is_frame_open = False
        if self._info.content_size is not None and self._info.content_size != self._content_len:
            raise ContentLengthError(self._info.content_size, self._content_len)
        self.w.write(b"\0\0\0\0")
        if self._info.content_checksum:
            self.w.write(self._hasher.digest().to_bytes(4, "little"))

    def _write_block(self):
        """compress.rs:261-371, minus the compression itself (queued)."""
        bs = self._info.block_size.get_size()
        if self._stream_offset + bs + WINDOW_SIZE >= _REPOSITION_LIMIT:       # compress.rs:266-271
            self._stream_offset = 0
        flags = _block.BLOCK_HASH5_ALWAYS | (_block.BLOCK_CONT if self._stream_offset else _block.BLOCK_FRESH)
        blk = bytes(self._src)
        self._src.clear()
        self._queue.append((blk, flags))
        self._queued_bytes += len(blk)
        if self._info.content_checksum:
            self._hasher.update(blk)
        self._content_len += len(blk)
        self._stream_offset += len(blk)
        if self._queued_bytes >= self._batch_bytes:
            self._drain()

    def _drain(self):
        if not self._queue:
            return
        ctx = self._ctx or _block.default_context()
        blocks = [b for b, _ in self._queue]
        flags = [f for _, f in self._queue]
        comp = _block.compress_blocks(blocks, flags, ctx)
        out = bytearray()
        for raw, c in zip(blocks, comp):
            if len(c) < len(raw):                                            # compress.rs:301-306
                out += len(c).to_bytes(4, "little"); payload = c
            else:
                out += (len(raw) | 0x80000000).to_bytes(4, "little"); payload = raw
            out += payload
            if self._info.block_checksums:
                out += xxh32(payload).to_bytes(4, "little")
        self.w.write(bytes(out))
        self._queue.clear()
        self._queued_bytes = 0


class AutoFinishEncoder:
    """frame::AutoFinishEncoder (compress.rs:417-438): finishes the stream when the scope ends."""

    def __init__(self, enc: FrameEncoder):
        self.encoder = enc

    def write(self, buf) -> int:
        return self.encoder.write(buf)

    def flush(self):
        self.encoder.flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.encoder is not None:
            try:
                self.encoder.try_finish()
            except Exception:
                if exc[0] is None:
                    raise
            self.encoder = None
        return False


# -----------------------------------------------------------------------------------------------------
# FrameDecoder
# -----------------------------------------------------------------------------------------------------

class FrameDecoder(io.RawIOBase):
    """frame::FrameDecoder<R> (decompress.rs:48-422): a reader that decompresses LZ4 frames from `r`.

    All concatenated frames of the underlying reader are decoded in one GPU batch at the first read();
    bytes before a corrupt block are still delivered, and the error surfaces on the read that reaches it,
    like the reference's block-at-a-time reader."""

    def __init__(self, r, ctx: _block.Context | None = None):
        super().__init__()
        self.r = r
        self._ctx = ctx
        self._buf: bytes | None = None
        self._pos = 0
        self._err: Exception | None = None

    @classmethod
    def new(cls, r, **kw) -> "FrameDecoder":
        return cls(r, **kw)

    def get_ref(self):
        return self.r

    def get_mut(self):
        return self.r

    def into_inner(self):
        return self.r

    def readable(self) -> bool:
        return True

    def _fill(self):
        if self._buf is None:
            data = self.r.read()
            self._buf, self._err = decompress_frame(data, self._ctx, partial=True) if data else (b"", None)

    def read(self, size: int = -1) -> bytes:
        self._fill()
        if self._pos >= len(self._buf):
            if self._err is not None:
                e, self._err = self._err, None
                raise e
            return b""
        end = len(self._buf) if size is None or size < 0 else min(len(self._buf), self._pos + size)
        out = self._buf[self._pos:end]
        self._pos = end
        if (size is None or size < 0) and self._err is not None:
            e, self._err = self._err, None
            raise e
        return out

    def readinto(self, b) -> int:
        data = self.read(len(b))
        b[: len(data)] = data
        return len(data)

    def read_to_end(self) -> bytes:
        return self.read(-1)

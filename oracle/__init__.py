"""ctypes view of the CPU parity oracle (oracle/lz4_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (lz4_flex_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblz4oracle.so")

OK = 0
ERR_COMPRESS_OUTPUT_TOO_SMALL = 1
ERR_OUTPUT_TOO_SMALL = 2
ERR_LITERAL_OOB = 3
ERR_EXPECTED_ANOTHER_BYTE = 4
ERR_OFFSET_ZERO = 5
ERR_OFFSET_OOB = 6

FERR_DECOMPRESSION = 101
FERR_WRONG_MAGIC = 102
FERR_RESERVED_BITS = 103
FERR_UNSUPPORTED_VERSION = 104
FERR_UNSUPPORTED_BLOCKSIZE = 105
FERR_HEADER_CHECKSUM = 106
FERR_BLOCK_CHECKSUM = 107
FERR_CONTENT_CHECKSUM = 108
FERR_CONTENT_LENGTH = 109
FERR_BLOCK_TOO_BIG = 110
FERR_SKIPPABLE = 111
FERR_DICTIONARY = 112
FERR_IO_EOF = 113
FERR_LINKED_UNSUPPORTED = 114
FERR_OUTPUT_FULL = 115

F_BLOCK_CHECKSUMS = 1
F_CONTENT_CHECKSUM = 2
F_CONTENT_SIZE = 4
F_LINKED = 8


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make -C oracle)."""
    srcs = [os.path.join(_HERE, f) for f in ("lz4_oracle.c", "lz4_cpu_baseline.c", "lz4_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, sz, i64 = C.c_void_p, C.c_size_t, C.c_int64
        L.lz4o_max_output_size.restype = sz
        L.lz4o_max_output_size.argtypes = [sz]
        L.lz4o_compress_block.restype = i64
        L.lz4o_compress_block.argtypes = [u8p, sz, u8p, sz]
        L.lz4o_compress_block_with_table.restype = i64
        L.lz4o_compress_block_with_table.argtypes = [u8p, sz, u8p, sz, C.c_void_p, C.c_uint64]
        L.lz4o_compress_block_dict.restype = i64
        L.lz4o_compress_block_dict.argtypes = [u8p, sz, u8p, sz, u8p, sz]
        L.lz4o_decompress_block_dict.restype = C.c_int
        L.lz4o_decompress_block_dict.argtypes = [u8p, sz, u8p, sz, u8p, sz] + [C.POINTER(sz)] * 3
        L.lz4o_compress_prepend_size.restype = i64
        L.lz4o_compress_prepend_size.argtypes = [u8p, sz, u8p, sz]
        L.lz4o_decompress_block.restype = C.c_int
        L.lz4o_decompress_block.argtypes = [u8p, sz, u8p, sz] + [C.POINTER(sz)] * 3
        L.lz4o_decompress_size_prepended.restype = C.c_int
        L.lz4o_decompress_size_prepended.argtypes = [u8p, sz, u8p, sz] + [C.POINTER(sz)] * 3
        L.lz4o_xxh32.restype = C.c_uint32
        L.lz4o_xxh32.argtypes = [u8p, sz, C.c_uint32]
        L.lz4o_frame_bound.restype = sz
        L.lz4o_frame_bound.argtypes = [sz, C.c_int]
        L.lz4o_frame_compress.restype = i64
        L.lz4o_frame_compress.argtypes = [u8p, sz, C.c_int, C.c_uint, sz, u8p, sz]
        L.lz4o_frame_decompress.restype = C.c_int
        L.lz4o_frame_decompress.argtypes = [u8p, sz, u8p, sz, C.POINTER(sz), C.POINTER(C.c_int)]
        for name in ("lz4o_compress_batch", "lz4o_decompress_batch"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [u8p] * 8 + [sz, C.c_int]
        # performance-oriented CPU baseline + persistent pool (oracle/lz4_cpu_baseline.c)
        L.lz4cpu_compress_block.restype = i64
        L.lz4cpu_compress_block.argtypes = [u8p, sz, u8p, sz, C.c_void_p]
        L.lz4cpu_decompress_block.restype = C.c_int
        L.lz4cpu_decompress_block.argtypes = [u8p, sz, u8p, sz] + [C.POINTER(sz)] * 3
        L.lz4cpu_pool_create.restype = C.c_void_p
        L.lz4cpu_pool_create.argtypes = [C.c_int]
        L.lz4cpu_pool_destroy.restype = None
        L.lz4cpu_pool_destroy.argtypes = [C.c_void_p]
        L.lz4cpu_pool_copy.restype = None
        L.lz4cpu_pool_copy.argtypes = [C.c_void_p, u8p, u8p, sz]
        L.lz4cpu_pool_run.restype = None
        L.lz4cpu_pool_run.argtypes = [C.c_void_p, C.c_int] + [u8p] * 8 + [sz]
        _lib = L
    return _lib


def _buf(b):
    """(keepalive, address, length) for bytes-like / numpy input."""
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    a = np.ascontiguousarray(a)
    return a, a.ctypes.data, a.size


def max_output_size(n: int) -> int:
    return lib().lz4o_max_output_size(n)


def compress_block(data) -> bytes:
    """block::compress (no size prefix): src/block/compress.rs:679."""
    a, p, n = _buf(data)
    out = np.empty(max_output_size(n), dtype=np.uint8)
    r = lib().lz4o_compress_block(p, n, out.ctypes.data, out.size)
    assert r >= 0
    return out[:r].tobytes()


def compress_into(data, cap: int):
    """block::compress_into with an explicit output capacity; returns bytes or None (OutputTooSmall)."""
    a, p, n = _buf(data)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    r = lib().lz4o_compress_block(p, n, out.ctypes.data, cap)
    return None if r < 0 else out[:r].tobytes()


def compress_with_dict(data, dict_data) -> bytes:
    """block::compress_with_dict / compress_into_with_dict: src/block/compress.rs:554-583, 610-616, 685-687."""
    a, p, n = _buf(data)
    d, dp, dn = _buf(dict_data)
    out = np.empty(max_output_size(n), dtype=np.uint8)
    r = lib().lz4o_compress_block_dict(p, n, dp, dn, out.ctypes.data, out.size)
    assert r >= 0
    return out[:r].tobytes()


def decompress_with_dict(data, cap: int, dict_data):
    """block::decompress_into_with_dict (src/block/decompress.rs:462-468).  Returns (status, bytes, expected, actual)."""
    a, p, n = _buf(data)
    d, dp, dn = _buf(dict_data)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    w, e1, e2 = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    st = lib().lz4o_decompress_block_dict(p, n, dp, dn, out.ctypes.data, cap, C.byref(w), C.byref(e1), C.byref(e2))
    return st, out[: w.value].tobytes(), e1.value, e2.value


def compress_prepend_size(data) -> bytes:
    a, p, n = _buf(data)
    out = np.empty(max_output_size(n) + 4, dtype=np.uint8)
    r = lib().lz4o_compress_prepend_size(p, n, out.ctypes.data, out.size)
    assert r >= 0
    return out[:r].tobytes()


class FrameTable:
    """The persistent HashTable4K + stream offset of a FrameEncoder (frame/compress.rs:77-83)."""

    def __init__(self):
        self.table = np.zeros(4096, dtype=np.uint32)
        self.offset = 0

    def compress(self, data, max_block_size: int) -> bytes:
        # frame/compress.rs:266-271
        if self.offset + max_block_size + 65536 >= 0x7FFFFFFF:
            self.table[:] = np.where(self.table > self.offset, self.table - self.offset, 0)
            self.offset = 0
        a, p, n = _buf(data)
        out = np.empty(max_output_size(n), dtype=np.uint8)
        r = lib().lz4o_compress_block_with_table(p, n, out.ctypes.data, out.size,
                                                 self.table.ctypes.data, self.offset)
        assert r >= 0
        self.offset += n
        return out[:r].tobytes()


def compress_block_cont(data) -> bytes:
    """A block as FrameEncoder compresses it when it is not the first (CONT mode): hash5 + u32 table,
    stream offset > 0, so nothing from earlier blocks is reachable (frame/compress.rs:357-367)."""
    t = FrameTable()
    t.offset = 1 << 20
    return t.compress(data, 65536)


def compress_block_fresh_h5(data) -> bytes:
    """First block of a frame: hash5 + u32 table regardless of the input length."""
    return FrameTable().compress(data, 65536)


def decompress_block(data, cap: int):
    """block::decompress_into.  Returns (status, bytes, expected, actual)."""
    a, p, n = _buf(data)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    w, e1, e2 = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    st = lib().lz4o_decompress_block(p, n, out.ctypes.data, cap, C.byref(w), C.byref(e1), C.byref(e2))
    return st, out[: w.value].tobytes(), e1.value, e2.value


def decompress_size_prepended(data, max_alloc: int = 1 << 27):
    a, p, n = _buf(data)
    if n < 4:
        return ERR_EXPECTED_ANOTHER_BYTE, b"", 0, 0
    want = int.from_bytes(bytes(a[:4]), "little")
    assert want <= max_alloc, "refusing a huge allocation (callers guard, cf. tests.rs:497-501)"
    out = np.zeros(max(want, 1), dtype=np.uint8)
    w, e1, e2 = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    st = lib().lz4o_decompress_size_prepended(p, n, out.ctypes.data, want, C.byref(w), C.byref(e1), C.byref(e2))
    return st, out[: w.value].tobytes(), e1.value, e2.value


def xxh32(data, seed: int = 0) -> int:
    a, p, n = _buf(data)
    return lib().lz4o_xxh32(p, n, seed)


def frame_compress(data, block_size_id: int = 0, flags: int = 0, flush_every: int = 0) -> bytes:
    a, p, n = _buf(data)
    bsid = block_size_id or 7
    cap = lib().lz4o_frame_bound(n, 4 if block_size_id == 0 else bsid) + 64
    if flush_every:
        cap += 8 * (n // flush_every + 2)
    out = np.empty(cap, dtype=np.uint8)
    r = lib().lz4o_frame_compress(p, n, block_size_id, flags, flush_every, out.ctypes.data, cap)
    assert r >= 0, r
    return out[:r].tobytes()


def frame_decompress(data, cap: int):
    """Returns (status, bytes, block_status)."""
    a, p, n = _buf(data)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    w, be = C.c_size_t(0), C.c_int(0)
    st = lib().lz4o_frame_decompress(p, n, out.ctypes.data, cap, C.byref(w), C.byref(be))
    return st, out[: w.value].tobytes(), be.value


def _batch(fn, src, in_off, in_len, dst, out_off, out_cap, nthreads):
    nb = len(in_len)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint32)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out_cap = np.ascontiguousarray(out_cap, dtype=np.uint32)
    out_len = np.zeros(nb, dtype=np.uint32)
    status = np.zeros(nb, dtype=np.int32)
    fn(src.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, dst.ctypes.data, out_off.ctypes.data,
       out_cap.ctypes.data, out_len.ctypes.data, status.ctypes.data, nb, nthreads)
    return out_len, status


def compress_batch(src, in_off, in_len, dst, out_off, out_cap, nthreads: int = 1):
    """Block-API compress of many blocks (table choice by length) on `nthreads` host threads."""
    return _batch(lib().lz4o_compress_batch, src, in_off, in_len, dst, out_off, out_cap, nthreads)


def decompress_batch(src, in_off, in_len, dst, out_off, out_cap, nthreads: int = 1):
    return _batch(lib().lz4o_decompress_batch, src, in_off, in_len, dst, out_off, out_cap, nthreads)


# ---- performance-oriented CPU baseline (oracle/lz4_cpu_baseline.c): same bytes as the functions above ----------

def fast_compress_block(data) -> bytes:
    a, p, n = _buf(data)
    out = np.empty(max_output_size(n), dtype=np.uint8)
    tab = np.empty(4096, dtype=np.uint32)
    r = lib().lz4cpu_compress_block(p, n, out.ctypes.data, out.size, tab.ctypes.data)
    assert r >= 0
    return out[:r].tobytes()


def fast_decompress_block(data, cap: int):
    """(status, bytes, expected, actual) like decompress_block."""
    a, p, n = _buf(data)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    w, e1, e2 = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    st = lib().lz4cpu_decompress_block(p, n, out.ctypes.data, cap, C.byref(w), C.byref(e1), C.byref(e2))
    return st, out[: w.value].tobytes(), e1.value, e2.value


class Pool:
    """Persistent worker threads for the CPU arm of bench.py: threads are created once, every run() hands out the
    blocks of one batch in chunks from an atomic counter."""

    def __init__(self, nthreads: int):
        self.nthreads = max(1, int(nthreads))
        self._h = lib().lz4cpu_pool_create(self.nthreads)

    def _run(self, decode, src, in_off, in_len, dst, out_off, out_cap):
        nb = len(in_len)
        out_len = np.zeros(nb, dtype=np.uint32)
        status = np.zeros(nb, dtype=np.int32)
        lib().lz4cpu_pool_run(self._h, decode, src.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, dst.ctypes.data,
                              out_off.ctypes.data, out_cap.ctypes.data, out_len.ctypes.data, status.ctypes.data, nb)
        return out_len, status

    def copy(self, dst, src=None):
        """dst[:] = src (or zeros) with the pool's threads doing the first touch of dst's pages."""
        lib().lz4cpu_pool_copy(self._h, dst.ctypes.data, src.ctypes.data if src is not None else None, dst.size)

    def compress(self, src, in_off, in_len, dst, out_off, out_cap):
        return self._run(0, src, in_off, in_len, dst, out_off, out_cap)

    def decompress(self, src, in_off, in_len, dst, out_off, out_cap):
        return self._run(1, src, in_off, in_len, dst, out_off, out_cap)

    def close(self):
        if self._h:
            lib().lz4cpu_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

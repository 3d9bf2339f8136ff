"""The C-ABI library loads on a CPU-only box, exports every symbol include/lz4b200.h declares, and fails
loudly (never falls back) when no CUDA device exists.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from lz4_flex_b200 import _native, block, frame
from lz4_flex_b200.errors import CudaError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lz4b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lz4b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    L = C.CDLL(_native.build())
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in lz4b200.h but not exported"
    assert set(_native.SIGNATURES) == set(syms), "ctypes table and header disagree"


def test_abi_basics():
    L = _native.lib()
    assert L.lz4b200_abi_version() == 1
    for n in (0, 1, 12, 13, 65535, 65536, 66675, 4 << 20):
        assert L.lz4b200_max_output_size(n) == 16 + 4 + n * 110 // 100 == block.get_maximum_output_size(n)
    assert L.lz4b200_xxh32(None, 0, 0) == 0x02CC5D05
    assert frame.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F
    h = frame._Xxh32(0)
    for part in (b"Nobody ins", b"", b"pects the spammish repe", b"tition"):
        h.update(part)
    assert h.digest() == 0xE2293B2F
    assert _native.status_string(4) == "expected another byte, found none"


def test_frame_header_writer():
    # fuzz_decomp_corrupt_frame.rs:26-27
    assert frame.FrameInfo(block_size=frame.BlockSize.Max64KB).header_bytes() == bytes([4, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x82])
    assert frame.FrameInfo(block_size=frame.BlockSize.Max4MB).header_bytes() == bytes([4, 0x22, 0x4D, 0x18, 0x60, 0x70, 0x73])
    h = frame.FrameInfo(block_size=frame.BlockSize.Max64KB, content_size=725).header_bytes()
    assert len(h) == 15 and int.from_bytes(h[6:14], "little") == 725          # tests/tests.rs:726-734
    assert frame.BlockSize.from_buf_length(0) == frame.BlockSize.Max64KB      # header.rs:57-67
    assert frame.BlockSize.from_buf_length(65536) == frame.BlockSize.Max64KB
    assert frame.BlockSize.from_buf_length(65537) == frame.BlockSize.Max256KB
    assert frame.BlockSize.from_buf_length(262145) == frame.BlockSize.Max4MB


def test_uncompressed_size():
    from lz4_flex_b200.errors import ExpectedAnotherByte
    assert block.uncompressed_size(b"\x05\x00\x00\x00abc") == (5, b"abc")
    with pytest.raises(ExpectedAnotherByte):
        block.uncompressed_size(b"\x05\x00")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_fails_loudly_without_gpu():
    with pytest.raises(CudaError):
        block.Context(0)
    with pytest.raises(CudaError):
        block.compress(b"hello hello hello hello")
    with pytest.raises(CudaError):
        frame.compress_frame(b"hello hello hello hello")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lz4_flex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "lz4_oracle" not in text and "lz4o_" not in text, f

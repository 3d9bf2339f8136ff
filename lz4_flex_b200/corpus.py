"""Benchmark/test corpus: the reference's bench fixtures (benches/*.txt, *.json; data, not code),
stored xz-compressed under data/ with their sha256 pinned, plus the synthetic generators named in
BASELINE.md §3."""
from __future__ import annotations

import hashlib
import lzma
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")

SHA256 = {
    "compression_1k.txt": "6b1904dbb0838f0aed3519fef5e4bbd0b3c44dafdc4d899eecdd102ecd4a0957",
    "compression_34k.txt": "2581db546d71049c5e1268b2f340e968770be9135e89b621a6d864e36f55cdff",
    "compression_65k.txt": "3a3e5c932e93dd0d9d72914477662c97866ba62c9e5c02e58e2964f8da2e6394",
    "compression_66k_JSON.txt": "88c4662e28fe439365f370ea86bb6da4ff7f8d192ff86c727c05120b881a2ec3",
    "dickens.txt": "ed4ba0a1380a45ea8ecf3ac472eec9968a92946e0a413b9a82b13fca794d9496",
    "hdfs.json": "443ec44e5f3de3ed05a259b5ecb751cc57dc7905c76fea5347d69d7fd6879035",
    "dickens.lz4": "8af73fbf03558df4f6d814323802ecd9872da4dee61f9b02ae19d0258fe4e49b",
}

_cache: dict[str, bytes] = {}


def load(name: str) -> bytes:
    """Fixture bytes by the reference's file name (sha256-verified)."""
    if name not in _cache:
        raw = os.path.join(_DATA, name)
        if os.path.exists(raw):
            b = open(raw, "rb").read()
        else:
            b = lzma.decompress(open(raw + ".xz", "rb").read())
        if hashlib.sha256(b).hexdigest() != SHA256[name]:
            raise RuntimeError(f"fixture {name} is corrupt")
        _cache[name] = b
    return _cache[name]


def tiled(name: str, total: int) -> np.ndarray:
    """buf[i] = file[i mod len(file)] for i < total."""
    src = np.frombuffer(load(name), dtype=np.uint8)
    reps = -(-total // src.size)
    return np.tile(src, reps)[:total].copy()


def xorshift64star_bytes(nbytes: int, seed: int = 0x9E3779B97F4A7C15) -> np.ndarray:
    """Incompressible filler of config 5: xorshift64* stream, 8 bytes per step, little endian."""
    n = -(-nbytes // 8)
    out = np.empty(n, dtype=np.uint64)
    x = seed & 0xFFFFFFFFFFFFFFFF
    m = 0x2545F4914F6CDD1D
    mask = 0xFFFFFFFFFFFFFFFF
    for i in range(n):
        x ^= x >> 12
        x ^= (x << 25) & mask
        x ^= x >> 27
        out[i] = (x * m) & mask
    return out.view(np.uint8)[:nbytes].copy()


def adversarial_blocks(nblocks: int, zero_fraction: float, block: int = 65536) -> np.ndarray:
    """Config 5: `nblocks` blocks, a `zero_fraction` share all-zero (evenly interleaved), the rest
    incompressible (distinct windows of one xorshift64* stream); deterministic."""
    extra = 8 * (nblocks // 16 + 1)
    rnd = xorshift64star_bytes(block * 16 + extra)
    out = np.zeros(nblocks * block, dtype=np.uint8)
    acc = 0.0
    k = 0
    for b in range(nblocks):
        acc += zero_fraction
        if acc >= 1.0 - 1e-9:
            acc -= 1.0
            continue  # zero block
        start = (k % 16) * block + (k // 16) * 8
        out[b * block:(b + 1) * block] = rnd[start:start + block]
        k += 1
    return out

"""Generates the BlockMode::Linked frame fixtures under tests/golden/linked/ with the system liblz4 (LZ4F API,
liblz4 1.9.4): the reference's own interop strategy (tests/tests.rs:109-147 uses lz4-sys / lzzzz the same way).
lz4_flex's encoder is not involved: these are valid-but-foreign streams for the DECODER (frame/decompress.rs:196-305).
Run from the repo root:  python tests/golden/make_linked_frames.py"""
import ctypes
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lz4_flex_b200 import corpus  # noqa: E402

L = ctypes.CDLL("liblz4.so.1")


class FrameInfo(ctypes.Structure):
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int), ("contentChecksumFlag", ctypes.c_int),
                ("frameType", ctypes.c_int), ("contentSize", ctypes.c_ulonglong), ("dictID", ctypes.c_uint),
                ("blockChecksumFlag", ctypes.c_int)]


class Prefs(ctypes.Structure):
    _fields_ = [("frameInfo", FrameInfo), ("compressionLevel", ctypes.c_int), ("autoFlush", ctypes.c_uint),
                ("favorDecSpeed", ctypes.c_uint), ("reserved", ctypes.c_uint * 3)]


L.LZ4F_compressFrameBound.restype = ctypes.c_size_t
L.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.POINTER(Prefs)]
L.LZ4F_compressFrame.restype = ctypes.c_size_t
L.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Prefs)]
L.LZ4F_isError.restype = ctypes.c_uint
L.LZ4F_isError.argtypes = [ctypes.c_size_t]


def lz4f(data: bytes, bsid: int, linked: bool, content_checksum=0, block_checksums=0, content_size=0, level=0) -> bytes:
    p = Prefs()
    p.frameInfo.blockSizeID = bsid
    p.frameInfo.blockMode = 0 if linked else 1                 # LZ4F_blockLinked = 0
    p.frameInfo.contentChecksumFlag = content_checksum
    p.frameInfo.blockChecksumFlag = block_checksums
    p.frameInfo.contentSize = len(data) if content_size else 0
    p.compressionLevel = level
    cap = L.LZ4F_compressFrameBound(len(data), ctypes.byref(p))
    buf = ctypes.create_string_buffer(cap)
    r = L.LZ4F_compressFrame(buf, cap, data, len(data), ctypes.byref(p))
    assert not L.LZ4F_isError(r)
    return buf.raw[:r]


def sources():
    """name -> bytes, rebuilt from the corpus fixtures by the tests (nothing but the recipe is stored)."""
    dick = corpus.load("dickens.txt")
    json66 = corpus.tiled("compression_66k_JSON.txt", 1 << 20).tobytes()
    noise = corpus.xorshift64star_bytes(70000).tobytes()
    mixed = dick[:100000] + noise + dick[100000:250000] + noise[:30000] + dick[90000:160000]
    return {"dickens_300k": dick[:300000], "json_1m": json66, "mixed_stored": mixed, "tiny": dick[:1000],
            "hc_dickens_200k": dick[200000:400000]}


RECIPES = [  # (file, source, block size id, kwargs)
    ("dickens_300k_64k_linked_all_checks.lz4", "dickens_300k", 4, dict(linked=True, content_checksum=1, block_checksums=1, content_size=1)),
    ("json_1m_256k_linked.lz4", "json_1m", 5, dict(linked=True)),
    ("mixed_stored_64k_linked.lz4", "mixed_stored", 4, dict(linked=True, content_checksum=1)),
    ("tiny_64k_linked.lz4", "tiny", 4, dict(linked=True)),
    ("hc_dickens_200k_64k_linked.lz4", "hc_dickens_200k", 4, dict(linked=True, level=9)),   # LZ4HC parse: long offsets, many cross-block matches
]

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "tests", "golden", "linked")
    src = sources()
    manifest = []
    for name, s, bsid, kw in RECIPES:
        f = lz4f(src[s], bsid, **kw)
        open(os.path.join(out_dir, name), "wb").write(f)
        manifest.append({"file": name, "source": s, "source_sha256": hashlib.sha256(src[s]).hexdigest(),
                         "frame_sha256": hashlib.sha256(f).hexdigest(), "frame_len": len(f), "source_len": len(src[s])})
        print(name, len(src[s]), "->", len(f))
    json.dump(manifest, open(os.path.join(out_dir, "manifest.json"), "w"), indent=1)

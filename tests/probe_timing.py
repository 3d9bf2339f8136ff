"""Timing-only probe for kernel variants whose output is intentionally invalid (no verification)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lz4_flex_b200 import block, corpus
nb, B, slot = 16384, 65536, 72112
data = corpus.tiled("compression_66k_JSON.txt", nb * B)
dev = torch.device("cuda", 0)
d_in = torch.from_numpy(data).to(dev)
d_comp = torch.zeros(nb * slot, dtype=torch.uint8, device=dev)
offs = np.arange(nb, dtype=np.uint64) * B; lens = np.full(nb, B, dtype=np.uint32)
enc = block.DeviceBatch(offs, lens, np.arange(nb, dtype=np.uint64) * slot, np.full(nb, slot, dtype=np.uint32), None, dev)
ctx = block.Context(0)
for _ in range(3): enc.compress(d_in, d_comp, ctx)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): enc.compress(d_in, d_comp, ctx)
e1.record(); torch.cuda.synchronize()
print("compress_ms", e0.elapsed_time(e1) / 3)

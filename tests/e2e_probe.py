import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from lz4_flex_b200 import block, _native, corpus
L=_native.lib()
ctx=block.Context(0)
nb=16384; B=65536
data=corpus.tiled("compression_66k_JSON.txt", nb*B)
h_in=torch.empty(nb*B,dtype=torch.uint8).pin_memory(); h_in.numpy()[:]=data
h_comp=torch.empty(400<<20,dtype=torch.uint8).pin_memory()
h_back=torch.empty(nb*B,dtype=torch.uint8).pin_memory()
print('kinds', L.lz4b200_host_pointer_kind(h_in.data_ptr()), L.lz4b200_host_pointer_kind(h_in.numpy().ctypes.data), L.lz4b200_host_pointer_kind(data.ctypes.data))
offs=np.arange(nb,dtype=np.uint64)*B; lens=np.full(nb,B,dtype=np.uint32)
for it in range(4):
    t0=time.perf_counter(); out,ooff,olen=block.compress_batch(h_in.numpy(),offs,lens,None,out=h_comp.numpy(),ctx=ctx); t1=time.perf_counter()
    block.decompress_batch(out,ooff,olen,h_back.numpy(),offs,lens,ctx=ctx); t2=time.perf_counter()
    print('compress call ms',1e3*(t1-t0),'decompress call ms',1e3*(t2-t1))
# raw copies through torch for reference
d=torch.empty(nb*B,dtype=torch.uint8,device='cuda')
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(h_in,non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    h_back.copy_(d,non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()
    print('torch H2D GB/s', nb*B/(t1-t0)/1e9, 'D2H GB/s', nb*B/(t2-t1)/1e9)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h_in,non_blocking=True)
d2=torch.empty_like(d)
with torch.cuda.stream(s2): h_back.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); t1=time.perf_counter()
print('duplex both 1GiB ms', 1e3*(t1-t0))

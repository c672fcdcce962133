"""The north_star's named kernel alone: tdiff_attn_aggregate_h (fused scatter_softmax -> scatter_sum, keys + values from HBM) on random
data of the cfg3 problem size (N = 204 800 nodes, k = 32), a few launches -- the target of
    ncu --set full --clock-control none -k regex:aggregate_h_kernel -s 2 -c 1 -o gpurun_out/<tag>_agg_h python tools/agg_h_standalone.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from targetdiff_b200 import _lib

lib = _lib.load()
dev = torch.device('cuda:0')
N, kk = 640 * 320, 32
E = N * kk
g = torch.Generator(device=dev).manual_seed(0)
gk = torch.randn(E, 128, device=dev, generator=g)
gv = torch.randn(E, 128, device=dev, generator=g)
gw = torch.rand(E, device=dev, generator=g)
gs = torch.randint(0, N, (N, kk), device=dev, dtype=torch.int32, generator=g)
gq = torch.randn(N, 128, device=dev, generator=g)
gh = torch.randn(N, 128, device=dev, generator=g)
go = torch.empty_like(gh)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5):
    if i == 2:
        e0.record()
    _lib.check(lib.tdiff_attn_aggregate_h(P(gk), P(gv), P(gw), P(gs), P(gq), P(gh), P(go), N, kk, st))
e1.record()
torch.cuda.synchronize()
b = E * 1028 + N * 1536
ms = e0.elapsed_time(e1) / 3
print('aggregate_h: %.3f ms per launch, %.1f GB/s of %d algorithmic bytes' % (ms, b / ms / 1e6, b))

"""Stand-alone graph / scatter operators of libtdiff.so with the call shapes of the reference's native seam
(torch_geometric.nn.knn_graph, torch_scatter.scatter_{softmax,sum,mean}).  CUDA tensors only."""
import ctypes

import torch

from . import _lib


def _need_cuda(*ts):
    for t in ts:
        if t is not None and t.device.type != 'cuda':
            raise RuntimeError('targetdiff_b200.ops run on CUDA tensors only (no CPU fallback)')


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _counts(batch, n):
    if batch is None:
        return [n]
    if batch.numel() > 1 and bool((batch[1:] < batch[:-1]).any()):
        raise ValueError('batch must be sorted ascending')
    return torch.bincount(batch).cpu().tolist()


def knn_slots(x, k, batch=None):
    """Fixed-degree neighbour slots [N,k] int32 (-1 padded) and the number of edges."""
    _need_cuda(x, batch)
    x = x.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    counts = _counts(batch, n)
    slots = torch.empty(n, k, dtype=torch.int32, device=x.device)
    ne = ctypes.c_int64(0)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().tdiff_knn_graph(_p(x), n, _lib.i32_array(counts), len(counts), k, _p(slots), None, ctypes.byref(ne), _st(x.device)))
    return slots, ne.value


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target'):
    """edge_index int64 [2,E] like torch_geometric.nn.knn_graph (reference models/uni_transformer.py:280)."""
    if loop:
        raise NotImplementedError('loop=True')
    _need_cuda(x, batch)
    x = x.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    counts = _counts(batch, n)
    slots = torch.empty(n, k, dtype=torch.int32, device=x.device)
    ei = torch.empty(2, n * k, dtype=torch.int64, device=x.device)
    ne = ctypes.c_int64(0)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().tdiff_knn_graph(_p(x), n, _lib.i32_array(counts), len(counts), k, _p(slots), _p(ei), ctypes.byref(ne), _st(x.device)))
    E = ne.value
    # the library wrote row 0 at [0,E) and row 1 at [E,2E) of the flat buffer
    ei = ei.view(-1)[:2 * E].view(2, E).clone()
    return ei if flow == 'source_to_target' else ei.flip(0)


def attn_aggregate_h(k, v, e_w, src_slots, q, h):
    """h + scatter_sum(scatter_softmax((q[dst]*k/sqrt(8)).sum(-1), dst)[..., None] * v * e_w, dst) on a slot list."""
    _need_cuda(k, v, e_w, src_slots, q, h)
    n, kk = src_slots.shape
    out = torch.empty_like(h)
    args = [t.contiguous() for t in (k.float(), v.float(), e_w.float(), src_slots.int(), q.float(), h.float())]
    with torch.cuda.device(h.device):
        _lib.check(_lib.load().tdiff_attn_aggregate_h(*[_p(t) for t in args], _p(out), n, kk, _st(h.device)))
        torch.cuda.current_stream(h.device).synchronize()
    return out


def attn_aggregate_x(k, v16, e_w, src_slots, q, x, mask):
    _need_cuda(k, v16, e_w, src_slots, q, x, mask)
    n, kk = src_slots.shape
    out = torch.empty_like(x)
    args = [t.contiguous() for t in (k.float(), v16.float(), e_w.float(), src_slots.int(), q.float(), x.float(), mask.to(torch.uint8))]
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().tdiff_attn_aggregate_x(*[_p(t) for t in args], _p(out), n, kk, _st(x.device)))
    return out


def scatter_mean3(src, batch):
    """scatter_mean(src [M,3], batch, dim=0) for a sorted batch vector (reference models/molopt_score_model.py:115)."""
    _need_cuda(src, batch)
    counts = _counts(batch, src.shape[0])
    out = torch.empty(len(counts), 3, device=src.device)
    s = src.detach().float().contiguous()
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().tdiff_scatter_mean3(_p(s), _lib.i32_array(counts), len(counts), _p(out), _st(src.device)))
    return out

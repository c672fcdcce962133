"""Pure-torch stand-in for `torch_geometric.nn.knn_graph` (PyG 2.2.0 -> torch_cluster 1.6.0 `knn`).

Test infrastructure only.  Reference call site: models/uni_transformer.py:280
(`knn_graph(x, k=self.k, batch=batch, flow='source_to_target')`).

The library's selection arithmetic is not under /root/reference ("parity unpinned").  Canonical
semantics fixed by this project:
  * per graph, for every query node i take the (k+1) nearest nodes by squared Euclidean distance
    computed in fp32 as ((dx*dx) + (dy*dy)) + (dz*dz), each op rounded (no FMA);
  * ascending distance, ties -> smaller node index first;
  * drop entries with src == i (PyG's `row != col` mask);
  * emit edge_index[0] = src (neighbour), edge_index[1] = dst (query), grouped by dst ascending.
"""
import torch


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target', cosine=False, num_workers=1):
    assert flow in ('source_to_target', 'target_to_source')
    assert not cosine
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    kk = k if loop else k + 1
    srcs, dsts = [], []
    num_graphs = int(batch.max()) + 1 if n > 0 else 0
    counts = torch.bincount(batch, minlength=num_graphs).tolist()
    start = 0
    for g in range(num_graphs):
        ng = counts[g]
        xg = x[start:start + ng]
        dx = xg[:, None, 0] - xg[None, :, 0]
        dy = xg[:, None, 1] - xg[None, :, 1]
        dz = xg[:, None, 2] - xg[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz          # [query, candidate]
        order = torch.sort(d2, dim=1, stable=True).indices[:, :min(kk, ng)]
        q = torch.arange(ng, device=x.device)[:, None].expand_as(order)
        if not loop:
            keep = order != q
            src, dst = order[keep], q[keep]
        else:
            src, dst = order.reshape(-1), q.reshape(-1)
        srcs.append(src + start)
        dsts.append(dst + start)
        start += ng
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.long)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.long)
    if flow == 'source_to_target':
        return torch.stack([src, dst], 0)
    return torch.stack([dst, src], 0)


def radius_graph(*args, **kwargs):
    raise NotImplementedError('radius_graph: dead path in the reference (models/uni_transformer.py:278 reads undefined self.r)')

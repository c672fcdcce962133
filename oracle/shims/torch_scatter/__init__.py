"""Pure-torch stand-in for the parts of `torch_scatter` 2.1.0 the reference's hot path calls.

Test infrastructure only.  Call sites in the reference: models/uni_transformer.py:73,78,135,139
(scatter_softmax / scatter_sum), models/molopt_score_model.py:115 and scripts/sample_diffusion.py:61
(scatter_mean).  The real wheel is not installable here (no network); these mirror its documented
Python-level composition: `scatter_sum` = broadcast index + `scatter_add_` (CPU: sequential in edge
order), `scatter_mean` = sum / clamp(count, 1), `scatter_softmax` = max-shift, exp, sum, divide.
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(0, dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = dim_size
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode='floor')
    return out


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    index_b = _broadcast(index, src, dim)
    size = _out_size(src, index_b, dim, dim_size)
    res = torch.zeros(size, dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, index_b, src, reduce='amax', include_self=False)
    return res, None


def scatter_softmax(src, index, dim=-1, dim_size=None):
    if not torch.is_floating_point(src):
        raise ValueError('`scatter_softmax` can only be computed over tensors with floating point data types.')
    index = _broadcast(index, src, dim)
    max_value_per_index = scatter_max(src, index, dim=dim, dim_size=dim_size)[0]
    max_per_src_element = max_value_per_index.gather(dim, index)
    recentered_scores = src - max_per_src_element
    recentered_scores_exp = recentered_scores.exp_()
    sum_per_index = scatter_sum(recentered_scores_exp, index, dim, dim_size=dim_size)
    normalizing_constants = sum_per_index.gather(dim, index)
    return recentered_scores_exp.div(normalizing_constants)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    """torch_scatter.scatter dispatcher -- only imported by reference datasets/protein_ligand.py:10 (SDF parsing, not on the sampling path)."""
    if reduce in ('sum', 'add'):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == 'mean':
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == 'max':
        return scatter_max(src, index, dim, out, dim_size)
    raise NotImplementedError(reduce)


def segment_coo(*args, **kwargs):
    raise NotImplementedError('torch_scatter.segment_coo: only referenced by reference datasets/pl_data.py:61 (bond matrices), not on the sampling path')

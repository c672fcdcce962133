"""Stand-in for `torch_geometric.loader.DataLoader` -- only imported (reference datasets/pl_data.py:5), never used on the sampling path."""
import torch.utils.data


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, follow_batch=None, exclude_keys=None, **kwargs):
        raise NotImplementedError('torch_geometric.loader.DataLoader shim: training-side loader, not part of the sampling path')

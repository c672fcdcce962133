"""Minimal stand-in for `torch_geometric.data.{Data, Batch}` (PyG 2.2.0) -- TEST INFRASTRUCTURE ONLY.

Just enough for the reference's sampling driver to run unmodified on CPU (reference scripts/sample_diffusion.py:42:
`Batch.from_data_list([data.clone() ...], follow_batch=FOLLOW_BATCH).to(device)`; reference datasets/pl_data.py:10-36):
attribute/item storage, `clone`, `to`, `__inc__`, and collation = concatenate tensors along dim 0 (dim -1 for keys containing
'index', shifted by `__inc__`), python objects gathered into lists, `<key>_batch` vectors for the followed keys.
"""
import copy

import torch


class Data:
    def __init__(self, **kwargs):
        object.__setattr__(self, '_store', {})
        for k, v in kwargs.items():
            self._store[k] = v

    # item / attribute access on the same store
    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __getattr__(self, key):
        store = object.__getattribute__(self, '_store')
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    @property
    def keys(self):
        return list(self._store.keys())

    def __inc__(self, key, value, *args, **kwargs):
        return 0

    def __cat_dim__(self, key, value, *args, **kwargs):
        return -1 if 'index' in key else 0

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        object.__setattr__(out, '_store', {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in self._store.items()})
        return out

    def to(self, device, *args, **kwargs):
        for k, v in self._store.items():
            if torch.is_tensor(v):
                self._store[k] = v.to(device)
        return self


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        follow_batch = tuple(follow_batch or ())
        out = cls()
        keys = data_list[0].keys
        for k in keys:
            vals = [d[k] for d in data_list]
            if torch.is_tensor(vals[0]):
                dim = data_list[0].__cat_dim__(k, vals[0])
                shifted, inc = [], 0
                for d, v in zip(data_list, vals):
                    shifted.append(v + inc if (inc != 0) else v)
                    step = d.__inc__(k, v)
                    inc = inc + (int(step) if not torch.is_tensor(step) else int(step.item()))
                out[k] = torch.cat(shifted, dim=dim) if vals[0].dim() > 0 else torch.stack(vals)
                if k in follow_batch:
                    sizes = torch.tensor([v.size(dim) for v in vals])
                    out[k + '_batch'] = torch.repeat_interleave(torch.arange(len(vals)), sizes)
            else:
                out[k] = vals
        out['num_graphs'] = len(data_list)
        return out

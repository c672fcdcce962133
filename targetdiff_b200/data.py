"""Minimal stand-ins for the PyG containers the sampling driver touches (reference datasets/pl_data.py:10-36 and
`Batch.from_data_list(..., follow_batch=FOLLOW_BATCH)` at scripts/sample_diffusion.py:42).  Plain attribute bags of tensors."""
import copy

import torch

FOLLOW_BATCH = ('protein_element', 'ligand_element', 'ligand_bond_type',)


class ProteinLigandData(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @staticmethod
    def from_protein_ligand_dicts(protein_dict=None, ligand_dict=None, **kwargs):
        inst = ProteinLigandData(**kwargs)
        for key, item in (protein_dict or {}).items():
            setattr(inst, 'protein_' + key, item)
        for key, item in (ligand_dict or {}).items():
            setattr(inst, 'ligand_' + key, item)
        return inst

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def clone(self):
        return ProteinLigandData(**{k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in self.__dict__.items()})

    def to(self, device):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Batch(ProteinLigandData):
    """Concatenates per-atom tensors of a list of ProteinLigandData and adds `protein_element_batch` /
    `ligand_element_batch` (what PyG's follow_batch produces)."""

    @staticmethod
    def from_data_list(data_list, follow_batch=FOLLOW_BATCH):
        out = Batch()
        keys = data_list[0].keys()
        for k in keys:
            vals = [d[k] for d in data_list]
            if torch.is_tensor(vals[0]) and vals[0].dim() >= 1 and (k.startswith('protein_') or k.startswith('ligand_')) \
                    and 'bond' not in k:
                setattr(out, k, torch.cat(vals, 0))
            else:
                setattr(out, k, vals)
        for prefix in ('protein', 'ligand'):
            key = prefix + '_pos' if hasattr(data_list[0], prefix + '_pos') else prefix + '_element'
            if hasattr(data_list[0], key):
                sizes = torch.tensor([d[key].shape[0] for d in data_list])
                setattr(out, prefix + '_element_batch', torch.repeat_interleave(torch.arange(len(data_list)), sizes))
        out.num_graphs = len(data_list)
        return out

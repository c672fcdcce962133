"""Stability screen of generated molecules on the GPU (reference utils/evaluation/analyze.py:106-143 `check_stability`;
caller scripts/evaluate_diffusion.py:78-84).  `check_stability` keeps the reference's signature and return tuple for one molecule;
`check_stability_batch` screens a whole result set in one launch (one warp per molecule)."""
import ctypes

import numpy as np
import torch

from . import _lib


def check_stability_batch(positions, atom_types, hs=False, device='cuda:0'):
    """positions: list of [n_i,3] arrays (or one [sum n_i,3] tensor with `atom_types` a list of arrays); atom_types: atomic numbers.
    Returns (molecule_stable [M] bool, nr_stable_atoms [M] int, n_atoms [M] int, nr_bonds [sum n_i] int) as numpy arrays."""
    counts = [len(a) for a in atom_types]
    pos = torch.as_tensor(np.concatenate([np.asarray(p, dtype=np.float64) for p in positions]) if isinstance(positions, (list, tuple))
                          else np.asarray(positions)).to(torch.float32)
    z = torch.as_tensor(np.concatenate([np.asarray(a).astype(np.int64) for a in atom_types])).to(torch.int32)
    if pos.shape[0] != sum(counts) or pos.dim() != 2 or pos.shape[1] != 3:
        raise ValueError('positions must be [n,3] per molecule')
    dev = torch.device(device)
    pos, z = pos.to(dev).contiguous(), z.to(dev).contiguous()
    M = len(counts)
    nr_bonds = torch.zeros(max(1, pos.shape[0]), dtype=torch.int32, device=dev)
    stable_atoms = torch.zeros(max(1, M), dtype=torch.int32, device=dev)
    mol_stable = torch.zeros(max(1, M), dtype=torch.uint8, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tdiff_check_stability(P(pos), P(z), _lib.i32_array(counts), M, int(bool(hs)), P(nr_bonds), P(stable_atoms), P(mol_stable),
                                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return (mol_stable[:M].cpu().numpy().astype(bool), stable_atoms[:M].cpu().numpy().astype(np.int64), np.asarray(counts, dtype=np.int64),
            nr_bonds[:pos.shape[0]].cpu().numpy().astype(np.int64))


def check_stability(positions, atom_type, debug=False, hs=False, return_nr_bonds=False, device='cuda:0'):
    """One molecule, the reference's signature: (molecule_stable, nr_stable_bonds, n_atoms[, nr_bonds])."""
    positions = np.asarray(positions)
    assert len(positions.shape) == 2 and positions.shape[1] == 3
    ms, ns, n, nb = check_stability_batch([positions], [np.asarray(atom_type)], hs=hs, device=device)
    if return_nr_bonds:
        return bool(ms[0]), int(ns[0]), int(n[0]), nb
    return bool(ms[0]), int(ns[0]), int(n[0])

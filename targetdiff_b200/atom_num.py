"""Ligand-size prior (reference utils/evaluation/atom_num.py:9-26): pocket "space size" = median of the 10 largest
pairwise distances; the number of atoms is drawn from an empirical per-bin table with numpy's global RNG (so that
`seed_all` reproduces the reference's draws).  CPU, once per batch -- not on the hot path."""
import json
import os

import numpy as np

_TABLE = None


def _table():
    global _TABLE
    if _TABLE is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'atom_num_prior.json')) as f:
            _TABLE = json.load(f)
    return _TABLE


def get_space_size(pocket_3d_pos):
    p = np.asarray(pocket_3d_pos, dtype=np.float64)
    n = p.shape[0]
    # condensed pairwise distances (same values scipy's pdist returns), top-10 by partial sort
    iu = np.triu_indices(n, k=1)
    d = np.sqrt(((p[iu[0]] - p[iu[1]]) ** 2).sum(-1))
    top = np.sort(d)[::-1][:10]
    return np.median(top)


def _get_bin_idx(space_size):
    bounds = _table()['bounds']
    for i, b in enumerate(bounds):
        if b > space_size:
            return i
    return len(bounds)


def sample_atom_num(space_size):
    b = _table()['bins'][_get_bin_idx(space_size)]
    return np.random.choice(b['num_atoms'], p=b['prob'])

"""sample_diffusion_ligand -- the sampling driver (reference scripts/sample_diffusion.py:31-116), same signature and
7-tuple result, running on the libtdiff engine.  Differences in mechanics, not in results: the batch of `n_data` clones
is assembled directly (no PyG Batch), trajectories come back from the device once per batch instead of 4 D2H copies per
step, and un-batching is done on stacked arrays.

Randomness.  `rng='device'` (default): initial state from torch's generator of `device`, per-step noise from the engine's
counter-based Philox stream keyed by a seed drawn from torch's CPU generator (so `seed_all` reproduces a run).
`rng='cpu'`: every draw comes from torch's global CPU generator in exactly the reference's order -- per batch
`randn_like(center)` (:63), `rand_like(uniform_logits)` (:69 via models/molopt_score_model.py:161), then per denoising step
`randn_like(ligand_pos)` and `rand_like(log_model_prob)` (models/molopt_score_model.py:678,685) -- pre-drawn as a noise tape.
A run seeded with `seed_all(s)` then consumes the same random numbers as the unmodified reference run on CPU with the same
seed, which is what "identical RNG seeds" parity needs (tests/golden/pocket_1h36_*.npz were produced that way)."""
import time

import numpy as np
import torch

from . import atom_num
from .score_model import log_sample_categorical


def seed_all(seed):
    """reference utils/misc.py:58-61"""
    import random
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def _split(arr, cum, n_data):
    return [arr[..., cum[k]:cum[k + 1], :] if arr.ndim == 3 else arr[..., cum[k]:cum[k + 1]] for k in range(n_data)]


def sample_diffusion_ligand(model, data, num_samples, batch_size=16, device='cuda:0', num_steps=None, pos_only=False,
                            center_pos_mode='protein', sample_num_atoms='prior', rng='device'):
    if rng not in ('device', 'cpu'):
        raise ValueError("rng must be 'device' or 'cpu'")
    all_pred_pos, all_pred_v = [], []
    all_pred_pos_traj, all_pred_v_traj = [], []
    all_pred_v0_traj, all_pred_vt_traj = [], []
    time_list = []
    num_batch = int(np.ceil(num_samples / batch_size))
    current_i = 0
    device = torch.device(device)
    protein_pos_cpu = data.protein_pos.detach().cpu().float()
    n_prot = protein_pos_cpu.shape[0]
    protein_pos_dev = protein_pos_cpu.to(device)
    protein_feat_dev = data.protein_atom_feature.detach().to(device).float()
    for i in range(num_batch):
        n_data = batch_size if i < num_batch - 1 else num_samples - batch_size * (num_batch - 1)
        t1 = time.time()
        with torch.no_grad():
            batch_protein = torch.repeat_interleave(torch.arange(n_data, device=device), n_prot)
            if sample_num_atoms == 'prior':
                pocket_size = atom_num.get_space_size(protein_pos_cpu.numpy())
                ligand_num_atoms = [int(atom_num.sample_atom_num(pocket_size)) for _ in range(n_data)]
            elif sample_num_atoms == 'range':
                ligand_num_atoms = list(range(current_i + 1, current_i + n_data + 1))
            elif sample_num_atoms == 'ref':
                ligand_num_atoms = [int(data.ligand_element.size(0))] * n_data
            else:
                raise ValueError
            batch_ligand = torch.repeat_interleave(torch.arange(n_data), torch.tensor(ligand_num_atoms)).to(device)
            protein_pos = protein_pos_dev.repeat(n_data, 1)
            protein_v = protein_feat_dev.repeat(n_data, 1)

            # init ligand pos: pocket centre + N(0, 1)   (reference :61-63; scatter_mean = sequential sum / count, every clone
            # of the pocket has the same centre)
            n_lig = len(batch_ligand)
            center = (torch.zeros(1, 3).index_add_(0, torch.zeros(n_prot, dtype=torch.long), protein_pos_cpu) / max(n_prot, 1)).to(device)
            draw_dev = device if rng == 'device' else 'cpu'
            init_ligand_pos = center.expand(n_lig, 3) + torch.randn(n_lig, 3, device=draw_dev).to(device)
            # init ligand v (reference :66-70)
            if pos_only:
                init_ligand_v = data.ligand_atom_feature_full.to(device).repeat(n_data)
            else:
                uniform_logits = torch.zeros(n_lig, model.num_classes, device=draw_dev)
                init_ligand_v = log_sample_categorical(uniform_logits).to(device)
            tape = None
            if rng == 'cpu':
                S = model.num_timesteps if num_steps is None else int(num_steps)
                pn = torch.empty(S, n_lig, 3)
                vu = torch.zeros(S, n_lig, model.num_classes)
                for st in range(S):                                     # the reference's interleaved draw order
                    pn[st] = torch.randn(n_lig, 3)
                    if not pos_only:
                        vu[st] = torch.rand(n_lig, model.num_classes)
                tape = (pn, vu)

            r = model.sample_diffusion(protein_pos=protein_pos, protein_v=protein_v, batch_protein=batch_protein,
                                       init_ligand_pos=init_ligand_pos, init_ligand_v=init_ligand_v, batch_ligand=batch_ligand,
                                       num_steps=num_steps, pos_only=pos_only, center_pos_mode=center_pos_mode, stack_traj=True,
                                       noise_tape=tape)
            cum = np.cumsum([0] + ligand_num_atoms)
            pos = r['pos'].cpu().numpy().astype(np.float64)
            all_pred_pos += [pos[cum[k]:cum[k + 1]] for k in range(n_data)]
            pos_traj = r['pos_traj'].numpy().astype(np.float64)                   # [S, Nl, 3]
            all_pred_pos_traj += _split(pos_traj, cum, n_data)                    # n_data * [S, n_k, 3]
            v = r['v'].cpu().numpy()
            all_pred_v += [v[cum[k]:cum[k + 1]] for k in range(n_data)]
            all_pred_v_traj += _split(r['v_traj'].numpy(), cum, n_data)
            if not pos_only:
                all_pred_v0_traj += _split(r['v0_traj'].numpy(), cum, n_data)
                all_pred_vt_traj += _split(r['vt_traj'].numpy(), cum, n_data)
        t2 = time.time()
        time_list.append(t2 - t1)
        current_i += n_data
    return all_pred_pos, all_pred_v, all_pred_pos_traj, all_pred_v_traj, all_pred_v0_traj, all_pred_vt_traj, time_list

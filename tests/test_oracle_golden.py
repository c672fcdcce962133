"""Oracle restatement vs the committed golden vectors (produced by the reference itself, oracle/make_golden.py).
Runs anywhere (CPU).  Bit-exact: same torch build, same ops."""
import os

import numpy as np
import pytest
import torch

from oracle import restate, synth
from oracle.make_golden import CASES, GOLDEN


def _load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + '.npz')).items()}


def _close(a, b, name):
    if a.dtype.is_floating_point:
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6, msg=lambda m: name + ': ' + m)
    else:
        assert torch.equal(a, b), name


def test_forward_small_golden():
    case, g = CASES['forward_small'], _load('forward_small')
    sd = synth.make_state_dict(case['weight_seed'], schedules=restate.make_schedules())
    b = synth.make_batch(**case['batch'])
    pp, lp, _ = restate.center_pos(b['protein_pos'], b['init_ligand_pos'], b['batch_protein'], b['batch_ligand'])
    tr = {}
    out = restate.forward(sd, None, pp, b['protein_v'], b['batch_protein'], lp, b['init_ligand_v'], b['batch_ligand'], trace=tr)
    assert torch.equal(tr['edge_index'], g['edge_index'])
    _close(tr['e_w'], g['e_w'], 'e_w')
    _close(torch.stack(tr['all_x'][1:]), g['layer_x'], 'layer_x')
    _close(tr['all_h'][1], g['layer0_h'], 'layer0_h')
    _close(tr['all_h'][5], g['layer4_h'], 'layer4_h')
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'final_h'):
        _close(out[k], g[k], k)


@pytest.mark.parametrize('name', ['chain_trunc', 'chain_full_T20'])
def test_chain_golden(name):
    case, g = CASES[name], _load(name)
    sd = synth.make_state_dict(case['weight_seed'], case['cfg'], schedules=restate.make_schedules(case['cfg']))
    b = synth.make_batch(**case['batch'])
    S = case['num_steps'] or sd['betas'].shape[0]
    pn, vu = synth.make_tape(case['tape_seed'], S, len(b['batch_ligand']))
    r = restate.sample_diffusion(sd, case['cfg'], b['protein_pos'], b['protein_v'], b['batch_protein'],
                                 b['init_ligand_pos'], b['init_ligand_v'], b['batch_ligand'], pn, vu,
                                 num_steps=case['num_steps'])
    _close(r['pos'], g['pos'], 'pos')
    assert torch.equal(r['v'], g['v'])
    _close(torch.stack(r['pos_traj']), g['pos_traj'], 'pos_traj')
    assert torch.equal(torch.stack(r['v_traj']), g['v_traj'])
    _close(torch.stack(r['v0_traj']), g['v0_traj'], 'v0_traj')
    _close(torch.stack(r['vt_traj']), g['vt_traj'], 'vt_traj')


def _likelihood_inputs():
    case = CASES['likelihood']
    sd = synth.make_state_dict(case['weight_seed'], schedules=restate.make_schedules())
    b = synth.make_batch(**case['batch'])
    pn, vu = synth.make_tape(case['tape_seed'], 1, len(b['batch_ligand']))
    args = (b['protein_pos'], b['protein_v'], b['batch_protein'], b['init_ligand_pos'], b['init_ligand_v'], b['batch_ligand'])
    return case, sd, args, pn[0], vu[0]


def test_likelihood_golden():
    """oracle restatement of likelihood_estimation vs vectors written by the reference itself."""
    case, sd, args, pn, vu = _likelihood_inputs()
    g = _load('likelihood')
    kp, kv = restate.likelihood_estimation(sd, None, *args, torch.tensor(case['time_steps']), pn, vu)
    _close(kp, g['kl_pos'], 'kl_pos')
    _close(kv, g['kl_v'], 'kl_v')
    kp, kv = restate.likelihood_estimation(sd, None, *args, torch.full((3,), 1000))
    _close(kp, g['kl_pos_prior'], 'kl_pos_prior')
    _close(kv, g['kl_v_prior'], 'kl_v_prior')

"""CPU tests of the host-side mirror: config surface, data containers, size prior, sampling-driver bookkeeping."""
import numpy as np
import pytest
import torch

from targetdiff_b200 import atom_num
from targetdiff_b200.config import Config, check_supported, default_model_config, default_sampling_config, load_config
from targetdiff_b200.data import Batch, FOLLOW_BATCH, ProteinLigandData


def test_config_attribute_access_and_yaml(tmp_path):
    p = tmp_path / 'sampling.yml'
    p.write_text('model:\n  checkpoint: ./x.pt\nsample:\n  seed: 2021\n  num_samples: 100\n  num_steps: 1000\n  pos_only: False\n'
                 '  center_pos_mode: protein\n  sample_num_atoms: prior\n')
    c = load_config(str(p))
    assert c.sample.num_steps == 1000 and c.model.checkpoint == './x.pt' and c['sample']['seed'] == 2021
    d = default_sampling_config()
    assert dict(d.sample) == dict(c.sample)
    m = default_model_config()
    assert m.knn == 32 and m.num_layers == 9 and m.hidden_dim == 128 and m.ew_net_type == 'global'
    check_supported(m)
    m.knn = 65
    with pytest.raises(NotImplementedError):
        check_supported(m)
    assert isinstance(Config({'a': {'b': 1}}).a, Config)


def test_batch_container_matches_pyg_follow_batch_semantics():
    d0 = ProteinLigandData(protein_pos=torch.zeros(5, 3), protein_atom_feature=torch.ones(5, 27), protein_element=torch.ones(5),
                           ligand_pos=torch.zeros(2, 3), ligand_element=torch.ones(2), protein_filename='a.pdb')
    d1 = d0.clone()
    d1.protein_pos += 1
    b = Batch.from_data_list([d0, d1], follow_batch=FOLLOW_BATCH)
    assert b.protein_pos.shape == (10, 3) and b.protein_atom_feature.shape == (10, 27)
    assert b.protein_element_batch.tolist() == [0] * 5 + [1] * 5 and b.ligand_element_batch.tolist() == [0, 0, 1, 1]
    assert b.protein_filename == ['a.pdb', 'a.pdb'] and b.num_graphs == 2
    assert float(d0.protein_pos.sum()) == 0.0          # clone() is deep


def test_size_prior_table_and_space_size():
    from scipy.spatial.distance import pdist
    rng = np.random.RandomState(0)
    pos = rng.uniform(-15, 15, size=(200, 3))
    want = np.median(np.sort(pdist(pos))[::-1][:10])
    assert atom_num.get_space_size(pos) == pytest.approx(want, rel=0, abs=1e-12)
    t = atom_num._table()
    assert len(t['bounds']) == 9 and len(t['bins']) == 10
    for b in t['bins']:
        assert abs(sum(b['prob']) - 1.0) < 1e-3 and len(b['prob']) == len(b['num_atoms'])
    assert atom_num._get_bin_idx(10.0) == 0 and atom_num._get_bin_idx(100.0) == 9
    np.random.seed(3)
    a = [int(atom_num.sample_atom_num(30.0)) for _ in range(50)]
    np.random.seed(3)
    assert a == [int(atom_num.sample_atom_num(30.0)) for _ in range(50)]
    assert 2 <= min(a) and max(a) <= 86


class _FakeModel:
    """Stands in for ScorePosNet3D to test the driver's batching / un-batching without a GPU."""
    num_classes = 13

    def __init__(self):
        self.calls = []

    def sample_diffusion(self, protein_pos, protein_v, batch_protein, init_ligand_pos, init_ligand_v, batch_ligand, num_steps, pos_only,
                         center_pos_mode, stack_traj, noise_tape=None):
        S, nl = num_steps, len(batch_ligand)
        self.calls.append((int(batch_protein.max()) + 1, nl))
        ar = torch.arange(nl, dtype=torch.float32)
        return {'pos': ar[:, None].repeat(1, 3), 'v': batch_ligand.clone(), 'pos_traj': ar[None, :, None].repeat(S, 1, 3),
                'v_traj': batch_ligand[None].repeat(S, 1), 'v0_traj': torch.zeros(S, nl, 13), 'vt_traj': torch.zeros(S, nl, 13)}


def test_sample_diffusion_ligand_bookkeeping():
    from targetdiff_b200.sampling import sample_diffusion_ligand, seed_all
    seed_all(1)
    data = ProteinLigandData(protein_pos=torch.randn(40, 3) * 8, protein_atom_feature=torch.zeros(40, 27))
    m = _FakeModel()
    out = sample_diffusion_ligand(m, data, num_samples=5, batch_size=2, device='cpu', num_steps=4, sample_num_atoms='range')
    pos, v, pos_traj, v_traj, v0_traj, vt_traj, times = out
    assert [c[0] for c in m.calls] == [2, 2, 1] and len(times) == 3
    assert [len(p) for p in pos] == [1, 2, 3, 4, 5]                       # 'range': sample i has i+1 atoms (reference :52-53)
    assert all(p.dtype == np.float64 for p in pos) and pos_traj[1].shape == (4, 2, 3) and pos_traj[1].dtype == np.float64
    assert v_traj[4].shape == (4, 5) and v0_traj[2].shape == (4, 3, 13) and len(vt_traj) == 5
    assert (v[1] == 1).all()                                             # second sample of the first batch
    seed_all(2)
    out2 = sample_diffusion_ligand(m, data, num_samples=3, batch_size=16, device='cpu', num_steps=2, sample_num_atoms='prior')
    assert len(out2[0]) == 3 and all(len(p) >= 1 for p in out2[0])


def test_likelihood_estimation_host_math_against_oracle(monkeypatch):
    """The [Nl,13]-sized formulas around the network call (reference models/molopt_score_model.py:565-617).  The network itself
    is replaced by the oracle's CPU forward here (test infrastructure only -- the product path has no CPU execution); the GPU
    test `test_likelihood_estimation_vs_oracle` runs the real thing."""
    from oracle import restate, synth
    from targetdiff_b200 import ops
    from targetdiff_b200.score_model import ScorePosNet3D
    sd = synth.make_state_dict(5, schedules=restate.make_schedules())
    model = ScorePosNet3D(default_model_config(), synth.PROTEIN_FEATURE_DIM, synth.LIGAND_NUM_CLASSES)
    model.load_state_dict(sd, strict=True)
    b = synth.make_batch(4, 3, n_protein=40, ligand_sizes=[7, 10, 4])
    args = (b['protein_pos'], b['protein_v'], b['batch_protein'], b['init_ligand_pos'], b['init_ligand_v'], b['batch_ligand'])
    pn, vu = synth.make_tape(9, 1, len(b['batch_ligand']))

    def cpu_mean3(src, batch):
        n = int(batch.max()) + 1
        return torch.zeros(n, 3).index_add_(0, batch, src) / torch.bincount(batch, minlength=n).clamp(min=1)[:, None]

    def oracle_forward(pp, pv, bp, lp, lv, bl, time_step=None, **k):
        return restate.forward(sd, None, pp, pv, bp, lp, lv, bl)

    monkeypatch.setattr(ops, 'scatter_mean3', cpu_mean3)
    monkeypatch.setattr(model, 'forward', oracle_forward)
    for t in (torch.tensor([0, 999, 417]), torch.full((3,), 1000)):
        want = restate.likelihood_estimation(sd, None, *args, t, pn[0], vu[0])
        got = model.likelihood_estimation(*args, time_step=t, noise=(pn[0], vu[0]))
        for w, g in zip(want, got):
            torch.testing.assert_close(g, w, rtol=2e-5, atol=1e-6)
    with pytest.raises(ValueError):
        model.likelihood_estimation(*args, time_step=torch.tensor([0, 1000, 5]), noise=(pn[0], vu[0]))

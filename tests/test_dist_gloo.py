"""Multi-process host logic on CPU (gloo, world_size 2): pocket sharding and the start-up weight broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from targetdiff_b200 import dist as tdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from oracle import restate, synth
    from targetdiff_b200.config import default_model_config
    from targetdiff_b200.score_model import ScorePosNet3D
    r, w, _ = tdist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                        # different random init on every rank
    model = ScorePosNet3D(default_model_config(), synth.PROTEIN_FEATURE_DIM, synth.LIGAND_NUM_CLASSES)
    if rank == 0:
        model.load_state_dict(synth.make_state_dict(0, schedules=restate.make_schedules()), strict=True)
    tdist.broadcast_state_dict(model, 0)
    want = synth.make_state_dict(0, schedules=restate.make_schedules())
    same = all(torch.equal(v, want[k]) for k, v in model.state_dict().items())
    pockets = list(range(11))
    mine = tdist.shard_round_robin(pockets, rank, world)
    lpt = tdist.shard_longest_first([300, 700, 250, 640, 512, 333, 480], world)
    slow = tdist.max_over_ranks(1.0 + rank)
    q.put((rank, same, mine, lpt, slow))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_broadcast_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, same0, mine0, lpt0, slow0), (r1, same1, mine1, lpt1, slow1) = res
    assert same0 and same1                                   # rank 1 received rank 0's weights bit-exactly
    assert mine0 == [0, 2, 4, 6, 8, 10] and mine1 == [1, 3, 5, 7, 9]     # reference's i % world partition
    assert lpt0 == lpt1 and sorted(lpt0[0] + lpt0[1]) == list(range(7))  # identical, complete assignment on every rank
    loads = [sum([300, 700, 250, 640, 512, 333, 480][i] for i in part) for part in lpt0]
    assert abs(loads[0] - loads[1]) <= 250
    assert slow0 == slow1 == 2.0


def test_sharding_edge_cases():
    assert tdist.shard_round_robin([], 0, 8) == []
    assert tdist.shard_longest_first([], 4) == [[], [], [], []]
    assert tdist.shard_longest_first([5.0], 2) == [[0], []]


_POCKET_PDB = """HEADER    POCKET
ATOM      1  N   LEU A  36      36.155  52.241  55.687  1.00 30.88           N
ATOM      2  CA  LEU A  36      35.391  51.712  54.566  1.00 30.88           C
ATOM      3  C   LEU A  36      35.560  50.200  54.537  1.00 30.88           C
ATOM      4  O   LEU A  36      36.675  49.694  54.705  1.00 30.88           O
%sEND
"""


def _cli_worker(rank, world, port, tmp, q):
    """`python -m targetdiff_b200.cli sample_pockets` under a 2-rank gloo group; the CUDA sampler is replaced by a stub that
    records what it was asked to do (the kernels are covered by the gpu tests, this is the host / IO / sharding logic)."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import numpy as np
    from targetdiff_b200 import cli
    tdist.init_from_env('gloo')
    calls = []

    def fake_model(config, device, rank=0):
        return 'model-on-%s' % device

    def fake_sampler(model, data, num_samples, **kw):
        calls.append((int(data.protein_pos.shape[0]), num_samples, kw['num_steps'], float(torch.rand(1))))
        pos = [np.zeros((3, 3)) for _ in range(num_samples)]
        return pos, [np.zeros(3, dtype=np.int64)] * num_samples, pos, pos, pos, pos, [0.5]

    cli._load_model, cli.sample_diffusion_ligand = fake_model, fake_sampler
    done = cli.sample_pockets([os.path.join(tmp, 'sampling.yml'), '--pocket_dir', os.path.join(tmp, 'pockets'), '--result_path',
                               os.path.join(tmp, 'out'), '--num_samples', '2', '--device', 'cpu', '--schedule', 'longest_first'])
    q.put((rank, [d[0] for d in done], calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sample_pockets_cli(tmp_path):
    """pockets -> result_{i}.pt, sharded over the ranks of one job, reference result schema."""
    os.makedirs(tmp_path / 'pockets')
    extra = 'ATOM      5  CB  LEU A  36      35.842  52.361  53.252  1.00 30.88           C\n'
    for i in range(5):                                       # pocket i has 4 + i atoms: costs differ for longest_first
        (tmp_path / 'pockets' / ('p%d.pdb' % i)).write_text(_POCKET_PDB % (extra * i))
    (tmp_path / 'sampling.yml').write_text('model:\n  checkpoint: none.pt\nsample:\n  seed: 2021\n  num_samples: 7\n  num_steps: 11\n'
                                           '  pos_only: false\n  center_pos_mode: protein\n  sample_num_atoms: prior\n')
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_cli_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids0, ids1 = res[0][1], res[1][1]
    assert sorted(ids0 + ids1) == [0, 1, 2, 3, 4] and ids0 and ids1
    loads = [sum(4 + i for i in ids) for ids in (ids0, ids1)]
    assert abs(loads[0] - loads[1]) <= 4                     # longest-first balance by atom count
    calls = res[0][2] + res[1][2]
    assert all(c[1] == 2 and c[2] == 11 for c in calls)      # --num_samples override, num_steps from the yml
    assert len({c[3] for c in calls}) == 1                   # every pocket starts from the same seed, like one reference process per pocket
    for i in range(5):
        r = torch.load(tmp_path / 'out' / ('result_%d.pt' % i), weights_only=False)
        assert sorted(r) == ['data', 'pred_ligand_pos', 'pred_ligand_pos_traj', 'pred_ligand_v', 'pred_ligand_v_traj', 'time']
        assert r['data'].protein_pos.shape[0] == 4 + i and len(r['pred_ligand_pos']) == 2
    assert os.path.isfile(tmp_path / 'out' / 'sample.yml')

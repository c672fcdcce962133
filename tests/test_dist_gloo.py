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

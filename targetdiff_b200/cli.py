"""Command line for the sampling path (same arguments / outputs as the reference's scripts).

    python -m targetdiff_b200.cli sample_for_pocket configs/sampling.yml --pdb_path pocket.pdb [--num_samples N] [--result_path DIR]
        (reference scripts/sample_for_pocket.py:34-93: sample ligands into one pocket given as a PDB file)

    [torchrun --nproc-per-node N -m] python -m targetdiff_b200.cli sample_pockets configs/sampling.yml --pocket_dir DIR | --pocket_list FILE
            [-i ID] [--schedule round_robin|longest_first] [--result_path DIR] [--num_samples N] [--batch_size B]
        (reference scripts/sample_diffusion.py:118-186 + scripts/batch_sample_diffusion.sh: pocket i -> `result_{i}.pt`, pockets
        assigned to workers round-robin; here the workers are the ranks of one torchrun job, one weight broadcast, no other collective)

Result file: `<result_path>/sample.pt` (sample_for_pocket) or `<result_path>/result_{i}.pt` (sample_pockets) = {'data', 'pred_ligand_pos', 'pred_ligand_v', 'pred_ligand_pos_traj', 'pred_ligand_v_traj', 'time'}
-- the schema scripts/sample_diffusion.py:175-182 writes and scripts/evaluate_diffusion.py:70-76 reads (positions float64, per-sample
lists; trajectories [steps, atoms, 3]).  Molecule reconstruction / SDF writing needs RDKit + OpenBabel and stays out of scope."""
import argparse
import os
import shutil
import sys

import torch

from .config import load_config
from .pocket import pdb_to_pocket_data
from .sampling import sample_diffusion_ligand, seed_all
from .score_model import ScorePosNet3D


PROTEIN_FEATURE_DIM = 27                                                    # reference utils/transforms.py:115-132
LIGAND_ATOM_MODE_CLASSES = {'basic': 8, 'add_aromatic': 13, 'full': 23}     # len(MAP_ATOM_TYPE_*_TO_INDEX), utils/transforms.py:11-66,143-149


def build_result(data, outputs):
    pred_pos, pred_v, pred_pos_traj, pred_v_traj, pred_v0_traj, pred_vt_traj, time_list = outputs
    return {'data': data, 'pred_ligand_pos': pred_pos, 'pred_ligand_v': pred_v, 'pred_ligand_pos_traj': pred_pos_traj,
            'pred_ligand_v_traj': pred_v_traj, 'time': time_list}


def _load_model(config, device, rank=0):
    """Checkpoint -> engine-backed model on `device`.  Under torchrun only rank 0's weights count: one flat broadcast
    (targetdiff_b200.dist.broadcast_state_dict) replaces the per-process checkpoint parsing of the reference's shell loop."""
    from . import dist as tdist
    ckpt = torch.load(config.model.checkpoint, map_location='cpu', weights_only=False)
    tc = ckpt['config']
    get = (lambda c, k: getattr(c, k)) if hasattr(tc, 'model') else (lambda c, k: c[k])
    # feature widths come from the checkpoint's featurisers like the reference (scripts/sample_diffusion.py:140-160):
    # protein = 6 elements + 20 residue types + backbone flag, ligand = class count of data.transform.ligand_atom_mode
    try:
        mode = get(get(get(tc, 'data'), 'transform'), 'ligand_atom_mode')
    except (AttributeError, KeyError, TypeError):
        mode = 'add_aromatic'
    if mode not in LIGAND_ATOM_MODE_CLASSES:
        raise NotImplementedError('checkpoint ligand_atom_mode=%r (known: %s)' % (mode, sorted(LIGAND_ATOM_MODE_CLASSES)))
    model = ScorePosNet3D(get(tc, 'model'), PROTEIN_FEATURE_DIM, LIGAND_ATOM_MODE_CLASSES[mode])
    if rank == 0:
        model.load_state_dict(ckpt['model'])
    model = model.to(device)
    tdist.broadcast_state_dict(model, src=0)
    return model


def list_pockets(pocket_dir=None, pocket_list=None):
    """Task list in a sharding-independent order: sorted *.pdb of a directory, or the lines of a list file."""
    if (pocket_dir is None) == (pocket_list is None):
        raise ValueError('give exactly one of --pocket_dir / --pocket_list')
    if pocket_dir is not None:
        paths = sorted(os.path.join(pocket_dir, f) for f in os.listdir(pocket_dir) if f.lower().endswith('.pdb'))
    else:
        base = os.path.dirname(os.path.abspath(pocket_list))
        with open(pocket_list) as f:
            paths = [ln.strip() for ln in f if ln.strip() and not ln.startswith('#')]
        paths = [p if os.path.isabs(p) else os.path.join(base, p) for p in paths]
    if not paths:
        raise ValueError('no pockets found')
    return paths


def assign_pockets(paths, rank, world, schedule='round_robin', data_id=None):
    """Pocket ids this rank works on.  round_robin = the reference's `i % NODE_ALL == NODE_THIS`
    (scripts/batch_sample_diffusion.sh:15-21); longest_first balances by ATOM-record count (cost ~ nodes x k)."""
    from . import dist as tdist
    ids = list(range(len(paths)))
    if data_id is not None:
        if not 0 <= data_id < len(paths):
            raise ValueError('data_id %d outside 0..%d' % (data_id, len(paths) - 1))
        return [data_id] if data_id % world == rank else []
    if schedule == 'round_robin':
        return tdist.shard_round_robin(ids, rank, world)
    if schedule == 'longest_first':
        costs = []
        for p in paths:
            with open(p) as f:
                costs.append(sum(1 for ln in f if ln.startswith('ATOM')))
        return tdist.shard_longest_first(costs, world)[rank]
    raise ValueError('schedule %r' % (schedule,))


def sample_pockets(argv):
    from . import dist as tdist
    ap = argparse.ArgumentParser(prog='targetdiff_b200.cli sample_pockets')
    ap.add_argument('config', type=str)
    ap.add_argument('--pocket_dir', type=str)
    ap.add_argument('--pocket_list', type=str)
    ap.add_argument('-i', '--data_id', type=int)
    ap.add_argument('--schedule', type=str, default='round_robin', choices=('round_robin', 'longest_first'))
    ap.add_argument('--device', type=str)
    ap.add_argument('--batch_size', type=int, default=100)
    ap.add_argument('--result_path', type=str, default='./outputs')
    ap.add_argument('--num_samples', type=int)
    a = ap.parse_args(argv)
    config = load_config(a.config)
    rank, world, local_rank = tdist.init_from_env()
    device = a.device or 'cuda:%d' % local_rank
    paths = list_pockets(a.pocket_dir, a.pocket_list)
    mine = assign_pockets(paths, rank, world, a.schedule, a.data_id)
    model = _load_model(config, device, rank)
    n = a.num_samples if a.num_samples is not None else config.sample.num_samples
    os.makedirs(a.result_path, exist_ok=True)
    if rank == 0:
        shutil.copyfile(a.config, os.path.join(a.result_path, 'sample.yml'))
    done = []
    for i in mine:
        seed_all(config.sample.seed)          # the reference starts one process per pocket, each seeded the same way (:133)
        data = pdb_to_pocket_data(paths[i])
        outputs = sample_diffusion_ligand(model, data, n, batch_size=a.batch_size, device=device, num_steps=config.sample.num_steps,
                                          pos_only=config.sample.pos_only, center_pos_mode=config.sample.center_pos_mode,
                                          sample_num_atoms=config.sample.sample_num_atoms)
        torch.save(build_result(data, outputs), os.path.join(a.result_path, 'result_%d.pt' % i))
        done.append((i, len(outputs[0]), sum(outputs[-1])))
        print('[rank %d/%d] pocket %d (%s): %d molecules, %.1f s' % (rank, world, i, os.path.basename(paths[i]), done[-1][1], done[-1][2]))
    return done


def sample_for_pocket(argv):
    ap = argparse.ArgumentParser(prog='targetdiff_b200.cli sample_for_pocket')
    ap.add_argument('config', type=str)
    ap.add_argument('--pdb_path', type=str, required=True)
    ap.add_argument('--device', type=str, default='cuda:0')
    ap.add_argument('--batch_size', type=int, default=100)
    ap.add_argument('--result_path', type=str, default='./outputs_pdb')
    ap.add_argument('--num_samples', type=int)
    a = ap.parse_args(argv)
    config = load_config(a.config)
    seed_all(config.sample.seed)
    model = _load_model(config, a.device)
    data = pdb_to_pocket_data(a.pdb_path)
    n = a.num_samples if a.num_samples is not None else config.sample.num_samples
    outputs = sample_diffusion_ligand(model, data, n, batch_size=a.batch_size, device=a.device, num_steps=config.sample.num_steps,
                                      pos_only=config.sample.pos_only, center_pos_mode=config.sample.center_pos_mode,
                                      sample_num_atoms=config.sample.sample_num_atoms)
    os.makedirs(a.result_path, exist_ok=True)
    shutil.copyfile(a.config, os.path.join(a.result_path, 'sample.yml'))
    torch.save(build_result(data, outputs), os.path.join(a.result_path, 'sample.pt'))
    print('Sample done! %d molecules, %.1f s' % (len(outputs[0]), sum(outputs[-1])))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    commands = {'sample_for_pocket': sample_for_pocket, 'sample_pockets': sample_pockets}
    if not argv or argv[0] not in commands:
        raise SystemExit(__doc__)
    commands[argv[0]](argv[1:])


if __name__ == '__main__':
    main()

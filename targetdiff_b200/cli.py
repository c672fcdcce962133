"""Command line for the sampling path (same arguments / outputs as the reference's scripts).

    python -m targetdiff_b200.cli sample_for_pocket configs/sampling.yml --pdb_path pocket.pdb [--num_samples N] [--result_path DIR]
        (reference scripts/sample_for_pocket.py:34-93: sample ligands into one pocket given as a PDB file)

Result file: `<result_path>/sample.pt` = {'data', 'pred_ligand_pos', 'pred_ligand_v', 'pred_ligand_pos_traj', 'pred_ligand_v_traj', 'time'}
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


def build_result(data, outputs):
    pred_pos, pred_v, pred_pos_traj, pred_v_traj, pred_v0_traj, pred_vt_traj, time_list = outputs
    return {'data': data, 'pred_ligand_pos': pred_pos, 'pred_ligand_v': pred_v, 'pred_ligand_pos_traj': pred_pos_traj,
            'pred_ligand_v_traj': pred_v_traj, 'time': time_list}


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
    ckpt = torch.load(config.model.checkpoint, map_location='cpu', weights_only=False)
    model = ScorePosNet3D(ckpt['config'].model if hasattr(ckpt['config'], 'model') else ckpt['config']['model'], 27, 13)
    model.load_state_dict(ckpt['model'])
    model = model.to(a.device)
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
    if not argv or argv[0] not in ('sample_for_pocket',):
        raise SystemExit(__doc__)
    sample_for_pocket(argv[1:])


if __name__ == '__main__':
    main()

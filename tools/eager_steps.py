"""Run a few denoising steps of the bench workload (cfg3) WITHOUT CUDA-graph replay so that ncu sees the individual launches.

    TDIFF_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/eager_steps.py 3
    python tools/launch_shares.py launches.csv > shares.csv        # one steady-state step (between the last two knn launches)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('TDIFF_NO_GRAPH', '1')
import torch
from oracle import synth, restate
from targetdiff_b200.config import default_model_config
from targetdiff_b200.score_model import ScorePosNet3D

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model = ScorePosNet3D(default_model_config(), synth.PROTEIN_FEATURE_DIM, synth.LIGAND_NUM_CLASSES)
model.load_state_dict(synth.make_state_dict(0, None, schedules=restate.make_schedules(None)))
model = model.to('cuda')
b = synth.make_batch(1, 640, n_protein=300, n_ligand=20, distinct_pockets=64)
args = tuple(b[k].to('cuda') for k in ('protein_pos', 'protein_v', 'batch_protein', 'init_ligand_pos', 'init_ligand_v', 'batch_ligand'))
out = model.sample_diffusion(*args, num_steps=steps, center_pos_mode='protein', seed=1)
torch.cuda.synchronize()
print('ok', out['pos'].shape)

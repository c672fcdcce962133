"""Print the clock64 timeline of one edge_mlp_v3 launch (CTA 0, tiles 4..11) on the bench workload.

    TDIFF_V3_TS=1 python tools/v3_timeline.py     # 1 = first big launch (layer 0 key MLP), 2 = layer 0 value MLP (+aggregation), ...

Columns per role are the `stamp(role, tile, ev)` events in targetdiff_b200/csrc/edge_mlp_v3.cu, in SM clocks relative to tile 4.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, restate
from targetdiff_b200.config import default_model_config
from targetdiff_b200.score_model import ScorePosNet3D

os.environ.setdefault('TDIFF_V3_TS', '1')
cfg = default_model_config()
model = ScorePosNet3D(cfg, synth.PROTEIN_FEATURE_DIM, synth.LIGAND_NUM_CLASSES)
model.load_state_dict(synth.make_state_dict(0, None, schedules=restate.make_schedules(None)))
model = model.to('cuda')
b = synth.make_batch(1, 640, n_protein=300, n_ligand=20, distinct_pockets=64)
args = tuple(b[k].to('cuda') for k in ('protein_pos', 'protein_v', 'batch_protein', 'init_ligand_pos', 'init_ligand_v', 'batch_ligand'))
model(*args)
torch.cuda.synchronize()

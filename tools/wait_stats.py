"""In-situ wait accounting of the edge kernel's warp roles (no profiler attached: ordinary sampling steps of the bench workload).

    TDIFF_VARIANT=waitstats TDIFF_NVCC_EXTRA=-DTDIFF_WAIT_STATS python -m targetdiff_b200.build
    TDIFF_LIB=$PWD/targetdiff_b200/libtdiff_waitstats.so python tools/wait_stats.py [steps]

Prints, per role, the share of its tile loop spent in each mbarrier wait (SM clock cycles summed over every edge_mlp_v4 launch of the
timed steps; see the slot list in edge_mlp_v4.cu).  The instrumented build is for diagnosis only (clock reads in the hot loops)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, restate
from targetdiff_b200 import _lib
from targetdiff_b200.config import default_model_config
from targetdiff_b200.score_model import ScorePosNet3D

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
if not hasattr(raw, 'tdiff_debug_wait_stats'):
    raise SystemExit('%s is not a TDIFF_WAIT_STATS build' % _lib.LIB_PATH)
model = ScorePosNet3D(default_model_config(), synth.PROTEIN_FEATURE_DIM, synth.LIGAND_NUM_CLASSES)
model.load_state_dict(synth.make_state_dict(0, None, schedules=restate.make_schedules(None)))
model = model.to('cuda')
b = synth.make_batch(1, 640, n_protein=300, n_ligand=20, distinct_pockets=64)
args = tuple(b[k].to('cuda') for k in ('protein_pos', 'protein_v', 'batch_protein', 'init_ligand_pos', 'init_ligand_v', 'batch_ligand'))
model.sample_diffusion(*args, num_steps=3, center_pos_mode='protein', seed=1)          # warm-up (also builds the ligand-free cache)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
raw.tdiff_debug_wait_stats(None, 1)
model.sample_diffusion(*args, num_steps=steps, center_pos_mode='protein', seed=2)
torch.cuda.synchronize()
assert raw.tdiff_debug_wait_stats(buf, 0) == 0
v = [int(x) for x in buf]
out = {
    'steps': steps,
    'row_warps': {'loop_cycles_per_warp_launch': v[0] / max(v[11], 1), 'wait_S_FULL': v[1] / max(v[0], 1), 'wait_DPRE_FULL': v[2] / max(v[0], 1),
                  'wait_A_EMPTY': v[3] / max(v[0], 1)},
    'gather_warps': {'loop_cycles_per_warp_launch': v[4] / max(v[12], 1), 'wait_S_EMPTY': v[5] / max(v[4], 1), 'copy_issue_and_completion': v[6] / max(v[4], 1),
                     'mma_warp_wait_G_FULL (of all 4 gather warps\' time)': v[7] / max(v[4], 1), 'mma_warp_wait_D_EMPTY_A_FULL (same)': v[8] / max(v[4], 1)},
    'epilogue_warps': {'loop_cycles_per_warp_launch': v[9] / max(v[13], 1), 'wait_D_FULL': v[10] / max(v[9], 1)},
    'raw': v,
}
print(json.dumps(out, indent=1))

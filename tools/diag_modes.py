import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import synth
from test_gpu_configs import _model, _args
b = synth.make_batch(12, 5, n_protein=90, ligand_sizes=[20, 1, 33, 7, 45])
S = 8
pn, vu = synth.make_tape(3, S, int(b['init_ligand_pos'].shape[0]))
def run(env):
    for k in ('TDIFF_KNN_FULL','TDIFF_NO_SLOT_KEEP','TDIFF_FREE_DEPTH','TDIFF_NO_RESTRICT','TDIFF_NO_GRAPH'):
        os.environ.pop(k, None)
    os.environ.update(env)
    model, _ = _model(2)
    out = model.sample_diffusion(*_args(b), num_steps=S, center_pos_mode='protein', noise_tape=(pn, vu), stack_traj=True)
    return out['pos_traj']
base = run({})
for env in [{}, {}, {'TDIFF_KNN_FULL':'1'}, {'TDIFF_NO_SLOT_KEEP':'1'}, {'TDIFF_FREE_DEPTH':'0'}, {'TDIFF_NO_RESTRICT':'1'}, {'TDIFF_NO_GRAPH':'1'},
            {'TDIFF_NO_SLOT_KEEP':'1','TDIFF_KNN_FULL':'1'}, {'TDIFF_FREE_DEPTH':'0','TDIFF_NO_SLOT_KEEP':'1'}]:
    r = run(env)
    d = (r - base).abs().flatten(1).max(1).values
    print(env, 'max diff per step:', ['%.1e' % x for x in d.tolist()])

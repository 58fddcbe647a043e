"""Stage-by-stage comparison of a math mode against the oracle (debugging aid; prints every stage's error)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_case, load_gnn
from gnn_branching_b200 import GraphNet
from oracle import graphnet_oracle as O

math = sys.argv[1] if len(sys.argv) > 1 else 'tc'
arch = sys.argv[2] if len(sys.argv) > 2 else 'base'
weights = sys.argv[3] if len(sys.argv) > 3 else 'random'
fr, ref = load_case(arch, 'fr')
sd = load_gnn(weights)
stages = {}
s_or, _ = O.gnn_forward(sd, fr, stages=stages)
m = GraphNet(2, 64, math=math); m.load_state_dict(sd); m = m.eval().cuda()
sc = m.scorer(0); sc.set_option('snapshot', 1)
print('scoring...', flush=True)
best, idx, scores = m.score_frontier(fr.to('cuda'))
torch.cuda.synchronize()
print('done', flush=True)
for name, want in stages.items():
    key = name
    if '_relax' in name:
        if not name.startswith('t0_') or math == 'tc': continue
        key = name.replace('t0_fwd_relax', 'relax_f').replace('t0_bwd_relax', 'relax_b')
    got = sc.snapshot(key).reshape(want.shape)
    err = float((got - want).abs().max()) / max(float(want.abs().max()), 1e-20)
    print(f'{name:18s} max|ref| {float(want.abs().max()):10.4f}  norm err {err:.3e}  nan {int(torch.isnan(got).sum())}')
print(O.parity_report(scores.cpu(), ref[f'scores_{weights}'], fr.mask, idx.cpu()))

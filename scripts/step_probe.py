"""Timing probe of FrontierStep: python scripts/step_probe.py <arch> <parents per step>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from golden_io import load_root, load_gnn, GOLDEN
from gnn_branching_b200 import GraphNet, FrontierStep

arch, PB = sys.argv[1], int(sys.argv[2])
net, lbs, ubs, wp, bp = load_root(arch)
x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(-1)
model = GraphNet(2, 64, math='tc'); model.load_state_dict(load_gnn('random')); model = model.eval().cuda()
fs = FrontierStep(model, net, x, 0.145, wp, bp, capacity=1 << 15, decision_bound=float('inf'))
fs.seed_root(lbs, ubs)
while len(fs.queue) < PB:
    st = fs.step(PB)
print('grown to', len(fs.queue), st, flush=True)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    st = fs.step(PB)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f'step {arch}: {st.picked} parents -> {st.children} children in {dt * 1e3:.1f} ms = {st.children / dt:.0f} children/s; second pass {st.second_pass}, infeasible {st.infeasible}, added {st.added}, queue {len(fs.queue)}', flush=True)

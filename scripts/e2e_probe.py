"""Host-buffer (end-to-end) vs device-resident scoring time for several wave sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_root, load_gnn
from gnn_branching_b200 import GraphNet, synthetic_frontier

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net, lbs, ubs, wp, bp = load_root('base')
fr = synthetic_frontier(net, lbs, ubs, wp, bp, B, seed=7, device='cuda')
hf = fr.cpu().pin()
for chunk in (128, 256, 512, 1024):
    m = GraphNet(2, 64, chunk=chunk); m.load_state_dict(load_gnn('random')); m = m.eval().cuda()
    for mode, f in (('device', fr), ('host', hf)):
        for _ in range(2):
            m.score_frontier(f, return_scores=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            m.score_frontier(f, return_scores=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f'chunk {chunk:5d} {mode:6s} {dt * 1e3:7.2f} ms  {B / dt:9.0f} /s', flush=True)

import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_root
from gnn_branching_b200 import synthetic_frontier, babsr_frontier, Scorer
sc = Scorer(0)
for arch in ('base', 'wide', 'deep'):
    net, lbs, ubs, wp, bp = load_root(arch)
    for B in (64, 1024, 8192):
        fr = synthetic_frontier(net, lbs, ubs, wp, bp, B, seed=5, device='cuda')
        for _ in range(3): babsr_frontier(fr, scorer=sc)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): babsr_frontier(fr, scorer=sc)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f'{arch} B={B:5d} {dt*1e3:8.3f} ms/call {B/dt:10.0f} /s', flush=True)

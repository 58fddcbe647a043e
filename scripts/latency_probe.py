"""Small-batch latency of the scoring call (the reference's native usage is B = 1, graph_score.py:21-56)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_root, load_gnn
from gnn_branching_b200 import GraphNet, synthetic_frontier

for arch in ('base', 'wide', 'deep'):
    net, lbs, ubs, wp, bp = load_root(arch)
    m = GraphNet(2, 64); m.load_state_dict(load_gnn('random')); m = m.eval().cuda()
    for B in (1, 4, 16, 64):
        fr = synthetic_frontier(net, lbs, ubs, wp, bp, B, seed=3, device='cuda')
        sc = m.scorer(0); sc.set_network(fr.net, key=fr.net.key)
        for _ in range(3): sc.score(fr, return_scores=False, check=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): sc.score(fr, return_scores=False, check=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f'{arch} B={B:3d} {dt * 1e3:7.3f} ms per call  ({B / dt:8.0f} subdomains/s)', flush=True)

"""Do pinned H2D copies on a side stream overlap the scoring kernels?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_root, load_gnn
from gnn_branching_b200 import GraphNet, synthetic_frontier

net, lbs, ubs, wp, bp = load_root('base')
fr = synthetic_frontier(net, lbs, ubs, wp, bp, 1024, seed=7, device='cuda')
m = GraphNet(2, 64, chunk=256); m.load_state_dict(load_gnn('random')); m = m.eval().cuda()
src = torch.empty(142 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
dst = torch.empty_like(src, device='cuda')
side = torch.cuda.Stream()
def t(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def score(): m.score_frontier(fr, return_scores=False)
def copy():
    with torch.cuda.stream(side): dst.copy_(src, non_blocking=True)
def copy_many():
    with torch.cuda.stream(side):
        n = src.numel() // 140
        for i in range(140): dst[i * n:(i + 1) * n].copy_(src[i * n:(i + 1) * n], non_blocking=True)
def both(): copy(); score()
def both_many(): copy_many(); score()
print('score %.2f ms  copy %.2f ms  copy_many %.2f ms  both %.2f ms  both_many %.2f ms' % (t(score), t(copy), t(copy_many), t(both), t(both_many)))

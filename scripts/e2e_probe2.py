"""Where the end-to-end time goes: H2D copy alone, device-resident scoring alone, host-buffer call, for a few wave sizes."""
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
tens = []
for v in hf.tensors().values():
    tens += v if isinstance(v, list) else [v]
dst = [torch.empty_like(t, device='cuda') for t in tens]
nbytes = sum(t.numel() * 4 for t in tens)
for _ in range(2):
    for d, t in zip(dst, tens): d.copy_(t, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for d, t in zip(dst, tens): d.copy_(t, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f'H2D alone: {len(tens)} copies, {nbytes/1e6:.1f} MB in {dt*1e3:.2f} ms = {nbytes/dt/1e9:.1f} GB/s', flush=True)
big = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device='cuda')
dbig.copy_(big, non_blocking=True); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f'H2D one block: {nbytes/dt/1e9:.1f} GB/s', flush=True)
for chunk in (512, 1024):
    m = GraphNet(2, 64, chunk=chunk); m.load_state_dict(load_gnn('random')); m = m.eval().cuda()
    for mode, f in (('device', fr), ('host', hf)):
        for _ in range(3): m.score_frontier(f, return_scores=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): m.score_frontier(f, return_scores=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f'chunk {chunk:5d} {mode:6s} {dt * 1e3:7.2f} ms  {B / dt:9.0f} /s', flush=True)
for Bs in (64, 256, 704, 2048):
    f2 = synthetic_frontier(net, lbs, ubs, wp, bp, Bs, seed=9, device='cuda')
    m = GraphNet(2, 64, chunk=Bs); m.load_state_dict(load_gnn('random')); m = m.eval().cuda()
    for _ in range(3): m.score_frontier(f2, return_scores=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): m.score_frontier(f2, return_scores=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f'device-resident wave of {Bs:5d}: {dt*1e3:7.3f} ms  {Bs/dt:9.0f} /s', flush=True)

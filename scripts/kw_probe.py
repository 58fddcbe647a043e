"""Timing probe of gnnb_child_bounds / gnnb_kw_bounds: python scripts/kw_probe.py <arch> <B>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from golden_io import load_root, GOLDEN
from gnn_branching_b200 import Scorer

arch, B = sys.argv[1], int(sys.argv[2])
net, lbs, ubs, wp, bp = load_root(arch)
x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(1, -1)
sc = Scorer(0); sc.set_network(net, key=net.key)
L = net.L
g = torch.Generator().manual_seed(3)
plb = [t.reshape(1, -1).repeat(B, 1).cuda() for t in lbs]
pub = [t.reshape(1, -1).repeat(B, 1).cuda() for t in ubs]
lay = torch.randint(0, L, (B,), generator=g)
idx = torch.zeros(B, dtype=torch.long)
for b in range(B):
    amb = ((lbs[int(lay[b]) + 1] < 0) & (ubs[int(lay[b]) + 1] > 0)).nonzero().view(-1)
    idx[b] = amb[int(torch.randint(0, len(amb), (1,), generator=g))]
ch = torch.randint(0, 2, (B,), generator=g)
W = wp.reshape(1, -1).repeat(B, 1).cuda(); bb = torch.full((B,), float(bp)).cuda(); xx = x.cuda()
for name, fn in (('child_bounds', lambda: sc.child_bounds(xx, 0.145, W, bb, plb, pub, lay, idx, ch)),
                 ('kw_bounds', lambda: sc.kw_bounds(xx, 0.145, W, bb))):
    fn(); torch.cuda.synchronize()
    t0 = time.time(); n = 3
    for _ in range(n): fn()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    print(f'{name} {arch} B={B}: {dt * 1e3:.1f} ms/call, {B / dt:.0f} domains/s', flush=True)

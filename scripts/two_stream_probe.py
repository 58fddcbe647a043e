"""Would two waves on two streams hide each other's launch fill and drain?  Two contexts (own workspaces), each scoring half of
the frontier on its own stream, against one context scoring all of it: python scripts/two_stream_probe.py <arch> <B>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_root, load_gnn
from gnn_branching_b200 import Scorer, synthetic_frontier

arch, B = sys.argv[1], int(sys.argv[2])
net, lbs, ubs, wp, bp = load_root(arch)
sd = load_gnn('random')
fr = synthetic_frontier(net, lbs, ubs, wp, bp, B, seed=7, device='cuda')
scs = []
for i in range(2):
    sc = Scorer(0); sc.set_gnn(sd); sc.set_network(net, key=net.key); scs.append(sc)
halves = [fr.slice(0, B // 2).contiguous(), fr.slice(B // 2, B).contiguous()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def one():
    scs[0].score(fr, return_scores=False, check=False)

def two():
    ev = torch.cuda.Event(); ev.record()
    for i in range(2):
        streams[i].wait_event(ev)
        with torch.cuda.stream(streams[i]):
            scs[i].score(halves[i], return_scores=False, check=False)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)

for name, fn in (('one stream x %d' % B, one), ('two streams x %d' % (B // 2), two), ('one stream x %d' % B, one), ('two streams x %d' % (B // 2), two)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{arch} {name}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} subdomains/s', flush=True)

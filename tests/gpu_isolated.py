"""Checks executed in their own process (own CUDA context, time-out) by tests/test_gpu_parity.py (`_run_isolated`):
exit code 0 = pass.  Usage: python tests/gpu_isolated.py <check> <arg>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                                                     # noqa: E402

from golden_io import load_case, load_gnn, load_root             # noqa: E402
from gnn_branching_b200 import GraphNet, synthetic_frontier      # noqa: E402


def check_gather_prefetch(arch):
    """The propagation kernel variants k_tc_prop_pf (option value 1) and k_tc_prop_pf25 (2) must give bit-identical scores,
    winners and indices."""
    fr, _ = load_case(arch, 'fr')
    model = GraphNet(2, 64, math='tc')
    model.load_state_dict(load_gnn('random'))
    model = model.eval().cuda()
    model.scorer(0).set_option('fuse', 0)          # the variants belong to the stand-alone propagation kernel
    for f in (fr.to('cuda'), synthetic_frontier(*load_root(arch), 37, seed=5, device='cuda')):
        model.scorer(0).set_option('gather_prefetch', 0)
        b0, i0, s0 = model.score_frontier(f)
        for variant in (1, 2):
            model.scorer(0).set_option('gather_prefetch', variant)
            b1, i1, s1 = model.score_frontier(f)
            torch.cuda.synchronize()
            assert torch.equal(s0, s1) and torch.equal(i0, i1) and torch.equal(b0, b1), f'gather_prefetch={variant} changes the results'


def check_fused(arch):
    """k_tc_fused (default) against the two-launch path: both split the same fp32 propagation accumulator into the same fp16
    hi / lo operand, but the first chain GEMM reads it from tensor memory instead of shared memory and the hardware's
    accumulation order differs (measured 1.3e-6 of the largest score) — held to 1e-5, winners equal wherever the top-2 margin
    exceeds that; B chosen so that CTAs walk several items, the last group of 4 subdomains is ragged, and (B = 1) most of the
    grid is idle."""
    model = GraphNet(2, 64, math='tc')
    model.load_state_dict(load_gnn('random'))
    model = model.eval().cuda()
    sc = model.scorer(0)
    for B in (3, 1, 301, 1024):
        f = synthetic_frontier(*load_root(arch), B, seed=5 + B, device='cuda')
        sc.set_option('fuse', 1)
        b1, i1, s1 = model.score_frontier(f)
        torch.cuda.synchronize()
        sc.set_option('fuse', 0)
        b0, i0, s0 = model.score_frontier(f)
        torch.cuda.synchronize()
        _close(s0, s1, i0, i1, f.mask, f'fused {arch} B={B}')


def _close(s0, s1, i0, i1, mask, what, rtol=1e-5):
    scale = float(s0.abs().max())
    err = float((s0 - s1).abs().max()) / scale
    masked = torch.where(mask != 0, s0, torch.full_like(s0, float('-inf')))
    top2 = masked.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * rtol * scale
    print(f'{what}: identical={torch.equal(s0, s1)} max rel diff {err:.3e}, winners compared {int(safe.sum())}/{len(safe)}', flush=True)
    assert err <= rtol, f'{what}: scores differ by {err:.3e} of the largest score'
    assert torch.equal(i0[safe], i1[safe]), f'{what}: winners differ'


def check_kw_bounds(arch):
    """gnnb_kw_bounds against the reference's own DualNetwork: root bounds of the net (tests/golden/nets.npz) and child domains
    with one ReLU fixed, all children in one batched call (tests/golden/kw_children.npz); 5e-5 of the largest bound."""
    import numpy as np
    from golden_io import GOLDEN
    from gnn_branching_b200 import Scorer
    from oracle import kw_bounds_oracle as KW
    net, lbs, ubs, wp, bp = load_root(arch)
    x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(1, -1)
    sc = Scorer(0)
    sc.set_network(net, key=net.key)

    def err(a, b):
        return float((a.reshape(-1).cpu() - b.reshape(-1)).abs().max()) / max(1.0, float(b.abs().max()))

    gl, gu = sc.kw_bounds(x, 0.145, wp.reshape(1, -1), torch.tensor([bp]))
    for k in range(net.L + 2):
        assert err(gl[k], lbs[k]) <= 5e-5 and err(gu[k], ubs[k]) <= 5e-5, ('root', k, err(gl[k], lbs[k]), err(gu[k], ubs[k]))
    z = dict(np.load(os.path.join(GOLDEN, 'kw_children.npz')))
    nc = int(z[f'{arch}_ncases'])
    plbs, pubs = [], []
    for c in range(nc):
        lay, idx, choice = z[f'{arch}_c{c}_decision'].tolist()
        a, b = KW.split_bounds(lbs, ubs, (lay, idx), choice)
        plbs.append(a); pubs.append(b)
    plb = [torch.stack([plbs[c][k] for c in range(nc)]) for k in range(net.L)]
    pub = [torch.stack([pubs[c][k] for c in range(nc)]) for k in range(net.L)]
    gl, gu = sc.kw_bounds(x, 0.145, wp.reshape(1, -1).repeat(nc, 1), torch.full((nc,), float(bp)), plb, pub)
    for c in range(nc):
        for k in range(1, net.L + 2):
            rl, ru = torch.from_numpy(z[f'{arch}_c{c}_lb{k}']), torch.from_numpy(z[f'{arch}_c{c}_ub{k}'])
            l, u = gl[k][c].cpu(), gu[k][c].cpu()
            if k == net.L + 1:                     # init_kw_bounds :285-286 intersects the output bounds with the parent's
                l, u = torch.max(l, lbs[-1]), torch.min(u, ubs[-1])
            assert err(l, rl) <= 5e-5 and err(u, ru) <= 5e-5, ('child', c, k, err(l, rl), err(u, ru))


if __name__ == '__main__':
    {'gather_prefetch': check_gather_prefetch, 'kw_bounds': check_kw_bounds, 'fused': check_fused}[sys.argv[1]](sys.argv[2])
    print('ok')

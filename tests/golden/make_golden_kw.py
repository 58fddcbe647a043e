#!/usr/bin/env python
"""Golden KW bounds of child domains, produced by the reference's own (vendored) ``DualNetwork``.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_kw.py

For the CIFAR base and deep nets (weights and root bounds from tests/golden/nets.npz, themselves written by make_golden.py
with the same class) it fixes a few ambiguous ReLUs of the root domain, one at a time, to blocked / passing and recomputes
the bounds exactly as ``init_kw_bounds`` does for given parent bounds (plnn/dual_network_linear_approximation.py:252-288):
``DualNetwork(net, x, eps, bounded_input=False, provided_zl=..., provided_zu=...)``, pre-ReLU bounds = ``DualReLU.zl / zu``,
output bounds = ``max(dual(1), parent)`` / ``min(-dual(-1), parent)``.  Written to tests/golden/kw_children.npz.
"""
import os
import sys
import warnings

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'convex_adversarial'))
warnings.filterwarnings('ignore')

from plnn.modules import Flatten as RefFlatten      # noqa: E402
from convex_adversarial import DualNetwork          # noqa: E402  (the reference's vendored dependency)
from convex_adversarial.dual_layers import DualReLU  # noqa: E402

from golden_io import load_root                      # noqa: E402

EPS = 0.145


def modules_of(net, wp, bp):
    mods, flat = [], False
    for a in net.affine:
        if a.kind == 'conv':
            m = nn.Conv2d(a.in_shape[0], a.out_shape[0], a.weight.shape[2], stride=a.stride, padding=a.padding)
        else:
            if not flat:
                mods.append(RefFlatten()); flat = True
            m = nn.Linear(a.n_in, a.n_out)
        with torch.no_grad():
            m.weight.copy_(a.weight); m.bias.copy_(a.bias)
        mods += [m, nn.ReLU()]
    prop = nn.Linear(wp.numel(), 1)
    with torch.no_grad():
        prop.weight.copy_(wp.reshape(1, -1)); prop.bias.fill_(bp)
    mods.append(prop)
    for m in mods:
        for q in m.parameters():
            q.requires_grad = False
    return nn.Sequential(*mods)


def main():
    nets = dict(np.load(os.path.join(HERE, 'nets.npz')))
    out = {}
    for arch in ('base', 'deep', 'wide'):
        net, lbs, ubs, wp, bp = load_root(arch)
        x = torch.from_numpy(nets[f'{arch}_x'].copy())
        seq = modules_of(net, wp, bp)
        shapes = [a.out_shape for a in net.affine]
        g = torch.Generator().manual_seed(11)
        case = 0
        for lay in range(net.L):
            amb = ((lbs[lay + 1] < 0) & (ubs[lay + 1] > 0)).nonzero().view(-1)
            if amb.numel() == 0:
                continue
            for choice in (0, 1):
                idx = int(amb[int(torch.randint(0, amb.numel(), (1,), generator=g))])
                plb = [lbs[k + 1].clone().reshape(1, *shapes[k]) for k in range(net.L)]
                pub = [ubs[k + 1].clone().reshape(1, *shapes[k]) for k in range(net.L)]
                if choice == 0:
                    pub[lay].view(-1)[idx] = 0
                else:
                    plb[lay].view(-1)[idx] = 0
                dual = DualNetwork(seq, x, EPS, bounded_input=False, provided_zl=plb, provided_zu=pub)
                k = 0
                for layer in dual.dual_net:
                    if type(layer) is DualReLU:
                        out[f'{arch}_c{case}_lb{k + 1}'] = layer.zl.reshape(-1).numpy().copy()
                        out[f'{arch}_c{case}_ub{k + 1}'] = layer.zu.reshape(-1).numpy().copy()
                        k += 1
                lo = torch.max(dual(torch.ones(1, 1, 1)).view(-1), lbs[-1].view(-1))
                hi = torch.min(-dual(-torch.ones(1, 1, 1)).view(-1), ubs[-1].view(-1))
                out[f'{arch}_c{case}_lb{net.L + 1}'], out[f'{arch}_c{case}_ub{net.L + 1}'] = lo.numpy().copy(), hi.numpy().copy()
                out[f'{arch}_c{case}_decision'] = np.array([lay, idx, choice], dtype=np.int64)
                print(arch, 'case', case, 'decision', [lay, idx], 'choice', choice, 'out bounds', float(lo), float(hi))
                case += 1
        out[f'{arch}_ncases'] = np.int64(case)
    path = os.path.join(HERE, 'kw_children.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()

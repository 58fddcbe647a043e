#!/usr/bin/env python
"""Golden fixture of the online fine-tuning step, produced by the UNMODIFIED reference class.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_online.py

``graphnet.graph_score_online.GraphChoice`` (graph_score_online.py:8-90) is imported from /root/reference with the
three shims of make_golden.py (sys.path, ``.cuda()`` no-op, ``map_location``).  For the committed CIFAR-base
frontier case it makes a decision on subdomain 0, fine-tunes against a KW decision (``online_learning``), makes a
decision on subdomain 1 with the updated parameters and fine-tunes again.  Written to tests/golden/online_base.npz:
the decisions, the loss terms, the parameter gradients of the first step (``p.grad`` after ``loss.backward()``) for the
random and the shipped GNN, and the parameters after two Adam steps (random GNN).
"""
import contextlib
import io
import os
import sys
import tempfile
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
warnings.filterwarnings('ignore')

torch.Tensor.cuda = lambda self, *a, **k: self
nn.Module.cuda = lambda self, *a, **k: self
_orig_load = torch.load
torch.load = lambda f, *a, **k: _orig_load(f, *a, **{**k, 'map_location': 'cpu'})

from graphnet.graph_score_online import GraphChoice   # noqa: E402  (the reference)
from plnn.modules import Flatten as RefFlatten        # noqa: E402

from golden_io import load_case, load_gnn             # noqa: E402

LR, WD = 1e-3, 1e-4      # lr larger than the reference default (1e-4) so that two steps move the decision scores visibly


def ref_fixed_layers(net):
    """nn modules of the verified net in the reference's layer list form (conv, relu, ..., Flatten, linear, relu)."""
    mods, flat = [], False
    for a in net.affine:
        if a.kind == 'conv':
            m = nn.Conv2d(a.in_shape[0], a.out_shape[0], a.weight.shape[2], stride=a.stride, padding=a.padding)
        else:
            if not flat:
                mods.append(RefFlatten())
                flat = True
            m = nn.Linear(a.n_in, a.n_out)
        with torch.no_grad():
            m.weight.copy_(a.weight)
            m.bias.copy_(a.bias)
        for q in m.parameters():
            q.requires_grad = False
        mods += [m, nn.ReLU()]
    return mods


def bab_mask(fr, b):
    out, off = [], 0
    for n in fr.net.hidden_sizes:
        m = fr.mask[b, off:off + n]
        out.append(torch.where(m != 0, torch.full_like(m, -1), torch.ones_like(m)).int())
        off += n
    return out


def pick_kw(fr, b, gnn_dec):
    """A deterministic stand-in for the KW heuristic's decision: the last candidate that is not the GNN's."""
    sizes = fr.net.hidden_sizes
    cand = fr.mask[b].nonzero().view(-1).tolist()
    gnn_flat = sum(sizes[:gnn_dec[0]]) + gnn_dec[1]
    flat = [c for c in cand if c != gnn_flat][-1]
    lay = 0
    while flat >= sizes[lay]:
        flat -= sizes[lay]
        lay += 1
    return [lay, flat]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    fr, _ = load_case('base', 'fr')
    fixed = ref_fixed_layers(fr.net)
    out = {'lr': np.float32(LR), 'wd': np.float32(WD)}
    tmp = tempfile.mkdtemp()
    for wname in ('random', 'shipped'):
        sd = load_gnn(wname)
        path = os.path.join(tmp, wname + '.pt')
        torch.save(sd, path)
        gc = GraphChoice(bab_mask(fr, 0), path, lr=LR, wd=WD)
        for step, b in enumerate((0, 1)):
            one = fr.slice(b, b + 1)
            lbs, ubs, duals, primals, pin, layers, _ = one.to_reference_args()
            layers['fixed_layers'] = fixed
            with contextlib.redirect_stdout(io.StringIO()):
                dec = gc.decision(lbs, ubs, duals, pin, [q.tolist() for q in primals], layers, bab_mask(fr, b))
                kw = pick_kw(fr, b, dec)
                gnn_score = float(gc.gnn_score)
                # kw_score exactly as online_learning reads it (graph_score_online.py:63-69)
                partial = (0 if kw[0] == 0 else int(gc.trans_len[kw[0] - 1])) + kw[1]
                kw_score = float(gc.scores[0][len(gc.mask_1d[0][:partial].nonzero())])
                gc.online_learning(kw, 1)
            out[f'{wname}_dec{step}'] = np.array(dec, dtype=np.int64)
            out[f'{wname}_kw{step}'] = np.array(kw, dtype=np.int64)
            out[f'{wname}_gnn_score{step}'] = np.float32(gnn_score)
            out[f'{wname}_kw_score{step}'] = np.float32(kw_score)
            print(wname, 'step', step, 'gnn', dec, gnn_score, 'kw', kw, kw_score)
            if step == 0:
                for k, q in gc.model.named_parameters():
                    out[f'{wname}_grad0_{k}'] = q.grad.detach().numpy().copy()
            gc.del_score()
            if wname == 'shipped':
                break
        if wname == 'random':
            for k, v in gc.model.state_dict().items():
                out[f'{wname}_sd2_{k}'] = v.detach().numpy().copy()
    path = os.path.join(HERE, 'online_base.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Golden vectors of the BaBSR / KW branching heuristic from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_babsr.py

``choose_node_conv`` (plnn/kw_score_conv.py:41-156) is run one subdomain at a time, as the reference does
(plnn/relu_conv_gnnkwthreshold.py:155-157), on the frontier cases already committed in case_<arch>.npz, for three
scenarios that reach its three branches: the score decision (defaults), the intercept-score fallback
(decision_threshold = 1e9, counter 0) and the preference-ordered fallback (decision_threshold = 1e9, counter 2).
Written: babsr_<arch>.npz with, per case name, dense scores [B, sum n_k] and per scenario decisions [B, 2] and
counters [B].  Nothing under /root/reference is copied.
"""
import contextlib
import io
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
warnings.filterwarnings('ignore')

from plnn.kw_score_conv import choose_node_conv          # noqa: E402  (the reference)
from plnn.modules import Flatten as RefFlatten           # noqa: E402
from golden_io import ARCHS, load_case                   # noqa: E402

SCENARIOS = {'score': (0.001, 0), 'intercept': (1e9, 0), 'order': (1e9, 2)}
SPARSEST = 0


def reference_layers(net, wp_row, bp_val):
    """The reference's ``net.layers``: conv / relu / flatten / linear modules + the folded property layer."""
    layers, flat = [], False
    for a in net.affine:
        if a.kind == 'conv':
            m = nn.Conv2d(a.in_shape[0], a.out_shape[0], a.weight.shape[2], stride=a.stride, padding=a.padding)
        else:
            if not flat:
                layers.append(RefFlatten())
                flat = True
            m = nn.Linear(a.n_in, a.n_out)
        with torch.no_grad():
            m.weight.copy_(a.weight)
            m.bias.copy_(a.bias)
        layers += [m, nn.ReLU()]
    prop = nn.Linear(net.hidden_sizes[-1], 1)
    with torch.no_grad():
        prop.weight.copy_(wp_row.reshape(1, -1))
        prop.bias.fill_(float(bp_val))
    layers.append(prop)
    return layers


def run_case(fr):
    net = fr.net
    out = {'scores': [], **{f'{s}_{w}': [] for s in SCENARIOS for w in ('decisions', 'counters')}}
    random_order = list(range(net.L))
    random_order.remove(SPARSEST)
    random_order = [SPARSEST] + random_order                      # relu_conv_gnnkwthreshold.py:98-101
    for b in range(fr.B):
        layers = reference_layers(net, fr.Wp[b], fr.bp[b])
        # bounds list indexed like the reference's: one entry per layer output, shaped; pre-ReLU entries hold the data
        lbs, ubs, pre_relu = [fr.lb[0][b].reshape(net.input_shape)], [fr.ub[0][b].reshape(net.input_shape)], []
        k, shape = 0, net.input_shape
        for m in layers:
            if isinstance(m, (nn.Conv2d, nn.Linear)) and k < net.L:
                shape = net.affine[k].out_shape
                lbs.append(fr.lb[k + 1][b].reshape(shape)); ubs.append(fr.ub[k + 1][b].reshape(shape))
                pre_relu.append(len(lbs) - 1)
                k += 1
            elif isinstance(m, RefFlatten):
                n = int(np.prod(shape))
                lbs.append(torch.zeros(n)); ubs.append(torch.zeros(n))
                shape = (n,)
            else:
                lbs.append(torch.zeros(shape)); ubs.append(torch.zeros(shape))
        orig_mask, off = [], 0
        for n in net.hidden_sizes:
            m = fr.mask[b, off:off + n]
            orig_mask.append(torch.where(m != 0, torch.full_like(m, -1), torch.ones_like(m)).int())
            off += n
        for name, (thr, counter) in SCENARIOS.items():
            with contextlib.redirect_stdout(io.StringIO()):
                dec, cnt, score = choose_node_conv(lbs, ubs, orig_mask, layers, pre_relu, counter, random_order, SPARSEST,
                                                   decision_threshold=thr, gt=True)
            out[f'{name}_decisions'].append(dec)
            out[f'{name}_counters'].append(cnt)
            if name == 'score':
                out["scores"].append(torch.cat([s.reshape(-1) for s in score]).detach().numpy())
    return {k: np.asarray(v) for k, v in out.items()}


def main():
    for arch in ARCHS:
        blob = {}
        for case in ('fr', 'root'):
            fr, _ = load_case(arch, case)
            for k, v in run_case(fr).items():
                blob[f'{case}_{k}'] = v
        path = os.path.join(HERE, f'babsr_{arch}.npz')
        np.savez_compressed(path, **blob)
        print(arch, {k: v.shape for k, v in blob.items()}, os.path.getsize(path))


if __name__ == '__main__':
    main()

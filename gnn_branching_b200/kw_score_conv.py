"""``choose_node_conv`` with the reference's signature (plnn/kw_score_conv.py:41-156), computed by libgnnb on the GPU,
plus the batched entry ``babsr_frontier``.  SURVEY §8f rank 1: the hand-written BaBSR / KW score the reference
falls back to when the GNN's decision did not improve the bound (plnn/relu_conv_gnnkwthreshold.py:155-157)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .engine import Scorer
from .frontier import Frontier
from .networks import netspec_from_modules

_scorers = {}


def _scorer(device_index: int) -> Scorer:
    if device_index not in _scorers:
        _scorers[device_index] = Scorer(device_index)
    return _scorers[device_index]


def babsr_frontier(fr: Frontier, icp_score_counter=None, random_order=None, sparsest_layer: int = 0,
                   decision_threshold: float = 0.001, return_scores: bool = False, scorer: Optional[Scorer] = None):
    """Decisions for every subdomain of a frontier: (decision [B, 2], counters [B], kind [B], scores or None)."""
    if scorer is None:
        scorer = _scorer(fr.device.index if fr.device.type == 'cuda' else torch.cuda.current_device())
    scorer.set_network(fr.net, key=fr.net.key)
    return scorer.babsr(fr, icp_score_counter, random_order, sparsest_layer, decision_threshold, return_scores)


def choose_node_conv(lower_bounds, upper_bounds, orig_mask, layers, pre_relu_indices, icp_score_counter, random_order,
                     sparsest_layer, decision_threshold=0.001, gt=False):
    """Drop-in for plnn.kw_score_conv.choose_node_conv: one subdomain, the reference's argument lists.

    ``layers`` is the verified network's module list ending with the (property-folded) last ``nn.Linear``;
    ``lower_bounds[pre_relu_indices[k]]`` are the pre-ReLU bounds of ReLU layer k; ``orig_mask`` holds -1 for
    undecided ReLUs.  Returns ``(decision, icp_score_counter)`` (and the per-layer scores when ``gt``)."""
    fixed, prop = list(layers[:-1]), layers[-1]
    if not isinstance(prop, nn.Linear) or prop.out_features != 1:
        raise NotImplementedError('the last layer must be the folded property layer Linear(n_L, 1)')
    dev = prop.weight.device if prop.weight.is_cuda else torch.device('cuda', torch.cuda.current_device())
    net = netspec_from_modules(fixed, tuple(torch.as_tensor(lower_bounds[0]).shape), name='kw')
    L = net.L
    f = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(1, -1).to(dev)
    zeros = lambda n: torch.zeros(1, n, device=dev)
    lb = [f(lower_bounds[0])] + [f(lower_bounds[i]) for i in pre_relu_indices] + [zeros(1)]
    ub = [f(upper_bounds[0])] + [f(upper_bounds[i]) for i in pre_relu_indices] + [zeros(1)]
    mask = torch.cat([(torch.as_tensor(m) == -1).float().reshape(-1) for m in orig_mask]).reshape(1, -1).to(dev)
    hs = net.hidden_sizes
    fr = Frontier(net=net.to(dev), lb=lb, ub=ub, dual=[zeros(n * 3).reshape(1, n, 3) for n in hs],
                  prim_pre=[zeros(n) for n in hs], prim_post=[zeros(n) for n in hs], prim_out=zeros(1).reshape(1),
                  primal_input=zeros(net.n0), Wp=prop.weight.detach().float().reshape(1, -1).to(dev),
                  bp=prop.bias.detach().float().reshape(1).to(dev), mask=mask)
    dec, cnt, _, scores = babsr_frontier(fr, [int(icp_score_counter)], list(random_order), int(sparsest_layer),
                                         float(decision_threshold), return_scores=bool(gt))
    decision = [int(dec[0, 0]), int(dec[0, 1])]
    if decision[0] < 0:
        raise RuntimeError('no undecided ReLU (mask has no -1 entry)')
    if not gt:
        return decision, int(cnt[0])
    per_layer, off = [], 0
    for n in hs:
        per_layer.append(scores[0, off:off + n])
        off += n
    return decision, int(cnt[0]), per_layer

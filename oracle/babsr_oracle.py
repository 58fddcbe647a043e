"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the BaBSR / KW branching heuristic.

Only ``tests/`` and ``bench.py``'s cpu_baseline leg may import this module, as the checker / the timed CPU baseline.  What it restates: ``choose_node_conv`` of
oval-group/GNN_branching (plnn/kw_score_conv.py:41-156; ``compute_ratio`` :23-37), the hand-written score the
reference falls back to when the GNN decision did not improve the bound enough
(plnn/relu_conv_gnnkwthreshold.py:155-157) — SURVEY §8f rank 1.  Batched over subdomains; the reference runs
one subdomain per call.

Parity pin: ``tests/golden/make_golden_babsr.py`` runs the UNMODIFIED reference function (imported from
/root/reference in the build container) on the committed frontier cases and stores its scores, decisions and
intercept counters in ``tests/golden/babsr_<arch>.npz``; ``tests/test_oracle.py`` holds this oracle to them.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F


def compute_ratio(l: torch.Tensor, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """kw_score_conv.py:23-37 -> (slope_ratio, intercept), same operation order."""
    lt = l - F.relu(l)
    ut = F.relu(u)
    slope = ut / (ut - lt)
    intercept = -1 * lt * slope
    return slope, intercept


def babsr_scores(fr) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (score [B, sum n_k], intercept_tb [B, sum n_k]) in flat hidden order (kw_score_conv.py:72-121).

    ``fr`` is a ``gnn_branching_b200.Frontier``: ``fr.net.affine`` are the conv / linear layers that precede a
    ReLU, ``fr.Wp`` the per-subdomain property layer (the last ``nn.Linear`` of the reference's ``layers``)."""
    net, B = fr.net, fr.B
    L = net.L
    ratio = fr.Wp.float().clone()                                         # :81-85  W^T @ ones(1) of the property layer
    scores: List[torch.Tensor] = [None] * L
    icps: List[torch.Tensor] = [None] * L
    off = [0]
    for a in net.affine:
        off.append(off[-1] + a.n_out)
    for k in range(L - 1, -1, -1):
        a = net.affine[k]
        l, u = fr.lb[k + 1].float(), fr.ub[k + 1].float()
        m = fr.mask[:, off[k]:off[k + 1]].float()
        r0, icp = compute_ratio(l, u)                                     # :90
        intercept_candidate = torch.clamp(ratio, max=0) * icp             # :92-93
        icps[k] = intercept_candidate * m                                 # :94
        if a.kind == 'conv':                                              # :97-99  bias per channel, broadcast over H, W
            b = a.bias.float().reshape(1, -1, 1).expand(1, a.out_shape[0], a.out_shape[1] * a.out_shape[2]).reshape(1, -1)
        else:
            b = a.bias.float().reshape(1, -1)
        ratio_1 = ratio * (r0 - 1)                                        # :100
        bias_candidate_1 = b * ratio_1
        ratio = ratio * r0                                                # :102
        bias_candidate_2 = b * ratio
        bias_candidate = torch.max(bias_candidate_1, bias_candidate_2)    # :104
        scores[k] = (bias_candidate + intercept_candidate).abs() * m      # :109-110
        # through A_k^T to the previous layer's nodes (:81-86 linear, :115-119 conv); not needed after the first layer
        if k > 0:
            if a.kind == 'conv':
                ratio = F.conv_transpose2d(ratio.reshape(B, *a.out_shape), a.weight.float(), stride=a.stride,
                                           padding=a.padding).reshape(B, -1)
            else:
                ratio = ratio @ a.weight.float()
    return torch.cat(scores, 1), torch.cat(icps, 1)


def babsr_decide(score: torch.Tensor, intercept: torch.Tensor, mask: torch.Tensor, hidden_sizes: Sequence[int],
                 icp_score_counter: Sequence[int], random_order: Sequence[int], sparsest_layer: int,
                 decision_threshold: float = 0.001):
    """kw_score_conv.py:123-152 per subdomain -> (decisions [B, 2], counters [B], kinds [B]);
    kind 0 = score, 1 = intercept score, 2 = preference-ordered choice."""
    B = score.shape[0]
    offs = [0]
    for n in hidden_sizes:
        offs.append(offs[-1] + n)
    decisions, counters, kinds = [], [], []
    for b in range(B):
        counter = int(icp_score_counter[b])
        sc = [score[b, offs[k]:offs[k + 1]] for k in range(len(hidden_sizes))]
        ic = [intercept[b, offs[k]:offs[k + 1]] for k in range(len(hidden_sizes))]
        mk = [mask[b, offs[k]:offs[k + 1]] for k in range(len(hidden_sizes))]
        max_info = [torch.max(s, 0) for s in sc]
        # `max(max_info)` compares (value, index) tuples and `.index` returns the first equal entry: on equal maxima the
        # layer whose in-layer argmax index is larger wins, the first layer among fully equal pairs
        vals = [float(v) for v, _ in max_info]
        pairs = [(float(v), int(i)) for v, i in max_info]
        decision_layer = pairs.index(max(pairs))
        decision_index = int(max_info[decision_layer][1])
        if decision_layer != sparsest_layer and vals[decision_layer] > decision_threshold:
            decision, kind = [decision_layer, decision_index], 0
        else:
            min_info = [[i, torch.min(ic[i], 0)] for i in range(len(ic)) if float(torch.min(ic[i])) < -1e-4]
            if len(min_info) != 0 and counter < 2:
                intercept_layer = min_info[-1][0]
                intercept_index = int(min_info[-1][1][1])
                counter += 1
                decision, kind = [intercept_layer, intercept_index], 1
                if intercept_layer != 0:
                    counter = 0
            else:
                choice = list(random_order)
                decision, kind = None, 2
                while decision is None:
                    preferred = choice.pop(-1)
                    nz = mk[preferred].nonzero()
                    if len(nz) != 0:
                        decision = [preferred, int(nz[0])]
                counter = 0
        decisions.append(decision)
        counters.append(counter)
        kinds.append(kind)
    return torch.tensor(decisions), torch.tensor(counters), torch.tensor(kinds)

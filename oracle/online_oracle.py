"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the online fine-tuning step of the reference.

Only ``tests/`` (and the golden generator) may import this module; the product path never does.

What it restates: ``GraphChoice.online_learning`` of oval-group/GNN_branching
(graphnet/graph_score_online.py:62-77): ``loss = gnn_score - kw_score + improvement``, ``loss.backward()`` through
``GraphNet.forward`` and one ``torch.optim.Adam(lr, weight_decay=wd)`` step (:15).  The derivative is PyTorch
autograd over the forward restatement in ``graphnet_oracle.gnn_forward``; the optimiser is ``torch.optim.Adam``
itself (third-party arithmetic of the reference, torch 2.11.0 in this image).

Parity pin: ``tests/golden/make_golden_online.py`` runs the UNMODIFIED reference class from /root/reference and
commits its gradients / updated parameters / decisions as ``tests/golden/online_base.npz``;
``tests/test_oracle.py`` holds this module to them.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import graphnet_oracle as O


def kw_flat_index(mask_row: torch.Tensor, hidden_sizes: Sequence[int], kw_decision) -> int:
    """graph_score_online.py:63-69: the KW decision [layer, idx] -> position in the ragged score vector
    (= number of candidates before it) -> flat index of that candidate."""
    partial_len = (0 if kw_decision[0] == 0 else int(sum(hidden_sizes[:kw_decision[0]]))) + int(kw_decision[1])
    kw_index = int(mask_row[:partial_len].nonzero().numel())
    cand = mask_row.nonzero().view(-1)
    return int(cand[kw_index])          # IndexError when there is no candidate at or after it, as in the reference


def score_grads(state_dict: Dict[str, torch.Tensor], fr, terms: List[Tuple[int, int, float]], T: int = 2
                ) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """d(sum_i coeff_i * score[domain_i][flat_i]) / d(parameters) by autograd.  Returns (grads, term scores)."""
    params = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in state_dict.items()}
    scores, _ = O.gnn_forward(params, fr, T=T, keep_graph=True)
    vals = torch.stack([scores[b, i] for b, i, _ in terms])
    loss = sum(c * scores[b, i] for b, i, c in terms)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach() for k, v in params.items()}
    return grads, vals.detach()


class OnlineOracle:
    """Stateful twin of the reference's online ``GraphChoice``: parameters + Adam state."""

    def __init__(self, state_dict, lr=1e-4, wd=1e-4, T=2):
        self.T = T
        self.params = {k: torch.nn.Parameter(v.detach().float().cpu().clone()) for k, v in state_dict.items()}
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr, weight_decay=wd)

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.params.items()}

    def decision(self, fr):
        with torch.no_grad():
            scores, _ = O.gnn_forward(self.state_dict(), fr, T=self.T)
        _, flat, dec = O.decide(scores, fr.mask, fr.net.hidden_sizes)
        return int(flat[0]), dec[0]

    def online_learning(self, fr, gnn_flat: int, kw_decision, improvement: float):
        kw_flat = kw_flat_index(fr.mask[0], fr.net.hidden_sizes, kw_decision)
        self.opt.zero_grad()
        scores, _ = O.gnn_forward(self.params, fr, T=self.T, keep_graph=True)
        loss = scores[0, gnn_flat] - scores[0, kw_flat] + improvement
        loss.backward()
        grads = {k: v.grad.detach().clone() for k, v in self.params.items()}
        self.opt.step()
        return grads, float(loss.detach())

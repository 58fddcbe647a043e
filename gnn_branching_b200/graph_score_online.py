"""``GraphChoice`` of graphnet/graph_score_online.py:8-90: the decision path of graph_score.py plus the online
fine-tuning step (``loss = gnn_score - kw_score + improvement``; ``loss.backward()``; Adam, :62-77).

The reference differentiates through ``GraphNet.forward`` with PyTorch autograd; here the gradient and the Adam step
run in the CUDA library (``gnnb_score_grad`` / ``gnnb_adam_step``, csrc/gnnb_train.cu).  There is no PyTorch fallback."""
from __future__ import annotations

import time

import torch

from .engine import flat_to_layer_index
from .frontier import Frontier
from .graph_score import GraphChoice as _GraphChoice


class GraphChoice(_GraphChoice):
    def __init__(self, init_mask, model_name, lr=1e-4, wd=1e-4, linear=False, math=None):
        super().__init__(init_mask, model_name, linear=linear, math=math)
        self.lr, self.wd = lr, wd                                    # torch.optim.Adam(lr=lr, weight_decay=wd), :15
        self._fr = None

    def decision(self, lower_bounds_all, upper_bounds_all, dual_vars, primal_input, primals, layers, mask):
        """As graph_score.GraphChoice.decision; keeps the subdomain, its candidate mask and the winning score for
        ``online_learning`` (the reference keeps ``self.scores``, ``self.gnn_score``, ``self.mask_1d``, :24-41)."""
        mask = [(i == -1).float() for i in mask]
        self.mask_1d = torch.cat([i for i in mask], 0).unsqueeze(0)
        start = time.time()
        dev = next(self.model.parameters()).device
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).to(dev)
        with torch.no_grad():
            fr = Frontier.from_reference_args([f(i) for i in lower_bounds_all], [f(i) for i in upper_bounds_all],
                                              [f(i) for i in dual_vars], [f(i) for i in primals], f(primal_input), layers,
                                              self.mask_1d.to(dev))
            best, idx, _ = self.model.score_frontier(fr, return_scores=False)
        flat = int(idx[0].item())
        if self.verbose:
            print(f'graph requires: {time.time() - start}')
        if flat < 0:
            raise RuntimeError('max(): no undecided ReLU (mask has no -1 entry)')
        self._fr, self._gnn_flat, self.gnn_score = fr, flat, float(best[0].item())
        return flat_to_layer_index(flat, self.hidden_sizes)

    def online_learning(self, kw_decision, improvement):
        """graph_score_online.py:62-77.  Returns the loss value (the reference returns None)."""
        if self._fr is None:
            raise RuntimeError('online_learning needs a preceding decision() (the reference reads self.scores)')
        partial_len = (0 if kw_decision[0] == 0 else int(self.trans_len[kw_decision[0] - 1])) + int(kw_decision[1])
        cand = self.mask_1d[0].nonzero().view(-1)
        kw_index = int(self.mask_1d[0][:partial_len].nonzero().numel())       # position in the ragged score list, :68
        kw_flat = int(cand[kw_index])
        sc = self.model.scorer(self._fr.device.index)
        sc.set_network(self._fr.net, key=self._fr.net.key)
        vals = sc.score_grad(self._fr, [(0, self._gnn_flat, 1.0), (0, kw_flat, -1.0)])
        sc.adam_step(self.lr, weight_decay=self.wd)
        self.model.adopt_weights(sc)
        return float(vals[0] - vals[1]) + float(improvement)

    def del_score(self):
        self._fr = None
        self.mask_1d = None
        self.gnn_score = None

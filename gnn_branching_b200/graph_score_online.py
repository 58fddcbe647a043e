"""``GraphChoice`` of graphnet/graph_score_online.py:8-90: same decision path; the online fine-tuning step
(Adam on ``gnn_score - kw_score + improvement``, :62-77) needs autograd through the fused forward, which is a
"next" row of the scope table and is not built."""
from __future__ import annotations

from .graph_score import GraphChoice as _GraphChoice


class GraphChoice(_GraphChoice):
    def __init__(self, init_mask, model_name, lr=1e-4, wd=1e-4, linear=False, math=None):
        super().__init__(init_mask, model_name, linear=linear, math=math)
        self.lr, self.wd = lr, wd

    def online_learning(self, kw_decision, improvement):
        raise NotImplementedError('online fine-tuning (graph_score_online.py:62-77) is out of scope of the scoring path')

    def del_score(self):
        pass

"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's domain list.

Only ``tests/`` (and the golden generator) may import this module; the product path never does.

What it restates: ``add_domain`` / ``pick_out`` / ``prune_domains`` of oval-group/GNN_branching
(plnn/branch_and_bound.py:159-184, 264-281) on the sorted list ``domains`` — plain Python, ``bisect`` like the
reference.  Domains are identified by an integer id; the payload does not matter for the order.

Parity pin: ``tests/golden/make_golden_queue.py`` replays seeded operation traces through the reference's own three
functions (imported from /root/reference) and commits the traces with the reference's results
(``tests/golden/queue_traces.npz``); ``tests/test_oracle.py`` replays them through this module.
"""
from __future__ import annotations

import bisect
from typing import List, Tuple


class _Dom:
    __slots__ = ('lower_bound', 'ident')

    def __init__(self, lb, ident):
        self.lower_bound, self.ident = lb, ident

    def __lt__(self, other):                     # ReLUDomain.__lt__ (relu_conv_gnnkwthreshold.py:43-44)
        return self.lower_bound < other.lower_bound


class QueueOracle:
    def __init__(self):
        self.domains: List[_Dom] = []

    def add(self, lb: float, ident: int) -> None:
        bisect.insort_left(self.domains, _Dom(lb, ident))           # branch_and_bound.py:164

    def pick(self, threshold: float):
        """-> ident of the picked domain, or None when no domain is below the threshold (the reference asserts; every
        domain has been popped by then, branch_and_bound.py:177-182)."""
        while self.domains:
            d = self.domains.pop(0)
            if d.lower_bound < threshold:
                return d.ident
        return None

    def prune(self, threshold: float) -> None:
        for i, d in enumerate(self.domains):                        # branch_and_bound.py:276-279
            if d.lower_bound >= threshold:
                self.domains = self.domains[:i]
                break

    def state(self) -> List[Tuple[float, int]]:
        return [(d.lower_bound, d.ident) for d in self.domains]


def replay(ops) -> Tuple[List[int], List[int]]:
    """ops: rows (kind, value, ident): kind 0 = add(lb=value, ident), 1 = pick(threshold=value), 2 = prune(threshold=value).
    Returns (ident picked by every pick op, -1 for none; idents left in the queue, in order)."""
    q, picked = QueueOracle(), []
    for kind, value, ident in ops:
        kind = int(kind)
        if kind == 0:
            q.add(float(value), int(ident))
        elif kind == 1:
            r = q.pick(float(value))
            picked.append(-1 if r is None else r)
        else:
            q.prune(float(value))
    return picked, [i for _, i in q.state()]

#!/usr/bin/env python
"""Golden traces of the domain list, produced by the UNMODIFIED reference functions.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_queue.py

``add_domain`` / ``pick_out`` / ``prune_domains`` are imported from /root/reference/plnn/branch_and_bound.py and driven
with seeded operation traces on ``ReLUDomain``-like objects (lower bound + id): random adds (with many equal lower bounds,
signed zeros, negative and positive values), picks with thresholds inside and outside the range of the bounds, prunes.
Written to tests/golden/queue_traces.npz: per trace the operations ``[n, 3]`` (kind, value, id) as float64, the id every
pick returned (-1 where the reference's assert fired: nothing below the threshold, everything popped) and the ids left.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from plnn.branch_and_bound import add_domain, pick_out, prune_domains    # noqa: E402  (the reference)


class Dom:          # the ordering protocol of ReLUDomain (relu_conv_gnnkwthreshold.py:43-50)
    def __init__(self, lb, ident):
        self.lower_bound, self.ident = lb, ident

    def __lt__(self, other):
        return self.lower_bound < other.lower_bound

    def __le__(self, other):
        return self.lower_bound <= other.lower_bound

    def __eq__(self, other):
        return self.lower_bound == other.lower_bound


def make_trace(seed, n_ops, grid, thr_mu=-0.5):
    rng = np.random.default_rng(seed)
    ops, nid = [], 0
    for _ in range(n_ops):
        r = rng.random()
        if r < 0.6:
            lb = float(np.float32(rng.normal(-1.0, 1.0)))
            if grid:                                   # many ties
                lb = float(np.float32(np.round(lb * 4) / 4))
                if lb == 0.0 and rng.random() < 0.5:
                    lb = -0.0
            ops.append((0, lb, nid)); nid += 1
        elif r < 0.9:
            ops.append((1, float(np.float32(rng.normal(thr_mu, 1.0))), -1))
        else:
            ops.append((2, float(np.float32(rng.normal(thr_mu + 0.5, 1.0))), -1))
    return np.array(ops, dtype=np.float64)


def run_reference(ops):
    domains, picked = [], []
    for kind, value, ident in ops:
        kind = int(kind)
        if kind == 0:
            add_domain(Dom(float(value), int(ident)), domains)
        elif kind == 1:
            if len(domains) == 0:
                picked.append(-1)
                continue
            try:
                picked.append(pick_out(domains, float(value)).ident)
            except AssertionError:                     # 'No domain left to pick from.': the list is empty now
                picked.append(-1)
        else:
            domains = prune_domains(domains, float(value))
    return np.array(picked, dtype=np.int64), np.array([d.ident for d in domains], dtype=np.int64)


def main():
    out = {}
    for t, (seed, n_ops, grid, mu) in enumerate([(1, 400, False, -0.5), (2, 400, True, -0.5), (3, 1500, True, -0.5), (4, 60, True, -0.5),
                                                 (5, 2000, True, 2.5), (6, 2000, False, 1.5)]):
        ops = make_trace(seed, n_ops, grid, mu)
        picked, left = run_reference(ops)
        out[f't{t}_ops'], out[f't{t}_picked'], out[f't{t}_left'] = ops, picked, left
        print(f'trace {t}: {n_ops} ops, {len(picked)} picks ({int((picked < 0).sum())} empty), {len(left)} left')
    path = os.path.join(HERE, 'queue_traces.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()

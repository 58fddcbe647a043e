"""Checks of GPU code paths that have not run on a GPU yet, executed in their own process by tests/test_gpu_parity.py
(`_run_isolated`): exit code 0 = pass.  Usage: python tests/gpu_isolated.py <check> <arg>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                                                     # noqa: E402

from golden_io import load_case, load_gnn, load_root             # noqa: E402
from gnn_branching_b200 import GraphNet, synthetic_frontier      # noqa: E402


def check_gather_prefetch(arch):
    """The propagation kernel variant k_tc_prop_pf must give bit-identical scores, winners and indices."""
    fr, _ = load_case(arch, 'fr')
    model = GraphNet(2, 64, math='tc')
    model.load_state_dict(load_gnn('random'))
    model = model.eval().cuda()
    for f in (fr.to('cuda'), synthetic_frontier(*load_root(arch), 37, seed=5, device='cuda')):
        model.scorer(0).set_option('gather_prefetch', 0)
        b0, i0, s0 = model.score_frontier(f)
        model.scorer(0).set_option('gather_prefetch', 1)
        b1, i1, s1 = model.score_frontier(f)
        torch.cuda.synchronize()
        assert torch.equal(s0, s1) and torch.equal(i0, i1) and torch.equal(b0, b1), 'gather_prefetch changes the results'


if __name__ == '__main__':
    {'gather_prefetch': check_gather_prefetch}[sys.argv[1]](sys.argv[2])
    print('ok')

#!/usr/bin/env python
"""Golden bounds of child domains, produced by the reference's own ``KWConvGen.update_the_model`` (bounds part).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_child.py

``update_the_model`` (plnn/conv_kwinter_gen.py:558-660) computes a child's intermediate bounds in three steps — KW bounds
from the parent's bounds with one ReLU fixed (``update_kw_bounds``), interval bounds of the layers behind the split
intersected with them, and a second KW pass when the interval bounds tightened anything — and then hands them to Gurobi.
Gurobi is not installed, so the UNMODIFIED method is executed up to its first access to the LP model: a stub ``gurobipy``
lets the modules import, the method raises at ``grb.Model()`` / ``self.gurobi_vars``, and the bounds it had computed by then
are read from the frame of the raised exception.  Nothing of the reference is copied or edited.

For base, wide and deep (weights, x, property layer from tests/golden/nets.npz): the root bounds of ``build_the_model``
(KW intersected with interval bounds), then chains of splits root -> child -> grandchild ... on ambiguous ReLUs of seeded
layers / sides, each child computed from the bounds of the case before it.  Written to tests/golden/child_bounds.npz.
"""
import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'convex_adversarial'))
warnings.filterwarnings('ignore')


class _NoGurobi(Exception):
    pass


def _raise(*a, **k):
    raise _NoGurobi()


grb = types.ModuleType('gurobipy')
grb.Model = _raise
grb.GRB = types.SimpleNamespace(MINIMIZE=1, BINARY='B', CONTINUOUS='C')
sys.modules['gurobipy'] = grb

from plnn.conv_kwinter_gen import KWConvGen      # noqa: E402  (the unmodified reference class)
from make_golden_kw import modules_of, EPS       # noqa: E402
from golden_io import load_root                  # noqa: E402


def frame_locals(exc, func_name):
    tb = exc.__traceback__
    found = None
    while tb is not None:
        if tb.tb_frame.f_code.co_name == func_name:
            found = tb.tb_frame.f_locals
        tb = tb.tb_next
    assert found is not None, func_name
    return found


def run_until_gurobi(fn, name, *args):
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            fn(*args)
    except (_NoGurobi, AttributeError) as e:
        loc = frame_locals(e, name)
        return loc, buf.getvalue()
    raise AssertionError('the reference did not reach its LP model')


def main():
    nets = dict(np.load(os.path.join(HERE, 'nets.npz')))
    out = {}
    for arch in ('base', 'deep', 'wide'):
        net, _, _, wp, bp = load_root(arch)
        x = torch.from_numpy(nets[f'{arch}_x'].copy())
        seq = modules_of(net, wp, bp)
        gen = KWConvGen(list(seq))
        dom = torch.stack([x[0] - EPS, x[0] + EPS], dim=-1)
        loc, log = run_until_gurobi(gen.build_the_model, 'build_the_model', dom, x, EPS, False)
        root_lb, root_ub = [t.clone() for t in loc['lower_bounds']], [t.clone() for t in loc['upper_bounds']]
        pre = list(loc['pre_relu_indices'])
        gen.replacing_bd_index = len(root_lb)                    # build_the_model sets it after the LP model (:526)
        print(arch, 'root', 'pre_relu_indices', pre, 'log:', log.strip().replace('\n', ' | ') or '-')
        hid = pre + [len(root_lb) - 1]

        def put(case, lbs, ubs):
            for k, i in enumerate(hid):
                out[f'{arch}_c{case}_lb{k + 1}'] = lbs[i].reshape(-1).numpy().copy()
                out[f'{arch}_c{case}_ub{k + 1}'] = ubs[i].reshape(-1).numpy().copy()

        put(0, root_lb, root_ub)
        out[f'{arch}_c0_parent'] = np.int64(-1)
        out[f'{arch}_c0_decision'] = np.array([-1, -1, -1], dtype=np.int64)
        g = torch.Generator().manual_seed(23)
        case = 1
        bounds_of = {0: (root_lb, root_ub)}
        # chains of splits: every chain starts at the root; depth 3
        for chain in range(4 if arch != 'wide' else 2):
            parent = 0
            for depth in range(3):
                plb, pub = bounds_of[parent]
                lay = int(torch.randint(0, net.L, (1,), generator=g)) if depth else chain % net.L
                flat_l, flat_u = plb[pre[lay]].reshape(-1), pub[pre[lay]].reshape(-1)
                amb = ((flat_l < 0) & (flat_u > 0)).nonzero().view(-1)
                if amb.numel() == 0:
                    break
                idx = int(amb[int(torch.randint(0, amb.numel(), (1,), generator=g))])
                choice = int(torch.randint(0, 2, (1,), generator=g))
                mask = [torch.zeros(plb[i].numel()) for i in pre]
                loc, log = run_until_gurobi(gen.update_the_model, 'update_the_model', mask, [t.clone() for t in plb],
                                            [t.clone() for t in pub], [lay, idx], choice)
                clb, cub = [t.clone() for t in loc['lower_bounds']], [t.clone() for t in loc['upper_bounds']]
                bounds_of[case] = (clb, cub)
                put(case, clb, cub)
                out[f'{arch}_c{case}_parent'] = np.int64(parent)
                out[f'{arch}_c{case}_decision'] = np.array([lay, idx, choice], dtype=np.int64)
                out[f'{arch}_c{case}_interval_better'] = np.int64('interval is better' in log)
                print(arch, 'case', case, 'parent', parent, 'decision', [lay, idx], 'choice', choice, 'out', float(clb[-1]), float(cub[-1]),
                      'interval pass' if 'interval is better' in log else '')
                parent = case
                case += 1
        out[f'{arch}_ncases'] = np.int64(case)
    path = os.path.join(HERE, 'child_bounds.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()

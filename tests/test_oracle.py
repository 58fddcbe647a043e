"""The oracle (oracle/graphnet_oracle.py) against the reference's own outputs (tests/golden/*.npz).

The goldens were produced by the unmodified reference GraphNet / GraphChoice in the build container
(tests/golden/make_golden.py).  Tolerance: the oracle restates the same fp32 PyTorch-CPU arithmetic in
batched form, so only fp32 re-association noise is allowed: 2e-6 * max|s| (observed <= 5e-7).
"""
import os

import pytest
import torch

from golden_io import ARCHS, load_case, load_gnn
from oracle import graphnet_oracle as O

TOL = 2e-6


@pytest.mark.parametrize('arch', ARCHS)
@pytest.mark.parametrize('case', ['fr', 'root'])
@pytest.mark.parametrize('weights', ['shipped', 'random'])
def test_oracle_matches_reference_scores_and_decisions(arch, case, weights):
    fr, ref = load_case(arch, case)
    scores, _ = O.gnn_forward(load_gnn(weights), fr)
    s_ref = ref[f'scores_{weights}']
    m = fr.mask != 0
    scale = float(s_ref[m].abs().max())
    err = float((scores - s_ref)[m].abs().max())
    assert err <= TOL * scale, (err, scale)
    _, flat, dec = O.decide(scores, fr.mask, fr.net.hidden_sizes)
    assert dec == ref[f'decisions_{weights}'].tolist()
    rep = O.parity_report(scores, s_ref, fr.mask, flat)
    assert rep['ok'], rep


@pytest.mark.parametrize('T', [1, 3])
def test_oracle_other_round_counts(T):
    fr, ref = load_case('base', 'fr')
    scores, _ = O.gnn_forward(load_gnn('random'), fr, T=T)
    s_ref = ref[f'scores_random_T{T}']
    m = fr.mask != 0
    assert float((scores - s_ref)[m].abs().max()) <= TOL * float(s_ref[m].abs().max())


def test_dead_input_update_is_dead():
    """the last round's input-layer update never reaches the scores."""
    fr, _ = load_case('base', 'fr')
    sd = load_gnn('random')
    a, _ = O.gnn_forward(sd, fr)
    b, _ = O.gnn_forward(sd, fr, dead_input_update=True)
    assert torch.equal(a, b)


def test_shipped_checkpoint_is_degenerate_in_forward_half():
    """with the shipped checkpoint T=1 and T=2 agree (forward-half weights are denormals),
    which is why every parity test also runs the non-degenerate random GraphNet."""
    fr, _ = load_case('base', 'fr')
    sd = load_gnn('shipped')
    a, _ = O.gnn_forward(sd, fr, T=1)
    b, _ = O.gnn_forward(sd, fr, T=2)
    assert torch.equal(a, b)
    r1, _ = O.gnn_forward(load_gnn('random'), fr, T=1)
    r2, _ = O.gnn_forward(load_gnn('random'), fr, T=2)
    assert not torch.allclose(r1, r2)


def test_compute_ratio_cases():
    l = torch.tensor([-1.0, 0.5, -2.0, 0.0, -3.0])
    u = torch.tensor([1.0, 2.0, -1.0, 4.0, 0.0])
    r0, r1, beta, amb = O.compute_ratio(l, u)
    assert r0.tolist() == [0.5, 1.0, 0.0, 1.0, 0.0]
    assert amb.tolist() == [1.0, 0.0, 0.0, 0.0, 0.0]
    assert r1.tolist() == [0.5, 1.0, 0.0, 1.0, 0.0]
    assert beta.tolist() == [0.5, 0.0, 0.0, 0.0, 0.0]
    r0, _, _, _ = O.compute_ratio(torch.zeros(1), torch.zeros(1))
    assert torch.isnan(r0).all()          # l = u = 0 is 0/0 in the reference too


# ---- BaBSR / KW heuristic -------------------------------------------------------------------
BABSR_SCENARIOS = {'score': (0.001, 0), 'intercept': (1e9, 0), 'order': (1e9, 2)}


def _babsr_golden(arch):
    import numpy as np
    import os
    from golden_io import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, f'babsr_{arch}.npz')))


@pytest.mark.parametrize('case', ['fr', 'root'])
@pytest.mark.parametrize('arch', ARCHS)
def test_babsr_oracle_matches_reference(arch, case):
    """oracle/babsr_oracle.py against the reference's choose_node_conv outputs (tests/golden/make_golden_babsr.py):
    scores to 1e-6, decisions and intercept counters exactly, in the three branches of the decision rule."""
    from oracle import babsr_oracle as BO
    z = _babsr_golden(arch)
    fr, _ = load_case(arch, case)
    sc, ic = BO.babsr_scores(fr)
    ref = torch.from_numpy(z[f'{case}_scores'])
    assert float((sc - ref).abs().max()) <= 1e-6 * max(1.0, float(ref.abs().max()))
    order = [0] + [k for k in range(fr.net.L) if k != 0]
    for name, (thr, cnt) in BABSR_SCENARIOS.items():
        d, c, _ = BO.babsr_decide(sc, ic, fr.mask, fr.net.hidden_sizes, [cnt] * fr.B, order, 0, thr)
        assert d.tolist() == z[f'{case}_{name}_decisions'].tolist()
        assert c.tolist() == z[f'{case}_{name}_counters'].tolist()


# ---- online fine-tuning -----------------------------------------------------------------------
def test_online_oracle_matches_reference():
    """oracle/online_oracle.py against the reference's own online_learning (tests/golden/make_golden_online.py): same
    decisions, same loss terms, gradients of the first step to 2e-5 of each tensor's largest entry (autograd over a
    batched restatement re-associates sums), parameters after two Adam steps to 2 % of one step (lr)."""
    from golden_io import load_online
    from oracle.online_oracle import OnlineOracle, kw_flat_index
    z = load_online()
    fr, _ = load_case('base', 'fr')
    lr, wd = float(z['lr']), float(z['wd'])
    for w in ('random', 'shipped'):
        oo = OnlineOracle(load_gnn(w), lr=lr, wd=wd)
        for step in ((0, 1) if w == 'random' else (0,)):
            one = fr.slice(step, step + 1)
            flat, dec = oo.decision(one)
            assert dec == z[f'{w}_dec{step}'].tolist()
            kw = z[f'{w}_kw{step}'].tolist()
            grads, loss = oo.online_learning(one, flat, kw, 1.0)
            ref_loss = float(z[f'{w}_gnn_score{step}']) - float(z[f'{w}_kw_score{step}']) + 1.0
            assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
            assert kw_flat_index(one.mask[0], fr.net.hidden_sizes, kw) == sum(fr.net.hidden_sizes[:kw[0]]) + kw[1]
            if step == 0:
                for k, g in grads.items():
                    ref = torch.from_numpy(z[f'{w}_grad0_{k}'])
                    assert float((g - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-12, k
        if w == 'random':
            for k, v in oo.state_dict().items():
                ref = torch.from_numpy(z[f'{w}_sd2_{k}'])
                assert float((v - ref).abs().max()) <= 0.02 * lr, k


# ---- domain queue -------------------------------------------------------------------------------
def test_queue_oracle_matches_reference_traces():
    """oracle/queue_oracle.py against the reference's own add_domain / pick_out / prune_domains on seeded traces
    (tests/golden/make_golden_queue.py): the same domain from every pick (ties, signed zeros, empty picks), the same list left."""
    import numpy as np
    import os
    from golden_io import GOLDEN
    from oracle import queue_oracle as QO
    z = dict(np.load(os.path.join(GOLDEN, 'queue_traces.npz')))
    for t in range(6):
        picked, left = QO.replay(z[f't{t}_ops'])
        assert picked == z[f't{t}_picked'].tolist()
        assert left == z[f't{t}_left'].tolist()


# ---- KW intermediate bounds -----------------------------------------------------------------------
def _kw_err(a, b):
    return float((a.reshape(-1) - b.reshape(-1)).abs().max()) / max(1.0, float(b.abs().max()))


@pytest.mark.parametrize('arch', ARCHS)
def test_kw_bounds_oracle_matches_reference_root_bounds(arch):
    """oracle/kw_bounds_oracle.py against the reference's own DualNetwork on the root domain of the three CIFAR nets
    (tests/golden/nets.npz, written by make_golden.py with convex_adversarial.DualNetwork): every layer's pre-ReLU bounds and
    the output bounds to 2e-5 of the largest bound (fp32 sums over up to 3 072 + |I| columns in a different order)."""
    import numpy as np
    from golden_io import GOLDEN, load_root
    from oracle import kw_bounds_oracle as KW
    net, lbs, ubs, wp, bp = load_root(arch)
    x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(-1)
    got_l, got_u = KW.kw_bounds(net, x, 0.145, wp, bp)
    for k in range(net.L + 2):
        assert _kw_err(got_l[k], lbs[k]) <= 2e-5 and _kw_err(got_u[k], ubs[k]) <= 2e-5, k
        assert bool((got_l[k] <= got_u[k] + 1e-6).all())


@pytest.mark.parametrize('arch', ['base', 'deep', 'wide'])
def test_kw_bounds_oracle_matches_reference_children(arch):
    """Child domains: one ambiguous ReLU of the root fixed to blocked / passing, bounds recomputed with the parent's bounds
    provided (tests/golden/make_golden_kw.py, the DualNetwork(provided_zl, provided_zu) path of init_kw_bounds)."""
    import numpy as np
    from golden_io import GOLDEN, load_root
    from oracle import kw_bounds_oracle as KW
    z = dict(np.load(os.path.join(GOLDEN, 'kw_children.npz')))
    net, lbs, ubs, wp, bp = load_root(arch)
    x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(-1)
    for c in range(int(z[f'{arch}_ncases'])):
        lay, idx, choice = z[f'{arch}_c{c}_decision'].tolist()
        plb, pub = KW.split_bounds(lbs, ubs, (lay, idx), choice)
        got_l, got_u = KW.kw_bounds(net, x, 0.145, wp, bp, plb, pub)
        got_l[-1], got_u[-1] = torch.max(got_l[-1], lbs[-1]), torch.min(got_u[-1], ubs[-1])     # init_kw_bounds :285-286
        for k in range(1, net.L + 2):
            rl, ru = torch.from_numpy(z[f'{arch}_c{c}_lb{k}']), torch.from_numpy(z[f'{arch}_c{c}_ub{k}'])
            assert _kw_err(got_l[k], rl) <= 2e-5 and _kw_err(got_u[k], ru) <= 2e-5, (c, k)
        fixed_l, fixed_u = got_l[lay + 1][idx], got_u[lay + 1][idx]
        assert (float(fixed_u) == 0.0) if choice == 0 else (float(fixed_l) == 0.0)


def _child_cases(arch):
    """(net, x, wp, bp, get(case) -> (lbs, ubs), z, ncases) of tests/golden/child_bounds.npz."""
    import numpy as np
    from golden_io import GOLDEN, load_root
    z = dict(np.load(os.path.join(GOLDEN, 'child_bounds.npz')))
    net, _, _, wp, bp = load_root(arch)
    x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(-1)

    def get(c):
        lbs = [x - 0.145] + [torch.from_numpy(z[f'{arch}_c{c}_lb{k}'].copy()) for k in range(1, net.L + 2)]
        ubs = [x + 0.145] + [torch.from_numpy(z[f'{arch}_c{c}_ub{k}'].copy()) for k in range(1, net.L + 2)]
        return lbs, ubs
    return net, x, wp, bp, get, z, int(z[f'{arch}_ncases'])


@pytest.mark.parametrize('arch', ARCHS)
def test_child_bounds_oracle_matches_reference_update_the_model(arch):
    """oracle child_bounds / root_bounds against the bounds computed by the UNMODIFIED KWConvGen.build_the_model /
    update_the_model of the reference, executed up to their first Gurobi access (tests/golden/make_golden_child.py): root and
    chains of three splits, including the cases that take the interval-triggered second KW pass."""
    from oracle import kw_bounds_oracle as KW
    net, x, wp, bp, get, z, nc = _child_cases(arch)
    rl, ru = KW.root_bounds(net, x, 0.145, wp, bp)
    gl, gu = get(0)
    for k in range(1, net.L + 2):
        assert _kw_err(rl[k], gl[k]) <= 2e-5 and _kw_err(ru[k], gu[k]) <= 2e-5, ('root', k)
    seconds = 0
    for c in range(1, nc):
        lay, idx, choice = z[f'{arch}_c{c}_decision'].tolist()
        pl, pu = get(int(z[f'{arch}_c{c}_parent']))
        cl, cu, second = KW.child_bounds(net, x, 0.145, wp, bp, pl, pu, (lay, idx), choice)
        gl, gu = get(c)
        assert bool(second) == bool(int(z[f'{arch}_c{c}_interval_better'])), c
        seconds += int(second)
        for k in range(1, net.L + 2):
            assert _kw_err(cl[k], gl[k]) <= 2e-5 and _kw_err(cu[k], gu[k]) <= 2e-5, (c, k)
    if arch == 'deep':
        assert seconds >= 2

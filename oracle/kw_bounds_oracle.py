"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the KW (Wong & Kolter) intermediate bounds of the reference.

Only ``tests/`` (and the golden generators) may import this module; the product path never does.  It is the checker for
SURVEY §8f rank 3 (the batched bound producer, the step before the scoring path): ``gnnb_kw_bounds`` and
``gnnb_child_bounds`` (csrc/gnnb_kw.cu) are held to it, and it is held to the reference's own outputs (pins below).

What it restates: ``DualNetwork.__init__`` of the reference's vendored third-party dependency
(convex_adversarial/convex_adversarial/dual_network.py:15-101, dual_layers.py:40-330 ``DualLinear`` / ``DualConv2d`` /
``DualReLU``, dual_inputs.py:24-70 ``InfBall``; un-versioned, vendored in the reference tree) as the reference calls it from
``init_kw_bounds`` (plnn/dual_network_linear_approximation.py:205-270): pre-ReLU bounds of every layer, optionally
intersected with provided bounds (the parent's bounds with one ReLU fixed: ``update_kw_bounds`` :313-319), and the bounds
of the property output.

Algorithm (l-infinity ball of radius eps around x, ``bounded_input=False``).  With D_j = diag(d_j), d_j = [zl_j >= 0] +
I_j * zu_j / (zu_j - zl_j), I_j = [zl_j < 0 < zu_j], and M_{j->k} = A_k D_{k-1} A_{k-1} ... D_{j+1} A_{j+1}:

    centre_k = M_{0->k} x + sum_{j<=k} M_{j->k} b_j
    zl_k = centre_k - eps * rowsum |M_{0->k}| + sum_{j<k} sum_{i in I_j} zl_{j,i} * relu(-d_{j,i} M_{j->k}[:, i])
    zu_k = centre_k + eps * rowsum |M_{0->k}| - sum_{j<k} sum_{i in I_j} zl_{j,i} * relu(+d_{j,i} M_{j->k}[:, i])

evaluated like the reference by pushing columns (x, the identity, the biases, the scaled unit vectors of every I-set)
forward through the layers.  Parity pins: the KW root bounds of the three CIFAR nets computed by the reference's own
``DualNetwork`` (tests/golden/nets.npz, make_golden.py) and child domains with one ReLU fixed (tests/golden/kw_children.npz,
make_golden_kw.py); ``tests/test_oracle.py`` holds this module to both.

``child_bounds`` / ``root_bounds`` restate the bounds part of ``KWConvGen.update_the_model`` / ``build_the_model``
(plnn/conv_kwinter_gen.py:558-660, 199-270): KW bounds from the parent's bounds with one ReLU fixed, interval bounds of the
layers behind the split intersected with them, a second KW pass when the interval bounds tightened a hidden layer.  Pinned
to the UNMODIFIED reference methods executed up to their first Gurobi access (tests/golden/child_bounds.npz,
make_golden_child.py): root + chains of three splits for base, wide and deep, two of which take the second KW pass.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def _apply_affine(a, cols: torch.Tensor, bias: bool = False) -> torch.Tensor:
    """cols [m, n_in] -> [m, n_out] through the affine layer ``a`` (no bias unless asked)."""
    if a.kind == 'conv':
        y = F.conv2d(cols.reshape(cols.shape[0], *a.in_shape), a.weight, a.bias if bias else None, stride=a.stride, padding=a.padding)
        return y.reshape(cols.shape[0], -1)
    return F.linear(cols, a.weight, a.bias if bias else None)


def _bias_row(a) -> torch.Tensor:
    if a.kind == 'conv':
        return a.bias.reshape(-1, 1).expand(a.out_shape[0], a.out_shape[1] * a.out_shape[2]).reshape(1, -1)
    return a.bias.reshape(1, -1)


def kw_bounds(net, x: torch.Tensor, eps: float, wp: torch.Tensor, bp: float,
              provided_lb: Optional[Sequence[torch.Tensor]] = None, provided_ub: Optional[Sequence[torch.Tensor]] = None
              ) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """One domain.  ``x`` flat input [n0]; ``wp`` [n_L], ``bp``: the folded property layer; ``provided_*``: per hidden layer
    k = 1..L a tensor [n_k] (the parent's pre-ReLU bounds with the split applied) or None for the root.
    Returns (lbs, ubs): L+2 tensors — input box, pre-ReLU bounds of the L hidden layers, bounds of the property output."""
    L = net.L
    x = x.reshape(1, -1).float()
    lbs, ubs = [(x - eps).reshape(-1)], [(x + eps).reshape(-1)]
    nu_x = x                                   # [1, n]   x pushed through the linearised network
    nu_1 = torch.eye(x.shape[1])               # [n0, n]  the identity pushed through it
    biases: List[torch.Tensor] = []            # [1, n] each
    isets: List[Tuple[torch.Tensor, torch.Tensor]] = []   # (zl of the I-set [m], columns [m, n])
    affs = list(net.affine) + [None]
    for k in range(1, L + 2):
        if k <= L:
            a = net.affine[k - 1]
            push = lambda c: _apply_affine(a, c)
            brow = _bias_row(a)
        else:
            push = lambda c: c @ wp.reshape(-1, 1)
            brow = torch.tensor([[float(bp)]])
        nu_x, nu_1 = push(nu_x), push(nu_1)
        biases = [push(b) for b in biases] + [brow]
        isets = [(zl_i, push(c)) for zl_i, c in isets]
        centre = nu_x + sum(biases)
        l1 = nu_1.abs().sum(0, keepdim=True)
        zl, zu = centre - eps * l1, centre + eps * l1
        for zl_i, c in isets:                   # dual_layers.py:271-283
            zl = zl + (zl_i.reshape(-1, 1) * (-c).clamp(min=0)).sum(0, keepdim=True)
            zu = zu - (zl_i.reshape(-1, 1) * c.clamp(min=0)).sum(0, keepdim=True)
        zl, zu = zl.reshape(-1), zu.reshape(-1)
        if k <= L and provided_lb is not None:   # dual_network.py:85-86
            zu = torch.min(zu, provided_ub[k - 1].reshape(-1))
            zl = torch.max(zl, provided_lb[k - 1].reshape(-1))
        lbs.append(zl)
        ubs.append(zu)
        if k > L:
            break
        # through the ReLU: d-scaling of every column, new I-set columns (dual_layers.py:208-232, 297-312)
        I = (zu > 0) & (zl < 0)
        d = (zl >= 0).float()
        d[I] = d[I] + zu[I] / (zu[I] - zl[I])
        nu_x, nu_1 = nu_x * d, nu_1 * d
        biases = [b * d for b in biases]
        isets = [(zl_i, c * d) for zl_i, c in isets]
        idx = I.nonzero().reshape(-1)
        if idx.numel() > 0:
            cols = torch.zeros(idx.numel(), zl.numel())
            cols[torch.arange(idx.numel()), idx] = d[idx]
            isets.append((zl[idx], cols))
    return lbs, ubs


def split_bounds(lbs: Sequence[torch.Tensor], ubs: Sequence[torch.Tensor], decision, choice: int):
    """update_kw_bounds' first step (plnn/dual_network_linear_approximation.py:313-319): fix ReLU ``decision`` = (layer,
    index) of the parent's pre-ReLU bounds to blocked (choice 0: u = 0) or passing (choice 1: l = 0).
    Returns (provided_lb, provided_ub) for ``kw_bounds``: the hidden layers' bounds."""
    plb = [t.clone() for t in lbs[1:-1]]
    pub = [t.clone() for t in ubs[1:-1]]
    if choice == 0:
        pub[decision[0]].view(-1)[decision[1]] = 0
    else:
        plb[decision[0]].view(-1)[decision[1]] = 0
    return plb, pub


def interval_layer(a, lo_post: torch.Tensor, hi_post: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Interval bounds of affine layer ``a`` from the box of its input (plnn/conv_kwinter_gen.py:599-628): W+ l + W- u + b."""
    if a.kind == 'conv':
        wp_, wn_ = a.weight.clamp(min=0), a.weight.clamp(max=0)
        l4, u4 = lo_post.reshape(1, *a.in_shape), hi_post.reshape(1, *a.in_shape)
        lo = F.conv2d(l4, wp_, a.bias, stride=a.stride, padding=a.padding) + F.conv2d(u4, wn_, None, stride=a.stride, padding=a.padding)
        hi = F.conv2d(u4, wp_, a.bias, stride=a.stride, padding=a.padding) + F.conv2d(l4, wn_, None, stride=a.stride, padding=a.padding)
        return lo.reshape(-1), hi.reshape(-1)
    wp_, wn_ = a.weight.clamp(min=0), a.weight.clamp(max=0)
    return wp_ @ lo_post + wn_ @ hi_post + a.bias, wp_ @ hi_post + wn_ @ lo_post + a.bias


def _interval_tighten(net, wp, bp, lbs, ubs, first_layer: int) -> Tuple[bool, bool]:
    """Interval pass over hidden layers ``first_layer`` .. L (1-based) and the property output, in place.  Returns
    (a hidden layer changed, anything changed) — the reference re-runs KW only in the first case (:652-656)."""
    L = net.L
    hidden_changed = any_changed = False
    for k in range(first_layer, L + 2):
        lo_post, hi_post = lbs[k - 1].clamp(min=0), ubs[k - 1].clamp(min=0)
        if k <= L:
            lo, hi = interval_layer(net.affine[k - 1], lo_post, hi_post)
        else:
            w = wp.reshape(-1)
            lo = (w.clamp(min=0) * lo_post).sum() + (w.clamp(max=0) * hi_post).sum() + bp
            hi = (w.clamp(min=0) * hi_post).sum() + (w.clamp(max=0) * lo_post).sum() + bp
            lo, hi = lo.reshape(1), hi.reshape(1)
        if not bool((lbs[k] >= lo).all()):
            lbs[k] = torch.max(lbs[k], lo)
            any_changed = True
            hidden_changed |= k <= L
        if not bool((ubs[k] <= hi).all()):
            ubs[k] = torch.min(ubs[k], hi)
            any_changed = True
            hidden_changed |= k <= L
    return hidden_changed, any_changed


def _kw_pass(net, x, eps, wp, bp, lbs, ubs, keep_upto: int):
    """update_kw_bounds (plnn/dual_network_linear_approximation.py:296-439): hidden layers 1 .. keep_upto keep the given
    bounds, later layers get KW bounds (every ReLU linearised with the current bounds) intersected with the given ones, the
    property output likewise."""
    L = net.L
    k_lbs, k_ubs = kw_bounds(net, x, eps, wp, bp, [t for t in lbs[1:L + 1]], [t for t in ubs[1:L + 1]])
    out_l = [lbs[0]] + [lbs[k] if k <= keep_upto else k_lbs[k] for k in range(1, L + 1)] + [torch.max(k_lbs[L + 1], lbs[L + 1])]
    out_u = [ubs[0]] + [ubs[k] if k <= keep_upto else k_ubs[k] for k in range(1, L + 1)] + [torch.min(k_ubs[L + 1], ubs[L + 1])]
    return out_l, out_u


def child_bounds(net, x: torch.Tensor, eps: float, wp: torch.Tensor, bp: float, parent_lbs: Sequence[torch.Tensor],
                 parent_ubs: Sequence[torch.Tensor], decision, choice: int):
    """Bounds part of ``update_the_model`` (plnn/conv_kwinter_gen.py:558-660).  ``parent_*``: L + 2 flat tensors (input box,
    pre-ReLU bounds, property output); ``decision`` = (hidden layer 0-based, index).  Returns (lbs, ubs, second_pass)."""
    lay, idx = int(decision[0]), int(decision[1])
    lbs, ubs = [t.clone().reshape(-1) for t in parent_lbs], [t.clone().reshape(-1) for t in parent_ubs]
    if choice == 0:
        ubs[lay + 1][idx] = 0
    else:
        lbs[lay + 1][idx] = 0
    lbs, ubs = _kw_pass(net, x, eps, wp, bp, lbs, ubs, keep_upto=lay + 1)
    hidden_changed, _ = _interval_tighten(net, wp, bp, lbs, ubs, first_layer=lay + 2)
    if hidden_changed:
        lbs, ubs = _kw_pass(net, x, eps, wp, bp, lbs, ubs, keep_upto=lay + 1)
    return lbs, ubs, hidden_changed


def root_bounds(net, x: torch.Tensor, eps: float, wp: torch.Tensor, bp: float):
    """Bounds part of ``build_the_model`` (plnn/conv_kwinter_gen.py:199-270): KW root bounds, intersected layer by layer with
    interval bounds; if a pre-ReLU layer moved by more than 1e-4, one KW pass from the first such layer."""
    L = net.L
    k_lbs, k_ubs = kw_bounds(net, x, eps, wp, bp)
    lbs, ubs = [k_lbs[0]], [k_ubs[0]]
    for k in range(1, L + 2):
        lo_post, hi_post = (lbs[k - 1], ubs[k - 1]) if k == 1 else (lbs[k - 1].clamp(min=0), ubs[k - 1].clamp(min=0))
        if k <= L:
            lo, hi = interval_layer(net.affine[k - 1], lo_post, hi_post)
        else:
            w = wp.reshape(-1)
            lo = ((w.clamp(min=0) * lo_post).sum() + (w.clamp(max=0) * hi_post).sum() + bp).reshape(1)
            hi = ((w.clamp(min=0) * hi_post).sum() + (w.clamp(max=0) * lo_post).sum() + bp).reshape(1)
        lbs.append(torch.max(k_lbs[k], lo))
        ubs.append(torch.min(k_ubs[k], hi))
    for k in range(1, L + 1):
        if bool(((lbs[k] - k_lbs[k]).abs() > 1e-4).any()) or bool(((ubs[k] - k_ubs[k]).abs() > 1e-4).any()):
            return _kw_pass(net, x, eps, wp, bp, lbs, ubs, keep_upto=k)
    return lbs, ubs

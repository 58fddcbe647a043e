"""Frontier pack: a batch of BaB subdomains laid out for batched scoring.

The reference hands ``GraphNet.forward`` Python lists of per-layer tensors with the batch folded
into the leading dimension (graphnet/graph_conv.py:77-83, 479).  ``Frontier``
holds the same data as per-layer ``[B, n_k]`` struct-of-arrays so a whole BaB frontier is one
contiguous block per field, and converts losslessly (views, no copies for contiguous inputs) to
and from the reference argument lists.

It also holds the seeded synthetic frontier generator of SURVEY §8(d) (configs 2-5): root bounds
of one property, per-domain forced ReLU splits, seeded tightening, surrogate LP primals / duals
(Gurobi is out of scope, so LP values are surrogates by construction).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .networks import NetSpec, netspec_from_modules


@dataclass
class Frontier:
    net: NetSpec
    lb: List[torch.Tensor]            # L+2 tensors [B, n_k]; k=0 input, 1..L hidden (pre-ReLU), L+1 output
    ub: List[torch.Tensor]
    dual: List[torch.Tensor]          # L tensors [B, n_k, 3] (LP duals of the 3 ReLU constraints)
    prim_pre: List[torch.Tensor]      # L tensors [B, n_k]: LP value of the pre-activation  (primals[i_k])
    prim_post: List[torch.Tensor]     # L tensors [B, n_k]: LP value of the post-activation (primals[i_k+1])
    prim_out: torch.Tensor            # [B]    LP value of the property output (primals[-1])
    primal_input: torch.Tensor        # [B, n0] LP argmin input point
    Wp: torch.Tensor                  # [B, n_L] per-domain property layer weight
    bp: torch.Tensor                  # [B]      per-domain property layer bias
    mask: torch.Tensor                # [B, sum n_k] float 0/1: 1 = ReLU still undecided (a branching candidate)

    @property
    def B(self) -> int:
        return int(self.lb[0].shape[0])

    @property
    def device(self):
        return self.lb[0].device

    def tensors(self) -> Dict[str, object]:
        return dict(lb=self.lb, ub=self.ub, dual=self.dual, prim_pre=self.prim_pre, prim_post=self.prim_post,
                    prim_out=self.prim_out, primal_input=self.primal_input, Wp=self.Wp, bp=self.bp, mask=self.mask)

    def _map(self, fn) -> 'Frontier':
        m = lambda v: [fn(t) for t in v] if isinstance(v, list) else fn(v)
        return Frontier(net=self.net, **{k: m(v) for k, v in self.tensors().items()})

    def to(self, device, non_blocking: bool = False) -> 'Frontier':
        out = self._map(lambda t: t.to(device, non_blocking=non_blocking))
        out.net = self.net.to(device)
        return out

    def cpu(self) -> 'Frontier':
        return self.to('cpu')

    def pin(self) -> 'Frontier':
        return self._map(lambda t: t.contiguous().pin_memory())

    def slice(self, start: int, stop: int) -> 'Frontier':
        return self._map(lambda t: t[start:stop])

    def contiguous(self) -> 'Frontier':
        return self._map(lambda t: t.contiguous().float())

    def input_bytes(self) -> int:
        n = 0
        for v in self.tensors().values():
            for t in (v if isinstance(v, list) else [v]):
                n += t.numel() * t.element_size()
        return n

    # ---- reference argument lists -------------------------------------------------------------
    @staticmethod
    def from_reference_args(lower_bounds_all, upper_bounds_all, dual_vars, primals, primal_inputs,
                            layers, masks, net: Optional[NetSpec] = None) -> 'Frontier':
        """Arguments exactly as ``GraphNet.forward`` receives them (graph_conv.py:479)."""
        B = len(lower_bounds_all[0])
        if net is None:
            net = netspec_from_modules(layers['fixed_layers'], tuple(lower_bounds_all[0].shape[1:]))
        L = net.L
        if len(lower_bounds_all) != L + 2 or len(upper_bounds_all) != L + 2:
            raise ValueError(f'expected {L + 2} bound tensors, got {len(lower_bounds_all)}')
        if len(dual_vars) < L:
            raise ValueError(f'expected {L} dual tensors, got {len(dual_vars)}')
        dev = torch.as_tensor(lower_bounds_all[0]).device     # everything follows the bounds' device
        f32 = lambda t: torch.as_tensor(t).to(dev).float()
        lb = [f32(t).reshape(B, -1) for t in lower_bounds_all]
        ub = [f32(t).reshape(B, -1) for t in upper_bounds_all]
        dual = [f32(dual_vars[k]).reshape(B, net.affine[k].n_out, 3) for k in range(L)]
        prim_pre = [f32(primals[a.layer_index]).reshape(B, a.n_out) for a in net.affine]
        prim_post = [f32(primals[a.layer_index + 1]).reshape(B, a.n_out) for a in net.affine]
        prim_out = f32(primals[-1]).reshape(B)
        props = layers['prop_layers']
        if len(props) != B:
            raise ValueError('one property layer per domain is required (graph_conv.py:196)')
        Wp = torch.stack([f32(m.weight.detach()).reshape(-1) for m in props], 0)
        bp = torch.stack([f32(m.bias.detach()).reshape(()) for m in props], 0)
        mask = f32(masks).reshape(B, -1)
        return Frontier(net, lb, ub, dual, prim_pre, prim_post, prim_out, f32(primal_inputs).reshape(B, -1),
                        Wp, bp, mask)

    def to_reference_args(self):
        """-> (lower_bounds_all, upper_bounds_all, dual_vars, primals, primal_inputs, layers, masks)."""
        net, B = self.net, self.B
        shapes = [net.input_shape] + [a.out_shape for a in net.affine] + [(1,)]
        lbs = [t.reshape(B, *s) for t, s in zip(self.lb, shapes)]
        ubs = [t.reshape(B, *s) for t, s in zip(self.ub, shapes)]
        duals = [d.reshape(-1, 3) for d in self.dual]
        sizes = net.primal_sizes()
        primals = [torch.zeros(B * n, device=self.device) for n in sizes]
        for k, a in enumerate(net.affine):
            primals[a.layer_index] = self.prim_pre[k].reshape(-1)
            primals[a.layer_index + 1] = self.prim_post[k].reshape(-1)
        # Flatten outputs repeat the value that flows through them
        for j in range(1, len(sizes) - 1):
            if all(a.layer_index != j and a.layer_index + 1 != j for a in net.affine):
                primals[j] = primals[j - 1]
        primals[-1] = self.prim_out.reshape(-1)
        props = []
        for b in range(B):
            m = nn.Linear(self.Wp.shape[1], 1)
            with torch.no_grad():
                m.weight.copy_(self.Wp[b:b + 1])
                m.bias.copy_(self.bp[b:b + 1])
            for q in m.parameters():
                q.requires_grad = False
            props.append(m.to(self.device))
        layers = {'fixed_layers': [m.to(self.device) for m in net.modules()], 'prop_layers': props}
        return lbs, ubs, duals, primals, self.primal_input.reshape(B, *net.input_shape), layers, self.mask


# ------------------------------------------------------------------------------------------------
# Synthetic frontier generator

def net_forward_activations(net: NetSpec, x: torch.Tensor):
    """Pre- and post-activation of every hidden layer and the last hidden post-activation, for [B,n0] inputs."""
    B = x.shape[0]
    cur = x.reshape(B, *net.input_shape)
    pre, post = [], []
    for a in net.affine:
        if a.kind == 'conv':
            cur = F.conv2d(cur.reshape(B, *a.in_shape), a.weight, a.bias, stride=a.stride, padding=a.padding)
        else:
            cur = F.linear(cur.reshape(B, -1), a.weight, a.bias)
        pre.append(cur.reshape(B, -1))
        cur = F.relu(cur)
        post.append(cur.reshape(B, -1))
    return pre, post


def interval_root_bounds(net: NetSpec, x: torch.Tensor, eps: float, wp: torch.Tensor, bp: float):
    """Interval (box) bounds through the verified net; used only when no KW root bounds are supplied."""
    lo, hi = (x - eps).reshape(1, *net.input_shape), (x + eps).reshape(1, *net.input_shape)
    lbs, ubs = [lo.reshape(-1)], [hi.reshape(-1)]
    for a in net.affine:
        mid, rad = (lo + hi) / 2, (hi - lo) / 2
        if a.kind == 'conv':
            c = F.conv2d(mid.reshape(1, *a.in_shape), a.weight, a.bias, stride=a.stride, padding=a.padding)
            r = F.conv2d(rad.reshape(1, *a.in_shape), a.weight.abs(), None, stride=a.stride, padding=a.padding)
        else:
            c = F.linear(mid.reshape(1, -1), a.weight, a.bias)
            r = F.linear(rad.reshape(1, -1), a.weight.abs())
        lbs.append((c - r).reshape(-1))
        ubs.append((c + r).reshape(-1))
        lo, hi = F.relu(c - r), F.relu(c + r)
    mid, rad = ((lo + hi) / 2).reshape(-1), ((hi - lo) / 2).reshape(-1)
    c, r = (wp * mid).sum() + bp, (wp.abs() * rad).sum()
    lbs.append((c - r).reshape(1))
    ubs.append((c + r).reshape(1))
    return lbs, ubs


def synthetic_frontier(net: NetSpec, root_lb: Sequence[torch.Tensor], root_ub: Sequence[torch.Tensor],
                       wp: torch.Tensor, bp: float, B: int, seed: int, device='cpu',
                       max_splits: int = 32, dual_density: float = 0.1) -> Frontier:
    """B synthetic subdomains of one property:

    * start from the root bounds; per domain force d ~ U{1..max_splits} root-ambiguous ReLUs to one side
      (u=0 or l=0, what update_kw_bounds does first: plnn/dual_network_linear_approximation.py:313-319);
    * tighten every other hidden / output bound toward its midpoint by a seeded factor in [0.9, 1]
      (keeps l <= u and never produces l = u = 0);
    * mask = 1 where l < 0 < u after the edits;
    * surrogate LP point: seeded uniform input in the box, primals = activations of that point;
    * surrogate duals: sparse signed values on ambiguous nodes (cols 0,1 >= 0, col 2 <= 0).
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    net = net.to(dev)
    L = net.L
    rl = [t.to(dev).float().reshape(-1) for t in root_lb]
    ru = [t.to(dev).float().reshape(-1) for t in root_ub]
    rand = lambda *s: torch.rand(*s, generator=g, device=dev)

    lb = [t.unsqueeze(0).repeat(B, 1) for t in rl]
    ub = [t.unsqueeze(0).repeat(B, 1) for t in ru]
    # forced splits on root-ambiguous ReLUs
    amb_root = torch.cat([(rl[k] < 0) & (ru[k] > 0) for k in range(1, L + 1)]).nonzero().reshape(-1)
    n_hidden = net.n_hidden
    force_u0 = torch.zeros(B, n_hidden, dtype=torch.bool, device=dev)
    force_l0 = torch.zeros(B, n_hidden, dtype=torch.bool, device=dev)
    if amb_root.numel() > 0:
        d = torch.randint(1, max_splits + 1, (B, 1), generator=g, device=dev)
        pick = amb_root[torch.randint(0, amb_root.numel(), (B, max_splits), generator=g, device=dev)]
        side = torch.randint(0, 2, (B, max_splits), generator=g, device=dev)
        use = torch.arange(max_splits, device=dev).unsqueeze(0) < d
        rows = torch.arange(B, device=dev).unsqueeze(1).expand_as(pick)
        force_u0[rows[use & (side == 0)], pick[use & (side == 0)]] = True
        force_l0[rows[use & (side == 1)], pick[use & (side == 1)]] = True
        force_l0 &= ~force_u0
    # tightening toward the midpoint (hidden layers and output); a forced node keeps its exact zero
    # and only its free side shrinks, so l < u always holds and l = u = 0 never appears
    off = 0
    for k in range(1, L + 2):
        n = rl[k].numel()
        f = 0.9 + 0.1 * rand(B, n)
        mid = (lb[k] + ub[k]) * 0.5
        lo = mid - f * (mid - lb[k])
        hi = mid + f * (ub[k] - mid)
        if k <= L:
            u0, l0 = force_u0[:, off:off + n], force_l0[:, off:off + n]
            lo = torch.where(u0, lb[k] * f, torch.where(l0, torch.zeros_like(lo), lo))
            hi = torch.where(l0, ub[k] * f, torch.where(u0, torch.zeros_like(hi), hi))
            off += n
        lb[k], ub[k] = lo, hi
    mask = torch.cat([((lb[k] < 0) & (ub[k] > 0)).float() for k in range(1, L + 1)], 1)

    x = lb[0] + (ub[0] - lb[0]) * rand(B, net.n0)
    pre, post = net_forward_activations(net, x)
    wp_b = wp.to(dev).float().reshape(1, -1).repeat(B, 1)
    bp_b = torch.full((B,), float(bp), device=dev)
    prim_out = (wp_b * post[-1]).sum(1) + bp_b

    dual = []
    for k in range(1, L + 1):
        n = rl[k].numel()
        amb = ((lb[k] < 0) & (ub[k] > 0)).float().unsqueeze(-1)
        keep = (rand(B, n, 3) < dual_density).float()
        mag = torch.randn(B, n, 3, generator=g, device=dev).abs()
        sign = torch.tensor([1.0, 1.0, -1.0], device=dev)
        dual.append(mag * keep * amb * sign)
    return Frontier(net, lb, ub, dual, pre, post, prim_out, x, wp_b, bp_b, mask)

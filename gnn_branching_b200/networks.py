"""Verified-network descriptions: the graph the GNN runs on.

The GNN's graph *is* the verified network (reference: graphnet/graph_conv.py:107-137, 222-249):
nodes are the input pixels, every hidden ReLU and one output node; edges are the verified
network's own conv / linear weights.  This module turns the reference's ``layers['fixed_layers']``
list of ``nn.Module`` into a plain shape/weight description (``NetSpec``) that the C-ABI library
and the oracle both consume, and holds the three CIFAR workload architectures
(reference: exp_utils/model_utils.py:120-166 — base ``cifar_model_m2``, wide ``cifar_model``,
deep ``cifar_model_deep``).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn


class Flatten(nn.Module):
    """Marker module with the reference's name and behaviour (plnn/modules.py:4-6)."""

    def forward(self, x):
        return x.view(x.size(0), -1)


_NET_IDS = itertools.count()


@dataclass
class AffineSpec:
    """One edge set A_k of the graph (a conv or linear layer followed by a ReLU)."""

    kind: str                      # 'conv' | 'linear'
    weight: torch.Tensor           # conv: [C_out, C_in, kh, kw]; linear: [out, in]
    bias: torch.Tensor             # [C_out] or [out]
    in_shape: Tuple[int, ...]      # (C, H, W) for conv, (n,) for linear
    out_shape: Tuple[int, ...]
    stride: int = 1
    padding: int = 0
    layer_index: int = 0           # i_k: position of A_k in the verified net's layer list

    @property
    def n_in(self) -> int:
        n = 1
        for s in self.in_shape:
            n *= s
        return n

    @property
    def n_out(self) -> int:
        n = 1
        for s in self.out_shape:
            n *= s
        return n

    @property
    def macs_per_channel(self) -> int:
        """MAC(A_k) of SURVEY §8(d): n_out*C_in*kh*kw (conv) or in*out (linear)."""
        if self.kind == 'conv':
            co, ci, kh, kw = self.weight.shape
            return self.n_out * ci * kh * kw
        return self.weight.shape[0] * self.weight.shape[1]


@dataclass
class NetSpec:
    """Shape + weights of a verified network, property layer excluded."""

    name: str
    input_shape: Tuple[int, int, int]
    affine: List[AffineSpec] = field(default_factory=list)
    n_layers_total: int = 0        # number of modules in fixed_layers
    key: object = None             # identity of the network's weights; survives .to(); lets the engine skip re-uploads

    def __post_init__(self):
        if self.key is None:
            self.key = ('netspec', next(_NET_IDS))

    @property
    def L(self) -> int:
        return len(self.affine)

    @property
    def n0(self) -> int:
        c, h, w = self.input_shape
        return c * h * w

    @property
    def hidden_sizes(self) -> List[int]:
        return [a.n_out for a in self.affine]

    @property
    def n_hidden(self) -> int:
        return sum(self.hidden_sizes)

    @property
    def hidden_offsets(self) -> List[int]:
        """Start of each hidden layer in the flat ReLU index (graph_score.py:14-19 trans_len)."""
        off, out = 0, []
        for n in self.hidden_sizes:
            out.append(off)
            off += n
        return out

    def primal_sizes(self) -> List[int]:
        """Sizes of the reference's ``primals`` list: one entry per verified-net layer output
        (every module of fixed_layers, then the property layer); plnn/conv_kwinter_gen.py:549-554."""
        sizes = [0] * (self.n_layers_total + 1)
        cur = self.n0
        k = 0
        for j in range(self.n_layers_total):
            if k < self.L and self.affine[k].layer_index == j:
                cur = self.affine[k].n_out
                k += 1
            sizes[j] = cur
        sizes[self.n_layers_total] = 1
        return sizes

    def to(self, device) -> 'NetSpec':
        out = NetSpec(self.name, self.input_shape, [], self.n_layers_total, key=self.key)
        for a in self.affine:
            out.affine.append(AffineSpec(a.kind, a.weight.to(device), a.bias.to(device), a.in_shape,
                                         a.out_shape, a.stride, a.padding, a.layer_index))
        return out

    def modules(self) -> List[nn.Module]:
        """Rebuild the ``fixed_layers`` module list (what the reference API is handed)."""
        mods: List[nn.Module] = []
        flat = False
        for a in self.affine:
            while len(mods) < a.layer_index:
                mods.append(Flatten())
                flat = True
            if a.kind == 'conv':
                co, ci, kh, kw = a.weight.shape
                m = nn.Conv2d(ci, co, (kh, kw), stride=a.stride, padding=a.padding)
            else:
                m = nn.Linear(a.weight.shape[1], a.weight.shape[0])
            with torch.no_grad():
                m.weight.copy_(a.weight)
                m.bias.copy_(a.bias)
            for p in m.parameters():
                p.requires_grad = False
            mods.append(m)
            mods.append(nn.ReLU())
        del flat
        return mods

    def flops_per_domain(self, T: int = 2, p: int = 64) -> float:
        """Algorithmic FLOPs per subdomain, SURVEY §8(d) contract figure (dense, dead work excluded)."""
        n = self.hidden_sizes
        mac = [a.macs_per_channel for a in self.affine]
        nL = n[-1]
        macs = (T * sum(nk * (14 * p + 19 * p * p) for nk in n)
                + T * (4 * p + 3 * p * p)
                + (T - 1) * self.n0 * (2 * p + 4 * p * p)
                + self.n0 * (3 * p + p * p)
                + sum(nk * (p * p + p) for nk in n)
                + T * p * (sum(mac) + nL)
                + T * p * (sum(mac[1:]) + nL)
                + (T - 1) * p * mac[0])
        return 2.0 * macs


def netspec_from_modules(fixed_layers: Sequence[nn.Module], input_shape: Tuple[int, int, int],
                         name: str = 'net') -> NetSpec:
    """Walk ``layers['fixed_layers']`` the way the reference forward does
    (graph_conv.py:107-192: Conv2d / Linear / ReLU / Flatten dispatch, anything else raises)."""
    spec = NetSpec(name=name, input_shape=tuple(int(s) for s in input_shape),
                   n_layers_total=len(fixed_layers))
    shape: Tuple[int, ...] = spec.input_shape
    pending: Optional[AffineSpec] = None
    for j, m in enumerate(fixed_layers):
        if isinstance(m, nn.Conv2d):
            if pending is not None:
                raise NotImplementedError('two affine layers without a ReLU in between')
            if len(shape) != 3:
                raise NotImplementedError('Conv2d after Flatten')
            kh, kw = m.kernel_size
            sh, sw = m.stride
            ph, pw = m.padding
            if kh != kw or sh != sw or ph != pw or m.dilation != (1, 1) or m.groups != 1:
                raise NotImplementedError('only square, undilated, ungrouped convolutions')
            c, h, w = shape
            ho = (h + 2 * ph - kh) // sh + 1
            wo = (w + 2 * pw - kw) // sw + 1
            out_shape = (m.out_channels, ho, wo)
            pending = AffineSpec('conv', m.weight.detach(), m.bias.detach(), shape, out_shape,
                                 stride=sh, padding=ph, layer_index=j)
            shape = out_shape
        elif isinstance(m, nn.Linear):
            if pending is not None:
                raise NotImplementedError('two affine layers without a ReLU in between')
            n = 1
            for s in shape:
                n *= s
            if n != m.in_features:
                raise ValueError('Linear in_features does not match the running shape')
            pending = AffineSpec('linear', m.weight.detach(), m.bias.detach(), (n,), (m.out_features,),
                                 layer_index=j)
            shape = (m.out_features,)
        elif isinstance(m, nn.ReLU):
            if pending is None:
                raise NotImplementedError('ReLU without a preceding affine layer')
            spec.affine.append(pending)
            pending = None
        elif type(m).__name__ == 'Flatten':
            n = 1
            for s in shape:
                n *= s
            shape = (n,)
        else:
            raise NotImplementedError(f'unsupported layer type {type(m).__name__}')
    if pending is not None:
        raise NotImplementedError('fixed_layers must end with a ReLU (the property layer is separate)')
    # the same module list scored again (every BaB step) maps to the same key, edited weights to a new one
    spec.key = ('modules',) + tuple((a.kind, a.weight.data_ptr(), a.weight._version, a.bias.data_ptr(), a.bias._version,
                                     a.stride, a.padding, tuple(a.in_shape)) for a in spec.affine)
    return spec


# ------------------------------------------------------------------------------------------------
# CIFAR workload shapes (exp_utils/model_utils.py:120-166).  Layers listed without the last Linear(100,10):
# the reference folds it with the property into Linear(100,1) (model_utils.py:187-208).
_ARCH = {
    'base': [('conv', 3, 8, 4, 2, 1), ('conv', 8, 16, 4, 2, 1), ('flatten',), ('linear', 1024, 100)],
    'wide': [('conv', 3, 16, 4, 2, 1), ('conv', 16, 32, 4, 2, 1), ('flatten',), ('linear', 2048, 100)],
    'deep': [('conv', 3, 8, 4, 2, 1), ('conv', 8, 8, 3, 1, 1), ('conv', 8, 8, 3, 1, 1),
             ('conv', 8, 8, 4, 2, 1), ('flatten',), ('linear', 512, 100)],
}


def cifar_modules(arch: str, seed: int = 0) -> List[nn.Module]:
    """``fixed_layers`` of a CIFAR base / wide / deep verified net with seeded random weights."""
    g = torch.Generator().manual_seed(seed)
    mods: List[nn.Module] = []
    for item in _ARCH[arch]:
        if item[0] == 'conv':
            _, ci, co, k, s, p = item
            m = nn.Conv2d(ci, co, k, stride=s, padding=p)
            fan = ci * k * k
        elif item[0] == 'linear':
            _, fi, fo = item
            m = nn.Linear(fi, fo)
            fan = fi
        else:
            mods.append(Flatten())
            continue
        with torch.no_grad():
            bound = 1.0 / fan ** 0.5
            m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)
            m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
        for q in m.parameters():
            q.requires_grad = False
        mods.append(m)
        mods.append(nn.ReLU())
    return mods


def cifar_netspec(arch: str, seed: int = 0) -> NetSpec:
    return netspec_from_modules(cifar_modules(arch, seed), (3, 32, 32), name=f'cifar_{arch}')

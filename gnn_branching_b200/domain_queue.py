"""Device-resident domain queue of the branch-and-bound loop (``gnnb_queue_*``, csrc/gnnb_queue.cu).

Replaces the sorted Python list ``domains`` of ``ReLUDomain`` objects and ``add_domain`` / ``pick_out`` /
``prune_domains`` of the reference (plnn/branch_and_bound.py:159-184, 264-281; plnn/relu_conv_gnnkwthreshold.py:20-53,
126-244).  Two surfaces:

* batched: ``DomainQueue.add(batch)``, ``.pick(max_B, threshold)``, ``.prune(threshold)``, ``len(q)``, ``.global_lb`` — a
  ``DomainBatch`` holds the same per-layer ``[B, n_k]`` bound arrays as ``Frontier`` (so picked domains go straight to
  the bounding / scoring kernels), the int8 mask in the BaB convention (-1 undecided, 0 / 1 fixed), and the stored decision;
* the reference's function API on ``ReLUDomain`` objects, one domain at a time: ``add_domain(candidate, q)``,
  ``pick_out(q, threshold)``, ``prune_domains(q, threshold)``, ``q[0].lower_bound``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib
from .engine import Scorer
from .networks import NetSpec


@dataclass
class DomainBatch:
    lower_bound: torch.Tensor            # [B] f32   ReLUDomain.lower_bound
    upper_bound: torch.Tensor            # [B] f32   ReLUDomain.upper_bound
    lb: List[torch.Tensor]               # L+2 tensors [B, n_k] f32 (layers 0..L+1 as in Frontier.lb)
    ub: List[torch.Tensor]
    mask: torch.Tensor                   # [B, sum n_k] int8: -1 undecided, 0 / 1 fixed
    decision: Optional[torch.Tensor]     # [B, 2] int32 (layer, index) or None

    @property
    def B(self) -> int:
        return int(self.lower_bound.shape[0])

    @property
    def device(self):
        return self.lower_bound.device

    def slice(self, a: int, b: int) -> 'DomainBatch':
        return DomainBatch(self.lower_bound[a:b], self.upper_bound[a:b], [t[a:b] for t in self.lb], [t[a:b] for t in self.ub],
                           self.mask[a:b], None if self.decision is None else self.decision[a:b])

    def to(self, device) -> 'DomainBatch':
        return DomainBatch(self.lower_bound.to(device), self.upper_bound.to(device), [t.to(device) for t in self.lb],
                           [t.to(device) for t in self.ub], self.mask.to(device),
                           None if self.decision is None else self.decision.to(device))

    @staticmethod
    def empty(net: NetSpec, B: int, device) -> 'DomainBatch':
        sizes = [net.n0] + net.hidden_sizes + [1]
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        return DomainBatch(f(B), f(B), [f(B, n) for n in sizes], [f(B, n) for n in sizes],
                           torch.empty(B, net.n_hidden, dtype=torch.int8, device=device),
                           torch.empty(B, 2, dtype=torch.int32, device=device))


class ReLUDomain:
    """plnn/relu_conv_gnnkwthreshold.py:20-53 — what ``pick_out`` returns and ``add_domain`` takes (one domain)."""

    def __init__(self, mask, lb=-float('inf'), ub=float('inf'), lb_all=None, up_all=None, gnn_decision=None):
        self.mask, self.lower_bound, self.upper_bound = mask, lb, ub
        self.lower_all, self.upper_all, self.gnn_decision = lb_all, up_all, gnn_decision

    def __lt__(self, other):
        return self.lower_bound < other.lower_bound

    def __le__(self, other):
        return self.lower_bound <= other.lower_bound

    def __eq__(self, other):
        return self.lower_bound == other.lower_bound


class DomainQueue:
    def __init__(self, scorer: Scorer, capacity: int):
        if scorer.net is None:
            raise RuntimeError('set_network first')
        self.scorer, self.net, self.lib = scorer, scorer.net, scorer.lib
        self.sizes = [self.net.n0] + self.net.hidden_sizes + [1]
        h = C.c_void_p()
        scorer._ok(self.lib.gnnb_queue_create(scorer.h, int(capacity), C.byref(h)))
        self.h = h
        self.capacity = int(capacity)

    def __del__(self):
        h, self.h = getattr(self, 'h', None), None
        if h:
            try:
                self.lib.gnnb_queue_destroy(h)
            except Exception:
                pass

    # ---- plumbing ----
    def _desc(self, d: DomainBatch, B: int):
        f32 = lambda t, shape: Scorer._as_f32(t, shape)
        lb = [f32(d.lb[k], (B, self.sizes[k])) for k in range(len(self.sizes))]
        ub = [f32(d.ub[k], (B, self.sizes[k])) for k in range(len(self.sizes))]
        lower, upper = f32(d.lower_bound, (B,)), f32(d.upper_bound, (B,))
        mask = d.mask if (d.mask.dtype == torch.int8 and d.mask.is_contiguous()) else d.mask.to(torch.int8).contiguous()
        if tuple(mask.shape) != (B, self.net.n_hidden):
            raise ValueError(f'mask must be [{B}, {self.net.n_hidden}]')
        dec = None
        if d.decision is not None:
            dec = d.decision if (d.decision.dtype == torch.int32 and d.decision.is_contiguous()) else d.decision.to(torch.int32).contiguous()
        desc = _lib.DomainsDesc()
        desc.B, desc.mem = B, (_lib.MEM_HOST if d.device.type == 'cpu' else _lib.MEM_DEVICE)
        plb, k1 = _lib.fptr_array(lb)
        pub, k2 = _lib.fptr_array(ub)
        desc.lb, desc.ub, desc.lower_bound, desc.upper_bound = plb, pub, _lib.fptr(lower), _lib.fptr(upper)
        desc.mask = C.cast(mask.data_ptr(), C.POINTER(C.c_int8))
        desc.decision = C.cast(dec.data_ptr(), C.POINTER(C.c_int32)) if dec is not None else None
        return desc, (lb, ub, lower, upper, mask, dec, k1, k2)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- batched surface ----
    def add(self, d: DomainBatch, keep: Optional[torch.Tensor] = None) -> int:
        """``add_domain`` for every domain of ``d`` (those with ``keep[b]`` when given).  Returns how many were added."""
        if d.B == 0:
            return 0
        desc, hold = self._desc(d, d.B)
        kp = None
        if keep is not None:
            kp = keep.to(device=d.device, dtype=torch.uint8).contiguous()
        n = C.c_int32(0)
        with torch.cuda.device(self.scorer.device):
            self.scorer._ok(self.lib.gnnb_queue_add(self.h, C.byref(desc), C.cast(kp.data_ptr(), C.POINTER(C.c_uint8)) if kp is not None else None,
                                                    C.byref(n), self._stream()))
        del hold
        return int(n.value)

    def pick(self, max_B: int, threshold: float, device=None, discard_rest: bool = False) -> DomainBatch:
        """The next ``max_B`` results of ``pick_out(domains, threshold)``, in order (fewer when the queue runs out of domains
        below the threshold).  ``discard_rest=True`` drops every remaining domain in that case, as the reference's loop does."""
        device = torch.device(device) if device is not None else torch.device('cuda', self.scorer.device)
        out = DomainBatch.empty(self.net, max_B, device)
        if max_B == 0:
            return out
        desc, hold = self._desc(out, max_B)
        n = C.c_int32(0)
        with torch.cuda.device(self.scorer.device):
            self.scorer._ok(self.lib.gnnb_queue_pick(self.h, float(threshold), 1 if discard_rest else 0, C.byref(desc), C.byref(n),
                                                     self._stream()))
        del hold
        return out.slice(0, int(n.value))

    def prune(self, threshold: float) -> None:
        with torch.cuda.device(self.scorer.device):
            self.scorer._ok(self.lib.gnnb_queue_prune(self.h, float(threshold), self._stream()))

    def __len__(self) -> int:
        n = C.c_int64(0)
        self.scorer._ok(self.lib.gnnb_queue_stats(self.h, C.byref(n), None, None))
        return int(n.value)

    @property
    def global_lb(self) -> float:
        """``domains[0].lower_bound``."""
        if len(self) == 0:
            raise IndexError('the domain queue is empty')
        v = C.c_float(0)
        with torch.cuda.device(self.scorer.device):
            self.scorer._ok(self.lib.gnnb_queue_stats(self.h, None, C.cast(C.byref(v), C.POINTER(C.c_float)), self._stream()))
        return float(v.value)

    # ---- the reference's per-domain surface ----
    def __getitem__(self, i):
        if i != 0:
            raise IndexError('only domains[0] (the global lower bound) is addressable')
        return ReLUDomain(None, lb=self.global_lb)

    def _batch_of(self, c: ReLUDomain) -> DomainBatch:
        """ReLUDomain -> a batch of one.  ``lower_all`` / ``upper_all`` hold the bounds of the L+2 layers the GNN reads
        (the reference's ``[lower_bounds_all[i] for i in bounds_indices]``, relu_conv_gnnkwthreshold.py:111-114)."""
        dev = torch.device('cuda', self.scorer.device)
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(1, -1).to(dev)
        mask = torch.cat([torch.as_tensor(m).reshape(-1) for m in c.mask]).to(torch.int8).reshape(1, -1).to(dev)
        dec = None if c.gnn_decision is None else torch.tensor([list(c.gnn_decision)], dtype=torch.int32, device=dev)
        return DomainBatch(torch.tensor([float(c.lower_bound)], device=dev), torch.tensor([float(c.upper_bound)], device=dev),
                           [f(t) for t in c.lower_all], [f(t) for t in c.upper_all], mask, dec)


def add_domain(candidate: ReLUDomain, domains: DomainQueue) -> None:
    """plnn/branch_and_bound.py:159-164."""
    domains.add(domains._batch_of(candidate))


def pick_out(domains: DomainQueue, threshold: float) -> ReLUDomain:
    """plnn/branch_and_bound.py:167-184."""
    assert len(domains) > 0, 'The given domains list is empty.'
    b = domains.pick(1, threshold, discard_rest=True)
    assert b.B == 1, 'No domain left to pick from.'
    mask, off = [], 0
    for n in domains.net.hidden_sizes:
        mask.append(b.mask[0, off:off + n].to(torch.int32))
        off += n
    shapes = [domains.net.input_shape] + [a.out_shape for a in domains.net.affine] + [(1,)]
    dec = None if int(b.decision[0, 0]) < 0 else [int(b.decision[0, 0]), int(b.decision[0, 1])]
    return ReLUDomain(mask, lb=float(b.lower_bound[0]), ub=float(b.upper_bound[0]),
                      lb_all=[t[0].reshape(*s) for t, s in zip(b.lb, shapes)], up_all=[t[0].reshape(*s) for t, s in zip(b.ub, shapes)],
                      gnn_decision=dec)


def prune_domains(domains: DomainQueue, threshold: float) -> DomainQueue:
    """plnn/branch_and_bound.py:264-281."""
    domains.prune(threshold)
    return domains

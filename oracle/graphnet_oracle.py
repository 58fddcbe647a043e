"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference GNN branching-score forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module, and only as the checker / the timed CPU baseline.  The product path
(``gnn_branching_b200``) never imports it and fails loudly when its CUDA library is missing.

What it restates: ``GraphNet.forward`` + ``GraphChoice.decision`` of oval-group/GNN_branching
(graphnet/graph_conv.py:77-388 update, :442-470 score head, :487-514 init / ratio;
graphnet/graph_score.py:21-56 mask + index mapping) as batched fp32 PyTorch-CPU tensor algebra.
The arithmetic lives in PyTorch (addmm / conv2d / conv_transpose2d), unpinned by the reference
(README.md:6 asks for ``pytorch >= 0.4.1``); this image has torch 2.11.0.

Parity pin: the reference has no tests or golden vectors for this path, so the pin is
the reference module itself, imported from /root/reference in the build container by
``tests/golden/make_golden.py``; its outputs are committed under ``tests/golden/`` and this oracle
is checked against them in ``tests/test_oracle.py`` (max abs diff ~1e-6 on scores of magnitude 1-80).

Dead work of the reference is not restated: the unused ``ratio`` chain (graph_conv.py:214-216, 228,
243, 356) and the last round's input-layer update (graph_conv.py:360-385 when i == T-1; mu[0] is
read only by the next round).  ``dead_input_update=True`` re-enables the latter for cross-checks.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# state_dict key stems, in the order EmbedLayerUpdate.__init__ creates them (graph_conv.py:36-74)
UPDATE_LINEARS = ['inp_f', 'inp_f_1', 'inp_b', 'inp_b_1', 'inp_b2', 'inp_b2_2', 'fc1', 'fc1_1', 'fc3', 'fc3_2',
                  'fc4', 'fc4_2', 'out1', 'out2', 'out3', 'bc1', 'bc1_1', 'bc1_2', 'bc2', 'bc2_1', 'bc3', 'bc3_1',
                  'bc4', 'bc4_1']
SCORE_LINEARS = ['fnode', 'fscore']


def compute_ratio(l: torch.Tensor, u: torch.Tensor):
    """graph_conv.py:499-514, same operation order."""
    lt = l - F.relu(l)
    ut = F.relu(u)
    r0 = ut / (ut - lt)
    beta = -1 * lt * r0
    amb = (beta > 0).float()
    r1 = (1 - 2 * (r0 * amb)) * amb + r0
    return r0, r1, beta, amb


class _Params:
    def __init__(self, sd: Dict[str, torch.Tensor], keep_graph: bool = False, device='cpu'):
        # keep_graph: use the tensors as given (fp32 CPU leaves that require grad) so that autograd reaches them
        # device: 'cpu' everywhere except bench.py's "eager PyTorch on the GPU" baseline leg
        self.sd = dict(sd) if keep_graph else {k: v.detach().float().to(device) for k, v in sd.items()}

    def lin(self, name: str, x: torch.Tensor) -> torch.Tensor:
        pre = 'ComputeFinalScore.' if name in SCORE_LINEARS else 'EmbedUpdates.update.'
        return F.linear(x, self.sd[pre + name + '.weight'], self.sd[pre + name + '.bias'])

    def mlp2(self, a: str, b: str, x: torch.Tensor) -> torch.Tensor:
        return self.lin(b, F.relu(self.lin(a, x)))


def _conv_forward(a, mu_prev: torch.Tensor) -> torch.Tensor:
    """graph_conv.py:110-121: conv over the (B*p)-batched embeddings, bias not added."""
    B, n, p = mu_prev.shape
    x = mu_prev.permute(0, 2, 1).reshape(B * p, *a.in_shape)
    y = F.conv2d(x, a.weight, None, stride=a.stride, padding=a.padding)
    return y.reshape(B, p, -1).permute(0, 2, 1)


def _conv_backward(a, mu_next: torch.Tensor, normalise: bool) -> torch.Tensor:
    """graph_conv.py:299-318 (hidden layers, divided by the tap count ``freq``) and :361-372
    (input layer, not divided)."""
    B, n, p = mu_next.shape
    x = mu_next.permute(0, 2, 1).reshape(B * p, *a.out_shape)
    y = F.conv_transpose2d(x, a.weight, None, stride=a.stride, padding=a.padding)
    if tuple(y.shape[1:]) != tuple(a.in_shape):
        raise NotImplementedError('conv_transpose2d output does not match the layer input (output_padding needed)')
    if normalise:
        kh, kw = a.weight.shape[2:]
        freq = F.conv_transpose2d(torch.ones(1, 1, *a.out_shape[1:], device=y.device), torch.ones(1, 1, kh, kw, device=y.device), None,
                                  stride=a.stride, padding=a.padding)
        y = y / freq
    return y.reshape(B, p, -1).permute(0, 2, 1)


def _bias_per_node(a) -> torch.Tensor:
    """graph_conv.py:122-124, 133, 264-270."""
    if a.kind == 'conv':
        return a.bias.reshape(-1, 1).expand(a.out_shape[0], a.out_shape[1] * a.out_shape[2]).reshape(-1)
    return a.bias


def gnn_forward(state_dict: Dict[str, torch.Tensor], fr, T: int = 2, p: int = 64,
                stages: Optional[dict] = None, dead_input_update: bool = False, keep_graph: bool = False):
    """Batched restatement of GraphNet.forward (graph_conv.py:479-483).

    ``fr`` is a ``gnn_branching_b200.frontier.Frontier`` on the CPU.  Returns
    ``(scores [B, sum n_k], mu)`` where scores are computed for *every* hidden node (callers apply the
    mask; the reference evaluates the head only on rows with mask != 0, graph_conv.py:445-450).
    ``stages`` (optional dict) receives named intermediates for kernel-by-kernel debugging.
    """
    P = _Params(state_dict, keep_graph, device=fr.lb[0].device)
    net = fr.net
    L, B = net.L, fr.B
    rec = (lambda k, v: stages.__setitem__(k, v.clone())) if stages is not None else (lambda k, v: None)

    # init_mu (graph_conv.py:487-496)
    sizes = [net.n0] + net.hidden_sizes + [1]
    mu: List[torch.Tensor] = [torch.zeros(B, n, p, device=fr.lb[0].device) for n in sizes]
    l0, u0 = fr.lb[0], fr.ub[0]

    for t in range(T):
        if t == 0:   # graph_conv.py:90-95
            inp = torch.stack([l0, fr.primal_input, u0], -1)
            mu[0] = P.mlp2('inp_f', 'inp_f_1', inp)
            rec('mu0_embed', mu[0])
        # ---- forward sweep (graph_conv.py:107-192) ----
        for k in range(1, L + 1):
            a = net.affine[k - 1]
            if a.kind == 'conv':
                nb = _conv_forward(a, mu[k - 1])
            else:                                   # graph_conv.py:130-132
                nb = torch.matmul(a.weight, mu[k - 1])
            l, u = fr.lb[k], fr.ub[k]
            r0, r1, beta, amb = compute_ratio(l, u)
            d = fr.dual[k - 1]
            bias = _bias_per_node(a).unsqueeze(0).expand(B, -1)
            feat = torch.stack([beta, l, u, d[:, :, 1] - d[:, :, 2], fr.prim_pre[k - 1], fr.prim_post[k - 1], bias], -1)
            relax = P.mlp2('fc1', 'fc1_1', feat) * amb.unsqueeze(-1)             # :153-161
            e = P.mlp2('fc3', 'fc3_2', torch.cat([nb * r0.unsqueeze(-1), nb * r1.unsqueeze(-1)], -1))   # :169-170
            new = P.mlp2('fc4', 'fc4_2', torch.cat([relax, e], -1))             # :176-177
            mu[k] = new * (r0 != 0).float().unsqueeze(-1)                        # :178
            rec(f't{t}_fwd_nb{k}', nb); rec(f't{t}_fwd_relax{k}', relax); rec(f't{t}_fwd_mu{k}', mu[k])
        # ---- output node (graph_conv.py:196-210) ----
        nb = torch.einsum('bn,bnp->bp', fr.Wp, mu[L]).unsqueeze(1)
        feat = torch.stack([fr.lb[L + 1][:, 0], fr.ub[L + 1][:, 0], fr.prim_out, fr.bp], -1).unsqueeze(1)
        h = F.relu(P.lin('out1', feat))
        mu[L + 1] = P.lin('out3', F.relu(P.lin('out2', torch.cat([h, nb], -1))))
        rec(f't{t}_mu_out', mu[L + 1])
        # ---- backward sweep (graph_conv.py:222-350) ----
        for k in range(L, 0, -1):
            a = net.affine[k - 1]
            l, u = fr.lb[k], fr.ub[k]
            r0, r1, beta, amb = compute_ratio(l, u)
            d = fr.dual[k - 1]
            bias = _bias_per_node(a).unsqueeze(0).expand(B, -1)
            feat = torch.stack([l, u, beta, -d[:, :, 2] + d[:, :, 1], fr.prim_post[k - 1], fr.prim_pre[k - 1], bias], -1)
            s1 = P.lin('bc1_2', F.relu(P.lin('bc1_1', F.relu(P.lin('bc1', feat)))))           # :285
            s2 = torch.cat([s1, s1 * (-d[:, :, 2]).unsqueeze(-1), s1 * d[:, :, 1].unsqueeze(-1)], -1)   # :287-290
            relax = P.mlp2('bc2', 'bc2_1', s2) * amb.unsqueeze(-1)                            # :291-293
            if k == L:                               # next layer is the property layer, :324-326
                nb = fr.Wp.unsqueeze(-1) * mu[L + 1]             # [B,nL,1]*[B,1,p]
            else:
                nxt = net.affine[k]
                if nxt.kind == 'conv':               # :299-318
                    nb = _conv_backward(nxt, mu[k + 1], normalise=True)
                else:                                # :320-322
                    nb = torch.matmul(nxt.weight.t(), mu[k + 1])
            e = P.mlp2('bc3', 'bc3_1', torch.cat([nb * r0.unsqueeze(-1), nb * r1.unsqueeze(-1)], -1))   # :331-336
            new = P.mlp2('bc4', 'bc4_1', torch.cat([relax, e], -1))                           # :344-345
            mu[k] = new * (r0 != 0).float().unsqueeze(-1)                                      # :347
            rec(f't{t}_bwd_nb{k}', nb); rec(f't{t}_bwd_relax{k}', relax); rec(f't{t}_bwd_mu{k}', mu[k])
        # ---- input layer (graph_conv.py:360-385); dead on the last round ----
        if t < T - 1 or dead_input_update:
            a = net.affine[0]
            if a.kind == 'conv':
                nb = _conv_backward(a, mu[1], normalise=False)
            else:
                nb = torch.matmul(a.weight.t(), mu[1])
            inp_relax = P.mlp2('inp_b', 'inp_b_1', torch.stack([l0, u0], -1))
            mu[0] = P.mlp2('inp_b2', 'inp_b2_2', torch.cat([inp_relax, nb], -1))
            rec(f't{t}_mu0', mu[0])

    # ---- score head (graph_conv.py:442-450), evaluated densely ----
    H = torch.cat(mu[1:L + 1], 1)
    scores = P.lin('fscore', F.relu(P.lin('fnode', H)))[..., 0]
    return scores, mu


def ragged_scores(scores: torch.Tensor, mask: torch.Tensor) -> List[torch.Tensor]:
    """What GraphNet.forward returns: per domain, scores of rows with mask != 0 (graph_conv.py:445-450)."""
    return [scores[b][mask[b].nonzero().view(-1)] for b in range(scores.shape[0])]


def decide(scores: torch.Tensor, mask: torch.Tensor, hidden_sizes) -> Tuple[torch.Tensor, torch.Tensor, List[List[int]]]:
    """GraphChoice.decision index mapping (graph_score.py:41-47), batched.

    Returns best score [B], best flat ReLU index [B] (-1 when no candidate) and [layer, idx] per domain.
    torch.max on CPU returns the first maximal element; masked argmax over the dense vector with
    lowest-index tie-break selects the same flat index.
    """
    B = scores.shape[0]
    best, flat, dec = torch.full((B,), float('-inf')), torch.full((B,), -1, dtype=torch.long), []
    trans_len = torch.tensor(hidden_sizes).cumsum(0)
    for b in range(B):
        idxs = mask[b].nonzero().view(-1)
        if idxs.numel() == 0:
            dec.append([-1, -1])
            continue
        s = scores[b][idxs]
        v, choice = torch.max(s, 0)
        idx = int(idxs[choice])
        lay = int((trans_len > idx).nonzero()[0])
        dec.append([lay, idx if lay == 0 else idx - int(trans_len[lay - 1])])
        best[b], flat[b] = v, idx
    return best, flat, dec


def parity_report(s_new: torch.Tensor, s_ref: torch.Tensor, mask: torch.Tensor, flat_new: torch.Tensor,
                  rtol: float = 1e-4) -> dict:
    """The parity criteria of BASELINE.json / SURVEY §8(c):
    per domain ``max|s_new - s_ref| <= rtol * max|s_ref|`` over candidate rows, and the same argmax
    wherever the reference's top-2 margin exceeds ``2*rtol*max|s_ref|``."""
    B = s_ref.shape[0]
    worst, bad_dec, checked_dec, nan = 0.0, 0, 0, 0
    for b in range(B):
        idxs = mask[b].nonzero().view(-1)
        if idxs.numel() == 0:
            if int(flat_new[b]) != -1:
                bad_dec += 1
            continue
        r, n = s_ref[b][idxs].double(), s_new[b][idxs].double()
        if not torch.isfinite(n).all():
            nan += 1
            continue
        scale = float(r.abs().max())
        worst = max(worst, float((r - n).abs().max()) / max(scale, 1e-30))
        top = torch.topk(r, min(2, r.numel()))
        margin = float(top.values[0] - top.values[1]) if r.numel() > 1 else float('inf')
        if margin > 2 * rtol * scale:
            checked_dec += 1
            if int(flat_new[b]) != int(idxs[top.indices[0]]):
                bad_dec += 1
    return dict(max_norm_err=worst, decisions_checked=checked_dec, decisions_wrong=bad_dec, nonfinite_domains=nan,
                ok=(worst <= rtol and bad_dec == 0 and nan == 0))

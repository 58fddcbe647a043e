"""Diagnostic: gnnb_score_grad vs an fp64 autograd oracle, per tensor (worst first), next to the fp32 CPU oracle's own error."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_case, load_gnn
from gnn_branching_b200 import GraphNet
from oracle import graphnet_oracle as O, online_oracle as OO


def f64_grads(sd, fr, terms, T):
    fr64 = fr._map(lambda t: t.double())
    net64 = copy.deepcopy(fr.net)
    for a in net64.affine:
        a.weight = a.weight.double(); a.bias = a.bias.double()
    fr64.net = net64
    params = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    torch.set_default_dtype(torch.float64)
    scores, _ = O.gnn_forward(params, fr64, T=T, keep_graph=True)
    torch.set_default_dtype(torch.float32)
    sum(c * scores[b, i] for b, i, c in terms).backward()
    return {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}


sd = load_gnn('random')
for arch, B, T, where in [('base', 1, 2, 'last'), ('base', 1, 2, 'first'), ('base', 1, 1, 'first'), ('deep', 1, 2, 'last'), ('deep', 1, 2, 'first'),
                          ('deep', 1, 1, 'first'), ('deep', 2, 2, 'mixed'), ('base', 3, 2, 'mixed')]:
    fr, _ = load_case(arch, 'fr'); fr = fr.slice(0, B)
    terms = []
    for b in range(B):
        cand = fr.mask[b].nonzero().view(-1).tolist()
        if where == 'last': terms += [(b, cand[-1], 1.0), (b, cand[-2], -1.0)]
        elif where == 'first': terms += [(b, cand[0], 1.0), (b, cand[1], -1.0)]
        else: terms += [(b, cand[0], 1.0), (b, cand[len(cand) // 2], -0.5 - b), (b, cand[-1], 0.25)]
    ref = f64_grads(sd, fr, terms, T)
    cpu, _ = OO.score_grads(sd, fr, terms, T=T)
    model = GraphNet(T, 64); model.load_state_dict(sd); model = model.eval().cuda()
    sc = model.scorer(0); sc.set_network(fr.net, key=fr.net.key)
    sc.score_grad(fr.to('cuda'), terms)
    got = sc.gradients()
    rows = []
    for k, r in ref.items():
        s = float(r.abs().max()) + 1e-30
        rows.append((float((got[k].double() - r).abs().max()) / s, float((cpu[k].double() - r).abs().max()) / s, s, k.split('.')[-2] + '.' + k.split('.')[-1]))
    rows.sort(reverse=True)
    print(arch, 'B', B, 'T', T, where, ' | '.join(f'{n} gpu {g:.1e} cpu {c:.1e} scale {s:.1e}' for g, c, s, n in rows[:4]))

#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference (oval-group/GNN_branching) has no tests or golden vectors for the GNN scoring path
(SURVEY §4), so the pin is the reference module itself executed here on CPU with three shims
(SURVEY §8c): sys.path, ``.cuda()`` as a no-op, ``torch.load(map_location='cpu')``.  Nothing under
/root/reference is copied: the fixtures hold tensors only (verified-net weights, the shipped GNN
checkpoint's tensors, KW root bounds computed with the vendored convex_adversarial, seeded
frontier inputs, and the reference's scores / decisions on them).

Fixtures written:
  nets.npz          verified nets base/wide/deep (real checkpoints), folded property layer, KW root bounds
  gnn_shipped.npz   the 52 tensors of models/cifar_trained_gnn/*.pt
  gnn_random.npz    a non-degenerate GraphNet(2,64): every parameter ~ N(0, 0.15), seed 1234
  case_<arch>.npz   frontier inputs + reference dense scores + reference [layer, idx] decisions
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'convex_adversarial'))
warnings.filterwarnings('ignore')

# shim 2: .cuda() is a no-op on this CPU-only container (graph_conv.py:308-309, graph_score.py:13,26-30)
torch.Tensor.cuda = lambda self, *a, **k: self
nn.Module.cuda = lambda self, *a, **k: self
# shim 3: the checkpoint was saved from CUDA tensors
_orig_load = torch.load
torch.load = lambda f, *a, **k: _orig_load(f, *a, **{**k, 'map_location': 'cpu'})

from graphnet.graph_conv import GraphNet            # noqa: E402  (the reference)
from graphnet.graph_score import GraphChoice        # noqa: E402
from plnn.modules import Flatten as RefFlatten      # noqa: E402
from convex_adversarial import DualNetwork          # noqa: E402
from convex_adversarial.dual_layers import DualReLU  # noqa: E402

from gnn_branching_b200.networks import netspec_from_modules   # noqa: E402
from gnn_branching_b200.frontier import Frontier, synthetic_frontier, net_forward_activations  # noqa: E402

GNN_CKPT = os.path.join(REF, 'models/cifar_trained_gnn/best_snapshot_None_0_val_acc_0.826_loss_val_0.1036_epoch_57.pt')
EPS = 0.145          # base_easy.pkl row 0: Idx=5115, Eps=0.145, prop=8
PROP_CLS = 8

ARCH_LAYERS = {      # exp_utils/model_utils.py:120-166 (shapes only; weights come from the checkpoints)
    'base': lambda: [nn.Conv2d(3, 8, 4, stride=2, padding=1), nn.ReLU(), nn.Conv2d(8, 16, 4, stride=2, padding=1),
                     nn.ReLU(), RefFlatten(), nn.Linear(1024, 100), nn.ReLU(), nn.Linear(100, 10)],
    'wide': lambda: [nn.Conv2d(3, 16, 4, stride=2, padding=1), nn.ReLU(), nn.Conv2d(16, 32, 4, stride=2, padding=1),
                     nn.ReLU(), RefFlatten(), nn.Linear(2048, 100), nn.ReLU(), nn.Linear(100, 10)],
    'deep': lambda: [nn.Conv2d(3, 8, 4, stride=2, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, stride=1, padding=1),
                     nn.ReLU(), nn.Conv2d(8, 8, 3, stride=1, padding=1), nn.ReLU(),
                     nn.Conv2d(8, 8, 4, stride=2, padding=1), nn.ReLU(), RefFlatten(), nn.Linear(512, 100),
                     nn.ReLU(), nn.Linear(100, 10)],
}


def load_verified_net(arch):
    model = nn.Sequential(*ARCH_LAYERS[arch]())
    sd = torch.load(os.path.join(REF, f'models/cifar_{arch}_kw.pth'))['state_dict'][0]
    model.load_state_dict(sd)
    for q in model.parameters():
        q.requires_grad = False
    return model


def fold_property(model, x):
    """add_single_prop (exp_utils/model_utils.py:187-208): Linear(100,10) o Linear(10,1) -> Linear(100,1)."""
    layers = list(model.children())
    y = int(torch.max(model(x)[0], 0)[1])
    cls = PROP_CLS if y != PROP_CLS else (y + 1) % 10
    c = torch.zeros(1, 10)
    c[0, cls], c[0, y] = -1, 1
    last = layers[-1]
    prop = nn.Linear(100, 1)
    with torch.no_grad():
        prop.weight.copy_(c @ last.weight)
        prop.bias.copy_(c @ last.bias)
    for q in prop.parameters():
        q.requires_grad = False
    return layers[:-1], prop


def kw_root_bounds(fixed, prop, x, eps):
    """init_kw_bounds (plnn/dual_network_linear_approximation.py:223-251): pre-ReLU bounds = DualReLU.zl/zu,
    output bounds = dual(+-1)."""
    dual = DualNetwork(nn.Sequential(*fixed, prop), x, eps, bounded_input=False)
    lbs, ubs = [(x - eps).reshape(-1)], [(x + eps).reshape(-1)]
    for layer in dual.dual_net:
        if type(layer) is DualReLU:
            lbs.append(layer.zl.reshape(-1).clone())
            ubs.append(layer.zu.reshape(-1).clone())
    lbs.append(dual(torch.ones(1, 1, 1)).view(-1))
    ubs.append(-dual(-torch.ones(1, 1, 1)).view(-1))
    return lbs, ubs


def run_reference(state_dict, fr: Frontier, ref_fixed, T=2):
    """Reference GraphNet.forward on the whole batch; returns dense scores (0 where mask == 0)."""
    model = GraphNet(T, 64)
    model.load_state_dict(state_dict)
    model.eval()
    lbs, ubs, duals, primals, pin, layers, masks = fr.to_reference_args()
    layers['fixed_layers'] = ref_fixed          # the reference dispatches on plnn.modules.Flatten identity
    with torch.no_grad():
        ragged = model(lbs, ubs, duals, primals, pin, layers, masks)
    dense = torch.zeros_like(fr.mask)
    for b, s in enumerate(ragged):
        dense[b][fr.mask[b].nonzero().view(-1)] = s
    return dense


def run_reference_decisions(ckpt_path, fr: Frontier, ref_fixed):
    """Reference GraphChoice.decision, one domain at a time (its native B=1 usage)."""
    net = fr.net
    decs = []
    for b in range(fr.B):
        one = fr.slice(b, b + 1)
        lbs, ubs, duals, primals, pin, layers, masks = one.to_reference_args()
        layers['fixed_layers'] = ref_fixed
        # GraphChoice wants the BaB mask convention: -1 = undecided (graph_score.py:22)
        init_mask, off = [], 0
        for n in net.hidden_sizes:
            m = masks[0, off:off + n]
            init_mask.append(torch.where(m != 0, torch.full_like(m, -1), torch.ones_like(m)).int())
            off += n
        if int(sum((m == -1).sum() for m in init_mask)) == 0:
            decs.append([-1, -1])
            continue
        gc = GraphChoice(init_mask, ckpt_path)
        primals_lists = [q.tolist() for q in primals]   # decision() receives python-float lists (graph_score.py:30)
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            decs.append(gc.decision(lbs, ubs, duals, pin, primals_lists, layers, init_mask))
    return np.array(decs, dtype=np.int64)


def frontier_to_npz(fr: Frontier, prefix=''):
    out = {}
    for name, v in fr.tensors().items():
        if isinstance(v, list):
            for i, t in enumerate(v):
                out[f'{prefix}{name}_{i}'] = t.numpy()
        else:
            out[f'{prefix}{name}'] = v.numpy()
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    shipped = torch.load(GNN_CKPT)
    np.savez_compressed(os.path.join(HERE, 'gnn_shipped.npz'), **{k: v.numpy() for k, v in shipped.items()})
    g = torch.Generator().manual_seed(1234)
    random_sd = {k: torch.randn(v.shape, generator=g) * 0.15 for k, v in shipped.items()}
    np.savez_compressed(os.path.join(HERE, 'gnn_random.npz'), **{k: v.numpy() for k, v in random_sd.items()})
    tmpdir = tempfile.mkdtemp()
    random_path = os.path.join(tmpdir, 'gnn_random.pt')
    torch.save(random_sd, random_path)

    nets = {}
    for ai, arch in enumerate(['base', 'wide', 'deep']):
        model = load_verified_net(arch)
        x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(0))
        fixed, prop = fold_property(model, x)
        spec = netspec_from_modules(fixed, (3, 32, 32), name=f'cifar_{arch}_kw')
        lbs, ubs = kw_root_bounds(fixed, prop, x, EPS)
        nets[f'{arch}_x'] = x.numpy()
        nets[f'{arch}_wp'] = prop.weight.reshape(-1).numpy()
        nets[f'{arch}_bp'] = prop.bias.reshape(()).numpy()
        for k, a in enumerate(spec.affine):
            nets[f'{arch}_w{k}'] = a.weight.numpy()
            nets[f'{arch}_b{k}'] = a.bias.numpy()
        for k, (l, u) in enumerate(zip(lbs, ubs)):
            nets[f'{arch}_lb{k}'] = l.numpy()
            nets[f'{arch}_ub{k}'] = u.numpy()
        namb = [int(((l < 0) & (u > 0)).sum()) for l, u in zip(lbs[1:-1], ubs[1:-1])]
        print(arch, 'hidden', spec.hidden_sizes, 'root ambiguous', namb, 'out bounds', float(lbs[-1]), float(ubs[-1]))

        # --- cases: synthetic frontier domains (configs 2-4 style) + the root domain (config 1 style) ---
        B = 3 if arch == 'base' else 2
        fr = synthetic_frontier(spec, lbs, ubs, prop.weight.reshape(-1), float(prop.bias), B, seed=1000 * (ai + 2))
        root = synthetic_frontier(spec, lbs, ubs, prop.weight.reshape(-1), float(prop.bias), 2, seed=7 + ai,
                                  max_splits=1)
        # root domain: untouched KW bounds, variant 0 zero duals, variant 1 sparse duals
        for k in range(len(lbs)):
            root.lb[k] = lbs[k].reshape(1, -1).repeat(2, 1)
            root.ub[k] = ubs[k].reshape(1, -1).repeat(2, 1)
        root.mask = torch.cat([((l < 0) & (u > 0)).float() for l, u in zip(root.lb[1:-1], root.ub[1:-1])], 1)
        for k in range(spec.L):
            amb = ((root.lb[k + 1] < 0) & (root.ub[k + 1] > 0)).float().unsqueeze(-1)
            root.dual[k] = root.dual[k] * amb
            root.dual[k][0] = 0
        # give the two property layers of the root pair different weights to exercise the per-domain path
        root.Wp[1] = root.Wp[1] * 0.5
        root.bp[1] = root.bp[1] + 0.25
        out = {}
        for name, f in (('fr', fr), ('root', root)):
            out.update(frontier_to_npz(f, prefix=name + '_'))
            for wname, sd, path in (('shipped', shipped, GNN_CKPT), ('random', random_sd, random_path)):
                dense = run_reference(sd, f, fixed)
                out[f'{name}_scores_{wname}'] = dense.numpy()
                out[f'{name}_decisions_{wname}'] = run_reference_decisions(path, f, fixed)
                print(f'  {arch} {name} {wname}: |s|max {float(dense.abs().max()):.4f}  decisions '
                      f'{out[f"{name}_decisions_{wname}"].tolist()}')
            if arch == 'base' and name == 'fr':
                for T in (1, 3):
                    out[f'{name}_scores_random_T{T}'] = run_reference(random_sd, f, fixed, T=T).numpy()
        np.savez_compressed(os.path.join(HERE, f'case_{arch}.npz'), **out)
    np.savez_compressed(os.path.join(HERE, 'nets.npz'), **nets)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KB')


if __name__ == '__main__':
    main()

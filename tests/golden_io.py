"""Loaders for the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from gnn_branching_b200.networks import AffineSpec, NetSpec
from gnn_branching_b200.frontier import Frontier

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ARCHS = ('base', 'wide', 'deep')

_ARCH_META = {   # (kind, stride, padding) per affine layer and the layer index i_k in fixed_layers
    'base': [('conv', 2, 1, 0), ('conv', 2, 1, 2), ('linear', 1, 0, 5)],
    'wide': [('conv', 2, 1, 0), ('conv', 2, 1, 2), ('linear', 1, 0, 5)],
    'deep': [('conv', 2, 1, 0), ('conv', 1, 1, 2), ('conv', 1, 1, 4), ('conv', 2, 1, 6), ('linear', 1, 0, 9)],
}
_N_LAYERS = {'base': 7, 'wide': 7, 'deep': 11}

_cache = {}


def _npz(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, name)))
    return _cache[name]


def load_gnn(which: str):
    """'shipped' (models/cifar_trained_gnn/*.pt tensors) or 'random' (N(0, 0.15), non-degenerate)."""
    return {k: torch.from_numpy(v.copy()) for k, v in _npz(f'gnn_{which}.npz').items()}


def load_net(arch: str) -> NetSpec:
    z = _npz('nets.npz')
    spec = NetSpec(name=f'cifar_{arch}_kw', input_shape=(3, 32, 32), n_layers_total=_N_LAYERS[arch])
    shape = (3, 32, 32)
    for k, (kind, stride, pad, idx) in enumerate(_ARCH_META[arch]):
        w = torch.from_numpy(z[f'{arch}_w{k}'].copy())
        b = torch.from_numpy(z[f'{arch}_b{k}'].copy())
        if kind == 'conv':
            co, ci, kh, kw = w.shape
            out = (co, (shape[1] + 2 * pad - kh) // stride + 1, (shape[2] + 2 * pad - kw) // stride + 1)
            spec.affine.append(AffineSpec('conv', w, b, shape, out, stride, pad, idx))
        else:
            n = int(np.prod(shape))
            out = (w.shape[0],)
            spec.affine.append(AffineSpec('linear', w, b, (n,), out, 1, 0, idx))
        shape = out
    return spec


def load_root(arch: str):
    """KW root bounds (convex_adversarial, eps = 0.145) + folded property layer of the real checkpoints."""
    z = _npz('nets.npz')
    net = load_net(arch)
    lbs = [torch.from_numpy(z[f'{arch}_lb{k}'].copy()) for k in range(net.L + 2)]
    ubs = [torch.from_numpy(z[f'{arch}_ub{k}'].copy()) for k in range(net.L + 2)]
    return net, lbs, ubs, torch.from_numpy(z[f'{arch}_wp'].copy()), float(z[f'{arch}_bp'])


def load_case(arch: str, name: str):
    """-> (Frontier, {'scores_shipped', 'scores_random', 'decisions_*', ...}) for name in {'fr', 'root'}."""
    z = _npz(f'case_{arch}.npz')
    net = load_net(arch)
    L = net.L
    t = lambda key: torch.from_numpy(z[f'{name}_{key}'].copy())
    fr = Frontier(net=net,
                  lb=[t(f'lb_{k}') for k in range(L + 2)], ub=[t(f'ub_{k}') for k in range(L + 2)],
                  dual=[t(f'dual_{k}') for k in range(L)],
                  prim_pre=[t(f'prim_pre_{k}') for k in range(L)], prim_post=[t(f'prim_post_{k}') for k in range(L)],
                  prim_out=t('prim_out'), primal_input=t('primal_input'), Wp=t('Wp'), bp=t('bp'), mask=t('mask'))
    ref = {k[len(name) + 1:]: torch.from_numpy(v.copy()) for k, v in z.items()
           if k.startswith(name + '_scores') or k.startswith(name + '_decisions')}
    return fr, ref


def load_online():
    """tests/golden/online_base.npz (make_golden_online.py): the reference's online fine-tuning steps on the base case."""
    return _npz('online_base.npz')

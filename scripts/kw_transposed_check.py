"""CPU emulation of the transposed recursion that csrc/gnnb_kw.cu implements (same steps, same order of operations per layer),
checked against oracle/kw_bounds_oracle.py and the reference's golden bounds.  It validates the algorithm and the indexing of
the CUDA code's design on a machine without a GPU; it is not a substitute for running the kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, torch.nn.functional as F
from golden_io import GOLDEN, load_root
from oracle import kw_bounds_oracle as KW


def At(a, cols):            # cols [m, n_out] -> A^T cols [m, n_in]   (prop_backward without normalisation)
    if a.kind == 'conv':
        y = F.conv_transpose2d(cols.reshape(cols.shape[0], *a.out_shape), a.weight, None, stride=a.stride, padding=a.padding)
        return y.reshape(cols.shape[0], -1)
    return cols @ a.weight


def bias_node(a):
    return KW._bias_row(a).reshape(-1)


def kw_transposed(net, x, eps, wp, bp, plb=None, pub=None):
    L = net.L
    n = [net.n0] + net.hidden_sizes + [1]
    lbs, ubs = [x - eps], [x + eps]
    for k in range(1, L + 2):
        out_layer = k == L + 1
        ncols = 1 if out_layer else n[k]
        bias = torch.zeros(ncols); low = torch.zeros(ncols); up = torch.zeros(ncols)
        if out_layer:
            t = wp.reshape(1, -1).clone(); have_t = True
        else:
            t = torch.eye(n[k]); have_t = False          # s_k (rows = columns of the CUDA code)
        for j in range(k - 1, 0, -1):
            if not have_t:
                t = At(net.affine[j], t)
            have_t = False
            l, u = lbs[j], ubs[j]
            I = (u > 0) & (l < 0)
            d = (l >= 0).float()
            d[I] = d[I] + u[I] / (u[I] - l[I])
            t = t * d
            bias += t @ bias_node(net.affine[j - 1])
            low += ((-t).clamp(min=0) * (l * I.float())).sum(1)
            up += (t.clamp(min=0) * (l * I.float())).sum(1)
        t0 = At(net.affine[0], t)
        cx, l1 = t0 @ x, t0.abs().sum(1)
        own = torch.tensor([float(bp)]) if out_layer else bias_node(net.affine[k - 1])
        centre = cx + (bias + own)
        zl, zu = centre - eps * l1 + low, centre + eps * l1 - up
        if not out_layer and plb is not None:
            zl, zu = torch.max(zl, plb[k - 1]), torch.min(zu, pub[k - 1])
        lbs.append(zl); ubs.append(zu)
    return lbs, ubs


worst = 0.0
for arch in ('base', 'deep', 'wide'):
    net, lbs, ubs, wp, bp = load_root(arch)
    x = torch.from_numpy(np.load(os.path.join(GOLDEN, 'nets.npz'))[f'{arch}_x'].copy()).reshape(-1)
    gl, gu = kw_transposed(net, x, 0.145, wp, bp)
    for k in range(net.L + 2):
        e = max(float((gl[k] - lbs[k]).abs().max()), float((gu[k] - ubs[k]).abs().max())) / max(1.0, float(ubs[k].abs().max()))
        worst = max(worst, e)
    print(arch, 'root: worst normalised error vs the reference', worst)
    if arch == 'wide':
        continue
    z = dict(np.load(os.path.join(GOLDEN, 'kw_children.npz')))
    for c in range(int(z[f'{arch}_ncases'])):
        lay, idx, choice = z[f'{arch}_c{c}_decision'].tolist()
        plb, pub = KW.split_bounds(lbs, ubs, (lay, idx), choice)
        gl, gu = kw_transposed(net, x, 0.145, wp, bp, plb, pub)
        for k in range(1, net.L + 1):
            rl, ru = torch.from_numpy(z[f'{arch}_c{c}_lb{k}']), torch.from_numpy(z[f'{arch}_c{c}_ub{k}'])
            worst = max(worst, float((gl[k] - rl).abs().max()) / max(1.0, float(rl.abs().max())), float((gu[k] - ru).abs().max()) / max(1.0, float(ru.abs().max())))
    print(arch, 'children: worst normalised error vs the reference', worst)
assert worst <= 5e-5, worst
print('ok')

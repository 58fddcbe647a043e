"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/gnnb.h
declares, the GraphNet mirror keeps the reference's state_dict keys, the frontier pack round-trips the reference
argument lists, and the product fails loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

from golden_io import ARCHS, load_case, load_gnn, load_net, load_root
from gnn_branching_b200 import GraphNet, Frontier, Scorer, STATE_DICT_KEYS, cifar_netspec, synthetic_frontier, _lib
from gnn_branching_b200.networks import netspec_from_modules
from gnn_branching_b200.engine import flat_to_layer_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'gnnb.h')).read()
    declared = set(re.findall(r'\b(gnnb_[a-z_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gnnb_abi_version() == 1


def test_state_dict_keys_match_the_shipped_checkpoint():
    model = GraphNet(2, 64)
    sd = load_gnn('shipped')
    assert list(sd.keys()) == STATE_DICT_KEYS
    assert list(model.state_dict().keys()) == STATE_DICT_KEYS
    res = model.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in model.state_dict().items():
        assert v.shape == sd[k].shape
    assert sum(v.numel() for v in sd.values()) == 117825


def test_other_embedding_sizes_are_rejected():
    with pytest.raises(NotImplementedError):
        GraphNet(2, 32)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        Scorer(0)
    fr, _ = load_case('base', 'fr')
    model = GraphNet(2, 64)
    with pytest.raises(RuntimeError):
        model(*fr.to_reference_args())


@pytest.mark.parametrize('arch', ARCHS)
def test_frontier_round_trips_reference_arguments(arch):
    fr, _ = load_case(arch, 'fr')
    args = fr.to_reference_args()
    lbs, ubs, duals, primals, pin, layers, masks = args
    net = fr.net
    assert len(lbs) == net.L + 2 and lbs[0].shape == (fr.B, 3, 32, 32) and lbs[-1].shape == (fr.B, 1)
    assert [p.numel() // fr.B for p in primals] == net.primal_sizes()
    assert duals[0].shape == (fr.B * net.hidden_sizes[0], 3)
    back = Frontier.from_reference_args(*args)
    assert back.net.hidden_sizes == net.hidden_sizes
    for a, b in zip(back.net.affine, net.affine):
        assert a.kind == b.kind and a.layer_index == b.layer_index and a.stride == b.stride and a.padding == b.padding
        assert torch.equal(a.weight, b.weight)
    for name, v in fr.tensors().items():
        w = back.tensors()[name]
        if isinstance(v, list):
            assert all(torch.equal(x, y) for x, y in zip(v, w)), name
        else:
            assert torch.equal(v, w), name


def test_netspec_shapes_and_flops():
    # SURVEY §8 table and §8(d) contract figures
    want = {'base': ([2048, 1024, 100], 1.3104e9), 'wide': ([4096, 2048, 100], 2.5947e9),
            'deep': ([2048, 2048, 2048, 512, 100], 2.5599e9)}
    for arch, (sizes, flops) in want.items():
        net = cifar_netspec(arch)
        assert net.hidden_sizes == sizes
        assert abs(net.flops_per_domain() - flops) / flops < 1e-3
        assert load_net(arch).hidden_sizes == sizes
    with pytest.raises(NotImplementedError):
        netspec_from_modules([torch.nn.Conv2d(3, 8, 4), torch.nn.Sigmoid()], (3, 32, 32))


def test_flat_index_mapping():
    sizes = [2048, 1024, 100]
    assert flat_to_layer_index(0, sizes) == [0, 0]
    assert flat_to_layer_index(2047, sizes) == [0, 2047]
    assert flat_to_layer_index(2048, sizes) == [1, 0]
    assert flat_to_layer_index(3171, sizes) == [2, 99]
    with pytest.raises(IndexError):
        flat_to_layer_index(3172, sizes)


def test_synthetic_frontier_is_well_formed():
    net, lbs, ubs, wp, bp = load_root('base')
    fr = synthetic_frontier(net, lbs, ubs, wp, bp, 16, seed=3)
    assert fr.B == 16 and fr.mask.shape == (16, 3172)
    for k in range(1, net.L + 1):
        l, u = fr.lb[k], fr.ub[k]
        assert bool((l <= u).all())
        assert not bool(((l == 0) & (u == 0)).any())          # never 0/0 in compute_ratio
    amb = torch.cat([((fr.lb[k] < 0) & (fr.ub[k] > 0)).float() for k in range(1, net.L + 1)], 1)
    assert torch.equal(amb, fr.mask)
    assert bool((fr.mask.sum(1) > 0).all())
    again = synthetic_frontier(net, lbs, ubs, wp, bp, 16, seed=3)
    assert torch.equal(again.lb[1], fr.lb[1]) and torch.equal(again.dual[0], fr.dual[0])
    assert fr.input_bytes() // 16 == 4 * (2 * (3072 + 3172 + 1) + 3 * 3172 + 2 * 3172 + 1 + 3072 + 100 + 1 + 3172)


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of gnnb_layer_desc / gnnb_frontier / gnnb_domains against the C compiler's view of include/gnnb.h
    (sizes and field offsets), and the header compiles as plain C."""
    import ctypes as C
    import subprocess
    src = os.path.join(tmp_path, 'layout.c')
    fields = {'gnnb_layer_desc': [f[0] for f in _lib.LayerDesc._fields_], 'gnnb_frontier': [f[0] for f in _lib.FrontierDesc._fields_],
              'gnnb_domains': [f[0] for f in _lib.DomainsDesc._fields_]}
    with open(src, 'w') as f:
        f.write('#include <stdio.h>\n#include <stddef.h>\n#include "gnnb.h"\nint main(void) {\n')
        for s, names in fields.items():
            f.write(f'  printf("{s} %zu", sizeof({s}));\n')
            for n in names:
                f.write(f'  printf(" %zu", offsetof({s}, {n}));\n')
            f.write('  printf("\\n");\n')
        f.write('  return 0;\n}\n')
    exe = os.path.join(tmp_path, 'layout')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), src, '-o', exe])
    out = subprocess.check_output([exe], text=True).strip().splitlines()
    mirrors = {'gnnb_layer_desc': _lib.LayerDesc, 'gnnb_frontier': _lib.FrontierDesc, 'gnnb_domains': _lib.DomainsDesc}
    for line in out:
        parts = line.split()
        cls = mirrors[parts[0]]
        assert int(parts[1]) == C.sizeof(cls), parts[0]
        assert [int(x) for x in parts[2:]] == [getattr(cls, n).offset for n, _ in cls._fields_], parts[0]


def test_plain_c_client_links_and_fails_loudly_without_a_gpu(tmp_path):
    """tests/c/abi_client.c built with gcc against include/gnnb.h and libgnnb.so: the ABI is plain C, and without a B200
    gnnb_create returns a status code and no context (with one, the context works)."""
    import subprocess
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    exe = os.path.join(tmp_path, 'abi_client')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'c', 'abi_client.c'),
                           '-L', lib_dir, '-lgnnb', '-Wl,-rpath,' + lib_dir, '-o', exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith('abi 1 create ')

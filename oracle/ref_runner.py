#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE — runs the UNMODIFIED reference `graphnet.graph_conv.GraphNet.forward` staged under
oracle/_ref/ (git-ignored; `scripts/stage_ref.sh` copies graphnet/graph_conv.py, graph_score.py, graph_score_online.py and
plnn/modules.py there from /root/reference, nothing is edited).  Only tests/, __graft_entry__.smoke() and bench.py's
baseline legs may use it; the product path (gnn_branching_b200/) never imports it.

As a library:   ref = load(device);  ref.scores(state_dict, frontier)           -> dense [B, sum n_k] scores
As a script:    python oracle/ref_runner.py --device cpu|cuda --workload base --weights random --batches 1,8,32 --reps 3
                prints one JSON object with the B = 1 latency and the rate of every batch size (SURVEY §8d "CPU baseline").

Shims, applied to torch, never to the reference files:
  * CPU runs: `.cuda()` is a no-op (graph_conv.py:308-309 and graph_score.py:13,26-30 hard-code it) and the process must
    not see a GPU (the script sets CUDA_VISIBLE_DEVICES='' before importing torch when --device cpu);
  * the reference dispatches on the identity of its own `plnn.modules.Flatten` (graph_conv.py:188, 355), so the
    `fixed_layers` list is rebuilt with that class.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, 'graphnet', 'graph_conv.py')) and os.path.isfile(os.path.join(REF_DIR, 'plnn', 'modules.py'))


class _Ref:
    def __init__(self, device):
        import warnings
        import torch
        from torch import nn
        warnings.filterwarnings('ignore', category=SyntaxWarning)          # `next_layer is 'Linear'` in the reference
        if not available():
            raise FileNotFoundError('oracle/_ref is not staged (scripts/stage_ref.sh needs /root/reference)')
        self.device = torch.device(device)
        if self.device.type == 'cpu':
            if torch.cuda.is_available():
                raise RuntimeError('CPU runs of the reference need a process without a visible GPU (CUDA_VISIBLE_DEVICES="")')
            torch.Tensor.cuda = lambda s, *a, **k: s
            nn.Module.cuda = lambda s, *a, **k: s
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from graphnet.graph_conv import GraphNet                            # the reference
        from plnn.modules import Flatten
        self.GraphNet, self.Flatten, self.torch = GraphNet, Flatten, torch

    def model(self, state_dict, T=2, p=64):
        m = self.GraphNet(T, p)
        m.load_state_dict(state_dict)
        return m.eval().to(self.device)

    def args(self, fr):
        """Frontier -> the reference's argument lists, with its own Flatten class in fixed_layers."""
        lbs, ubs, duals, primals, pin, layers, masks = fr.to(self.device).to_reference_args()
        layers['fixed_layers'] = [self.Flatten() if type(m).__name__ == 'Flatten' else m for m in layers['fixed_layers']]
        return lbs, ubs, duals, primals, pin, layers, masks

    def forward(self, model, args):
        with self.torch.no_grad():
            return model(*args)

    def scores(self, state_dict, fr, T=2):
        """Dense [B, sum n_k] scores (0 where mask == 0), like tests/golden/make_golden.py:run_reference."""
        torch = self.torch
        args = self.args(fr)
        ragged = self.forward(self.model(state_dict, T), args)
        mask = args[-1]
        dense = torch.zeros_like(mask)
        for b, s in enumerate(ragged):
            dense[b][mask[b].nonzero().view(-1)] = s
        return dense


def load(device='cpu'):
    return _Ref(device)


def _main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--device', default='cpu', choices=['cpu', 'cuda'])
    ap.add_argument('--workload', default='base')
    ap.add_argument('--weights', default='random')
    ap.add_argument('--batches', default='1,8,32')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--seed', type=int, default=99)
    args = ap.parse_args()
    if args.device == 'cpu':
        os.environ['CUDA_VISIBLE_DEVICES'] = ''
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    from golden_io import load_gnn, load_root
    from gnn_branching_b200.frontier import synthetic_frontier
    threads = args.threads or os.cpu_count()
    torch.set_num_threads(threads)
    ref = load(args.device)
    net, lbs, ubs, wp, bp = load_root(args.workload)
    batches = [int(b) for b in args.batches.split(',')]
    fr = synthetic_frontier(net, lbs, ubs, wp, bp, max(batches), seed=args.seed)
    model = ref.model(load_gnn(args.weights))

    def sync():
        if args.device == 'cuda':
            torch.cuda.synchronize()

    out = {'device': args.device, 'workload': args.workload, 'cores': threads if args.device == 'cpu' else None, 'per_batch': {},
           'kind': 'reference', 'what': 'unmodified graphnet/graph_conv.py GraphNet.forward from oracle/_ref (torch %s)' % torch.__version__}
    for B in batches:
        a = ref.args(fr.slice(0, B))
        for _ in range(args.warmup):
            ref.forward(model, a)
        sync()
        times = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            ref.forward(model, a)
            sync()
            times.append(time.perf_counter() - t0)
        out['per_batch'][str(B)] = {'best_s': min(times), 'mean_s': sum(times) / len(times), 'rate': B / min(times), 'times_s': times}
    out['b1_latency_ms'] = out['per_batch'].get('1', {}).get('best_s', float('nan')) * 1e3
    best_B = max(out['per_batch'], key=lambda k: out['per_batch'][k]['rate'])
    out['best_batch'], out['value'] = int(best_B), out['per_batch'][best_B]['rate']
    print(json.dumps(out))


if __name__ == '__main__':
    _main()

"""Time one online fine-tuning step (gnnb_score_grad + gnnb_adam_step) per CIFAR net, B = 1, next to the reference
algorithm (oracle port: autograd + torch.optim.Adam) on the host cores.  Usage: python scripts/online_probe.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from golden_io import load_case, load_gnn
from gnn_branching_b200 import GraphNet
from oracle.online_oracle import OnlineOracle

out = {}
for arch in ('base', 'wide', 'deep'):
    fr, _ = load_case(arch, 'fr')
    one = fr.slice(0, 1)
    cand = one.mask[0].nonzero().view(-1).tolist()
    model = GraphNet(2, 64); model.load_state_dict(load_gnn('random')); model = model.eval().cuda()
    sc = model.scorer(0); sc.set_network(fr.net, key=fr.net.key)
    dev = one.to('cuda')
    terms = [(0, cand[0], 1.0), (0, cand[-1], -1.0)]
    for _ in range(3):
        sc.score_grad(dev, terms); sc.adam_step(1e-4, weight_decay=1e-4)
    torch.cuda.synchronize(); l0 = sc.launches; t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        sc.score_grad(dev, terms); sc.adam_step(1e-4, weight_decay=1e-4)
    torch.cuda.synchronize(); gpu_ms = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        sc.score_grad(dev, terms)
    torch.cuda.synchronize(); grad_ms = (time.perf_counter() - t0) / n * 1e3
    oo = OnlineOracle(load_gnn('random'))
    oo.online_learning(one, cand[0], [0, 0], 1.0)
    t0 = time.perf_counter()
    for _ in range(3):
        oo.online_learning(one, cand[0], [0, 0], 1.0)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    out[arch] = dict(gpu_step_ms=round(gpu_ms, 3), gpu_grad_only_ms=round(grad_ms, 3), launches_per_step=(sc.launches - l0) // (2 * n) ,
                     cpu_reference_algorithm_ms=round(cpu_ms, 1), cpu_threads=torch.get_num_threads())
print(json.dumps(out))

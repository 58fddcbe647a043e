import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from golden_io import load_case, load_gnn
from gnn_branching_b200 import GraphNet
fr, ref = load_case('base', 'root')
m = GraphNet(2, 64, math='simt'); m.load_state_dict(load_gnn('shipped')); m = m.eval().cuda()
which = sys.argv[1]
if which == 'a':
    print(m.score_frontier(fr.slice(0, 1).to('cuda'), return_scores=True)[1])
elif which == 'b':
    print(m.score_frontier(fr.slice(0, 1).to('cuda'), return_scores=False)[1])
elif which == 'c':
    print(m.score_frontier(fr.to('cuda'), return_scores=False)[1])
torch.cuda.synchronize()
print('ok', which)

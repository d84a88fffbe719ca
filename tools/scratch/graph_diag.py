import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as g
pkg = g.build()
sys.path.insert(0, 'tests')
from test_gpu_graph import _net, DEV
gen = torch.Generator().manual_seed(3)
xs = [torch.rand(32, 3, 16, 16, generator=gen).to(DEV) for _ in range(7)]
ys = [torch.randint(0, 4, (32,), generator=gen).to(DEV) for _ in range(7)]
def run(mode):
    net = _net(pkg); net.train()
    for c in pkg.engine._rng_counters.values(): c.zero_()
    losses, cur = [], {}
    for i, (x, y) in enumerate(zip(xs, ys)):
        ls, cur = net.train_step(x, y, batch=i, current_measures=cur, graph=(mode == 'graph'))
        losses.append(ls['total'].detach().clone())
    return losses
a = run('eager'); b = run('eager'); c = run('graph'); d = run('graph')
for i in range(7):
    print(i, 'ee', float((a[i]-b[i]).abs().max()), 'eg', float((a[i]-c[i]).abs().max()), 'gg', float((c[i]-d[i]).abs().max()))

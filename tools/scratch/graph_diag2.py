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
def run():
    net = _net(pkg); net.train()
    for c in pkg.engine._rng_counters.values(): c.zero_()
    snaps, cur = [], {}
    for i, (x, y) in enumerate(zip(xs, ys)):
        ls, cur = net.train_step(x, y, batch=i, current_measures=cur)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        gr = {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in net.named_parameters()}
        snaps.append((ls['total'].detach().clone(), sd, gr))
    return snaps
a = run(); b = run()
for i in range(7):
    la, sa, ga = a[i]; lb, sb, gb = b[i]
    print('step', i, 'loss diff', float((la - lb).abs().max()))
    for k in ga:
        if ga[k] is None: continue
        d = float((ga[k] - gb[k]).abs().max())
        if d > 0: print('   grad', k, tuple(ga[k].shape), 'maxdiff %.3e' % d, 'max %.3e' % float(ga[k].abs().max()))
    for k in sa:
        d = float((sa[k].float() - sb[k].float()).abs().max())
        if d > 0: print('   state', k, 'maxdiff %.3e' % d, 'max %.3e' % float(sa[k].float().abs().max()))
    if i >= 5: break

import sys, os, time, cProfile, pstats, io
sys.path.insert(0, '.')
import torch
import __graft_entry__ as g
import bench
pkg = g.build()
dev = torch.device('cuda:0')
for wname in ('c2', 'c4'):
    wl = bench.WORKLOADS[wname]
    torch.manual_seed(0)
    net = pkg.ClassificationVariationalNetwork(**bench.make_ctor(wl)).to(dev)
    net.train()
    B = wl['batch']
    x = torch.rand(B, *wl['ctor']['input_shape'], device=dev); y = torch.randint(0, wl['ctor']['num_labels'], (B,), device=dev)
    for i in range(5): net.train_step(x, y)
    torch.cuda.synchronize()
    # pure CPU launch time of a step: time the python call with the GPU queue empty at start, no sync inside
    ts = []
    for i in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); net.train_step(x, y); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    print(wname, 'cpu launch ms', sorted(t[0] for t in ts)[5] * 1e3, 'step wall ms (sync each step)', sorted(t[1] for t in ts)[5] * 1e3)
    pr = cProfile.Profile(); pr.enable()
    for i in range(10): net.train_step(x, y)
    torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])

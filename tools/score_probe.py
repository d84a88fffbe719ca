import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as g
pkg = g.build()
wl = bench.WORKLOADS['c2']
torch.manual_seed(0)
net = pkg.ClassificationVariationalNetwork(**bench.make_ctor(wl)).to('cuda:0')
x = torch.rand(512, 3, 32, 32, device='cuda:0')
y = torch.randint(0, 10, (512,), device='cuda:0')
net.train()
for i in range(2): net.train_step(x, y)
net.eval()
with torch.no_grad():
    for i in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, logits, losses, _ = net.evaluate(x)
        t1 = time.perf_counter()
        net.batch_dist_measures(logits, losses, list(net.ood_methods))
        net.predict_after_evaluate(logits, losses, method=net.predict_methods[0])
        e1.record(); t2 = time.perf_counter()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f'iter {i}: cpu evaluate {1e3*(t1-t0):.2f} ms, cpu scores {1e3*(t2-t1):.2f} ms, gpu span {e0.elapsed_time(e1):.2f} ms, wall {1e3*(t3-t0):.2f} ms', flush=True)

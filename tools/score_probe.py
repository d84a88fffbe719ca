import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as g
pkg = g.build()
wl = bench.WORKLOADS['c2']
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kw = bench.make_ctor(wl); kw.update(num_labels=C, latent_dim=K)
torch.manual_seed(0)
net = pkg.ClassificationVariationalNetwork(**kw).to('cuda:0')
x = torch.rand(512, 3, 32, 32, device='cuda:0')
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
        print(f'C={C} iter {i}: cpu evaluate {1e3*(t1-t0):.2f} ms, cpu scores {1e3*(t2-t1):.2f} ms, gpu span {e0.elapsed_time(e1):.2f} ms, wall {1e3*(t3-t0):.2f} ms', flush=True)

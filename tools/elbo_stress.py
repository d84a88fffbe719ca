import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.build(); nat = pkg._native
dev = 'cuda:0'
ready = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for (B, L, K, C, D) in [(512, 16, 128, 10, 3072), (2048, 16, 128, 10, 3072)]:
    x = torch.rand(B, D, device=dev); xr = torch.rand(L + 1, B, D, device=dev).bfloat16()
    mu, lv = torch.randn(B, K, device=dev), torch.randn(B, K, device=dev) * 0.1
    y = torch.randint(0, C, (B,), device=dev)
    means, T = torch.randn(C, K, device=dev), torch.ones(C, device=dev)
    sig = torch.zeros(1, device=dev)
    cfg = nat.make_cfg(B=B, L=L, K=K, C=C, D=D, x_reco=xr, logits=None, var_dim='scalar', prior_kind='gaussian',
                       conditional=True, sigma_is_log=True, sigma_is_rmse=False, beta=1.0, gamma_w=0.0, var_w=1.0)
    if ready:
        nat.elbo_prior_stats(cfg, means, T); cfg.prior_stats_ready = 1
    ref = None
    for i in range(200):
        out = nat.elbo_train_fwd(cfg, x, xr, mu, lv, None, y, means, T, sig)
        if i % 50 == 0:
            torch.cuda.synchronize()
            t = out['total'].clone()
            if ref is None: ref = t
            print(B, i, 'ok', float(t.sum()), bool(torch.equal(t, ref)), flush=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): nat.elbo_train_fwd(cfg, x, xr, mu, lv, None, y, means, T, sig)
    e1.record(); torch.cuda.synchronize()
    print(B, 'us per launch', e0.elapsed_time(e1) * 1e3 / 50, flush=True)

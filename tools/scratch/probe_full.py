import sys, os, json, time
sys.path.insert(0, '.'); sys.path.insert(0, './tests')
import numpy as np, torch
import __graft_entry__ as g
pkg = g.build()
import full_cases as fc
DEV = 'cuda:0'
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-6, np.abs(b).max()))
for name in (sys.argv[1:] or list(fc.CASES)):
    d = np.load(f'tests/golden/{name}.npz')
    torch.manual_seed(0)
    net = pkg.ClassificationVariationalNetwork(**fc.ctor_kwargs(name))
    fc.fill_state_(net, chaotic=name in fc.CHAOTIC)
    net = net.to(DEV)
    x, y, eps_tr, eps_te = [t.to(DEV) for t in fc.inputs(name)]
    net.eval()
    net.encoder.sampling.injected_eps = eps_te
    with torch.no_grad():
        xr, logits, losses, _, mu, lv, z = net.evaluate(x, z_output=True)
        print(name, 'eval', {k: round(rel(v.cpu().numpy(), d['eval.loss.' + k]), 4) for k, v in losses.items()},
              'logits', round(rel(logits.float().cpu().numpy(), d['eval.logits']), 4), 'mu', round(rel(mu.cpu().numpy(), d['eval.mu']), 4),
              'xr', round(rel(xr[:2, :2].float().cpu().numpy(), d['eval.x_reco2']), 4))
        methods = json.loads(str(d['eval.methods']))
        dm = net.batch_dist_measures(logits, losses, methods)
        for m in methods:
            got, want = dm[m].float().cpu().numpy(), d['eval.measure.' + m]
            ag, frac = fc.rank_agreement(got, want, 2e-2)
            print('   score', m, 'rel', round(rel(got, want), 4), 'rank agree', ag, 'pairs kept', round(frac, 3))
        for m in json.loads(str(d['eval.predict_methods'])):
            got = net.predict_after_evaluate(logits, losses, method=m).cpu().numpy()
            print('   pred', m, (got == d['eval.pred.' + m]).mean())
    net.train()
    net.encoder.sampling.injected_eps = eps_tr
    net.optimizer.zero_grad()
    xr, logits, losses, meas, mu, lv, z = net.evaluate(x, y, with_beta=True, z_output=True)
    print(name, 'train', {k: round(rel(v.detach().cpu().numpy(), d['train.loss.' + k]), 4) for k, v in losses.items()},
          'mu', round(rel(mu.detach().cpu().numpy(), d['train.mu']), 4), 'xr', round(rel(xr[:2, :2].detach().float().cpu().numpy(), d['train.x_reco2']), 4))
    losses['total'].mean().backward()
    t0 = time.time()
    o = fc.oracle_outputs(pkg, name)
    print('   oracle cpu s', round(time.time() - t0, 1))
    errs = []
    for k, p in net.named_parameters():
        if p.grad is None or 'train.gnorm.' + k not in d.files: continue
        gr = o['train']['grads'][k].astype(np.float64)
        gg = p.grad.detach().float().cpu().numpy().astype(np.float64)
        nr = np.linalg.norm(gr)
        e_or = np.linalg.norm(gg - gr) / max(nr, 1e-30)
        e_pr = fc.projected_error(k, gg, float(d['train.gnorm.' + k]), d['train.gproj.' + k])
        errs.append((round(float(e_or), 4), round(e_pr, 4), float('%.3g' % nr), k))
    errs.sort()
    print('   grad errs (vs oracle, projected vs reference, norm): worst', errs[-12:])
    print('   median', errs[len(errs) // 2], 'n', len(errs), 'n>2e-2', sum(e[0] > 2e-2 for e in errs), 'n>5e-2', sum(e[0] > 5e-2 for e in errs))

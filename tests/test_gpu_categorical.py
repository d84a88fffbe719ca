"""output_distribution='categorical' (256-way per-pixel cross-entropy head; losses.py:30-49, cvae.py:654-660, 683, 776).
Kernel level: the categorical pre-pass + fused ELBO against torch fp32 (F.cross_entropy / autograd) in both logit layouts
and both dtypes.  Model level: golden fixtures from the unmodified reference (tests/golden/make_categorical_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from test_gpu_model import build, rel

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('layout', ['channels_last', 'reference'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_kernels_against_torch(pkg, layout, dtype):
    nat = pkg._native
    torch.manual_seed(0)
    B, L, K, C, Ci, H, W = 5, 3, 8, 4, 3, 6, 5
    D = Ci * H * W
    x = torch.rand(B, Ci, H, W, device=DEV)
    x[0, 0, 0, 0], x[0, 1, 0, 0] = 1.0, 0.0                      # the ends of the target range
    logits = (2 * torch.randn(L + 1, B, 256, Ci, H, W, device=DEV)).to(dtype)      # the reference's layout
    mu, lv = torch.randn(B, K, device=DEV), 0.3 * torch.randn(B, K, device=DEV)
    y = torch.randint(0, C, (B,), device=DEV)
    means, T = torch.randn(C, K, device=DEV), 1 + 0.2 * torch.rand(C, device=DEV)
    sig = torch.tensor([0.7], device=DEV)
    if layout == 'channels_last':      # (N, 256 Ci, H, W) stored NHWC, channel = v * Ci + c; x in NHWC order
        xr_k = logits.reshape(-1, 256 * Ci, H, W).contiguous(memory_format=torch.channels_last)
        x_k = x.contiguous(memory_format=torch.channels_last)
        group = Ci
    else:
        xr_k, x_k, group = logits.reshape(-1, 256 * Ci, H, W).contiguous(), x.contiguous(), D
    cfg = nat.make_cfg(B=B, L=L, K=K, C=C, D=D, x_reco=xr_k, logits=None, var_dim='scalar', prior_kind='gaussian',
                       conditional=True, sigma_is_log=False, sigma_is_rmse=False, beta=0.7, gamma_w=0.0, var_w=1.0,
                       categorical=True, cat_group=group)
    out = nat.elbo_train_fwd(cfg, x_k, xr_k, mu, lv, None, y, means, T, sig)
    # torch restatement of losses.py:30-49 and cvae.py:657-660, 776
    lg = logits.float().requires_grad_(True)
    tgt = (x * 255).long()
    ce = F.cross_entropy(lg[1:].reshape(-1, 256, Ci, H, W), tgt.expand(L, B, Ci, H, W).reshape(-1, Ci, H, W),
                         reduction='none').view(L, B, -1).sum(-1)
    cross_x = ce.mean(0)
    wmse = ((lg[1:].argmax(2) / 255 - x) ** 2).flatten(2).mean(-1).mean(0)
    assert torch.allclose(out['cross_x'], cross_x, rtol=1e-4, atol=1e-2)
    assert torch.allclose(out['wmse'], wmse, rtol=1e-5, atol=1e-6)
    assert torch.allclose(out['total'], cross_x + 0.7 * out['kl'], rtol=1e-5, atol=1e-2)
    g = torch.full((B,), 1.0 / B, device=DEV)
    d_xr, *_ = nat.elbo_train_bwd(cfg, g, x_k, xr_k, mu, lv, None, y, means, T, sig, out['wmse'])
    (cross_x.detach() * 0 + cross_x).mul(g).sum().backward()
    want = lg.grad.reshape(-1, 256 * Ci, H, W)
    got = d_xr.float()
    assert got.shape == want.shape
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert float((got - want).abs().max()) <= tol * float(want.abs().max())
    assert float(got.reshape(L + 1, B, -1)[0].abs().max()) == 0.0      # draw 0 (the mean) carries no loss
    # eval: per-class totals use the same cross_x; the importance weights start from -CE (cvae.py:683)
    z = mu[None] + torch.randn(L + 1, B, K, device=DEV) * 0.1
    en = torch.rand(L, B, device=DEV)
    r = nat.elbo_eval_fwd(cfg, x_k, xr_k, mu, lv, z, en, None, means, T, sig)
    assert torch.allclose(r['cross_x'], cross_x.detach(), rtol=1e-4, atol=1e-2)
    logp = -0.5 * K * np.log(2 * np.pi) - 0.5 * ((z[1:, None] - means[None, :, None]) * T[None, :, None, None]).pow(2).sum(-1) \
        + K * T.log()[None, :, None]
    li = -ce.detach()[:, None] + logp + 0.5 * (en + lv.sum(-1))[:, None] + 0.5 * K * np.log(2 * np.pi)
    iws = (li - li.max(0)[0]).exp().mean(0) + li.max(0)[0]
    assert torch.allclose(r['iws'], iws, rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize('name', ['cat_mlp_cvae', 'cat_conv_cvae'])
def test_model_matches_reference(pkg, name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    cfg, net = build(pkg, d)
    x = torch.from_numpy(d['x']).to(DEV)
    y = torch.from_numpy(d['y']).to(DEV)
    tol = 3e-2
    net.train()
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_train']).to(DEV)
    net.optimizer.zero_grad()
    n0 = pkg._native.launch_count()
    x_reco, logits, losses, _ = net.evaluate(x, y, with_beta=True, kl_var_weighting=float(d['train.kl_var_weighting']),
                                             gamma_weighting=float(d['train.gamma_weighting']))
    assert pkg._native.launch_count() > n0
    L1, B = d['eps_train'].shape[0], x.shape[0]
    assert tuple(x_reco.shape) == (L1, B, 256) + tuple(cfg['input_shape'])
    agree = (x_reco.float().argmax(2).cpu().numpy() == d['train.x_reco_argmax']).mean()
    assert agree > 0.9, agree                 # bf16 logits: near-ties of the 256-way arg-max may flip
    keys = sorted(k[len('train.loss.'):] for k in d.files if k.startswith('train.loss.'))
    assert sorted(losses) == keys
    for k in keys:
        t = 0.15 if k == 'wmse' else tol      # wmse comes from the arg-max image
        assert rel(losses[k].detach().cpu().numpy(), d['train.loss.' + k]) < t, (k, rel(losses[k].detach().cpu().numpy(), d['train.loss.' + k]))
    losses['total'].mean().backward()
    gmax = max(float(np.linalg.norm(d[k])) for k in d.files if k.startswith('train.grad.'))
    checked = 0
    for k, p in net.named_parameters():
        gk = 'train.grad.' + k
        if gk not in d.files:
            continue
        g, gr = p.grad.detach().float().cpu().numpy().astype(np.float64), d[gk].astype(np.float64)
        assert np.isfinite(g).all(), k
        assert np.linalg.norm(g - gr) <= 0.15 * np.linalg.norm(gr) + 0.02 * gmax, (k, np.linalg.norm(g - gr), np.linalg.norm(gr))
        checked += 1
    assert checked >= 4
    net.eval()
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_eval']).to(DEV)
    with torch.no_grad():
        _, logits, el, _ = net.evaluate(x)
    keys = sorted(k[len('eval.loss.'):] for k in d.files if k.startswith('eval.loss.'))
    assert sorted(el) == keys
    for k in keys:
        assert tuple(el[k].shape) == d['eval.loss.' + k].shape, k
        t = 0.15 if k == 'wmse' else tol
        assert rel(el[k].cpu().numpy(), d['eval.loss.' + k]) < t, (k, rel(el[k].cpu().numpy(), d['eval.loss.' + k]))

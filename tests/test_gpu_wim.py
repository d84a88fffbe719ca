"""WIM caller (ft/wim.py:48-129, 215-259): `encoder.prior` swapped for the alternate single prior and
num_labels = 1, then swapped back -- against golden vectors from the unmodified reference
(tests/golden/make_wim_golden.py).  The fused kernels read the prior through `encoder.prior` at every call, so the swap
needs no other hook; each prior layout (5 priors / 1 prior) keeps its own workspace."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_gpu_model import build, rel

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_evaluate_under_swapped_priors(pkg):
    d = np.load(os.path.join(GOLDEN, 'wim_alternate_prior.npz'))
    cfg, net = build(pkg, d)
    alt_kw = json.loads(str(d['alt_params']))
    alt = pkg.module.priors.build_prior(**alt_kw)
    alt.load_state_dict({k[4:]: torch.from_numpy(np.asarray(d[k])) for k in d.files if k.startswith('alt.')})
    alt = alt.to(DEV)
    for p in alt.parameters():
        p.requires_grad_(False)
    x = torch.from_numpy(d['x']).to(DEV)
    B = x.shape[0]
    original, C = net.encoder.prior, net.num_labels
    tol = 3e-2
    # reference point under the original prior (per-class losses (C, B))
    net.eval()
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_eval']).to(DEV)
    with torch.no_grad():
        before = net.evaluate(x)[2]
    assert before['total'].shape == (C, B)
    # ---- swap (ft/wim.py:56-61), train-mode step of finetune_batch on the mix batch
    net.encoder.prior, net.num_labels = alt, 1
    net.train()
    net.optimizer.zero_grad()
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_train']).to(DEV)
    y_mix = torch.zeros(B, dtype=torch.long, device=DEV)
    _, _, losses, _ = net.evaluate(x, y_mix, with_beta=True)
    keys = sorted(k[len('train.loss.'):] for k in d.files if k.startswith('train.loss.'))
    assert sorted(losses) == keys
    for k in keys:
        assert rel(losses[k].detach().cpu().numpy(), d['train.loss.' + k]) < tol, k
    losses['total'].mean().backward()
    checked = 0
    gmax = max(float(np.linalg.norm(d[k])) for k in d.files if k.startswith('train.grad.'))
    for k, p in net.named_parameters():
        gk = 'train.grad.' + k
        if gk not in d.files:
            continue
        g, gr = p.grad.detach().float().cpu().numpy().astype(np.float64), d[gk].astype(np.float64)
        assert np.linalg.norm(g - gr) <= 0.1 * np.linalg.norm(gr) + 0.02 * gmax, k
        checked += 1
    assert checked >= 6
    # ---- eval under the alternate prior (evaluate_on_both_priors, ft/wim.py:114-129): every loss is (B,)
    net.eval()
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_eval']).to(DEV)
    with torch.no_grad():
        _, logits, el, _ = net.evaluate(x)
    keys = sorted(k[len('eval.loss.'):] for k in d.files if k.startswith('eval.loss.'))
    assert sorted(el) == keys
    for k in keys:
        assert tuple(el[k].shape) == d['eval.loss.' + k].shape, k
        assert rel(el[k].cpu().numpy(), d['eval.loss.' + k]) < tol, k
    # ---- swap back: the original prior gives its per-class losses again
    net.encoder.prior, net.num_labels = original, C
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_eval']).to(DEV)
    with torch.no_grad():
        after = net.evaluate(x)[2]
    for k in before:
        assert torch.equal(before[k], after[k]), k

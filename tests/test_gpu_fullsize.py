"""Full-size checks at BASELINE.json's configurations, through size-independent properties (the oracle cannot run these
sizes in seconds): loss identities, predictions = arg-min / arg-max of the returned per-class tensors, determinism under
injected noise, finite gradients, a descending loss over a few Adam steps; plus a smoke run of the ResNet / ivgg
configuration (c4), executed entirely by the native conv engine (there is no library path to fall back to)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
PRIOR = {'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 1}


def _c2(pkg, C=10, K=128, L=16):
    torch.manual_seed(0)
    return pkg.ClassificationVariationalNetwork(
        (3, 32, 32), C, type='cvae', features='vgg19', upsampler='deconv32', encoder=[], decoder=[], classifier=[],
        batch_norm='both', latent_dim=K, latent_sampling=L, test_latent_sampling=L, gamma=0, beta=1.0,
        output_activation='linear', sigma={'value': 1.0, 'learned': True}, prior=dict(PRIOR),
        optimizer={'optim_type': 'adam', 'lr': 1e-3, 'weight_decay': 3e-5, 'grad_clipping': 100}).to(DEV)


def test_c2_full_size_properties(pkg):
    B, L, K, C = 512, 16, 128, 10
    net = _c2(pkg)
    g = torch.Generator(device='cpu').manual_seed(3)
    x = torch.rand(B, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, C, (B,), generator=g).to(DEV)
    eps = torch.randn(L + 1, B, K, generator=g).to(DEV)
    net.train()
    net.encoder.sampling.injected_eps = eps
    net.optimizer.zero_grad()
    _, logits, losses, _ = net.evaluate(x, y, with_beta=True)
    for k, v in losses.items():
        assert v.shape == (B,) and torch.isfinite(v).all(), k
    # cvae.py:744, 791, 887-902: total = cross_x + beta * kl (gamma = 0); kl = (zdist + var_kl) / 2
    assert torch.allclose(losses['total'], losses['cross_x'] + losses['kl'], rtol=1e-5, atol=1e-3)
    assert torch.allclose(losses['kl'], 0.5 * (losses['zdist'] + losses['var_kl']), rtol=1e-5, atol=1e-3)
    assert (losses['wmse'] >= 0).all() and (losses['zdist'] >= 0).all()
    # determinism: the same inputs and noise give bit-identical losses
    net.encoder.sampling.injected_eps = eps
    _, _, again, _ = net.evaluate(x, y, with_beta=True)
    assert torch.equal(again['total'], losses['total'])
    losses['total'].mean().backward()
    n = 0
    for name, p in net.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
            n += 1
    assert n > 40
    # a few optimisation steps on the same batch reduce the loss
    first = float(losses['total'].mean())
    for _ in range(6):
        net.encoder.sampling.injected_eps = eps
        ls, _ = net.train_step(x, y)
    assert float(ls['total'].mean()) < first
    assert int(net._finite_flag.item()) != 0


@pytest.mark.parametrize('C,K', [(10, 128), (100, 256), (1000, 256)])
def test_scoring_full_size_properties(pkg, C, K):
    """per-class evaluation at B = 512 (BASELINE configs[1], [2], [4]): shapes, identities, fused predictions / scores"""
    B, L = 512, 16
    net = _c2(pkg, C=C, K=K, L=L)
    net.eval()
    x = torch.rand(B, 3, 32, 32, device=DEV)
    with torch.no_grad():
        _, logits, losses, _ = net.evaluate(x)
        for k in ('kl', 'zdist', 'var_kl', 'total', 'iws'):
            assert losses[k].shape == (C, B) and torch.isfinite(losses[k]).all(), k
        for k in ('wmse', 'cross_x', 'dzdist'):
            assert losses[k].shape == (B,), k
        # eval: beta = 1, total_cb = cross_x_b + kl_cb (cvae.py:560, 898)
        assert torch.allclose(losses['total'], losses['cross_x'][None] + losses['kl'], rtol=1e-5, atol=1e-2)
        for m, want in (('closest', losses['zdist'].argmin(0)), ('iws', losses['iws'].argmax(0)),
                        ('loss', losses['total'].argmin(0))):
            got = net.predict_after_evaluate(logits, losses, method=m)
            assert (got == want).float().mean() > 0.999, m       # ties between classes may be broken differently
        sc = net.batch_dist_measures(logits, losses, ['elbo', 'zdist', 'kl', 'mse', 'iws'])
        assert torch.allclose(sc['elbo'], (-losses['total']).max(0)[0], rtol=1e-5, atol=1e-3)
        assert torch.allclose(sc['zdist'], (-losses['zdist']).max(0)[0], rtol=1e-5, atol=1e-3)
        assert torch.allclose(sc['mse'], -losses['cross_x'], rtol=1e-6)
        ref_iws = torch.logsumexp(losses['iws'], 0) + torch.log(torch.tensor(float(C), device=DEV))
        assert torch.allclose(sc['iws'], ref_iws, rtol=1e-4, atol=1e-2)


def test_c4_resnet_ivgg_runs(pkg):
    """BASELINE configs[3]: resnet18 features (torchvision modules executed by the native conv engine: 7x7 stem, padded
    max pool, residual blocks, global average pool) + native ivgg imager; nothing falls back to the library path"""
    torch.manual_seed(0)
    net = pkg.ClassificationVariationalNetwork(
        (3, 64, 64), 20, type='cvae', features='resnet18', upsampler='ivgg', encoder=[], decoder=[], classifier=[],
        batch_norm='both', latent_dim=256, latent_sampling=8, test_latent_sampling=8, gamma=0,
        output_activation='linear', sigma={'value': 1.0, 'learned': True}, prior=dict(PRIOR)).to(DEV)
    B = 32
    x = torch.rand(B, 3, 64, 64, device=DEV)
    y = torch.randint(0, 20, (B,), device=DEV)
    net.train()
    n0 = pkg._native.launch_count()
    losses, _ = net.train_step(x, y)
    assert pkg._native.launch_count() > n0
    assert torch.isfinite(losses['total']).all() and losses['total'].shape == (B,)
    net.eval()
    with torch.no_grad():
        xr, logits, el, _ = net.evaluate(x)
    assert tuple(xr.shape) == (9, B, 3, 64, 64) and el['total'].shape == (20, B) and torch.isfinite(el['total']).all()

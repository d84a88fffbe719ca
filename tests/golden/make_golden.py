"""Generate golden fixtures by running the UNMODIFIED reference (moxime/joint-vae,
mounted read-only at /root/reference) on CPU with injected noise.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Each case writes tests/golden/<name>.npz holding: the constructor kwargs (json),
the full state_dict, inputs x / y / eps (train and eval), and the reference's
outputs: every key of batch_losses (train (B,), eval (C,B)), logits, x_reco,
mu / log_var / z, gradients of every parameter after total.mean().backward(),
every batch_dist_measures score and every predict_after_evaluate prediction.

Noise injection: the reference's Sampling.forward (module/vae_layers/layers.py:230-244)
calls torch.randn((L+1,B,K)); we swap torch.randn for the duration of one
evaluate() so that it returns our eps (slab 0 is zeroed by the reference itself).
"""
import contextlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('JVAE_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for m in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):
        sys.modules.setdefault(m, types.ModuleType(m))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings('ignore')
    import cvae  # noqa
    return cvae


@contextlib.contextmanager
def injected_noise(eps):
    """Make torch.randn / torch.rand return `eps` (a (L+1,B,K) tensor) once."""
    orig_randn, orig_rand = torch.randn, torch.rand

    def fake_randn(size, *a, **k):
        assert tuple(size) == tuple(eps.shape), (tuple(size), tuple(eps.shape))
        return eps.clone()

    def fake_rand(size, *a, **k):
        # uniform prior: reference does (rand - 0.5) * sqrt(12); we inject eps already
        # in that scale, so invert here.
        if isinstance(size, (tuple, list, torch.Size)) and tuple(size) == tuple(eps.shape):
            return eps.clone() / np.sqrt(12) + 0.5
        return orig_rand(size, *a, **k)
    torch.randn, torch.rand = fake_randn, fake_rand
    try:
        yield
    finally:
        torch.randn, torch.rand = orig_randn, orig_rand


CASES = {
    # name: (ctor kwargs, B, extra)
    'mlp_cvae': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32, 16], decoder=[16, 32],
        classifier=[], latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.1},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 1}),
    'mlp_cvae_gamma_diag': dict(
        input_shape=(1, 8, 8), num_labels=6, type='cvae', encoder=[24], decoder=[24],
        classifier=[12], latent_dim=8, latent_sampling=2, test_latent_sampling=3, gamma=0.5, beta=0.7,
        output_activation='linear', sigma={'value': 0.5, 'learned': True},
        prior={'init_mean': 1.5, 'learned_means': True, 'var_dim': 'diag', 'seed': 2}),
    'mlp_cvae_full': dict(
        input_shape=(1, 6, 6), num_labels=4, type='cvae', encoder=[20], decoder=[20],
        classifier=[], latent_dim=6, latent_sampling=2, test_latent_sampling=2, gamma=0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.3},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'full', 'seed': 3}),
    'mlp_cvae_softmax': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32], decoder=[32],
        classifier=['softmax'], latent_dim=8, latent_sampling=2, test_latent_sampling=2, gamma=1.0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.2},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 4}),
    'mlp_cvae_tilted': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32], decoder=[32],
        classifier=[], latent_dim=8, latent_sampling=2, test_latent_sampling=2, gamma=0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.2},
        prior={'distribution': 'tilted', 'tau': 3.0, 'init_mean': 1.0, 'learned_means': True, 'seed': 5}),
    'mlp_cvae_uniform': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32], decoder=[32],
        classifier=[], latent_dim=8, latent_sampling=2, test_latent_sampling=2, gamma=0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.2},
        prior={'distribution': 'uniform', 'tau': 2.0, 'init_mean': 1.0, 'learned_means': True, 'seed': 6}),
    'mlp_jvae': dict(
        input_shape=(1, 8, 8), num_labels=5, type='jvae', encoder=[32], decoder=[32],
        classifier=[10], latent_dim=8, latent_sampling=2, test_latent_sampling=3, gamma=2.0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.2}, prior={'var_dim': 'scalar'}),
    'mlp_xvae': dict(
        input_shape=(1, 8, 8), num_labels=5, type='xvae', encoder=[32], decoder=[32],
        classifier=[10], latent_dim=8, latent_sampling=2, test_latent_sampling=3, gamma=1.0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.2},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 7}),
    'mlp_vae': dict(
        input_shape=(1, 8, 8), num_labels=5, type='vae', encoder=[32], decoder=[32],
        classifier=[], latent_dim=8, latent_sampling=2, test_latent_sampling=3, gamma=0, beta=1.0,
        output_activation='sigmoid', sigma={'value': 0.2}, prior={'var_dim': 'scalar'}),
    'mlp_vib': dict(
        input_shape=(1, 8, 8), num_labels=5, type='vib', encoder=[32], decoder=[32],
        classifier=[10], latent_dim=8, latent_sampling=2, test_latent_sampling=3, gamma=1.0, beta=1e-2,
        output_activation='sigmoid', sigma={'value': 0.2}, prior={'var_dim': 'scalar'}),
    'conv_cvae_bn': dict(
        input_shape=(3, 16, 16), num_labels=4, type='cvae', features='[x3+1]8-8-M-16:2-16',
        upsampler='[x3+1]16x4+0-16-8:2++1-8:2++1-!3x3+1', batch_norm='both',
        encoder=[], decoder=[], classifier=[], latent_dim=16, latent_sampling=3, test_latent_sampling=2,
        gamma=0, beta=1.0, output_activation='linear', sigma={'value': 1.0, 'learned': True},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 8}),
    'conv_cvae_nobn_k5': dict(
        input_shape=(3, 16, 16), num_labels=4, type='cvae', features='[x5+2]8-8:2-16x8+0',
        upsampler='[x5+2]16x4+0-8:2++1-8:2++1-8-!3x5+2', batch_norm=False,
        encoder=[], decoder=[], classifier=[], latent_dim=16, latent_sampling=2, test_latent_sampling=2,
        gamma=0, beta=1.0, output_activation='linear', sigma={'value': 1.0, 'learned': True},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 9}),
}

BATCH = {'default': 6}


def t2n(t):
    return t.detach().cpu().numpy().copy()   # copy: state_dict tensors are updated in place later (BN)


def describe(seq):
    """JSON description of an nn.Sequential built by the reference (layer types + hyper-parameters);
    the oracle's torch restatement is built from this list, and the product's spec parser is
    tested against it."""
    from torch import nn
    out = []
    if seq is None:
        return out
    for m in seq:
        if isinstance(m, nn.ConvTranspose2d):
            out.append(dict(t='convT', cin=m.in_channels, cout=m.out_channels, k=m.kernel_size[0],
                            s=m.stride[0], p=m.padding[0], op=m.output_padding[0]))
        elif isinstance(m, nn.Conv2d):
            out.append(dict(t='conv', cin=m.in_channels, cout=m.out_channels, k=m.kernel_size[0],
                            s=m.stride[0], p=m.padding[0]))
        elif isinstance(m, nn.BatchNorm2d):
            out.append(dict(t='bn', n=m.num_features, eps=m.eps, momentum=m.momentum))
        elif isinstance(m, nn.Linear):
            out.append(dict(t='linear', cin=m.in_features, cout=m.out_features))
        elif isinstance(m, nn.MaxPool2d):
            out.append(dict(t='maxpool', k=m.kernel_size, s=m.stride, p=m.padding))
        elif isinstance(m, nn.AvgPool2d):
            out.append(dict(t='avgpool', k=m.kernel_size, s=m.stride, p=m.padding))
        elif isinstance(m, nn.UpsamplingNearest2d):
            out.append(dict(t='upsample', s=int(m.scale_factor)))
        elif isinstance(m, nn.ReLU):
            out.append(dict(t='relu'))
        elif isinstance(m, nn.Sigmoid):
            out.append(dict(t='sigmoid'))
        elif isinstance(m, nn.Identity):
            out.append(dict(t='identity'))
        elif isinstance(m, nn.LeakyReLU):
            out.append(dict(t='leaky'))
        elif type(m).__name__ == 'Reshape':      # categorical imager: channels -> (256, C), conv.py:228-230
            out.append(dict(t='reshape', shape=[int(v) for v in m.shape]))
        else:
            raise TypeError(str(m))
    return out


def describe_model(model):
    arch = {'features': describe(model.features),
            'features_out': list(model.encoder.input_shape) if model.features is not None else None,
            'dense_projs': describe(model.encoder.dense_projs),
            'classifier': describe(getattr(model, 'classifier', None)),
            'classifier_type': model.classifier_type,
            'sampling': bool(model.encoder.sampling.is_sampled),
            'y_is_decoded': bool(model.y_is_decoded),
            'sigma': {'is_log': bool(model.sigma.is_log), 'is_rmse': bool(model.sigma.is_rmse),
                      'learned': bool(model.sigma.learned), 'sdim': int(model.sigma.sdim)},
            'prior': {'conditional': bool(model.encoder.prior.conditional),
                      'var_dim': model.encoder.prior.var_dim,
                      'distribution': model.encoder.prior.params['distribution'],
                      'tau': getattr(model.encoder.prior, 'tau', None)}}
    if not model.is_vib:
        arch['decoder'] = describe(model.decoder)
        arch['imager'] = describe(model.imager)
        arch['imager_in'] = list(model.imager.input_shape)
    return arch


def run_case(cvae_mod, name, kw, eval_part=True):
    CVNet = cvae_mod.ClassificationVariationalNetwork
    kw = json.loads(json.dumps(kw))  # deep copy; the reference mutates prior/sigma dicts
    ctor = dict(kw)
    ctor['input_shape'] = tuple(ctor['input_shape'])
    torch.manual_seed(1234)
    model = CVNet(**ctor)
    B = BATCH.get(name, BATCH['default'])
    C, K = kw['num_labels'], kw['latent_dim']
    Ltr, Lte = kw['latent_sampling'], kw['test_latent_sampling']
    g = torch.Generator().manual_seed(99)
    x = torch.rand(B, *ctor['input_shape'], generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    eps_tr = torch.randn(Ltr + 1, B, K, generator=g)
    eps_te = torch.randn(Lte + 1, B, K, generator=g)
    if kw['prior'].get('distribution') == 'uniform':
        eps_tr = (torch.rand(Ltr + 1, B, K, generator=g) - 0.5) * np.sqrt(12)
        eps_te = (torch.rand(Lte + 1, B, K, generator=g) - 0.5) * np.sqrt(12)
    # move means away from init a bit so that nothing is accidentally symmetric
    out = {'cfg': np.array(json.dumps(kw)), 'x': t2n(x), 'y': t2n(y),
           'eps_train': t2n(eps_tr), 'eps_eval': t2n(eps_te)}
    for k, v in model.state_dict().items():
        out['sd.' + k] = t2n(v)
    out['arch'] = np.array(json.dumps(describe_model(model), default=lambda o: o.item()))

    # ---------------- train step (forward + backward), cvae.py:2440-2459
    model.train()
    model.optimizer.zero_grad()
    with injected_noise(eps_tr):
        o = model.evaluate(x, y, with_beta=True, kl_var_weighting=0.8, gamma_weighting=0.9, z_output=True)
    x_reco, logits, losses, measures, mu, log_var, z = o
    out['train.kl_var_weighting'] = np.array(0.8)
    out['train.gamma_weighting'] = np.array(0.9)
    for k, v in losses.items():
        out['train.loss.' + k] = t2n(v)
    out['train.logits'] = t2n(logits)
    out['train.x_reco'] = t2n(x_reco)
    out['train.mu'], out['train.log_var'], out['train.z'] = t2n(mu), t2n(log_var), t2n(z)
    for k, v in measures.items():
        out['train.measure.' + k] = np.array(float(v))
    losses['total'].mean().backward()
    for k, p in model.named_parameters():
        if p.grad is not None:
            out['train.grad.' + k] = t2n(p.grad)
    for k, v in model.state_dict().items():
        if 'running_' in k or 'num_batches' in k or k == 'sigma':
            out['train.sd_after.' + k] = t2n(v)

    # ---------------- eval / scoring step, cvae.py:1629-1677
    model.eval()
    if not eval_part:
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'ok (train step only)', '%.1f KB' % (os.path.getsize(path) / 1024))
        return
    with torch.no_grad():
        with injected_noise(eps_te):
            o = model.evaluate(x, z_output=True)
        x_reco, logits, losses, measures, mu, log_var, z = o
        for k, v in losses.items():
            out['eval.loss.' + k] = t2n(v)
        out['eval.logits'] = t2n(logits)
        out['eval.x_reco'] = t2n(x_reco)
        out['eval.mu'], out['eval.log_var'], out['eval.z'] = t2n(mu), t2n(log_var), t2n(z)
        methods = [m for m in model.ood_methods if not m.startswith('odin')]
        extra = {'cvae': ['softkl-2', 'kl', 'max', 'hyz', 'baseline', 'baseline-10', 'logits', 'sum', 'mean', 'std',
                          'softiws', 'softiws-5', 'softzdist-2', 'mag', 'nstd', 'IYx', 'wmse'],
                 'xvae': ['iws', 'sum', 'max', 'mean', 'std', 'mag', 'IYx', 'logits', 'baseline'],
                 # the reference crashes for jvae (cvae.py:990-997, `iws` unbound) unless a method naming
                 # 'iws' is requested, in which case it falls back to iws = -total (cvae.py:992-994)
                 'jvae': ['iws', 'sum', 'max', 'mean', 'std', 'logits', 'baseline', 'hyz'],
                 'vae': ['kl'], 'vib': ['baseline', 'logits', 'hyz', 'baseline-2']}[kw['type']]
        methods = list(dict.fromkeys(methods + extra))
        dm = model.batch_dist_measures(logits, losses, methods)
        out['eval.methods'] = np.array(json.dumps(methods))
        for m, v in dm.items():
            out['eval.measure.' + m] = t2n(v)
        preds = list(model.predict_methods)
        out['eval.predict_methods'] = np.array(json.dumps(preds))
        for m in preds:
            out['eval.pred.' + m] = t2n(model.predict_after_evaluate(logits, losses, method=m))
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, 'ok', '%.1f KB' % (os.path.getsize(path) / 1024), 'train keys',
          sorted(k[11:] for k in out if k.startswith('train.loss.')),
          'eval keys', sorted(k[10:] for k in out if k.startswith('eval.loss.')))


if __name__ == '__main__':
    os.chdir('/tmp')
    mod = import_reference()
    torch.set_num_threads(1)
    only = sys.argv[1:]
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        run_case(mod, name, kw)

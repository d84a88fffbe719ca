"""ORACLE (test infrastructure, NOT product code).

Plain-PyTorch fp32 restatement of the reference network of moxime/joint-vae's hot path
(features -> Encoder -> Sampling -> decoder -> imager -> classifier, cvae.py:426-521) and of the
train-mode total loss (cvae.py:626-902) so that autograd yields reference gradients.

It is built from the JSON `arch` description stored in tests/golden/*.npz (dumped from the
reference's own modules by tests/golden/make_golden.py), uses the reference's state_dict key
names, and is pinned against the golden outputs and gradients in tests/test_oracle_golden.py.
It is also the CPU baseline "port" that bench.py times (`cpu_baseline`, `--impl reference`):
the reference itself is PyTorch eager code and cannot travel to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import json
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

LOG2PI = math.log(2 * math.pi)


def _build_seq(desc):
    layers = []
    for d in desc:
        t = d['t']
        if t == 'conv':
            layers.append(nn.Conv2d(d['cin'], d['cout'], d['k'], stride=d['s'], padding=d['p']))
        elif t == 'convT':
            layers.append(nn.ConvTranspose2d(d['cin'], d['cout'], d['k'], stride=d['s'], padding=d['p'],
                                             output_padding=d['op']))
        elif t == 'bn':
            layers.append(nn.BatchNorm2d(d['n'], eps=d.get('eps', 1e-5), momentum=d.get('momentum', 0.1)))
        elif t == 'linear':
            layers.append(nn.Linear(d['cin'], d['cout']))
        elif t == 'maxpool':
            layers.append(nn.MaxPool2d(d['k'], stride=d['s'], padding=d['p']))
        elif t == 'avgpool':
            layers.append(nn.AvgPool2d(d['k'], stride=d['s'], padding=d['p']))
        elif t == 'upsample':
            layers.append(nn.UpsamplingNearest2d(scale_factor=d['s']))
        elif t == 'relu':
            layers.append(nn.ReLU())
        elif t == 'sigmoid':
            layers.append(nn.Sigmoid())
        elif t == 'identity':
            layers.append(nn.Identity())
        elif t == 'leaky':
            layers.append(nn.LeakyReLU())
        else:
            raise ValueError(t)
    return nn.Sequential(*layers)


class _Encoder(nn.Module):
    def __init__(self, arch, K, C, in_features):
        super().__init__()
        self.dense_projs = _build_seq(arch['dense_projs'])
        f = in_features
        for d in arch['dense_projs']:
            if d['t'] == 'linear':
                f = d['cout']
        self.dense_mean = nn.Linear(f, K)
        self.dense_log_var = nn.Linear(f, K)
        self.prior = _PriorParams(arch['prior'], K, C)


class _PriorParams(nn.Module):
    def __init__(self, p, K, C):
        super().__init__()
        n = C if p['conditional'] else 1
        self.mean = nn.Parameter(torch.zeros(n, K))
        per = {'scalar': (), 'diag': (K,), 'full': (K, K)}[p['var_dim']]
        shape = ((n,) + per) if p['conditional'] else per
        self._var_parameter = nn.Parameter(torch.ones(shape))


class OracleNet(nn.Module):
    """Same parameter names / shapes as the reference's ClassificationVariationalNetwork."""

    def __init__(self, cfg, arch):
        super().__init__()
        if isinstance(cfg, str):
            cfg = json.loads(cfg)
        if isinstance(arch, str):
            arch = json.loads(arch)
        self.cfg, self.arch = cfg, arch
        self.type = cfg['type']
        self.input_shape = tuple(cfg['input_shape'])
        self.C, self.K = cfg['num_labels'], cfg['latent_dim']
        s = arch['sigma']
        self.sigma = nn.Parameter(torch.zeros(s['sdim']), requires_grad=s['learned'])
        if arch['features']:
            self.features = _build_seq(arch['features'])
            in_f = int(np.prod(arch['features_out']))
        else:
            self.features = None
            in_f = int(np.prod(self.input_shape))
        self.encoder = _Encoder(arch, self.K, self.C, in_f)
        if self.type != 'vib':
            self.decoder = _build_seq(arch['decoder'])
            self.imager = _build_seq(arch['imager'])
        if arch['classifier_type'] in ('linear', None):
            self.classifier = _build_seq(arch['classifier'])

    def load_numpy_state(self, npz, prefix='sd.', after_train=False):
        """after_train=True also applies the BN running statistics recorded after the golden train
        step (the reference's eval pass ran after it)."""
        sd = {k[len(prefix):]: torch.from_numpy(np.asarray(npz[k])) for k in npz.files if k.startswith(prefix)}
        if after_train:
            p2 = 'train.sd_after.'
            sd.update({k[len(p2):]: torch.from_numpy(np.asarray(npz[k])) for k in npz.files if k.startswith(p2)})
        self.load_state_dict(sd)
        return self

    # cvae.py:426-521 + layers.py:350-403
    def forward(self, x, eps):
        """eps (L+1,B,K), slab 0 is zeroed here like layers.py:238. Returns
        x_reco (L+1,B,*shape) | None, y_est (L+1,B,C), mu, log_var, z, eps_norm (L,B)."""
        B = x.shape[0]
        t = x
        if self.features is not None:
            t = self.features(x)
        u = self.encoder.dense_projs(t.reshape(B, -1))
        mu = self.encoder.dense_mean(u)
        log_var = torch.clip(self.encoder.dense_log_var(u), -20, 20)
        eps = eps.clone()
        eps[0] = 0
        z = mu + torch.exp(0.5 * log_var) * eps * float(self.arch['sampling'])
        L1 = eps.shape[0]
        x_reco = None
        if self.type != 'vib':
            h = self.decoder(z)
            h = self.imager(h.reshape(-1, *self.arch['imager_in']))
            x_reco = h.reshape(L1, B, *self.input_shape)
        if self.arch['classifier_type'] == 'softmax':
            m = self.encoder.prior.mean
            y_est = F.linear(z, m, m.pow(2).sum(-1) / 2)          # cvae.py:499
        else:
            y_est = self.classifier(z)
        return x_reco, y_est, mu, log_var, z, (eps[1:] ** 2).sum(-1)

    # ---- train-mode total loss in torch (gaussian prior), for autograd gradients
    def prior_kl_train(self, mu, log_var, y, var_weighting):
        p = self.encoder.prior
        vd = self.arch['prior']['var_dim']
        cond = self.arch['prior']['conditional']
        T = p._var_parameter.tril() if vd == 'full' else p._var_parameter
        m = p.mean[y] if cond else p.mean.reshape(-1)
        d = mu - m
        Ty = T[y] if cond else T
        if vd == 'full':
            w = torch.matmul(Ty, d.unsqueeze(-1)).squeeze(-1)
            diag = (T ** 2).sum(-2)
            logdet = -2 * torch.diagonal(T, dim1=-2, dim2=-1).abs().log().sum(-1)
        elif vd == 'diag':
            w = d * Ty
            diag = T ** 2
            logdet = -2 * T.abs().log().sum(-1)
        else:
            w = d * (Ty.unsqueeze(-1) if cond else Ty)
            diag = T ** 2
            logdet = -2 * self.K * T.log()
        dist = w.pow(2).sum(-1)
        if cond:
            diag, logdet = diag[y], logdet[y]
        if vd == 'scalar':
            diag = diag.unsqueeze(-1) if cond else diag
        trace = (log_var.exp() * diag).sum(-1)
        var_kl = trace - log_var.sum(-1) + logdet - self.K
        return 0.5 * (dist + var_weighting * var_kl), dist, var_kl

    def train_losses(self, x, y, eps, *, beta, gamma, kl_var_weighting=1.0, gamma_weighting=1.0):
        """cvae.py:523-917 with y given, self.training, with_beta=True; gaussian prior only."""
        x_reco, y_est, mu, log_var, z, _ = self.forward(x, eps)
        out = {}
        kl, dist, var_kl = self.prior_kl_train(mu, log_var, y if self.arch['prior']['conditional'] else None,
                                               kl_var_weighting)
        out['kl'], out['zdist'], out['var_kl'] = kl, dist, var_kl
        total = torch.zeros_like(kl)
        if self.type != 'vib':
            D = int(np.prod(self.input_shape))
            s = self.sigma
            sg = self.arch['sigma']
            sigma_ = s.exp() if sg['is_log'] else s
            log_sigma = s.squeeze() if sg['is_log'] else s.log().squeeze()
            nd = len(self.input_shape)
            wl = F.mse_loss(x_reco[1:] / sigma_, (x / sigma_).expand_as(x_reco[1:]),
                            reduction='none').mean(tuple(range(-nd, 0)))
            out['wmse'] = wl.mean(0)
            out['cross_x'] = D * (2 * log_sigma + out['wmse'] + LOG2PI) / 2
            total = total + out['cross_x']
        if self.arch['y_is_decoded']:
            L1, B, C = y_est.shape
            ce = F.cross_entropy(y_est.reshape(-1, C), y.repeat(L1), reduction='none').reshape(L1, B).mean(0)
            out['cross_y'] = ce
            w = gamma_weighting * gamma
            if w:
                total = total + w * ce
        total = total + beta * kl
        out['total'] = total
        return out, (x_reco, y_est, mu, log_var, z)


def train_step(net, opt, x, y, eps, *, beta, gamma, clip=None):
    """One reference-style optimisation step (cvae.py:2429-2461): zero_grad, evaluate, backward,
    clip_grad_norm_, step.  Used as the CPU baseline workload."""
    opt.zero_grad()
    losses, _ = net.train_losses(x, y, eps, beta=beta, gamma=gamma)
    losses['total'].mean().backward()
    if clip:
        nn.utils.clip_grad_norm_(net.parameters(), clip)
    opt.step()
    return losses


# --------------------------------------------------------------------------- arch description of a live model
def describe_seq(seq):
    """JSON-able description of an nn.Sequential made of standard torch layers (same schema as
    tests/golden/make_golden.py:describe, which produced the fixtures' `arch`)."""
    out = []
    if seq is None:
        return out
    for m in seq:
        if isinstance(m, nn.ConvTranspose2d):
            out.append(dict(t='convT', cin=m.in_channels, cout=m.out_channels, k=m.kernel_size[0], s=m.stride[0],
                            p=m.padding[0], op=m.output_padding[0]))
        elif isinstance(m, nn.Conv2d):
            out.append(dict(t='conv', cin=m.in_channels, cout=m.out_channels, k=m.kernel_size[0], s=m.stride[0],
                            p=m.padding[0]))
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
        else:
            raise TypeError(str(m))
    return out


def describe_model(model):
    """(cfg, arch) for OracleNet from any model exposing the reference's attribute names."""
    arch = {'features': describe_seq(model.features),
            'features_out': list(model.encoder.input_shape) if model.features is not None else None,
            'dense_projs': describe_seq(model.encoder.dense_projs),
            'classifier': describe_seq(getattr(model, 'classifier', None)),
            'classifier_type': model.classifier_type,
            'sampling': bool(model.encoder.sampling.is_sampled),
            'y_is_decoded': bool(model.y_is_decoded),
            'sigma': {'is_log': bool(model.sigma.is_log), 'is_rmse': bool(model.sigma.is_rmse),
                      'learned': bool(model.sigma.learned), 'sdim': int(model.sigma.sdim)},
            'prior': {'conditional': bool(model.encoder.prior.conditional), 'var_dim': model.encoder.prior.var_dim,
                      'distribution': model.encoder.prior.params['distribution'],
                      'tau': getattr(model.encoder.prior, 'tau', None)}}
    if not model.is_vib:
        arch['decoder'] = describe_seq(model.decoder)
        arch['imager'] = describe_seq(model.imager)
        arch['imager_in'] = list(model.imager.input_shape)
    cfg = {'type': model.type, 'input_shape': list(model.input_shape), 'num_labels': model.num_labels,
           'latent_dim': model.latent_dim, 'beta': model.beta, 'gamma': model.gamma or 0.0}
    return cfg, arch

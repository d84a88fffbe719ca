"""Encoder / Sampling / Classifier / Sigma with the reference's interface (module/vae_layers/layers.py:73-483).

Parameter containers are plain torch modules with the reference's state_dict names; the arithmetic runs in
libjvae_sm100.so through ...engine (dense layers on tcgen05, sampler kernel with Philox or injected noise).
"""
import logging

import numpy as np
import torch
from torch import nn
from torch.nn import Parameter

from .misc import activation_layers
from ..priors import build_prior
from ... import engine


class Sigma(Parameter):
    """Output-noise scale (layers.py:73-213): constant, learned (stored as log), rmse-tracking, decaying or
    coded by the encoder (`input_dim`); `sdim` 1 or per-pixel."""

    @staticmethod
    def __new__(cls, value=None, sdim=1, input_dim=False, learned=False, is_rmse=False, is_log=False, **kw):
        assert value is not None or is_rmse or input_dim
        if is_rmse or (input_dim and value is None):
            value = 0
        learned = learned or bool(input_dim)
        is_log = is_log or learned
        with np.errstate(divide='ignore'):      # a coded sigma starts from value 0 -> log 0 = -inf, as in the reference (layers.py:84-87)
            v = np.log(value) if is_log else value
        return super().__new__(cls, torch.zeros(sdim).fill_(v), requires_grad=learned)

    def __init__(self, value=None, learned=False, is_rmse=False, sdim=1, input_dim=False, reach=1, decay=0,
                 max_step=None, sigma0=None, is_log=False):
        assert not learned or not is_rmse
        assert not decay or not learned
        self._rmse = np.nan
        self.is_rmse = is_rmse
        self.sigma0 = value if (sigma0 is None and not is_rmse) else sigma0
        self.learned = learned
        self.input_dim = input_dim
        self.is_log = bool(learned or is_log or input_dim)
        self.decay = decay if not is_rmse else 1
        self.reach = reach if decay or is_rmse else None
        self.max_step = max_step
        self.sdim = sdim
        if self.coded:
            self._output_dim = input_dim if self.per_dim else (1,) * len(input_dim)
        else:
            self._output_dim = None

    @property
    def value(self):
        with torch.no_grad():
            if self.is_log:
                return (self.data * 2).exp().mean().sqrt().item()
            return self.data.pow(2).mean().sqrt().item()

    @property
    def coded(self):
        return bool(self.input_dim)

    @property
    def per_dim(self):
        return self.sdim != 1

    @property
    def output_dim(self):
        return self._output_dim

    @property
    def params(self):
        d = {k: v for k, v in self.__dict__.items() if not k.startswith('_')}
        d['value'] = self.value
        return d

    def update(self, rmse=None, v=None):
        assert rmse is None or v is None
        if v is not None:
            lead = tuple(range(v.dim() - self.dim()))
            self.data = v.mean(lead) if lead else v
            return
        if rmse is None:
            return
        self._rmse = rmse
        if self.learned or not self.decay:
            return
        delta = self.decay * (self.reach * rmse - self.data)
        if self.max_step and abs(delta) > self.max_step:
            delta = self.max_step if delta > 0 else -self.max_step
        self.data += delta

    # ---- text forms (job listings / logs).  Only `params`, `value` and `update` above are on the hot path.
    def _kind(self):
        """(one-letter code, description) of how this sigma is obtained"""
        if self.is_rmse:
            seen = '' if self._rmse is np.nan else ' ({:g})'.format(float(self._rmse))
            return 'e', 'rmse' + seen
        if self.coded:
            return ('C', 'coded mask') if self.per_dim else ('c', 'coded scalar')
        if self.learned:
            return 'l', '{:g}->rmse[l] ({:g})'.format(self.sigma0, self.value)
        if self.decay:
            target = ('' if self.reach == 1 else '{:g}*'.format(self.reach)) + 'rmse'
            cap = '<{:g}'.format(self.max_step) if self.max_step else ''
            return None, '{:g}->{}[-{:g}*{}]'.format(self.sigma0, target, self.decay, cap)
        return None, '{:g}'.format(float(self.data.detach().reshape(-1)[0]))

    def __str__(self):
        return self._kind()[1]

    def __format__(self, spec):
        if spec and spec[-1] in 'fge':            # numeric format: the current value
            return format(self.value, spec)
        code, text = self._kind()
        return code if (spec.endswith('i') and code) else text

    def __repr__(self):
        if self.is_rmse:
            return 'Sigma will be RMSE'
        base = super().__repr__()
        if not self.decay:
            return base
        return '{}, decaying to {}*mse with rate {})'.format(base[:-1], self.reach, self.decay)


class Sampling(nn.Module):
    """z = mean + exp(log_var / 2) * eps for sampling_size + 1 draws, draw 0 being the mean (layers.py:216-250).
    `injected_eps` (L+1,...,K), when set, replaces the Philox draws (parity / reproducibility mode)."""

    def __init__(self, latent_dim, sampling_size=1, sampling=True, distribution='gaussian', **kwargs):
        assert distribution in ('gaussian', 'uniform'), '{} for sampling unknown'.format(distribution)
        super().__init__(**kwargs)
        self.distribution = distribution
        self.sampling_size = sampling_size
        self.is_sampled = sampling
        self.injected_eps = None

    def forward(self, z_mean, z_log_var):
        head = torch.cat([z_mean.reshape(-1, z_mean.shape[-1]), z_log_var.reshape(-1, z_mean.shape[-1])], -1)
        _, _, z, eps, _ = self.from_head(head)
        L = self.sampling_size
        return z.view(L + 1, *z_mean.shape), eps.view(L, *z_mean.shape)

    def from_head(self, head):
        """head (B, 2K) = [mean | raw log_var] -> mean, clipped log_var, z (L+1,B,K), eps (L,B,K), |eps|^2 (L,B)"""
        eps = self.injected_eps
        if eps is not None:
            eps = eps.reshape(self.sampling_size + 1, head.shape[0], head.shape[1] // 2)
        return engine.sample(head, self.sampling_size, eps_in=eps, is_sampled=bool(self.is_sampled),
                             uniform=self.distribution == 'uniform')

    def __repr__(self):
        if not self.is_sampled:
            return 'Deactivated, returns mean'
        return 'Sampling({}, L={})'.format(self.distribution, self.sampling_size)


class Encoder(nn.Module):
    """x (+ one-hot y) -> dense_projs -> (mean, log_var clipped to +-20) -> Sampling (layers.py:253-403)."""

    def __init__(self, input_shape, num_labels, representation='rgb', y_is_coded=False, latent_dim=32,
                 intermediate_dims=[64], name='encoder', dropout=False, activation='relu', sampling_size=10,
                 sampling=True, sigma_output_dim=0, forced_variance=False, prior={}, **kwargs):
        super().__init__(**kwargs)
        self.name = name
        self.y_is_coded = y_is_coded
        self.input_shape = input_shape
        self.num_labels = num_labels
        self.forced_variance = forced_variance
        self._sampling_size = sampling_size
        self.latent_dim = latent_dim
        self.activation = activation

        layers = []
        d_in = int(np.prod(input_shape)) + num_labels * y_is_coded
        for d in intermediate_dims:
            layers += [nn.Linear(d_in, d), activation_layers[activation]()]
            if dropout:
                layers.append(nn.Dropout(p=dropout))
            d_in = d
        self.dense_projs = nn.Sequential(*layers)
        self.dense_mean = nn.Linear(d_in, latent_dim)
        self.dense_log_var = nn.Linear(d_in, latent_dim)
        self.sigma_output_dim = sigma_output_dim
        if sigma_output_dim:
            self.sigma = nn.Linear(d_in, int(np.prod(sigma_output_dim)))
        dist = {'tilted': 'gaussian', 'gaussian': 'gaussian', 'uniform': 'uniform'}.get(
            prior.get('distribution', 'gaussian'))
        self.sampling = Sampling(latent_dim, sampling_size, sampling, distribution=dist)
        prior['dim'] = latent_dim      # the reference mutates the caller's dict too (layers.py:306)
        self.prior = build_prior(**prior)
        logging.debug('Built %s', self.prior)

    def eval(self, *a):     # layers.py:311-312: mode switches only come through the parent's train(mode)
        print('eval', *a)

    @property
    def sampling_size(self):
        return self._sampling_size

    @sampling_size.setter
    def sampling_size(self, v):
        self._sampling_size = v
        self.sampling.sampling_size = v

    def capacity(self, m=None):
        """upper bound of I(Z;Y) from the class means (layers.py:323-336); m: a snapshot of the means"""
        m = self.prior.mean if m is None else m
        C = self.num_labels
        cdm = torch.cdist(m, m)
        return np.log(C) - 1 / C * torch.exp(-cdm.pow(2) / 4).sum(0).log().sum()

    def dict_min_distance(self, m=None):
        m = self.prior.mean if m is None else m
        C = self.num_labels
        diag = 2 * m.norm(dim=1).max() * torch.eye(C, device=m.device)
        return (torch.cdist(m, m) + diag).min()

    def forward(self, x, y=None):
        """x (..., F), y (..., C) one-hot or None -> (mean, log_var, z (L+1,...,K), eps (L,...,K), sigma|None)"""
        u = x if y is None else torch.cat((x, y), dim=-1)
        lead = u.shape[:-1]
        u = engine.run_sequential(self.dense_projs, u.reshape(-1, u.shape[-1]))
        K = self.latent_dim
        # the two heads are one GEMM on the same input: [mean | log_var]
        w = torch.cat([self.dense_mean.weight, self.dense_log_var.weight], 0)
        b = torch.cat([self.dense_mean.bias, self.dense_log_var.bias], 0)
        head = engine.linear(u, w, b, act='linear', out_dtype=torch.float32)
        if self.forced_variance:
            head = torch.cat([head[:, :K], torch.full_like(head[:, K:], float(np.log(self.forced_variance)))], -1)
        mu, lv, z, eps, _ = self.sampling.from_head(head)
        L = self.sampling_size
        sigma = None
        if self.sigma_output_dim:
            sigma = engine.linear(u, self.sigma.weight, self.sigma.bias, act='linear',
                                  out_dtype=torch.float32).view(*lead, -1)
        return mu.view(*lead, K), lv.view(*lead, K), z.view(L + 1, *lead, K), eps.view(L, *lead, K), sigma


class Classifier(nn.Sequential):
    """K -> ... -> C logits, no softmax (layers.py:456-483)"""

    def __init__(self, latent_dim, num_labels, intermediate_dims=[], name='classifier', activation='relu', **kwargs):
        layers = []
        d_in = latent_dim
        for d in intermediate_dims:
            layers += [nn.Linear(d_in, d), activation_layers[activation]()]
            d_in = d
        layers.append(nn.Linear(d_in, num_labels))
        super().__init__(*layers, **kwargs)
        self.name = name

    def forward(self, z):
        lead = z.shape[:-1]
        out = engine.run_sequential(self, z.reshape(-1, z.shape[-1]), out_dtype=torch.float32)
        return out.view(*lead, -1)

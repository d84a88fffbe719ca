"""ClassificationVariationalNetwork with the reference's interface (cvae.py:60-2866), B200-native underneath.

Constructor keywords, attribute names, state_dict keys, the return tuples of forward / evaluate, the loss-dict
keys per model type, predict_after_evaluate and batch_dist_measures follow the reference, so a caller of the
reference (train.py:195-217,333; test.py:272-304) can switch modules.  Underneath:

  features / dense_projs / heads / decoder / imager / classifier -> engine (tcgen05 GEMM / implicit-GEMM kernels)
  Sampling            -> csrc/sampler.cu (Philox or injected noise, fused mean/log-var head)
  the ELBO step       -> csrc/elbo.cu: ONE fused kernel for train forward, one for its backward, one for the
                         per-class eval that also emits OOD scores and predictions (cvae.py:626-1085)
  clip + Adam         -> csrc/optim.cu on a flat buffer

Host syncs of the reference's step (8 x .item() in evaluate, the per-parameter isnan scan, cvae.py:689-724,
2454-2457) are gone: `total_measures` is a lazy mapping read with one sync when accessed, and non-finite losses
raise a device flag that train_step polls every `check_every` steps.
"""
import json
import logging
import math
import os
from collections.abc import Mapping

import numpy as np
import torch
from torch import nn

from . import _native as nat
from . import engine
from .module.optimizers import Optimizer
from .module.vae_layers import (Classifier, Encoder, Sigma, activation_layers, build_de_conv_layers, find_input_shape,
                                onehot_encoding)

DEFAULT_ACTIVATION = 'relu'
DEFAULT_OUTPUT_ACTIVATION = 'linear'
DEFAULT_LATENT_SAMPLING = 100
VERSION = 2.


class _LazyMeasures(Mapping):
    """total_measures of the reference (python floats, cvae.py:622-762) computed on first access with ONE sync.
    `make(batch, current)` builds the thunk for a batch index and the previous running measures, so the same device-side
    sources can be read again after a replayed step (rebind)."""

    def __init__(self, names, make, batch=0, current=None):
        self._names, self._make, self._vals = list(names), make, None
        self._thunk = make(batch, current or {})

    def rebind(self, batch, current):
        return _LazyMeasures(self._names, self._make, batch, current)

    def _force(self):
        if self._vals is None:
            self._vals = self._thunk()
        return self._vals

    def __getitem__(self, k):
        return self._force()[k]

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)

    def __repr__(self):
        return repr(dict(self._force()))


def _develop_starred(methods, methods_params):
    """utils/save_load/dictify.py:198-212 (in place): 'softkl*' -> softkl-1, softkl-2, ... (one entry per temperature)"""
    starred = []
    for m in list(methods):
        if m.endswith('*'):
            methods += methods_params.get(m[:-1], [])
            starred.append(m)
    for m in starred:
        methods.remove(m)
    return methods


def _make_list(o, default_for_all):
    """utils/misc.py:1-13"""
    if isinstance(o, str):
        o = [o]
    if o is None:
        return []
    if o and o[0] in ('all', 'default'):
        return type(default_for_all)(default_for_all)
    if o and o[0] == 'first':
        return [next(iter(default_for_all))]
    return o


class ClassificationVariationalNetwork(nn.Module):
    """X -- features -- encoder -- Z -- decoder -- imager -- X^ ;  Z -- classifier -- Y^   (cvae.py:60-81)"""

    loss_components_per_type = {'jvae': ('cross_x', 'kl', 'cross_y', 'total'),
                                'cvae': ('cross_x', 'kl', 'total', 'zdist', 'var_kl', 'dzdist', 'iws', 'sigma', 'wmse',
                                         'z_logdet', 'z_tr_inv_cov'),
                                'xvae': ('cross_x', 'kl', 'total', 'zdist', 'iws'),
                                'vae': ('cross_x', 'kl', 'zdist', 'var_kl', 'total', 'iws'),
                                'vib': ('cross_y', 'kl', 'total')}
    predict_methods_per_type = {'jvae': ['loss', 'esty'], 'cvae': ['iws', 'closest'], 'xvae': ['loss', 'closest'],
                                'vae': [], 'vib': ['esty']}
    metrics_per_type = {'jvae': ['rmse', 'dB', 'sigma'], 'cvae': ['rmse', 'dB', 'd-mind', 'ld-norm', 'sigma'],
                        'xvae': ['rmse', 'dB', 'zdist', 'd-mind', 'ld-norm', 'sigma'], 'vae': ['rmse', 'dB', 'sigma'],
                        'vib': ['sigma']}
    ood_methods_per_type = {'cvae': ['iws-2s', 'iws-a-1-1', 'iws-a-4-1', 'iws', 'mse', 'elbo', 'soft', 'elbo-2s',
                                     'elbo-a-1-1', 'elbo-a-4-1', 'zdist'],
                            'xvae': ['max', 'mean', 'std'], 'jvae': ['max', 'sum', 'std'],
                            'vae': ['iws', 'iws-2s', 'iws-a-1-1', 'iws-a-4-1', 'elbo', 'elbo-2s', 'elbo-a-1-1',
                                    'elbo-a-4-1', 'zdist'],
                            'vib': ['odin*', 'baseline', 'logits']}
    misclass_methods_per_type = {'cvae': ['softkl*', 'iws', 'softiws*', 'kl', 'max', 'zdist', 'softzdist*',
                                          'baseline*', 'hyz'],
                                 'xvae': [], 'jvae': [], 'vae': [], 'vib': ['odin*', 'baseline', 'logits', 'hyz']}
    ODIN_TEMPS = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000]
    ODIN_EPS = [0.0002 * i for i in range(21)]
    odin_params = []
    for _T in ODIN_TEMPS:
        for _e in ODIN_EPS:
            odin_params.append('odin-{:.0f}-{:.4f}'.format(_T, _e))
    methods_params = {}
    for _k in ('softkl', 'softzdist', 'baseline'):
        methods_params[_k] = []
        for _T in ODIN_TEMPS:
            methods_params[_k].append(f'{_k}-{_T:.0f}')
    methods_params['odin'] = odin_params

    def __init__(self, input_shape, num_labels, type='cvae', y_is_coded=False, output_distribution='gaussian',
                 job_number=0, features=None, pretrained_features=None, batch_norm=False, dropout=False,
                 encoder=[36], latent_dim=32, prior={}, beta=1., gamma=0., decoder=[36], upsampler=None,
                 pretrained_upsampler=None, classifier=[36], name='joint-vae', activation=DEFAULT_ACTIVATION,
                 latent_sampling=DEFAULT_LATENT_SAMPLING, test_latent_sampling=None, encoder_forced_variance=False,
                 output_activation=DEFAULT_OUTPUT_ACTIVATION, sigma={'value': 1}, optimizer={}, shadow=False,
                 representation='rgb', version=VERSION, *args, **kw):
        super().__init__(*args, **kw)
        assert type in ('jvae', 'cvae', 'xvae', 'vib', 'vae')
        self.name, self.job_number, self.type = name, job_number, type
        self.loss_components = self.loss_components_per_type[type]
        self.metrics = self.metrics_per_type[type]
        self.predict_methods = self.predict_methods_per_type[type].copy()
        self.ood_methods = self.ood_methods_per_type[type].copy()
        self.misclass_methods = self.misclass_methods_per_type[type].copy()
        self.is_jvae, self.is_vib, self.is_vae = type == 'jvae', type == 'vib', type == 'vae'
        self.is_cvae, self.is_xvae = type == 'cvae', type == 'xvae'
        assert not (y_is_coded and (self.is_vib or self.is_vae))
        self.y_is_coded = y_is_coded
        self.y_is_decoded = gamma if (self.is_cvae or self.is_vae) else True
        self.x_is_generated = not self.is_vib
        self.output_distribution = output_distribution if self.x_is_generated else None
        self.losses_might_be_computed_for_each_class = not self.is_vae and not self.is_vib
        self._test_losses, self._test_measures, self._measures = {}, {}, {}

        if self.y_is_decoded:
            self.classifier_type = 'linear'
            if self.is_cvae and classifier and isinstance(classifier[0], str):
                assert classifier[0] in ('softmax',)
                self.classifier_type = classifier[0]
        else:
            self.classifier_type = None
            classifier = []
        if not self.x_is_generated:
            decoder, upsampler = [], None
        if self.y_is_decoded and 'esty' not in self.predict_methods:
            self.predict_methods = self.predict_methods + ['esty']
        if self.y_is_decoded and 'cross_y' not in self.loss_components:
            self.loss_components += ('cross_y',)

        assert not upsampler or features      # no upsampler without features (cvae.py:233)
        if not features:
            batch_norm = False
            bn_enc = bn_dec = False
        else:
            bn_enc = batch_norm in ('encoder', 'both')
            bn_dec = batch_norm == 'both'
        if features:
            self.features = build_de_conv_layers(input_shape, features, activation=activation, batch_norm=bn_enc,
                                                 pretrained_dict=pretrained_features)
            encoder_input_shape = self.features.output_shape
        else:
            self.features = None
            encoder_input_shape = input_shape

        self.trained = 0
        if isinstance(sigma, Sigma):
            self.sigma = sigma
        elif isinstance(sigma, dict):
            self.sigma = Sigma(**sigma)
        else:
            self.sigma = Sigma(value=sigma)
        test_latent_sampling = test_latent_sampling or latent_sampling
        self.beta = beta
        self.gamma = gamma if self.y_is_decoded else None
        if type in ('cvae', 'xvae'):
            prior['num_priors'] = num_labels           # the reference mutates the caller's dict (cvae.py:273-274)
        sampling = latent_sampling > 1 or beta > 0

        self.encoder = Encoder(encoder_input_shape, num_labels, intermediate_dims=encoder, latent_dim=latent_dim,
                               y_is_coded=y_is_coded, dropout=dropout,
                               sigma_output_dim=self.sigma.output_dim if self.sigma.coded else 0,
                               forced_variance=encoder_forced_variance, sampling_size=latent_sampling, prior=prior,
                               activation=activation, sampling=sampling)

        if self.x_is_generated:
            layers, d_in = [], latent_dim
            for d_out in decoder:
                layers += [nn.Linear(d_in, d_out), activation_layers[activation]()]
                if dropout:
                    layers.append(nn.Dropout(p=dropout))
                d_in = d_out
            self.decoder = nn.Sequential(*layers)
            imager_input_dim = d_in
            if upsampler:
                hw = find_input_shape(upsampler, input_shape[1:])
                f = hw[0] * hw[1]
                assert not imager_input_dim % f, 'Could not go from {} to *, {} {}'.format(imager_input_dim, *hw)
                imager_input_dim = (imager_input_dim // f, *hw)
                self.imager = build_de_conv_layers(imager_input_dim, upsampler, batch_norm=bn_dec, activation=activation,
                                                   output_activation=output_activation,
                                                   output_distribution=self.output_distribution,
                                                   pretrained_dict=pretrained_upsampler, where='output')
            else:
                f = 1 if self.output_distribution == 'gaussian' else 256
                upsampler = None
                self.imager = nn.Sequential(nn.Linear(imager_input_dim, f * int(np.prod(input_shape))),
                                            activation_layers[output_activation]())
                self.imager.input_shape = (imager_input_dim,)

        if self.classifier_type in ('linear', None):
            self.classifier = Classifier(latent_dim, num_labels, classifier, activation=activation)

        self.input_shape = tuple(input_shape)
        self.num_labels = num_labels
        self.input_dim = len(input_shape)
        self.batch_norm, self.dropout = batch_norm, dropout
        self._sizes_of_layers = [input_shape, num_labels, encoder, latent_dim, decoder, upsampler, classifier]
        self.architecture = {'input_shape': input_shape, 'num_labels': num_labels,
                             'output_distribution': self.output_distribution, 'type': type,
                             'representation': representation, 'encoder': encoder, 'batch_norm': batch_norm,
                             'dropout': dropout, 'activation': activation,
                             'encoder_forced_variance': self.encoder.forced_variance, 'latent_dim': latent_dim,
                             'test_latent_sampling': test_latent_sampling, 'prior': self.encoder.prior.params,
                             'decoder': decoder, 'upsampler': upsampler, 'classifier': classifier,
                             'output_activation': output_activation, 'version': VERSION}
        lin = self.classifier_type == 'linear'
        self.depth = (len(encoder) + len(decoder) + len(classifier)) if lin else 0
        self.width = (sum(encoder) + sum(decoder) + sum(classifier)) if lin else 0
        if features:
            self.architecture['features'] = self.features.name
        self.training_parameters = {'sigma': self.sigma.params, 'beta': beta, 'gamma': self.gamma,
                                    'latent_sampling': latent_sampling, 'set': None, 'data_augmentation': [],
                                    'pretrained_features': getattr(pretrained_features, 'name', None),
                                    'pretrained_upsampler': getattr(pretrained_upsampler, 'name', None), 'epochs': 0,
                                    'batch_size': None, 'fine_tuning': []}
        self.testing = {0: {m: {'n': 0, 'epochs': 0, 'accuracy': 0} for m in self.predict_methods}}
        self.ood_results = {}
        self.optimizer = Optimizer(self.parameters(), **optimizer)
        self.training_parameters['optimizer'] = self.optimizer.params
        self.train_history = {'epochs': 0}
        self.latent_dim = latent_dim
        self.latent_sampling = latent_sampling
        self._latent_samplings = {'train': latent_sampling, 'eval': test_latent_sampling}
        self.encoder_layer_sizes, self.decoder_layer_sizes = encoder, decoder
        self.classifier_layer_sizes = classifier
        self.upsampler, self.activation, self.output_activation = upsampler, activation, output_activation
        self.z_output = False
        self._fused = None          # scores / predictions of the last per-class evaluate
        self._steps = 0
        self.eval()

    # ------------------------------------------------------------------------------------------ modes
    def train(self, *a, **k):
        super().train(*a, **k)
        self.latent_sampling = self._latent_samplings['train' if self.training else 'eval']
        return self

    @property
    def latent_sampling(self):
        return self._latent_sampling

    @latent_sampling.setter
    def latent_sampling(self, v):
        self._latent_sampling = v
        self.encoder.sampling_size = v

    def to(self, d):
        super().to(d)
        self.optimizer.to(d)
        return self

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x, y=None, x_features=None, **kw):
        """cvae.py:426-455.  x (N1..Ng, *input_shape), y (N1..Ng) -> see forward_from_features"""
        if y is None and self.y_is_coded:
            raise ValueError('y is supposed to be an input of the net')
        batch_shape = (1,) if x.dim() == self.input_dim else x.shape[:-self.input_dim]
        if not self.features:
            x_features = x
        if x_features is None:
            t = engine.run_sequential(self.features, x.reshape(-1, *self.input_shape))
            x_features = t.reshape(*batch_shape, *self.encoder.input_shape)
        return self.forward_from_features(x_features, None if y is None else y.view(*batch_shape), x, **kw)

    def forward_from_features(self, x_features, y, x, z_output=True, sampling_epsilon_norm_out=False, sigma_out=False,
                              decode=True):
        """cvae.py:455-521 -> (x_reco (L+1,..,*shape), y_out (L+1,..,C)[, mu, log_var, z][, |eps|^2][, sigma_coded])"""
        nf = len(self.encoder.input_shape)
        batch_size = x_features.shape[:-nf]
        reco_shape = tuple(batch_size) + ((256,) if self.output_distribution == 'categorical' else ()) + self.input_shape
        x_ = x_features.reshape(*batch_size, -1)
        y_onehot = onehot_encoding(y, self.num_labels).float() if (y is not None and self.y_is_coded) else None
        mu, log_var, z, eps, sigma = self.encoder(x_, y_onehot)
        L1 = self.latent_sampling + 1
        K = self.latent_dim
        z2 = z.reshape(-1, K)
        if not self.is_vib and decode:
            u = engine.run_sequential(self.decoder, z2)
            xr = engine.run_sequential(self.imager, u.reshape(-1, *self.imager.input_shape), image_out=True)
        if self.classifier_type in ('linear', None):
            y_output = self.classifier(z)
        else:   # 'softmax': z.m^T + |m|^2/2 (cvae.py:499)
            m = self.encoder.prior.mean
            y_output = engine.linear(z2, m, m.pow(2).sum(-1) / 2, out_dtype=torch.float32).view(*z.shape[:-1], -1)
        out = (x,) if self.is_vib else ((xr.view(L1, *reco_shape),) if decode else (None,))
        out += (y_output,)
        if z_output:
            out += (mu, log_var, z)
        if sampling_epsilon_norm_out:
            out += ((eps ** 2).sum(-1),)
        if sigma_out:
            out += (sigma,)
        return out

    # ------------------------------------------------------------------------------------------ the ELBO step
    def _elbo_cfg(self, B, L, x_reco, logits, beta, gamma_w, var_w):
        prior = self.encoder.prior
        if self.sigma.per_dim:
            # the reference adds the per-pixel log sigma (*input_shape) to the per-sample wmse (B,) (cvae.py:773-775): that only
            # broadcasts when the batch happens to match the image width, so there is no contract to reproduce
            raise NotImplementedError('per-pixel sigma (sdim != 1) is not supported')
        D = int(np.prod(self.input_shape)) if x_reco is not None else 0
        cat = self.output_distribution == 'categorical' and x_reco is not None
        cat_group = 0
        if cat:
            if self.sigma.is_rmse:
                raise NotImplementedError('categorical output with sigma=rmse')
            # 256 logits per pixel variable: channels_last conv output (channel = v * C + c) or the reference's (256, *shape)
            cat_group = self.input_shape[0] if (x_reco.dim() == 4 and not x_reco.is_contiguous()) else D
        return nat.make_cfg(categorical=cat, cat_group=cat_group, sigma_per_sample=bool(self.sigma.coded) and x_reco is not None,
                            B=B, L=L, K=self.latent_dim, C=self.num_labels, D=D, x_reco=x_reco, logits=logits,
                            var_dim=prior.var_dim, prior_kind=prior.distribution, conditional=prior.conditional,
                            sigma_is_log=self.sigma.is_log, sigma_is_rmse=self.sigma.is_rmse, beta=beta, gamma_w=gamma_w,
                            var_w=var_w, tau=getattr(prior, 'tau', 0.0), alpha=getattr(prior, '_alpha', 0.0))

    @staticmethod
    def _dense(t):
        """tensor usable by the kernels as a flat buffer: row-major or channels_last dense"""
        if t.is_contiguous():
            return t
        if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
            return t
        return t.contiguous()

    def evaluate(self, x, y=None, batch=0, current_measures=None, with_beta=False, kl_var_weighting=1.,
                 gamma_weighting=1, z_output=False, **kw):
        """cvae.py:523-917.  Returns (x_reco, logits (..,C), batch_losses, total_measures[, mu, log_var, z])."""
        if not x.is_cuda:
            raise nat.NativeError('joint-vae_b200 runs on CUDA devices only (there is no CPU fallback)')
        y_in_input = y is not None
        per_class = self.losses_might_be_computed_for_each_class and not y_in_input
        if self.y_is_coded and not y_in_input:
            # the reference itself raises here (cvae.py:451 reshapes the C-replicated y to x's batch shape): no contract to mirror
            raise NotImplementedError('y_is_coded with per-class evaluation (x replicated C times, cvae.py:589-591)')
        if x.dim() != self.input_dim + 1:
            raise NotImplementedError('evaluate expects one batch dimension: x of shape (B, *input_shape)')
        cross_y_weight = False
        if self.y_is_decoded:
            cross_y_weight = gamma_weighting * self.gamma
            if (self.is_cvae or self.is_vae) and not self.training:
                cross_y_weight = False
        B = x.shape[0]
        L = self.latent_sampling
        # prior statistics (dictionary mean / variance, log-dets) depend on the prior parameters only: launch them
        # before the network so the loss step itself is one kernel (include/jvae_b200.h: jvae_elbo_prior_stats)
        prior0 = self.encoder.prior
        means0, inv_trans0 = prior0.mean, prior0.inv_trans
        pcfg = nat.make_cfg(B=B, L=L, K=self.latent_dim, C=self.num_labels, D=0, x_reco=None, logits=None,
                            var_dim=prior0.var_dim, prior_kind=prior0.distribution, conditional=prior0.conditional,
                            sigma_is_log=False, sigma_is_rmse=False, beta=1.0, gamma_w=0.0, var_w=1.0)
        nat.elbo_prior_stats(pcfg, means0.detach().contiguous(), inv_trans0.detach().contiguous())
        prior_ready = True
        o = self.forward(x, y=y if self.y_is_coded else None, sampling_epsilon_norm_out=True, sigma_out=True, **kw)
        x_reco, y_est, mu, log_var, z, eps_norm, sigma_coded = o
        prior = self.encoder.prior
        means, inv_trans = means0, inv_trans0
        beta = self.beta if with_beta else 1.

        xr_k = x_k = None
        if self.x_is_generated:
            if self.output_distribution == 'categorical':      # (L+1, B, 256, C, H, W): 256 C channels per pixel
                xr4 = x_reco.reshape(-1, 256 * self.input_shape[0], *self.input_shape[1:])
            else:
                xr4 = x_reco.reshape(-1, *self.input_shape)
            xr_k = self._dense(xr4)
            if xr_k.dim() == 4 and not xr_k.is_contiguous():      # channels_last reconstruction: same order for x
                x_k = x.float().contiguous(memory_format=torch.channels_last)
            else:
                x_k = x.float().contiguous()
            if xr_k.dtype not in (torch.float32, torch.bfloat16):
                xr_k = xr_k.float()
        logits_k = y_est.float().contiguous() if self.y_is_decoded else None
        gw = float(cross_y_weight) if cross_y_weight else 0.0
        cfg = self._elbo_cfg(B, L, xr_k, logits_k, beta, gw, kl_var_weighting)
        cfg.prior_stats_ready = int(prior_ready)
        sig = self.sigma if self.x_is_generated else None
        # the `sigma` measure is read before any update (cvae.py:624); the decaying / rmse update is in place: snapshot then
        # ... and so is the optimizer's (flat parameter buffer): in training the lazy measures read snapshots
        sigma_seen = self.sigma.data.clone() if self.training else self.sigma.data
        if self.x_is_generated and self.sigma.coded:       # cvae.py:631-634: one log sigma per sample from the encoder's head
            sig = sigma_coded.reshape(-1).float().contiguous()
            self.sigma.update(v=sigma_coded.detach().reshape(-1, *self.sigma.output_dim))
        batch_losses = {}
        self._fused = None

        if y_in_input and self.training:
            yk = y.reshape(-1).long().contiguous()
            kl, zdist, var_kl, wmse, cross_x, cross_y, total, dzdist, finite = engine.elbo_train(
                cfg, x_k, xr_k, mu.contiguous(), log_var.contiguous(), logits_k, yk, means, inv_trans, sig)
            self._finite_flag = finite
            res = dict(kl=kl, zdist=zdist, var_kl=var_kl, total=total, dzdist=dzdist, wmse=wmse, cross_x=cross_x,
                       cross_y=cross_y)
        else:
            with torch.no_grad():
                r = engine.elbo_eval(cfg, x_k, xr_k, mu.contiguous(), log_var.contiguous(), z.contiguous(),
                                     eps_norm.contiguous(), logits_k, means, inv_trans, sig,
                                     want_iws='iws' in self.loss_components, want_scores=not y_in_input)
            res = dict(r)
            if y_in_input:      # eval mode with labels: keep the row of the given class (cvae.py:548-556 with y given)
                yk = y.reshape(1, -1).long()
                for k in ('kl', 'zdist', 'var_kl', 'iws', 'total', 'cross_y'):
                    if res.get(k) is not None and res[k].shape[0] > 1:
                        res[k] = res[k].gather(0, yk).squeeze(0)
                    elif res.get(k) is not None:
                        res[k] = res[k].squeeze(0)
            else:
                if not prior.conditional:
                    for k in ('kl', 'zdist', 'var_kl', 'iws'):
                        if res.get(k) is not None:
                            res[k] = res[k].squeeze(0)
                    if res['total'].shape[0] == 1:
                        res['total'] = res['total'].squeeze(0)
                self._fused = {'scores': r['scores'], 'preds': r['preds'], 'total': res['total']}
        mse_rmse = None
        if self.x_is_generated and self.sigma.is_rmse:
            # sigma^2 := per-sample mse (cvae.py:662-670); recovered from cross_x = D/2 (log mse + 1 + log 2 pi)
            D_ = float(np.prod(self.input_shape))
            mse_rmse = (2 * res['cross_x'].detach() / D_ - 1 - math.log(2 * math.pi)).exp()
        if self.training and self.x_is_generated and self.sigma.decay and not self.sigma.learned:
            # cvae.py:768-771: a decaying / rmse-tracking sigma moves towards reach * rmse of the batch after every training
            # evaluate (no sync: the update stays on the device unless a max_step is set)
            if mse_rmse is not None:
                mse_b = mse_rmse
            else:
                sd = self.sigma.data.float().reshape(-1)[:1]
                mse_b = res['wmse'].detach() * ((2 * sd).exp() if self.sigma.is_log else sd ** 2)
            self.sigma.update(rmse=mse_b.mean().sqrt())
        # same key order as the reference's dict (cvae.py:726-902)
        for k in ('kl', 'zdist', 'var_kl', 'total'):
            batch_losses[k] = res[k]
        if prior.conditional:
            batch_losses['dzdist'] = res['dzdist']
        if self.x_is_generated:
            batch_losses['wmse'] = res['wmse']
            batch_losses['cross_x'] = res['cross_x']
            if not self.training and 'iws' in self.loss_components and res.get('iws') is not None:
                batch_losses['iws'] = res['iws']
        if self.y_is_decoded:
            batch_losses['cross_y'] = res['cross_y']

        logits_out = res['logits'] if res.get('logits') is not None else y_est[1:].mean(0)
        measures = self._lazy_measures(x, batch_losses, batch, current_measures, sigma_seen,
                                       sig.detach() if (self.x_is_generated and self.sigma.coded) else None, mse_rmse,
                                       means0.detach().clone() if self.training else means0.detach())
        out = (x_reco, logits_out, batch_losses, measures)
        if z_output:
            out += (mu, log_var, z)
        return out

    def _lazy_measures(self, x, losses, batch, current, sigma_data=None, sigma_coded=None, mse_rmse=None, means=None):
        """the reference computes these floats inside evaluate (cvae.py:622-762), i.e. BEFORE the optimizer step that
        follows a training evaluate; read lazily they must come from snapshots of sigma and of the class means"""
        names = ['sigma']
        if self.x_is_generated:
            names += ['xpow', 'mse', 'rmse', 'dB']
        names += ['zdist', 'var_kl']
        conditional = self.encoder.prior.conditional
        if conditional:
            names += ['ld-norm', 'imut-zy', 'd-mind']
        def make(batch, current):
          def thunk():
            with torch.no_grad():
                sd = (self.sigma.data if sigma_data is None else sigma_data).float().reshape(-1)[:1]
                dev = [sd if not self.sigma.is_log else sd.exp(),
                       losses['zdist'].mean().reshape(1), losses['var_kl'].mean().reshape(1)]
                if self.x_is_generated:
                    s2 = 1.0 if self.sigma.is_rmse else dev[0] ** 2
                    if sigma_coded is not None:       # per-sample sigma^2 (cvae.py:644-645, 674)
                        s2 = (2 * sigma_coded).exp() if self.sigma.is_log else sigma_coded ** 2
                    if mse_rmse is not None:          # sigma = rmse: mse = wmse * sigma^2 with wmse = 1 (cvae.py:662-674)
                        s2 = mse_rmse
                    dev += [x.float().pow(2).mean().reshape(1), (losses['wmse'] * s2).mean().reshape(1)]
                if conditional:
                    m = self.encoder.prior.mean if means is None else means
                    dev += [m.pow(2).mean().reshape(1), self.encoder.capacity(m).reshape(1),
                            self.encoder.dict_min_distance(m).reshape(1)]
                v = torch.cat([d.float() for d in dev]).tolist()      # the single sync
            run = lambda k, val: (current.get(k, 0.) * batch + val) / (batch + 1)
            out = {'sigma': v[0], 'zdist': run('zdist', v[1]), 'var_kl': run('var_kl', v[2])}
            i = 3
            if self.x_is_generated:
                out['xpow'] = run('xpow', v[3])
                out['mse'] = run('mse', v[4])
                out['rmse'] = math.sqrt(out['mse'])
                out['dB'] = 10 * math.log10(out['xpow'] / out['mse']) if out['mse'] > 0 else float('inf')
                i = 5
            if conditional:
                out['ld-norm'], out['imut-zy'], out['d-mind'] = v[i], v[i + 1], v[i + 2]
            return out
          return thunk

        return _LazyMeasures(names, make, batch, current)

    # ------------------------------------------------------------------------------------------ predictions / scores
    def predict(self, x, method=None, **kw):
        _, logits, losses, _ = self.evaluate(x)
        return self.predict_after_evaluate(logits, losses, method=method or self.predict_methods[0])

    def predict_after_evaluate(self, logits, losses, method='default'):
        """cvae.py:938-970; reads the kernel's fused arg-min/arg-max when `losses` is the last evaluate's dict"""
        if method == 'default':
            method = self.predict_methods[0]
        f = self._fused
        if f is not None and f['preds'] is not None and losses.get('total') is f['total'] and method in nat.PRED_INDEX:
            return f['preds'][:, nat.PRED_INDEX[method]].long()
        if method is None:
            return logits.softmax(-1)
        if method == 'mean':
            return logits.softmax(-1).mean(0).argmax(-1)
        if method == 'loss':
            return losses['total'].argmin(0)
        if method == 'esty':
            return logits.argmax(-1)
        if method == 'foo':
            return logits.argmin(-1)
        if method == 'closest':
            return losses['zdist'].argmin(0)
        if method == 'iws':
            return losses['iws'].argmax(0)
        if method == 'already':
            return losses['y_est_already']
        raise ValueError(f'Unknown method {method}')

    _FUSED_SCORES = {'cvae': {'elbo', 'max', 'sum', 'mean', 'iws', 'soft', 'softkl', 'zdist', 'kl', 'mse', 'wmse',
                              'logits', 'baseline', 'hyz', 'std', 'softiws'}}

    def batch_dist_measures(self, logits, losses, methods, to_cpu=False):
        """cvae.py:972-1085: per-sample OOD / misclassification scores from the per-class losses.  For a cvae the
        default methods were already reduced inside the eval kernel (csrc/elbo.cu); the others are small tensor ops
        on the (C,B) outputs."""
        out = {}
        C = self.num_labels
        per_class = self.losses_might_be_computed_for_each_class
        f = self._fused if (self._fused is not None and losses.get('total') is self._fused['total']) else None
        fused_ok = self._FUSED_SCORES.get(self.type, set()) if f is not None and f['scores'] is not None else set()
        lazy = {}

        def get(name):
            if name not in lazy:
                logp = -losses['total']
                if name == 'logp':
                    lazy[name] = logp
                elif name == 'logp_max':
                    lazy[name] = logp.max(0)[0] if logp.dim() > 1 else logp
                elif name == 'd_logp':
                    lazy[name] = logp - get('logp_max')
                elif name == 'iws':
                    lazy[name] = losses['iws'] if 'iws' in losses else -losses['total']
                elif name == 'iws_max':
                    lazy[name] = get('iws').max(0)[0]
            return lazy[name]

        for m_ in methods:
            m = m_[:-3] if m_.endswith('-2s') else m_
            if '-a-' in m:
                m = m.split('-')[0]
            if m in fused_ok and (m not in ('logits', 'baseline', 'hyz') or self.y_is_decoded):
                v = f['scores'][:, nat.SCORE_INDEX[m]]
            elif m == 'elbo':
                v = get('logp_max') if per_class else get('logp')
            elif m == 'iws':
                if per_class:
                    v = (get('iws') - get('iws_max')).exp().sum(0).log() + get('iws_max')
                    if not self.is_jvae:
                        v = v + np.log(C)
                else:
                    v = get('iws')
            elif m == 'sum':
                v = get('d_logp').exp().sum(0).log() + get('logp_max')
            elif m == 'max':
                v = get('logp_max')
            elif m == 'softiws':
                v = losses['iws'].softmax(0).max(0)[0]
            elif m.startswith('softiws-'):
                v = (-losses['iws'] / float(m[8:])).softmax(0).max(0)[0]
            elif m in ('soft', 'softkl'):
                v = (-losses['kl']).softmax(0).max(0)[0]
            elif m.startswith('softkl-'):
                v = (-losses['kl'] / float(m[7:])).softmax(0).max(0)[0]
            elif m in ('zdist', 'kl', 'fisher_rao', 'mahala', 'kl_rec'):
                v = -losses[m] if self.is_vae else (-losses[m]).max(0)[0]
            elif m.startswith('soft') and '-' in m:
                v = (-losses[m.split('-')[0][4:]] / float(m.split('-')[-1])).softmax(0).max(0)[0]
            elif m == 'logits':
                v = logits.max(-1)[0]
            elif m.startswith('baseline'):
                T = float(m.split('-')[-1]) if '-' in m else 1
                v = (logits / T).softmax(-1).max(-1)[0]
            elif m == 'mag':
                v = get('logp_max') - get('logp').median(0)[0]
            elif m == 'std':
                v = get('logp').std(0)
            elif m == 'mean':
                v = get('d_logp').exp().mean(0).log() + get('logp_max')
            elif m == 'nstd':
                e = get('d_logp').exp()
                v = (e.std(0).log() - e.mean(0).log()).exp().pow(2)
            elif m == 'hyz':
                p = logits.softmax(-1)
                v = (p * p.log()).sum(-1)
            elif m == 'IYx':
                d = get('d_logp')
                dx = d.exp().mean(0).log()
                v = (d * d.exp()).sum(0) / (C * dx.exp()) - dx
            elif m == 'mse' and self.is_cvae:
                v = -losses['cross_x']
            elif m == 'wmse' and self.is_cvae:
                v = -losses['wmse']
            elif m.startswith('odin'):
                v = losses[m]
            else:
                raise ValueError(f'{m} is an unknown ood method')
            out[m_] = v.cpu() if to_cpu else v
        return out

    # ------------------------------------------------------------------------------------------ training
    def train_step(self, x, y, kl_var_weighting=1., gamma_weighting=1., check_every=0, batch=0, current_measures=None,
                   graph=False):
        """One optimisation step = the body of the reference's batch loop (cvae.py:2427-2461): zero_grad, evaluate with
        beta (batch index and the epoch's running measures passed through, cvae.py:2441-2449), backward of total.mean(),
        clip, Adam.  No host sync unless check_every divides the step count, in which case the device-side finite flag
        replaces the reference's per-parameter isnan scan (cvae.py:2454-2457).
        graph=True: the same step replayed from a CUDA graph (captured on the third call with a given batch shape /
        weighting / learning rate; the first two run eagerly), which removes the ~180 kernel launches' host cost per step.
        Every per-step quantity lives on the device (Adam step counts, Philox position, BatchNorm counters), so a replay
        is the same arithmetic as an eager step."""
        if graph and self._graph_ok(x):
            return self._train_step_graphed(x, y, kl_var_weighting, gamma_weighting, check_every, batch, current_measures)
        return self._train_step_eager(x, y, kl_var_weighting, gamma_weighting, check_every, batch, current_measures)

    def _train_step_eager(self, x, y, kl_var_weighting, gamma_weighting, check_every, batch, current_measures):
        self.optimizer.zero_grad()
        _, _, losses, measures = self.evaluate(x, y, batch=batch, current_measures=current_measures, with_beta=True,
                                               kl_var_weighting=kl_var_weighting, gamma_weighting=gamma_weighting)
        loss = losses['total'].mean()
        loss.backward()
        self.optimizer.clip(self.parameters())
        self.optimizer.step()
        self._steps += 1
        if check_every and self._steps % check_every == 0 and int(self._finite_flag.item()) == 0:
            raise FloatingPointError('non-finite loss at step {}'.format(self._steps))
        return losses, measures

    def _graph_ok(self, x):
        """steps a replayed graph reproduces: Philox noise (an injected tensor would be frozen into the graph), the fused
        Adam (its state is on the device), sigma not updated from the host; the data-parallel step is captured too when its
        all-reduce runs on NCCL (the collective becomes a node of the graph; every rank captures at the same call)"""
        opt = self.optimizer
        dp_ok = opt.allreduce is None or getattr(opt, 'allreduce_capturable', False)
        return (x.is_cuda and self.encoder.sampling.injected_eps is None and opt._opt is None and dp_ok
                and not (self.sigma.decay and not self.sigma.learned) and not self.sigma.coded)

    def _train_step_graphed(self, x, y, kl_var_weighting, gamma_weighting, check_every, batch, current_measures):
        key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype, float(kl_var_weighting), float(gamma_weighting),
               float(self.optimizer.lr), self.latent_sampling, str(x.device), self.optimizer.grad_clipping)
        graphs = self.__dict__.setdefault('_graphs', {})
        st = graphs.setdefault(key, {'calls': 0})
        st['calls'] += 1
        if 'graph' not in st:
            # The warm-up steps and the capture share ONE side stream: autograd's AccumulateGrad nodes remember the stream they
            # were created on, and a node created on the default stream would pull the capture across streams.
            side = self.__dict__.setdefault('_graph_stream', None) or torch.cuda.Stream(device=x.device)
            self._graph_stream = side
            cur = torch.cuda.current_stream(x.device)
            side.wait_stream(cur)
            if st['calls'] <= 2:       # eager: every lazy initialisation (flat buffers, packing tables, workspaces) happens here
                with torch.cuda.stream(side):
                    out = self._train_step_eager(x, y, kl_var_weighting, gamma_weighting, check_every, batch, current_measures)
                cur.wait_stream(side)
                return out
            sx, sy = x.clone(), y.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):      # records the step; nothing executes yet
                losses, measures = self._train_step_eager(sx, sy, kl_var_weighting, gamma_weighting, 0, 0, None)
            self._steps -= 1
            st.update(graph=g, x=sx, y=sy, losses=losses, measures=measures, flag=self._finite_flag)
        st['x'].copy_(x, non_blocking=True)
        st['y'].copy_(y, non_blocking=True)
        st['graph'].replay()
        self._steps += 1
        self._finite_flag = st['flag']
        engine.bump_params()           # the replay rewrote parameters and BatchNorm statistics behind every version counter
        engine.bump_stats()
        if check_every and self._steps % check_every == 0 and int(self._finite_flag.item()) == 0:
            raise FloatingPointError('non-finite loss at step {}'.format(self._steps))
        return st['losses'], st['measures'].rebind(batch, current_measures)

    @staticmethod
    def warmup_weighting(epoch, warmup):
        """cvae.py:2432-2433: the ramp of the KL variance term / of gamma at `epoch` (0-based) for warmup = (w0, w1)"""
        return max(0., min(1., (epoch + 1 - warmup[0]) / (warmup[1] + 1)))

    def train_model(self, batches, epochs=1, warmup=(0, 0), warmup_gamma=(0, 0), check_every=100, on_batch=None, graph=False):
        """Epoch loop over an iterable of (x, y) device batches with the reference's warm-up ramps for the KL variance
        term and gamma, the per-epoch running measures (batch index i and the previous batch's measures feed the next
        evaluate) and the per-epoch mean losses of the history (cvae.py:2293, 2402-2493).  The loss means accumulate on
        the device and are read once per epoch.  Dataset handling, checkpoints and console tables of the reference's
        train_model stay outside the hot path."""
        history = self.train_history
        for epoch in range(history['epochs'], epochs):
            self.encoder.prior.thaw_means(epoch)
            self.train()
            kw = self.warmup_weighting(epoch, warmup)
            gw = self.warmup_weighting(epoch, warmup_gamma)
            measures, sums, n = {}, None, 0
            for i, (x, y) in enumerate(batches):
                losses, measures = self.train_step(x, y, kl_var_weighting=kw, gamma_weighting=gw, check_every=check_every,
                                                   batch=i, current_measures=measures, graph=graph)
                vec = torch.stack([v.detach().float().mean() for v in losses.values()])
                sums = vec if sums is None else sums + vec
                keys, n = list(losses), n + 1
                if on_batch is not None:
                    on_batch(epoch, i, losses, measures)
            self.eval()
            history['epochs'] = epoch + 1
            self.trained = epoch + 1
            history[epoch] = {}
            if n:
                mean = (sums / n).tolist()          # the epoch's single read-back
                history[epoch]['train_loss'] = dict(zip(keys, mean))
                history[epoch]['train_measures'] = dict(measures)
            history[epoch]['lr'] = self.optimizer.lr
            self.optimizer.update_lr()
        self.eval()
        return history

    # ------------------------------------------------------------------------------------------ scoring loops
    @torch.no_grad()
    def score_batches(self, batches, methods=None, predict_methods=None, recorder=None):
        """The per-batch body of accuracy() / ood_detection_rates() (cvae.py:1629-1687, 1788-1833): per-class evaluate,
        OOD scores and predictions, accumulated on the device; nothing is pulled to the host per batch.
        recorder: a utils.save_load.LossRecorder that receives every loss tensor of the batch, the logits transposed to
        (C, N) and y_true when the batch carries labels (cvae.py:1332-1334, 1673-1675)."""
        methods = methods or [m for m in self.ood_methods if not m.startswith('odin')]
        predict_methods = predict_methods or self.predict_methods
        scores = {m: [] for m in methods}
        preds = {m: [] for m in predict_methods}
        with_odin = any(m.startswith('odin') for m in methods)
        self.eval()
        for b in batches:
            x = b[0] if isinstance(b, (tuple, list)) else b
            _, logits, losses, _ = self.evaluate(x)
            if with_odin:          # cvae.py:1646-1663: the whole temperature x eps grid, merged into the loss dictionary
                losses = dict(losses, **self.odin_softmax(x))
            if recorder is not None:
                rec = dict(losses, logits=logits.T)
                if isinstance(b, (tuple, list)) and len(b) > 1:
                    rec['y_true'] = b[1]
                recorder.append_batch(**rec)
            for m, v in self.batch_dist_measures(logits, losses, methods).items():
                scores[m].append(v)
            for m in predict_methods:
                preds[m].append(self.predict_after_evaluate(logits, losses, method=m))
        return ({m: torch.cat(v) for m, v in scores.items() if v}, {m: torch.cat(v) for m, v in preds.items() if v})

    def odin_softmax(self, x, temps=None, eps=None):
        """ODIN scores of one batch, the inline loop of cvae.py:1627, 1646-1663 and 1798-1815: with x requiring a gradient,
        for every temperature T the sum over the batch of max softmax(logits / T) (logits = mean over the L draws) is
        back-propagated to x -- x.grad is NOT reset between temperatures, as in the reference -- and for every eps the
        network is evaluated again at x + eps * sign(x.grad).  Returns {'odin-T-eps': (B,) max softmax}.  The input
        gradient runs through the native data-gradient kernels (eval-mode BatchNorm folded into the weights); the
        reconstruction, which the reference also computes here and discards, is skipped."""
        temps = self.ODIN_TEMPS if temps is None else list(temps)
        eps = self.ODIN_EPS if eps is None else list(eps)
        x = x.detach().clone().requires_grad_(True)
        out = {}
        score = lambda logits, T: (logits[1:].float().mean(0) / T).softmax(-1).max(-1)[0]
        with torch.no_grad():
            for T in temps:
                with torch.enable_grad():
                    X = score(self.forward(x, z_output=False, decode=False)[1], T).sum()
                X.backward()
                dx = x.grad.sign()
                for e in eps:
                    out['odin-{:.0f}-{:.4f}'.format(T, e)] = score(self.forward(x + e * dx, z_output=False, decode=False)[1], T)
        return out

    def accuracy(self, batches, method='all'):
        """cvae.py:1187-1453 reduced to its arithmetic: accuracy per predict method over (x, y) batches"""
        methods = self.predict_methods if method == 'all' else [method]
        batches = list(batches)
        _, preds = self.score_batches(batches, methods=[], predict_methods=methods)
        y = torch.cat([b[1] for b in batches])
        return {m: float((preds[m] == y).float().mean()) for m in methods}

    def ood_detection_rates(self, ind_batches, ood_batches, ood_methods='all', kept_tpr=None, update_self_ood=True,
                            epoch='last', set_name='ood'):
        """The arithmetic of cvae.py:1455-1911 for one OOD set: scores of the in-distribution and of the OOD batches
        (per-class evaluate + batch_dist_measures, device resident, no per-batch host copies), then per method the ROC
        table on the device (utils/roc_curves.py of this package: '-2s' -> around-mean, '-a-p-q' -> (p, q), else
        one-sided).  Returns {method: {epochs, n, mean, std, auc, tpr, fpr, thresholds}} with the reference's keys
        (cvae.py:1883-1890); fpr at TPR 95 % is `utils.roc_curves.fpr_at_tpr(r['fpr'], r['tpr'], 0.95)`."""
        from .utils.roc_curves import roc_curve
        methods = [m for m in self.ood_methods if not m.startswith('odin')] if ood_methods == 'all' else list(ood_methods)
        kept_tpr = kept_tpr or [pc / 100 for pc in range(90, 100)]
        ind, _ = self.score_batches(ind_batches, methods=methods, predict_methods=[])
        ood, _ = self.score_batches(ood_batches, methods=methods, predict_methods=[])
        epoch = self.trained if epoch == 'last' else epoch
        results = {}
        for m in methods:
            two_sided = False
            if m.endswith('-2s'):
                two_sided = 'around-mean'
            if '-a-' in m:
                two_sided = tuple(int(_) for _ in m.split('-')[-2:])
            auc, fpr, tpr, thr = roc_curve(ind[m], ood[m], *kept_tpr, two_sided=two_sided)
            results[m] = {'epochs': epoch, 'n': int(ood[m].numel()), 'mean': float(ood[m].mean()),
                          'std': float(ood[m].std()), 'auc': auc, 'tpr': kept_tpr, 'fpr': list(fpr),
                          'thresholds': thr}      # the reference stores list(dict) = the two key names here
        if update_self_ood:
            self.ood_results.setdefault(epoch, {})[set_name] = results
        return results

    def misclassification_detection_rates(self, recorder, predict_methods='all', misclass_methods='all', epoch='last',
                                          update_self_results=True):
        """The arithmetic of cvae.py:1913-2079 on a LossRecorder of the test set (the reference loads `record-<set>.pth`
        through its result registry; here the recorder is passed in): per predict method the correct / missed split and the
        accuracy, per misclassification score the ROC table of correct against missed samples (device sort / searchsorted,
        utils/roc_curves.py of this package) and the precision at the kept thresholds.  Results use the reference's keys
        and, with update_self_results, land in self.testing[epoch][predict_method][score] as there."""
        from .utils.roc_curves import roc_curve
        methods = {}
        for which, asked, all_methods in (('predict', predict_methods, self.predict_methods),
                                          ('miss', misclass_methods, self.misclass_methods)):
            developed = _develop_starred(all_methods, self.methods_params)          # in place, as the reference does
            methods[which] = _make_list(asked, developed)
            for m in methods[which]:
                assert m in all_methods, m
        tensors = {k: recorder[k] for k in recorder}      # narrowed to the recorded samples (the buffers grow by doubling)
        logits = tensors.pop('logits').T
        y = tensors.pop('y_true')
        losses = tensors
        kept_tpr = [pc / 100 for pc in range(90, 100)]
        epoch = self.trained if epoch == 'last' else epoch
        sampling = self._latent_samplings['eval']
        results = {}
        for pm in methods['predict']:
            if not methods['miss']:
                continue
            y_ = self.predict_after_evaluate(logits, losses, method=pm)
            correct, missed = (y_ == y), (y_ != y)
            n_correct, n_missed = int(correct.sum()), int(missed.sum())
            acc = n_correct / (n_correct + n_missed)
            scores = self.batch_dist_measures(logits, losses, methods['miss'])
            results[pm] = {'n': int(y.numel()), 'epochs': epoch, 'sampling': sampling, 'accuracy': acc}
            for m in methods['miss']:
                v = scores[m].reshape(-1).double()
                auc, fpr, tpr, thr = roc_curve(v[correct], v[missed], *kept_tpr)
                t_low = torch.as_tensor(np.asarray(thr['low'], dtype=np.float64), device=v.device)
                pos = v[None, :] >= t_low[:, None]                      # (thresholds, samples), cvae.py:2010-2016
                tp = (pos & correct[None]).sum(1).double()
                fp = (pos & missed[None]).sum(1).double()
                precision = (tp / (tp + fp)).tolist()
                results[pm][m] = {'n': int(y.numel()), 'epochs': epoch, 'sampling': sampling, 'tpr': list(tpr), 'fpr': list(fpr),
                                  'auc': auc, 'precision': precision}
        if update_self_results:
            self.testing.setdefault(epoch, {}).update(results)
        return results

    # ------------------------------------------------------------------------------------------ persistence
    def save(self, dir_name):
        """state.pth / optimizer.pth / params.json / train_params.json as cvae.py:2650-2675"""
        os.makedirs(dir_name, exist_ok=True)
        json.dump({k: v for k, v in self.architecture.items()}, open(os.path.join(dir_name, 'params.json'), 'w'), default=str)
        tp = dict(self.training_parameters)
        tp['sigma'] = self.sigma.params
        json.dump(tp, open(os.path.join(dir_name, 'train_params.json'), 'w'), default=str)
        json.dump(self.train_history, open(os.path.join(dir_name, 'history.json'), 'w'), default=str)
        torch.save(self.state_dict(), os.path.join(dir_name, 'state.pth'))
        torch.save(self.optimizer.state_dict(), os.path.join(dir_name, 'optimizer.pth'))

    def load_state(self, dir_name):
        self.load_state_dict(torch.load(os.path.join(dir_name, 'state.pth'), map_location='cpu'))
        p = os.path.join(dir_name, 'optimizer.pth')
        if os.path.exists(p) and next(self.parameters()).is_cuda:
            self.optimizer.load_state_dict(torch.load(p, map_location=next(self.parameters()).device))
        return self

"""Latent priors with the reference's interface (module/priors.py:35-499): GaussianPrior (scalar / diag / full
inverse-Cholesky variance, C class means), TiltedGaussianPrior, UniformWithGaussianTailPrior, build_prior.

The methods below are the tensor-level API other callers use (aggregation, WIM swap priors in and out of
`encoder.prior`); ClassificationVariationalNetwork.evaluate reads `.mean` / `.inv_trans` and goes through the
fused ELBO kernels instead (csrc/elbo.cu), never materialising the (C,B,K) / (L,C,B,K) broadcasts.
"""
import logging
from math import erf, log, pi, sqrt

import torch
from torch import nn
from torch.nn import Parameter

LOG2PI = log(2 * pi)


def build_prior(dim, distribution='gaussian', **kw):
    """priors.py:35-52"""
    if kw.get('num_priors', 1) == 1:
        kw.pop('learned_means', False)
    valid = ('gaussian', 'tilted', 'uniform')
    assert distribution in valid, '{} unknown (try one of: {})'.format(distribution, ', '.join(valid))
    if distribution == 'gaussian':
        kw.pop('tau', None)
        return GaussianPrior(dim, **kw)
    kw.pop('var_dim', None)
    cls = TiltedGaussianPrior if distribution == 'tilted' else UniformWithGaussianTailPrior
    return cls(dim, **kw)


class GaussianPrior(nn.Module):
    """N(mean_c, (T_c^T T_c)^-1) per class; `_var_parameter` is the INVERSE std / inverse Cholesky factor."""

    distribution = 'gaussian'

    def __init__(self, dim, var_dim='scalar', num_priors=1, init_mean=0, mean_shift=0, learned_means=False,
                 freeze_means=0, force_conditional=False, seed=None):
        assert not learned_means or num_priors > 1
        super().__init__()
        gen = torch.Generator()
        gen.seed()
        if seed is not None:
            gen.manual_seed(seed)
        self.num_priors, self.var_dim, self.dim = num_priors, var_dim, dim
        self.learned_var = var_dim != 'scalar'
        self.learned_means, self.freeze_means = learned_means, freeze_means
        if num_priors == 1:
            self.conditional = bool(force_conditional)
            mean = init_mean * torch.randn(1, dim, generator=gen) + mean_shift
        else:
            self.conditional = True
            if isinstance(init_mean, str) and init_mean == 'onehot':
                assert dim >= num_priors, 'K={}<C={}'.format(dim, num_priors)
                mean = torch.eye(num_priors, dim)
            elif torch.is_tensor(init_mean):
                mean = init_mean.squeeze()
            else:
                mean = float(init_mean) * torch.randn(num_priors, dim, generator=gen).squeeze() + mean_shift
        self._frozen_means = not learned_means or freeze_means > 0
        self.mean = Parameter(mean, requires_grad=not self._frozen_means)
        per_class = {'scalar': torch.tensor(1.), 'diag': torch.ones(dim), 'full': torch.eye(dim)}
        if var_dim not in per_class:
            raise ValueError('var_dim {} unknown'.format(var_dim))
        v = per_class[var_dim]
        if self.conditional:
            v = torch.stack([v for _ in range(num_priors)])
        self._var_parameter = Parameter(v, requires_grad=self.learned_var)
        self.params = {'distribution': 'gaussian', 'dim': dim, 'init_mean': init_mean, 'var_dim': var_dim,
                       'num_priors': num_priors}
        if self.conditional:
            self.params.update({'learned_means': learned_means, 'freeze_means': freeze_means})

    # ------------------------------------------------------------------ parameters
    def thaw_means(self, epoch=None):
        if not self.learned_means or not self._frozen_means:
            return
        if epoch is None or epoch >= self.freeze_means:
            self.mean.requires_grad_()
            self._frozen_means = True       # sic, priors.py:139-140

    @property
    def inv_trans(self):
        return self._var_parameter.tril() if self.var_dim == 'full' else self._var_parameter

    @property
    def inv_var(self):
        T = self.inv_trans
        return T.transpose(-1, -2) @ T if self.var_dim == 'full' else T ** 2

    def log_det_per_class(self):
        """log det Sigma_c (priors.py:173-186)"""
        T = self.inv_trans
        if self.var_dim == 'full':
            return -2 * T.diagonal(dim1=-2, dim2=-1).abs().log().sum(-1)
        if self.var_dim == 'diag':
            return -2 * T.abs().log().sum(-1)
        return -2 * self.dim * T.log()

    def _select(self, t, y):
        return t.index_select(0, y.reshape(-1)).view(*y.shape, *t.shape[1:])

    # ------------------------------------------------------------------ tensor-level API
    def mahala(self, x, y=None):
        """|T_y (x - m_y)|^2, x (...,K), y (...) -> (...)  (priors.py:188-224)"""
        assert self.conditional ^ (y is None)
        m = self._select(self.mean, y) if self.conditional else self.mean.view(-1)
        d = x - m
        T = self.inv_trans
        if self.conditional:
            T = self._select(T, y)
        if self.var_dim == 'full':
            w = (T @ d.unsqueeze(-1)).squeeze(-1)
        elif self.var_dim == 'diag':
            w = d * T
        else:
            w = d * (T.unsqueeze(-1) if self.conditional else T)
        return w.pow(2).sum(-1)

    def trace_prod_by_var(self, var, y=None):
        """sum_k var_k diag(T^T T)_k (priors.py:226-250)"""
        assert self.conditional ^ (y is None)
        T = self.inv_trans
        diag = T.pow(2).sum(-2) if self.var_dim == 'full' else T.pow(2)
        if self.conditional:
            diag = self._select(diag, y)
        if self.var_dim == 'scalar':
            diag = diag.unsqueeze(-1)
        return (var * diag).sum(-1)

    def _expand(self, mu, log_var, y):
        if y is not None and y.ndim == mu.ndim:
            shape = (y.shape[0],) + tuple(mu.shape)
            return mu.expand(*shape), log_var.expand(*shape)
        return mu, log_var

    def kl(self, mu, log_var, y=None, output_dict=True, var_weighting=1.):
        """KL(N(mu, e^log_var) || prior_y); y (B,) or (C,B) for all classes (priors.py:252-326)."""
        mu, log_var = self._expand(mu, log_var, y)
        trace = self.trace_prod_by_var(log_var.exp(), y)
        ldp = self.log_det_per_class()
        if self.conditional:
            ldp = self._select(ldp, y)
        log_det = log_var.sum(-1)
        distance = self.mahala(mu, y)
        var_kl = trace - log_det + ldp - self.dim
        kl = 0.5 * (distance + var_weighting * var_kl)
        if torch.isnan(kl).any():
            logging.error('nan in kl')
            return None
        out = {'trace': trace, 'log_det_prior': ldp, 'log_det': log_det, 'distance': distance, 'var_kl': var_kl,
               'kl': kl}
        return out if output_dict else kl

    def log_density(self, z, y=None):
        """log p(z | y) (priors.py:328-342)"""
        assert self.conditional ^ (y is None)
        ld = self.log_det_per_class()
        if self.conditional:
            ld = self._select(ld, y)
        return -LOG2PI * self.dim / 2 - self.mahala(z, y) / 2 - ld / 2

    def __repr__(self):
        pre = 'conditional ' if self.conditional else ''
        var = ('learned ' if self.learned_var else '') + self.var_dim + ' variance'
        if self.conditional:
            mean = '{} {}means and '.format(self.num_priors, 'learned ' if self.learned_means else '')
        else:
            mean = 'mean centered on {} '.format(self.params['init_mean']) if self.params['init_mean'] else ''
        return 'gaussian {p}prior of dim {K} with {m}{v}'.format(p=pre, m=mean, v=var, K=self.dim)


class TiltedGaussianPrior(GaussianPrior):
    """priors.py:356-408: kl = (|mu - m_y| - tau)^2 / 2, log p -= |z|"""

    distribution = 'tilted'

    def __init__(self, dim, num_priors=1, init_mean=0, learned_means=False, tau=25, **kw):
        super().__init__(dim, num_priors=num_priors, init_mean=init_mean, learned_means=learned_means,
                         var_dim='scalar', **kw)
        self.tau = tau
        self._mu_star = tau
        self.params['distribution'] = 'tilted'
        self.params['tau'] = tau

    @property
    def mu_star(self):
        return self._mu_star

    def log_density(self, z, y=None):
        return super().log_density(z, y) - z.norm(dim=-1)

    def kl(self, mu, log_var, y=None, output_dict=True, var_weighting=1.):
        mu, log_var = self._expand(mu, log_var, y)
        distance = self.mahala(mu, y)
        mu_norm = distance.sqrt()
        out = {'distance': distance, 'mu_norm': mu_norm, 'var_kl': torch.zeros_like(mu_norm),
               'kl': 0.5 * (mu_norm - self.mu_star) ** 2}
        return out if output_dict else out['kl']

    def __repr__(self):
        m = ' with {} {}means'.format(self.num_priors, 'learned ' if self.learned_means else '') \
            if self.num_priors > 1 else ''
        return 'tilted gaussian {c}prior{m}, tau={tau}'.format(c='conditional ' if self.conditional else '', m=m,
                                                               tau=self.tau)


class UniformWithGaussianTailPrior(GaussianPrior):
    """priors.py:411-499: density flat on [-tau, tau]^K around m_y with gaussian tails"""

    distribution = 'uniform'

    def __init__(self, dim, num_priors=1, init_mean=0, learned_means=False, tau=5, **kw):
        super().__init__(dim, num_priors=num_priors, init_mean=init_mean, learned_means=learned_means,
                         var_dim='scalar')
        self.tau = tau
        phi_tau = 0.5 * (1 + erf(float(torch.tensor(float(tau), dtype=torch.float32)) / sqrt(2)))
        self._alpha = log(2 * tau) - log(2 * phi_tau - 1)
        self.params['distribution'] = 'uniform'
        self.params['tau'] = tau

    def kl(self, mu, log_var, y=None, output_dict=True, var_weighting=1.0):
        mu, log_var = self._expand(mu, log_var, y)
        assert self.conditional ^ (y is None)
        tau, alpha, c = self.tau, self._alpha, LOG2PI
        means = self._select(self.mean, y) if self.conditional else self.mean.unsqueeze(-1)
        span = 2 * sqrt(3) * (0.5 * log_var).exp()
        d = mu - means
        dist = d.square()
        lo, hi = d - 0.5 * span, d + 0.5 * span
        lo_ = tau * torch.clamp(lo / tau, -1, 1)
        hi_ = tau * torch.clamp(hi / tau, -1, 1)
        elogq = -0.5 * log_var - 0.5 * log(12)
        neg = (c + dist + span.square() / 12) / 2
        neg = neg + (alpha - c / 2) * (hi_ - lo_) / span
        neg = neg - (hi_.pow(3) - lo_.pow(3)) / span / 6
        var_kl = (elogq + alpha).sum(-1)
        kl = torch.max(elogq.sum(-1) + neg.sum(-1), var_kl)
        if var_weighting != 1.0:
            kl = kl + (var_weighting - 1) * var_kl
        out = {'distance': dist.sum(-1), 'var_kl': 2 * var_kl, 'kl': kl}
        return out if output_dict else kl

    def log_density(self, z, y=None):
        assert self.conditional ^ (y is None)
        if self.conditional:
            z = z - self._select(self.mean, y)
        inside = -self._alpha * torch.ones_like(z)
        tail = -LOG2PI / 2 - z.square() / 2
        return torch.where(z.abs() > self.tau, tail, inside).sum(-1)

    def __repr__(self):
        m = ' with {} {}means'.format(self.num_priors, 'learned ' if self.learned_means else '') \
            if self.num_priors > 1 else ''
        return 'uniform {c}prior{m}, tau={tau}'.format(c='conditional ' if self.conditional else '', m=m,
                                                       tau=self.tau)

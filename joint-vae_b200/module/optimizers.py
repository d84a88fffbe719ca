"""Optimizer wrapper with the reference's interface (module/optimizers.py:14-133): Adam (L2 weight decay in the
gradient, not AdamW) or SGD, global-norm clipping, exponential lr decay per epoch.

Adam runs as ONE kernel over a flat fp32 parameter buffer (csrc/optim.cu): the trainable parameters are re-pointed
to views of that buffer, their .grad to views of a flat gradient buffer, so that the data-parallel all-reduce, the
global norm and the update each touch one contiguous array.  The clip coefficient is computed on the device.
"""
import logging

import torch
from torch import nn, optim

from .. import _native as nat
from .. import engine

default_lr = {'sgd': 0.01, 'adam': 0.001}
params_by_type = {'sgd': ('momentum', 'nesterov', 'weight_decay'), 'adam': ('betas', 'weight_decay', 'amsgrad')}


class Optimizer:

    def __init__(self, parameters, optim_type='adam', lr=0, lr_decay=0, weight_decay=0, grad_clipping=None, epoch=0,
                 **kw):
        self.kind = optim_type
        lr = lr or default_lr[optim_type]
        self.params = {'optim_type': optim_type, 'lr': lr, 'lr_decay': lr_decay, 'weight_decay': weight_decay,
                       'grad_clipping': grad_clipping}
        self.params.update(kw)
        self.grad_clipping = grad_clipping
        self.init_lr = lr
        self.lr_decay = lr_decay
        self.weight_decay = weight_decay
        self._params = list(parameters)
        self._lr = lr
        self._clip_now = False
        self.grad_dtype = torch.float32      # torch.bfloat16: the DP bucket is reduced and consumed in bf16
        self.allreduce = None                # set by distributed.DataParallel: callable(flat_grad) -> flat_grad
        if optim_type == 'adam':
            self.betas = tuple(kw.get('betas', (0.9, 0.999)))
            self.eps = kw.get('eps', 1e-8)
            if kw.get('amsgrad'):
                raise NotImplementedError('amsgrad is not implemented by the fused Adam kernel')
            self._opt = None
            self._flat = None
            self._pending_state = {}          # id(param) -> loaded Adam state waiting for the flat buffers
        elif optim_type == 'sgd':
            # secondary optimizer of the reference: plain torch.optim.SGD (library), not on the measured path
            self._opt = optim.SGD(self._params, lr=lr, weight_decay=weight_decay, **kw)
        else:
            raise ValueError(optim_type)

    # ------------------------------------------------------------------ flat buffers (adam)
    def _trainable(self):
        return [p for p in self._params if p.requires_grad]

    def _flatten(self):
        """Re-point the trainable parameters (and their .grad) to slices of flat buffers; every slice starts on a multiple of
        JVAE_OPT_CHUNK elements so that a chunk belongs to one parameter (include/jvae_b200.h: chunk_seg)."""
        ps = self._trainable()
        if not ps:
            raise RuntimeError('no trainable parameter')
        dev = ps[0].device
        if dev.type != 'cuda':
            raise nat.NativeError('the fused Adam kernel needs the parameters on a CUDA device (no CPU fallback)')
        old = self._flat
        CH = nat.OPT_CHUNK
        sizes = [(p.numel() + CH - 1) // CH * CH for p in ps]
        n, nseg = sum(sizes), len(ps)
        aux = torch.zeros(1 + nseg, device=dev)                  # [norm2 | seg_active (int32 view)]: one zero_() per step
        f = {'ids': [id(p) for p in ps], 'n': n, 'p': torch.zeros(n, device=dev), 'm': torch.zeros(n, device=dev),
             'v': torch.zeros(n, device=dev), 'g': torch.zeros(n, device=dev), 'aux': aux, 'norm2': aux[:1],
             'seg_active': aux[1:].view(torch.int32), 'seg_step': torch.zeros(nseg, dtype=torch.int32, device=dev),
             'seg_bc': torch.zeros(2 * nseg, device=dev), 'views': {}, 'seg': {}}
        f['chunk_seg'] = torch.repeat_interleave(torch.arange(nseg, dtype=torch.int32),
                                                 torch.tensor([sz // CH for sz in sizes])).to(dev)
        off = 0
        steps = [0] * nseg
        for i, (p, sz) in enumerate(zip(ps, sizes)):
            sl = slice(off, off + p.numel())
            f['p'][sl].copy_(p.data.reshape(-1))
            if id(p) in self._pending_state:                      # Adam state from load_state_dict
                st = self._pending_state.pop(id(p))
                f['m'][sl].copy_(st['exp_avg'].reshape(-1))
                f['v'][sl].copy_(st['exp_avg_sq'].reshape(-1))
                steps[i] = int(st['step'])
            elif old is not None and id(p) in old['views']:
                osl = old['views'][id(p)]
                f['m'][sl].copy_(old['m'][osl])
                f['v'][sl].copy_(old['v'][osl])
                steps[i] = int(old['seg_step'][old['seg'][id(p)]])
            if p.grad is not None:
                f['g'][sl].copy_(p.grad.reshape(-1))
            p.data = f['p'][sl].view(p.shape)
            p.grad = f['g'][sl].view(p.shape)
            f['views'][id(p)] = sl
            f['seg'][id(p)] = i
            off += sz
        f['seg_step'].copy_(torch.tensor(steps, dtype=torch.int32))
        self._flat = f
        engine.bump_params()

    def _ensure_flat(self):
        ps = self._trainable()
        f = self._flat
        if f is None or f['ids'] != [id(p) for p in ps] or f['p'].device != ps[0].device or \
                any(p.data.data_ptr() != f['p'][f['views'][id(p)]].data_ptr() for p in ps[:1]):
            self._flatten()
        else:
            for p in ps:        # autograd may have replaced .grad by a fresh tensor
                sl = f['views'][id(p)]
                if p.grad is None:
                    p.grad = f['g'][sl].view(p.shape)
                elif p.grad.data_ptr() != f['g'][sl].data_ptr():
                    f['g'][sl].copy_(p.grad.reshape(-1))
                    p.grad = f['g'][sl].view(p.shape)

    @property
    def flat_grad(self):
        self._ensure_flat()
        return self._flat['g']

    # ------------------------------------------------------------------ reference interface
    @property
    def lr(self):
        return self._opt.param_groups[0]['lr'] if self._opt is not None else self._lr

    def zero_grad(self, *a, **kw):
        if self._opt is not None:
            return self._opt.zero_grad(*a, **kw)
        if self._trainable() and self._trainable()[0].is_cuda:
            self._ensure_flat()
            self._flat['g'].zero_()
        else:
            for p in self._params:
                p.grad = None

    def clip(self, parameters=None):
        """module/optimizers.py:79-81.  With the fused Adam the norm and the clip coefficient are computed inside
        step() on the device; this call only arms them."""
        if not self.grad_clipping:
            return
        if self._opt is not None:
            nn.utils.clip_grad_norm_(parameters if parameters is not None else self._params, self.grad_clipping)
        else:
            self._clip_now = True

    def step(self):
        if self._opt is not None:
            r = self._opt.step()
            engine.bump_params()
            return r
        self._ensure_flat()
        f = self._flat
        g = f['g']
        if self.grad_dtype == torch.bfloat16:
            g = nat.cast_f32_bf16(g, f.setdefault('g16', torch.empty(f['n'], dtype=torch.bfloat16, device=g.device)))
        if self.allreduce is not None:
            g = self.allreduce(g)
        max_norm = self.grad_clipping if self._clip_now else 0.0
        f['aux'].zero_()
        # one pass over the gradient: squared norm for the clip AND the per-parameter "has a gradient" flags
        nat.grad_sqnorm(g, f['norm2'], f['chunk_seg'], f['seg_active'])
        nat.adam_step(f['p'], f['m'], f['v'], g, f['norm2'], max_norm=max_norm, lr=self._lr, beta1=self.betas[0],
                      beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, chunk_seg=f['chunk_seg'],
                      seg_active=f['seg_active'], seg_step=f['seg_step'], seg_bc=f['seg_bc'])
        self._clip_now = False
        engine.bump_params()         # parameters changed in place behind autograd's version counters

    def update_lr(self):
        if not self.lr_decay:
            return
        if self._opt is not None:
            for gp in self._opt.param_groups:
                gp['lr'] *= 1 - self.lr_decay
        else:
            self._lr *= 1 - self.lr_decay

    def update_scheduler_from_epoch(self, n):
        for _ in range(n):
            self.update_lr()

    def to(self, device):
        if self._opt is not None:
            for state in self._opt.state.values():
                for k, v in state.items():
                    if isinstance(v, torch.Tensor):
                        state[k] = v.to(device)
        # the flat buffers follow the parameters lazily (see _ensure_flat)

    def state_dict(self, *a, **k):
        """torch.optim.Adam's layout: `state` indexed by the position of the parameter in the FULL parameter list the optimizer
        was given (frozen ones included, as torch does), one {step, exp_avg, exp_avg_sq} per parameter that was ever updated."""
        if self._opt is not None:
            return self._opt.state_dict(*a, **k)
        state = {}
        if self._flat is not None:
            f = self._flat
            steps = f['seg_step'].tolist()
            for i, p in enumerate(self._params):
                if id(p) not in f['views'] or steps[f['seg'][id(p)]] == 0:
                    continue
                sl = f['views'][id(p)]
                state[i] = {'step': torch.tensor(float(steps[f['seg'][id(p)]])), 'exp_avg': f['m'][sl].view(p.shape).clone(),
                            'exp_avg_sq': f['v'][sl].view(p.shape).clone()}
        group = {'lr': self._lr, 'betas': self.betas, 'eps': self.eps, 'weight_decay': self.weight_decay,
                 'amsgrad': False, 'params': list(range(len(self._params)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd, *a, **k):
        if self._opt is not None:
            return self._opt.load_state_dict(sd, *a, **k)
        self._lr = sd['param_groups'][0]['lr']
        self._pending_state = {id(self._params[int(i)]): st for i, st in sd['state'].items() if int(i) < len(self._params)}
        if self._pending_state and self._trainable() and self._trainable()[0].is_cuda:
            self._flatten()            # the loaded moments and step counts enter the flat buffers

    def __str__(self):
        return self.__format__('10')

    def __format__(self, format_spec):
        try:
            level = int(format_spec.rstrip('x'))
        except ValueError:
            level = 0
        if not level:
            level = 10
        s = [self.kind, f'lr={self.init_lr}']
        if self.lr_decay:
            s.append(f'decay={self.lr_decay}')
        else:
            level -= 1
        if self.weight_decay:
            s.append(f'weight_decay={self.weight_decay}')
        return '--'.join(s[:level])

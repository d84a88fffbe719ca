"""Golden vectors for the WIM caller (ft/wim.py:48-129, 215-259): the UNMODIFIED reference model
evaluated with `encoder.prior` swapped for the alternate (single, unconditional) prior and num_labels = 1, as
WIMJob._switch_to_alternate_prior does: the train-mode evaluate(x_mix, y_mix = 0, with_beta=True) of finetune_batch with
its gradients, and the eval-mode evaluate(x) of evaluate_on_both_priors.  Alternate prior parameters as ft/__main__.py:163-175
builds them.

    python tests/golden/make_wim_golden.py        # build container only
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import describe_model, import_reference, injected_noise, t2n  # noqa: E402

CTOR = dict(input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32, 16], decoder=[16, 32], classifier=[],
            latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0, output_activation='sigmoid',
            sigma={'value': 0.1}, prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 1})


def main():
    os.chdir('/tmp')
    cvae_mod = import_reference()
    from module.priors import build_prior
    torch.set_num_threads(1)
    kw = json.loads(json.dumps(CTOR))
    ctor = dict(kw)
    ctor['input_shape'] = tuple(ctor['input_shape'])
    torch.manual_seed(99)
    model = cvae_mod.ClassificationVariationalNetwork(**ctor)
    alt = model.encoder.prior.params.copy()                    # ft/__main__.py:163-175
    alt.update(learned_means=False, mean_shift=0., init_mean=0.5, num_priors=1, seed=11, tau=None)
    alt_prior = build_prior(**alt)                             # ft/wim.py:95-98
    for p in alt_prior.parameters():
        p.requires_grad_(False)
    B, K = 6, kw['latent_dim']
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, *ctor['input_shape'], generator=g)
    eps_tr = torch.randn(kw['latent_sampling'] + 1, B, K, generator=g)
    eps_te = torch.randn(kw['test_latent_sampling'] + 1, B, K, generator=g)
    out = {'cfg': np.array(json.dumps(kw)), 'alt_params': np.array(json.dumps(alt)), 'x': t2n(x), 'eps_train': t2n(eps_tr),
           'eps_eval': t2n(eps_te), 'arch': np.array(json.dumps(describe_model(model), default=lambda o: o.item()))}
    for k, v in model.state_dict().items():
        out['sd.' + k] = t2n(v)
    for k, v in alt_prior.state_dict().items():
        out['alt.' + k] = t2n(v)
    # ft/wim.py:56-61: swap
    model.encoder.prior = alt_prior
    model.num_labels = 1
    model.train()
    model.optimizer.zero_grad()
    y_mix = torch.zeros(B, dtype=int)                          # ft/wim.py:246
    with injected_noise(eps_tr):
        _, logits, losses, _ = model.evaluate(x, y_mix, with_beta=True)
    for k, v in losses.items():
        out['train.loss.' + k] = t2n(v)
    losses['total'].mean().backward()
    for k, p in model.named_parameters():
        if p.grad is not None:
            out['train.grad.' + k] = t2n(p.grad)
    model.eval()
    with torch.no_grad(), injected_noise(eps_te):
        _, logits, losses, _ = model.evaluate(x)
    for k, v in losses.items():
        out['eval.loss.' + k] = t2n(v)
    out['eval.logits'] = t2n(logits)
    np.savez_compressed(os.path.join(HERE, 'wim_alternate_prior.npz'), **out)
    print('ok; train keys', {k[11:]: out[k].shape for k in out if k.startswith('train.loss.')},
          'eval keys', {k[10:]: out[k].shape for k in out if k.startswith('eval.loss.')}, 'alt', alt)


if __name__ == '__main__':
    main()

"""Golden vectors for the ODIN caller, from the UNMODIFIED reference on CPU.

The reference has no callable for ODIN: the loop sits inline in ood_detection_rates (cvae.py:1627, 1646-1663).  This
script drives the reference MODEL (its forward / autograd) through exactly that loop -- x.requires_grad_(True) once,
for every temperature T: softmax = (logits[1:].mean(0) / T).softmax(-1).max(-1)[0]; softmax.sum().backward() (x.grad is
NOT zeroed between temperatures, as in the reference); dx = x.grad.sign(); for every eps: forward(x + eps * dx) -- with
the sampling noise pinned (torch.randn patched to return the same eps at every call).

    python tests/golden/make_odin_golden.py        # build container only
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import describe_model, import_reference, t2n  # noqa: E402

CASES = {
    'odin_mlp_vib': dict(
        input_shape=(1, 8, 8), num_labels=5, type='vib', encoder=[32], decoder=[32],
        classifier=[10], latent_dim=8, latent_sampling=2, test_latent_sampling=3, gamma=1.0, beta=1e-2,
        output_activation='sigmoid', sigma={'value': 0.2}, prior={'var_dim': 'scalar'}),
    'odin_conv_vib_bn': dict(
        input_shape=(3, 16, 16), num_labels=4, type='vib', features='[x3+1]8-8-M-16:2-16', batch_norm='encoder',
        encoder=[24], decoder=[], classifier=[12], latent_dim=8, latent_sampling=2, test_latent_sampling=2,
        gamma=1.0, beta=1e-2, output_activation='linear', sigma={'value': 1.0}, prior={'var_dim': 'scalar'}),
}
TEMPS = [1, 10, 1000]
EPS = [0.0, 0.002, 0.004]


def run(cvae_mod, name, kw):
    kw = json.loads(json.dumps(kw))
    ctor = dict(kw)
    ctor['input_shape'] = tuple(ctor['input_shape'])
    torch.manual_seed(4321)
    model = cvae_mod.ClassificationVariationalNetwork(**ctor)
    g = torch.Generator().manual_seed(7)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):       # eval mode must not be the identity
            m.running_mean.copy_(0.3 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.3 * torch.randn(m.num_features, generator=g))
    model.eval()
    B, K, L = 6, kw['latent_dim'], kw['test_latent_sampling']
    x = torch.rand(B, *ctor['input_shape'], generator=g)
    eps = torch.randn(L + 1, B, K, generator=g)
    out = {'cfg': np.array(json.dumps(kw)), 'x': t2n(x), 'eps_eval': t2n(eps), 'temps': np.array(TEMPS), 'eps_list': np.array(EPS),
           'arch': np.array(json.dumps(describe_model(model), default=lambda o: o.item()))}
    for k, v in model.state_dict().items():
        out['sd.' + k] = t2n(v)
    orig = torch.randn
    torch.randn = lambda size, *a, **k: eps.clone() if tuple(size) == tuple(eps.shape) else orig(size, *a, **k)
    try:
        x.requires_grad_(True)                                   # cvae.py:1627
        with torch.no_grad():
            for T in TEMPS:                                      # cvae.py:1647-1663
                with torch.enable_grad():
                    _, logits = model.forward(x, z_output=False)
                    softmax = (logits[1:].mean(0) / T).softmax(-1).max(-1)[0]
                    X = softmax.sum()
                X.backward()
                out[f'grad.{T}'] = t2n(x.grad)
                dx = x.grad.sign()
                for e in EPS:
                    _, odin_logits = model.forward(x + e * dx, z_output=False)
                    out['odin-{:.0f}-{:.4f}'.format(T, e)] = t2n((odin_logits[1:].mean(0) / T).softmax(-1).max(-1)[0])
    finally:
        torch.randn = orig
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'ok', sorted(k for k in out if k.startswith('odin-'))[:3], '...')


if __name__ == '__main__':
    os.chdir('/tmp')
    mod = import_reference()
    torch.set_num_threads(1)
    for name, kw in CASES.items():
        run(mod, name, kw)

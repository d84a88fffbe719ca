"""ODIN caller (cvae.py:1627, 1646-1663) against golden vectors produced by driving the UNMODIFIED
reference model through that loop (tests/golden/make_odin_golden.py): same weights, same x, pinned sampling noise.
Checked: the accumulated input gradient after every temperature (direction and norm) and every odin-T-eps score.
bf16 GEMM / conv operands: scores at 2e-2 relative (north_star tolerance); the sign of near-zero gradient entries may
differ, which moves x by at most 2 eps there."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_gpu_model import build

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('name', ['odin_mlp_vib', 'odin_conv_vib_bn'])
def test_odin_scores_match_reference(pkg, name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    cfg, net = build(pkg, d)
    net.eval()
    x = torch.from_numpy(d['x']).to(DEV)
    temps, eps = [int(t) for t in d['temps']], [float(e) for e in d['eps_list']]
    net.encoder.sampling.injected_eps = torch.from_numpy(d['eps_eval']).to(DEV)
    n0 = pkg._native.launch_count()

    def input_gradients():
        """accumulated x.grad after each temperature, as in the reference: (cosine, norm ratio, sign agreement) per T"""
        xg = x.clone().requires_grad_(True)
        out = []
        for T in temps:
            with torch.enable_grad():
                lg = net.forward(xg, z_output=False, decode=False)[1]
                (lg[1:].float().mean(0) / T).softmax(-1).max(-1)[0].sum().backward()
            got, want = xg.grad.detach().cpu().numpy().astype(np.float64), d[f'grad.{T}'].astype(np.float64)
            big = np.abs(want) > 0.05 * np.abs(want).max()
            out.append((float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want))),
                        float(np.linalg.norm(got) / np.linalg.norm(want)), float((np.sign(got[big]) == np.sign(want[big])).mean())))
        return out

    try:
        native = input_gradients()
        # noise floor of bf16 activations: the same network through cuDNN in bf16 (conv stacks with ReLU / max-pool
        # decisions on 8 channels amplify activation rounding in the input gradient, tests/test_gpu_conv.py)
        from library_arm import cudnn_bf16
        with cudnn_bf16(pkg):
            library = input_gradients()
        print('input gradient (cos, norm ratio, sign agreement): native', native, 'cudnn-bf16', library)
        for (c, r, sg), (cl, rl, sl) in zip(native, library):
            assert 1 - c < max(5e-3, 2 * (1 - cl)), (c, cl)
            assert abs(r - 1) < max(5e-2, 2 * abs(rl - 1)), (r, rl)
            assert 1 - sg < max(1e-2, 2 * (1 - sl)), (sg, sl)
        scores = net.odin_softmax(x, temps=temps, eps=eps)
    finally:
        net.encoder.sampling.injected_eps = None
    assert pkg._native.launch_count() > n0
    assert sorted(scores) == sorted(k for k in d.files if k.startswith('odin-'))
    for k, v in scores.items():
        want = d[k]
        assert tuple(v.shape) == want.shape
        err = float(np.abs(v.cpu().numpy() - want).max() / max(1e-6, np.abs(want).max()))
        assert err < 2e-2, (k, err)


def test_score_batches_merges_odin_scores(pkg):
    """'odin*' methods of the vib type (cvae.py:110): the temperature x eps grid is computed per batch and read back by
    batch_dist_measures like any other loss entry"""
    torch.manual_seed(0)
    net = pkg.ClassificationVariationalNetwork((1, 8, 8), 5, type='vib', encoder=[32], decoder=[32], classifier=[10],
                                               latent_dim=8, latent_sampling=2, test_latent_sampling=2, gamma=1.0, beta=1e-2,
                                               sigma={'value': 0.2}, prior={'var_dim': 'scalar'}).to(DEV)
    xs = [torch.rand(7, 1, 8, 8, device=DEV) for _ in range(2)]
    methods = ['odin-1-0.0000', 'odin-1000-0.0040', 'baseline']
    scores, _ = net.score_batches(xs, methods=methods, predict_methods=[])
    for m in methods:
        assert scores[m].shape == (14,) and torch.isfinite(scores[m]).all()
    assert float(scores['odin-1-0.0000'].min()) >= 1 / 5 - 1e-3 and float(scores['odin-1-0.0000'].max()) <= 1 + 1e-5

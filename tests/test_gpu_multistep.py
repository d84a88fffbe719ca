"""Multi-step training against the UNMODIFIED reference (tests/golden/make_multistep_golden.py: the body of cvae.py:2424-2461
driven for 3 batches with the reference's own Adam, then eval, one more step, eval).  Every step must read the parameters the
previous optimizer step wrote -- the fused Adam updates the flat buffer through a raw pointer, so the bf16 weight arrangements
and the BatchNorm-folded inference weights are re-derived under explicit epochs (engine.PARAM_EPOCH / STATS_EPOCH); the test
also shows that it is sensitive: with the epoch bump disabled (stale packed weights from step 2 on) it must NOT match."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(1e-6, np.abs(b).max()))


def _build(pkg, d):
    kw = json.loads(str(d['cfg']))
    kw['input_shape'] = tuple(kw['input_shape'])
    net = pkg.ClassificationVariationalNetwork(**kw)
    net.load_state_dict({k[3:]: torch.from_numpy(np.asarray(d[k])) for k in d.files if k.startswith('sd.')})
    return net.to(DEV)


def _drive(pkg, name):
    """-> per-step relative errors of the loss terms, errors of the two evaluations, the network"""
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    net = _build(pkg, d)
    T = int(d['steps'])
    t = lambda k: torch.from_numpy(d[k]).to(DEV)
    errs, meas = [], []

    def step(i, batch, current):
        net.encoder.sampling.injected_eps = t(f'eps.{i}')
        losses, measures = net.train_step(t(f'x.{i}'), t(f'y.{i}'), batch=batch, current_measures=current)
        e = {k: rel(v.detach().cpu().numpy(), d[f'step{i}.loss.{k}']) for k, v in losses.items()}
        errs.append(e)
        m = {k[len(f'step{i}.measure.'):]: (float(measures[k[len(f'step{i}.measure.'):]]), float(d[k]))
             for k in d.files if k.startswith(f'step{i}.measure.')}
        meas.append(m)
        return measures

    def evaluate(tag):
        net.eval()
        net.encoder.sampling.injected_eps = t('eps_eval')
        with torch.no_grad():
            _, logits, losses, _ = net.evaluate(t('x_eval'))
        e = {k: rel(v.cpu().numpy(), d[f'{tag}.loss.{k}']) for k, v in losses.items()}
        e['logits'] = rel(logits.float().cpu().numpy(), d[f'{tag}.logits'])
        preds = {m: (net.predict_after_evaluate(logits, losses, method=m).cpu().numpy(), d[f'{tag}.pred.{m}'],
                     losses) for m in ('iws', 'closest')}
        return e, preds

    net.train()
    cur = {}
    for i in range(T):
        cur = step(i, i, cur)
    sd = {k: v.detach().float().cpu().numpy() for k, v in net.state_dict().items()}
    ea, pa = evaluate('eval_a')
    net.train()
    step(T, 0, {})
    eb, pb = evaluate('eval_b')
    return d, errs, meas, sd, (ea, pa), (eb, pb)


@pytest.mark.parametrize('name', ['multistep_conv_cvae_bn', 'multistep_mlp_cvae'])
def test_three_adam_steps_then_eval_match_reference(pkg, name):
    d, errs, meas, sd, (ea, pa), (eb, pb) = _drive(pkg, name)
    tol = 2e-2
    for i, e in enumerate(errs):
        for k, v in e.items():
            assert v < tol, (f'step {i}', k, v, e)
    # running measures chained through current_measures / batch index as the reference does (cvae.py:2441-2449)
    for i, m in enumerate(meas):
        for k, (got, want) in m.items():
            assert abs(got - want) <= 3e-2 * max(1.0, abs(want)), (f'step {i}', k, got, want)
    # parameters after 3 steps: every tensor moved like the reference's (Adam: |update| ~ lr per element and step, so the
    # error is judged against the distance travelled)
    moved = 0
    for k, v in sd.items():
        want, start = d['sd_after.' + k].astype(np.float64), d['sd.' + k].astype(np.float64)
        if want.dtype.kind != 'f' or 'num_batches' in k:
            continue
        travelled = np.linalg.norm(want - start)
        if travelled < 1e-9:
            assert np.allclose(v, want, atol=1e-6), k
            continue
        moved += 1
        assert np.linalg.norm(v - want) <= 0.35 * travelled + 1e-4, (k, np.linalg.norm(v - want), travelled)
    assert moved >= 8
    # evaluation with the BatchNorm running statistics and weights of the moment (folded copies must be fresh)
    for tag, e in (('eval_a', ea), ('eval_b', eb)):
        for k, v in e.items():
            assert v < 3e-2, (tag, k, v)
    # predictions exact wherever the reference's decision margin exceeds the tolerance
    for preds, tag in ((pa, 'eval_a'), (pb, 'eval_b')):
        for m, (got, want, losses) in preds.items():
            key = 'iws' if m == 'iws' else 'zdist'
            ref = d[f'{tag}.loss.{key}'].astype(np.float64)
            srt = np.sort(ref if m == 'closest' else -ref, axis=0)
            margin = (srt[1] - srt[0]) / np.maximum(1.0, np.abs(srt[0]))
            clear = margin > 2 * 3e-2
            assert (got[clear] == want[clear]).all(), (tag, m)


def test_stale_weights_would_be_caught(pkg, monkeypatch):
    """the same drive with the parameter-epoch bump disabled = the packed bf16 weights of step 1 reused by later steps:
    the parity of steps 2+ must break, i.e. the test above really depends on fresh weights"""
    monkeypatch.setattr(pkg.engine, 'bump_params', lambda: None)
    from jointvae_b200 import conv_engine
    conv_engine._stacks.clear()
    d, errs, *_ = _drive(pkg, 'multistep_conv_cvae_bn')
    conv_engine._stacks.clear()
    worst = max(v for e in errs[1:] for k, v in e.items() if k in ('wmse', 'kl', 'zdist'))
    assert worst > 2e-2, errs

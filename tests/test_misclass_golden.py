"""misclassification_detection_rates (cvae.py:1913-2079) against the unmodified reference
(tests/golden/make_misclass_golden.py): same recorder tensors in, the reference's `model.testing[epoch]` out --
accuracy per predict method; AUC, kept FPR / TPR and precision at the kept thresholds per misclassification score.
Runs on CPU tensors (the method is tensor arithmetic on the recorder; on a GPU the same code keeps everything on the device)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _load(pkg):
    d = np.load(os.path.join(GOLDEN, 'misclass_cvae.npz'))
    kw = json.loads(str(d['cfg']))
    kw['input_shape'] = tuple(kw['input_shape'])
    net = pkg.ClassificationVariationalNetwork(**kw)
    tensors = {k[4:]: torch.from_numpy(np.asarray(d[k])) for k in d.files if k.startswith('rec.')}
    n = tensors['y_true'].numel()
    rec = pkg.utils.save_load.LossRecorder(n)
    rec.append_batch(**tensors)
    return net, rec, json.loads(str(d['results']))


@pytest.fixture()
def one_thread():
    """the golden was produced with one torch thread: ATen's CPU softmax splits its vectorised / scalar exp paths by the
    thread partition, and a one-ulp change reorders the near-tied soft*-500 scores (AUC moves by 1e-4)"""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_matches_reference(pkg, one_thread):
    net, rec, want = _load(pkg)
    got = net.misclassification_detection_rates(rec, epoch=7)
    assert sorted(got) == sorted(want)
    checked = 0
    for pm, wr in want.items():
        gr = got[pm]
        assert gr['n'] == wr['n'] and gr['sampling'] == wr['sampling']
        assert abs(gr['accuracy'] - wr['accuracy']) < 1e-12
        wm = sorted(k for k, v in wr.items() if isinstance(v, dict))
        assert sorted(k for k, v in gr.items() if isinstance(v, dict)) == wm
        for m in wm:
            a, b = gr[m], wr[m]
            assert abs(a['auc'] - b['auc']) < 1e-9, (pm, m)
            np.testing.assert_allclose(a['fpr'], b['fpr'], rtol=0, atol=1e-12, err_msg=f'{pm} {m}')
            np.testing.assert_allclose(a['tpr'], b['tpr'], rtol=0, atol=1e-12, err_msg=f'{pm} {m}')
            np.testing.assert_allclose(np.asarray(a['precision'], float), np.asarray(b['precision'], float), rtol=0, atol=1e-12,
                                       err_msg=f'{pm} {m}')
            checked += 1
    assert checked >= 20
    assert net.testing[7].keys() >= want.keys()


def test_selected_methods(pkg, one_thread):
    net, rec, want = _load(pkg)
    got = net.misclassification_detection_rates(rec, predict_methods=['iws'], misclass_methods=['kl', 'softkl-10'], epoch=3,
                                                update_self_results=False)
    assert list(got) == ['iws'] and sorted(k for k, v in got['iws'].items() if isinstance(v, dict)) == ['kl', 'softkl-10']
    assert abs(got['iws']['kl']['auc'] - want['iws']['kl']['auc']) < 1e-9
    assert 3 not in net.testing

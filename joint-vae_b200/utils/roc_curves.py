"""ROC / FPR-at-TPR tables of the OOD and misclassification sweeps on the device (utils/roc_curves.py:8-210 of the
reference; callers cvae.py:1857-1868, 2003-2035).

The reference walks ~len(ins) thresholds in a Python loop and advances four cursors into the sorted scores (about one
million iterations per method and OOD set at the BASELINE scoring sizes).  Here the scores are sorted once on the device
and every cursor position is a `searchsorted` count, so a table is a handful of vector operations; the scores never
leave the GPU.  The result reproduces the reference's conventions, quirks included:

  * cursors saturate at n-1 (the `idx < n - 1` guards), so the last sample of each set is never counted out;
  * the thresholds stored for a kept TPR are the NEXT pair of thresholds (the loop advances before it records);
  * when the TPR drops below a kept level the loop spends that iteration moving to the next level (no record);
  * `ins_are_higher=False` only flips the sign / side of the reported thresholds, not the scores;
  * AUC = trapezoid over the recorded (fpr, tpr) points plus (0, 0), as sklearn.metrics.auc.

`validation > 0` draws a random, unseeded split in the reference (roc_curves.py:54-57): it cannot be reproduced and is
not implemented.  Inputs may be torch tensors (any device) or array-likes; outputs are numpy like the reference's.
"""
import math

import numpy as np
import torch


def fpr_at_tpr(fpr, tpr, a, thresholds=None, return_threshold=False):
    """roc_curves.py:8-27 (fpr and tpr in ascending order)"""
    assert not return_threshold or thresholds is not None
    as_tpr, as_fpr = np.asarray(tpr), np.asarray(fpr)
    i_ = np.where(as_tpr >= a)[0].min()
    if not return_threshold:
        return as_fpr[i_]
    return as_fpr[i_], thresholds[i_]


def tpr_at_fpr(fpr, tpr, a):
    """roc_curves.py:30-35"""
    as_tpr, as_fpr = np.asarray(tpr), np.asarray(fpr)
    return as_tpr[np.where(as_fpr <= a)[0]].max()


def _as_tensor(v, device=None):
    if isinstance(v, torch.Tensor):
        return v.detach().reshape(-1).to(torch.float64)
    return torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(-1), device=device)


def roc_curve(ins, outs, *kept_tpr, two_sided=False, validation=0, debug=False, ins_are_higher=True):
    """-> (auroc, kept_fpr, kept_tpr, kept_thresholds={'low': ..., 'up': ...}) as roc_curves.py:38-210"""
    if validation:
        raise NotImplementedError('validation > 0 uses an unseeded random split in the reference (roc_curves.py:54-57)')
    sign = 1 if ins_are_higher else -1
    lowup = {'low': 'low', 'up': 'up'} if ins_are_higher else {'low': 'up', 'up': 'low'}
    ins = _as_tensor(ins)
    outs = _as_tensor(outs, ins.device).to(ins.device)
    dev = ins.device
    n_in, n_out = ins.numel(), outs.numel()
    s_in, s_out = torch.sort(ins)[0], torch.sort(outs)[0]
    inf = torch.tensor([math.inf], dtype=torch.float64, device=dev)

    # ---- threshold sequences in the order the loop visits them (roc_curves.py:66-92)
    if two_sided == 'around-mean':
        # the centre must carry the reference's exact bits (numpy's pairwise mean of the SORTED scores): samples sit
        # exactly on their own threshold center -/+ abs(x - center), so one ulp decides which side they are counted on
        center = torch.tensor(float(np.mean(s_in.cpu().numpy())), dtype=torch.float64, device=dev)
        delta = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), torch.sort((ins - center).abs())[0], inf])
        desc = torch.flip(delta, [0])
        low_seq, up_seq = center - desc, center + desc              # low[it], up[-1 - it]
    elif isinstance(two_sided, tuple):
        # roc_curves.py:73-83: thresholds = an interpolating cubic spline through the sorted scores, evaluated at its own
        # knots (validation = 0), i.e. the sorted scores up to the spline's rounding noise.  Samples sit exactly on these
        # thresholds, so the noise decides on which side they are counted: the thresholds are produced the reference's
        # way (scipy on the host, O(n)); the counting below stays on the device.
        from scipy.interpolate import UnivariateSpline
        f_lo, f_up = two_sided
        val = s_in.cpu().numpy()
        idx_old = np.arange(0, len(val))
        interp = UnivariateSpline(idx_old, val, k=3, s=0)(np.linspace(0, len(val) - 1, n_in))
        interp = torch.from_numpy(interp).to(dev)
        low_seq = torch.cat([-inf, interp[::f_lo], inf])
        up_seq = torch.flip(torch.cat([-inf, interp[::f_up], inf]), [0])
    else:
        low_seq = torch.cat([-inf, s_in])
        up_seq = torch.full_like(low_seq, math.inf)
    nt = min(low_seq.numel(), up_seq.numel())
    low_seq, up_seq = low_seq[:nt], up_seq[:nt]

    # ---- iterations 0 .. T-1: the loop runs while t_low < t_up and it < nt - 1 (roc_curves.py:132)
    go = low_seq[:nt - 1] < up_seq[:nt - 1]
    stop = torch.nonzero(~go)
    T = int(stop[0]) if stop.numel() else nt - 1
    original = sorted(kept_tpr)
    m = len(original)
    kept_tpr_out = np.zeros(m)
    kept_fpr = np.ones(m)
    kept_thr = {'low': -np.inf * np.ones(m), 'up': np.inf * np.ones(m)}
    if T <= 0:
        return 0.0, kept_fpr, kept_tpr_out, kept_thr
    t_lo, t_up = low_seq[:T], up_seq[:T]

    def rates(sorted_scores, n):
        below = torch.searchsorted(sorted_scores, t_lo, right=False).clamp(max=n - 1)          # scores < t_low
        above = (n - torch.searchsorted(sorted_scores, t_up, right=True)).clamp(max=n - 1)     # scores > t_up
        # the reference's cursors only ever advance: with thresholds that are not monotone (spline noise among tied
        # scores) a cursor keeps the furthest position reached so far
        below, above = torch.cummax(below, 0)[0], torch.cummax(above, 0)[0]
        return 1.0 - (below + above).to(torch.float64) / n

    tpr, fpr = rates(s_in, n_in), rates(s_out, n_out)
    next_lo, next_up = low_seq[1:T + 1], up_seq[1:T + 1]            # the loop advances before it records

    # ---- AUC over the recorded points (+ the closing (0, 0)), sklearn.metrics.auc on a non-increasing fpr
    ok = (fpr >= 0) & (tpr >= 0)
    zero = torch.zeros(1, dtype=torch.float64, device=dev)
    x, y = torch.cat([fpr[ok], zero]), torch.cat([tpr[ok], zero])
    auroc = float(-torch.trapezoid(y, x)) if x.numel() > 1 else 0.0

    # ---- kept TPR levels, highest first; one iteration is spent on every level change (roc_curves.py:176-186)
    neg_tpr = -tpr                                                   # ascending
    i_prev = -1                                                      # iteration at which the previous level was left
    for j in range(m - 1, -1, -1):
        first_below = int(torch.searchsorted(neg_tpr, torch.tensor(-original[j], dtype=torch.float64, device=dev),
                                             right=True))           # first iteration with tpr < level
        i_leave = max(first_below, i_prev + 1)
        last = min(i_leave, T) - 1                                   # last iteration that records for this level
        if last > i_prev:
            kept_fpr[j] = float(fpr[last])
            kept_tpr_out[j] = float(tpr[last])
            kept_thr[lowup['low']][j] = sign * float(next_lo[last])
            kept_thr[lowup['up']][j] = sign * float(next_up[last])
        if i_leave >= T:
            break                                                    # the loop ended before leaving this level
        i_prev = i_leave
    return auroc, kept_fpr, kept_tpr_out, kept_thr

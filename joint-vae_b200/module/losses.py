"""Loss terms with the reference's signatures (module/losses.py:8-110).

These are the stand-alone tensor-level functions of the reference API.  ClassificationVariationalNetwork
.evaluate does not call them: it uses the fused ELBO kernels (csrc/elbo.cu), which compute the same quantities
without materialising the intermediates.
"""
import torch
from torch.nn import functional as F


def mse_loss(x_output, x_target, ndim=3, batch_mean=True):
    """mean square error of x_output (L,...,*shape) against x_target (...,*shape): over everything if
    batch_mean, else over the last `ndim` dims only (losses.py:8-27)."""
    sq = (x_output - x_target.expand_as(x_output)).pow(2)
    return sq.mean() if batch_mean else sq.mean(tuple(range(-ndim, 0)))


def categorical_loss(x_output, x_target, ndim=3, batch_mean=True):
    """256-way cross entropy per pixel, summed over pixels (losses.py:30-49).
    x_output (L,...,256,*shape), x_target (...,*shape) in [0,1]."""
    tgt = (x_target * 255).long()
    lead = x_output.shape[:-ndim - 1]
    tgt = tgt.expand(*lead, *tgt.shape[-ndim:])
    logp = F.log_softmax(x_output, dim=-ndim - 1)
    ce = -logp.gather(-ndim - 1, tgt.unsqueeze(-ndim - 1)).squeeze(-ndim - 1)
    out = ce.flatten(-ndim).sum(-1)
    return out.mean() if batch_mean else out


def x_loss(y_target, logits, batch_mean=True):
    """y given: cross entropy averaged over all L+1 draws, (...)  (losses.py:73-86);
    y None: -log(softmax + 1e-6) averaged over draws 1..L and moved to (C, ...) (losses.py:62-71)."""
    if y_target is None:
        lp = (logits.softmax(-1) + 1e-6).log()
        lp = lp[1:].mean(0) if lp.shape[0] > 1 else lp[0]
        out = -lp.movedim(-1, 0)
        return out.mean() if batch_mean else out
    C = logits.shape[-1]
    L1 = logits.shape[0]
    tgt = y_target.expand(L1, *y_target.shape).reshape(-1)
    ce = F.cross_entropy(logits.reshape(-1, C), tgt, reduction='none').view(L1, *y_target.shape).mean(0)
    return ce.mean() if batch_mean else ce


def loss_mean(component, values, y=None, current_mean=0., n=0):
    """running mean of a loss component; for per-class (C,N) values the row of class y (given, or the
    arg-max for 'elbo'/'iws', else the arg-min) is the one averaged (losses.py:89-110)."""
    if values.ndim == 1:
        values = values.unsqueeze(0)
    bs = values.shape[-1]
    if values.shape[0] == 1:
        m = values.mean()
    else:
        if y is None:
            y = values.max(0)[1] if component in ('elbo', 'iws') else values.min(0)[1]
        m = values.index_select(0, y).mean()
    return (current_mean * n + m * bs) / (n + bs)

"""Small helpers of the layer package (mirrors module/vae_layers/misc.py:5-39 of the reference)."""
import torch
from torch import nn


def onehot_encoding(y, C):
    """y (...) int64 -> (..., C) float one-hot (misc.py:5-17)."""
    out = torch.zeros(tuple(y.shape) + (C,), device=y.device)
    return out.scatter_(-1, y.unsqueeze(-1), 1)


def _no_activation(a):
    return a


activation_layers = {'linear': nn.Identity, 'sigmoid': nn.Sigmoid, 'relu': nn.ReLU, 'leaky': nn.LeakyReLU}


class Reshape(nn.Module):
    def __init__(self, output_shape):
        super().__init__()
        self.shape = output_shape

    def forward(self, x):
        return x.view(-1, *self.shape)

"""Small helpers of the layer package: the names `module/vae_layers/misc.py` exports in the reference (one-hot coding of
labels for `y_is_coded` encoders, the activation table of the spec parser, the Reshape layer that ends categorical imagers),
plus the table that maps an activation module onto the fused-epilogue code of the native kernels."""
import torch
import torch.nn.functional as F
from torch import nn

#: activation name -> layer class, as the conv / dense builders look it up (reference: misc.py:24-27)
activation_layers = dict(linear=nn.Identity, sigmoid=nn.Sigmoid, relu=nn.ReLU, leaky=nn.LeakyReLU)

#: layer class -> jvae_act code of include/jvae_b200.h (epilogue of the GEMM / conv / BatchNorm kernels)
NATIVE_ACT_CODE = {nn.Identity: 0, nn.ReLU: 1, nn.Sigmoid: 2, nn.LeakyReLU: 3}


def onehot_encoding(y, C):
    """labels y (...) -> float one-hot (..., C) on y's device; what an encoder with y_is_coded concatenates to its input"""
    return F.one_hot(y.long(), num_classes=C).to(torch.float32)


def _no_activation(a):
    return a


class Reshape(nn.Module):
    """(N, prod(shape)) or (N, C', H, W) -> (N, *shape) as a VIEW (the categorical imager ends with it: channels -> (256, C));
    a view also of the channels_last image the native conv stack returns"""

    def __init__(self, output_shape):
        super().__init__()
        self.shape = tuple(int(s) for s in output_shape)

    def extra_repr(self):
        return 'shape={}'.format(self.shape)

    def forward(self, x):
        return x.view((-1,) + self.shape)

"""TEST INFRASTRUCTURE: the same conv stacks through cuDNN in bf16 (torch autocast, channels_last) as a NOISE-FLOOR arm for
GPU tests that judge gradients of bf16 networks.  The product has no library path; tests swap engine.run_conv_stack for the
duration of a `with cudnn_bf16(pkg):` block."""
import contextlib

import torch


def _stack(mods, x, image_out=False):
    with torch.autocast(device_type='cuda', dtype=torch.bfloat16):
        x = x.contiguous(memory_format=torch.channels_last)
        for m in mods:
            x = m(x)
    return x


@contextlib.contextmanager
def cudnn_bf16(pkg):
    orig = pkg.engine.run_conv_stack
    pkg.engine.run_conv_stack = _stack
    try:
        yield
    finally:
        pkg.engine.run_conv_stack = orig

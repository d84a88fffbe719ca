"""Single layers of the BASELINE stacks at their real spatial sizes (deconv32 / ivgg / vgg19 / resnet shapes), no BatchNorm:
forward, data gradient and weight gradient of the native kernels against torch fp32 on bf16-exact operands.  With exact
operands the only error sources are the bf16 rounding of the stored outputs (2^-9 relative per element, random) and fp32
accumulation order, so the tolerances are tight: 4e-3 of the tensor norm for bf16 outputs, 1e-3 for the fp32 weight gradient."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

# (module factory, input shape (C, H, W), batch)
LAYERS = {
    'head 32->3 k5 (separable)': (lambda: nn.Conv2d(32, 3, 5, padding=2), (32, 32, 32), 136),
    'head 32->3 k3 (ivgg)': (lambda: nn.Conv2d(32, 3, 3, padding=1), (32, 64, 64), 36),
    'convT 32->32 k5 s1': (lambda: nn.ConvTranspose2d(32, 32, 5, padding=2), (32, 32, 32), 136),
    'convT 32->32 k5 s2': (lambda: nn.ConvTranspose2d(32, 32, 5, stride=2, padding=2, output_padding=1), (32, 16, 16), 136),
    'convT 64->32 k5 s1': (lambda: nn.ConvTranspose2d(64, 32, 5, padding=2), (64, 16, 16), 136),
    'convT 64->64 k5 s2': (lambda: nn.ConvTranspose2d(64, 64, 5, stride=2, padding=2, output_padding=1), (64, 8, 8), 136),
    'convT 64->64 k5 s1': (lambda: nn.ConvTranspose2d(64, 64, 5, padding=2), (64, 8, 8), 136),
    'convT 128->64 k8 on 1x1 (GEMM)': (lambda: nn.ConvTranspose2d(128, 64, 8), (128, 1, 1), 544),
    'conv 3->64 k3 (vgg stem)': (lambda: nn.Conv2d(3, 64, 3, padding=1), (3, 32, 32), 64),
    'conv 128->256 k3 @8x8': (lambda: nn.Conv2d(128, 256, 3, padding=1), (128, 8, 8), 64),
    'conv 512->512 k3 @2x2': (lambda: nn.Conv2d(512, 512, 3, padding=1), (512, 2, 2), 128),
    'conv 64->128 k3 s2 (resnet)': (lambda: nn.Conv2d(64, 128, 3, stride=2, padding=1, bias=False), (64, 16, 16), 32),
    'conv 64->128 k1 s2 (shortcut)': (lambda: nn.Conv2d(64, 128, 1, stride=2, bias=False), (64, 16, 16), 32),
    'conv 3->64 k7 s2 (resnet stem)': (lambda: nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), (3, 64, 64), 32),
    'conv !128 k3 (ivgg)': (lambda: nn.Conv2d(16, 128, 3, padding=1), (16, 8, 8), 72),
}


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize('name', list(LAYERS))
def test_single_layer_matches_torch_fp32(pkg, name):
    from jointvae_b200 import conv_engine as ce
    make, shape, N = LAYERS[name]
    torch.manual_seed(0)
    layer = make().to(DEV)
    with torch.no_grad():
        layer.weight.copy_(layer.weight.to(torch.bfloat16).float())
    x = torch.randn(N, *shape, device=DEV).to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    want = layer(xr)
    go = torch.randn_like(want).to(torch.bfloat16).float()
    gx_w, gw_w = torch.autograd.grad(want, (xr, layer.weight), go)
    image_out = layer.out_channels == 3
    xin = x.clone().requires_grad_(True)
    ce._stacks.clear()
    n0 = pkg._native.launch_count()
    got = ce.run([layer], xin, image_out=image_out)
    assert pkg._native.launch_count() > n0
    g = go.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if image_out else go
    gx, gw = torch.autograd.grad(got, (xin, layer.weight), g)
    e_f, e_x, e_w = _rel(got.float(), want), _rel(gx, gx_w), _rel(gw, gw_w)
    print(f'{name}: forward {e_f:.5f}  data gradient {e_x:.5f}  weight gradient {e_w:.6f}')
    assert e_f < 4e-3, e_f
    assert e_x < 4e-3, e_x
    assert e_w < 1e-3, e_w

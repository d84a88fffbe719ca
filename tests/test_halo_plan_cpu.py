"""The planner of the halo convolution kernel (csrc/conv.cu: try_launch_halo) without a GPU: jvae_conv_halo_emulate executes the
plan it makes for a geometry on the host -- TMA box fills (zero outside the image, NaN where the kernel would find leftovers),
the MMAs in the order of the chunk records / tap table on addresses decoded from the descriptor words, fresh / accumulate flags,
the epilogue's block table, bounds and phase offsets -- and the result is compared with torch.nn on the same layer
(module/vae_layers/conv.py:189-219).  Covers the plain, tap-stacked (G > 1), parity-plane (stride 2) and streamed-weight plans
and the merged sub-pixel phases of jvae_conv_subpixel_gemm.  The device kernel is checked on the GPU (tests/test_gpu_conv.py)."""
import pytest
import torch


@pytest.fixture(scope='module')
def ce(pkg):
    from jointvae_b200 import conv_engine
    return conv_engine


def _nhwc(x, ld):
    n, c, h, w = x.shape
    t = torch.zeros(n, h, w, ld)
    t[..., :c] = x.permute(0, 2, 3, 1)
    return t.contiguous()


# N, H, W, Cin, Cout, k, pad, activation code -- stride-2 ConvTranspose2d with output 2H x 2W
SUBPIXEL = [
    (5, 16, 16, 32, 32, 5, 2, 0),      # c2 imager 16 -> 32
    (7, 8, 8, 64, 64, 5, 2, 0),        # c2 imager 8 -> 16: halved channel tile, stacked images, ragged batch
    (3, 16, 16, 32, 16, 4, 1, 1),      # even kernel, relu
    (2, 40, 24, 16, 24, 3, 1, 0),      # two row blocks per image, 24 channels in a 32-wide tile
    (20, 8, 8, 16, 16, 5, 2, 0),
    (2, 8, 12, 48, 3, 5, 2, 2),        # 3 output channels, sigmoid
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,p,act', SUBPIXEL)
def test_merged_subpixel_plan_matches_conv_transpose(pkg, ce, N, H, W, Cin, Cout, k, p, act):
    nat = pkg._native
    torch.manual_seed(N * 131 + k)
    conv = torch.nn.ConvTranspose2d(Cin, Cout, k, stride=2, padding=p, output_padding=2 + 2 * p - k)
    conv.weight.data = conv.weight.data.to(torch.bfloat16).float()      # the packed matrix is bf16
    x = torch.randn(N, Cin, H, W)
    ref = conv(x).detach()
    ref = {0: ref, 1: ref.relu(), 2: ref.sigmoid()}[act]
    Ho, Wo = 2 * H, 2 * W
    assert tuple(ref.shape[-2:]) == (Ho, Wo)
    ops = ce.deconv_form(k, p, 2, Ho, Wo)
    g_f = conv.weight.detach().permute(1, 2, 3, 0).reshape(Cout, k * k, Cin)
    wm = ce.pack_gather_weights(g_f, [i for op in ops for i in op['idx']], Cin).float().contiguous()
    out = torch.full((N, Ho, Wo, ce.r8(Cout)), float('nan'))
    info = nat.conv_halo_emulate(_nhwc(x, ce.r8(Cin)), Cin, wm, wm.shape[0], None, 1, H, W, out, Cout, (2, 2), (0, 0),
                                 conv.bias.detach().contiguous(), act, phases=nat.phases_arg(ops))
    assert info is not None, 'geometry not covered by the merged kernel'
    G, PH, MT, NBt, BN, nck, resident, stages = info
    assert (G, PH, resident) == (1, 4, 1) and stages >= 2 and 2 * MT * PH * BN <= 512
    assert nck <= 3 * len({t for op in ops for t in op['taps']})
    assert not torch.isnan(out).any(), 'an output element was not written, or a leftover reached a live row'
    got = out[..., :Cout].permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) < 1e-4
    assert float(out[..., Cout:].abs().max() if out.shape[-1] > Cout else 0) == 0


def test_c2_imager_phases_are_single_chunks_per_shift(pkg, ce):
    """k = 5, stride 2: the 25 (tap, phase) pairs sit on 9 input shifts; the planner's phase order makes each shift ONE MMA"""
    nat = pkg._native
    ops = ce.deconv_form(5, 2, 2, 32, 32)
    x = torch.zeros(1, 16, 16, 32)
    wm = torch.zeros(32, 25 * 32)
    out = torch.zeros(1, 32, 32, 32)
    info = nat.conv_halo_emulate(x, 32, wm, 32, None, 1, 16, 16, out, 32, (2, 2), (0, 0), None, 0, phases=nat.phases_arg(ops))
    assert info[5] == 9 and info[1] == 4 and info[4] == 32


# N, H, W, Cin, Cout, k, pad, stride, transposed
PLAIN = [
    (3, 32, 32, 32, 32, 5, 2, 1, True),      # c2 imager 32 -> 32: vertical tap stacking
    (5, 16, 16, 64, 32, 5, 2, 1, True),      # resident weights, one image per box
    (7, 8, 8, 64, 64, 5, 2, 1, True),        # streamed weights, stacked images
    (3, 32, 32, 32, 32, 5, 2, 2, False),     # stride 2: four parity planes
    (2, 40, 40, 8, 16, 3, 1, 1, False),      # two row blocks per image
    (2, 32, 32, 32, 3, 5, 2, 1, False),      # image head
    (4, 16, 16, 64, 64, 3, 1, 1, False),     # 64 input channels: one 64-channel chunk or two stacked 32-channel chunks
    (3, 16, 16, 48, 24, 3, 1, 1, False),     # 48 channels: the second chunk is half empty
    (6, 8, 8, 64, 32, 5, 2, 1, True),
    (3, 16, 16, 64, 32, 3, 1, 2, False),     # 64 channels, stride 2 (parity planes)
    (2, 20, 12, 40, 16, 5, 2, 1, False),     # ragged strip, 40 channels
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,p,s,transposed', PLAIN)
def test_halo_plan_matches_torch(pkg, ce, N, H, W, Cin, Cout, k, p, s, transposed):
    nat = pkg._native
    torch.manual_seed(7)
    conv = torch.nn.ConvTranspose2d(Cin, Cout, k, padding=p) if transposed else torch.nn.Conv2d(Cin, Cout, k, stride=s, padding=p)
    conv.weight.data = conv.weight.data.to(torch.bfloat16).float()
    x = torch.randn(N, Cin, H, W)
    ref = conv(x).detach()
    Ho, Wo = ref.shape[-2:]
    if transposed:
        taps = [(p - i, p - j) for i in range(k) for j in range(k)]
        g_f = conv.weight.detach().permute(1, 2, 3, 0).reshape(Cout, k * k, Cin)
    else:
        taps = [(i - p, j - p) for i in range(k) for j in range(k)]
        g_f = conv.weight.detach().permute(0, 2, 3, 1).reshape(Cout, k * k, Cin)
    wm = ce.pack_gather_weights(g_f, list(range(k * k)), Cin).float().contiguous()
    out = torch.full((N, Ho, Wo, ce.r8(Cout)), float('nan'))
    info = nat.conv_halo_emulate(_nhwc(x, ce.r8(Cin)), Cin, wm, wm.shape[0], nat.taps_arg(taps), s, Ho, Wo, out, Cout, (1, 1),
                                 (0, 0), conv.bias.detach().contiguous(), 0)
    assert info is not None
    assert not torch.isnan(out).any()
    assert float((out[..., :Cout].permute(0, 3, 1, 2) - ref).abs().max()) < 1e-4


# N, H, W, Cin, Cout, k, pad, stride -- Conv2d weight gradient: grid tensor dy, gathered tensor x
WGRAD = [
    (3, 16, 16, 32, 32, 5, 2, 1, 1),      # 32-channel gradient tile, five tap rows: vertical stacking on the N side
    (4, 8, 8, 64, 64, 5, 2, 1, 1),        # 64 channels: horizontal runs of two taps
    (2, 32, 32, 16, 32, 3, 1, 1, 1),
    (5, 16, 16, 32, 32, 5, 2, 2, 1),      # stride 2: four parity planes in one stage
    (3, 16, 16, 64, 64, 5, 2, 2, 4),      # stride 2, 64 channels: the planes do not fit together -> one launch per plane
    (2, 40, 24, 8, 16, 3, 1, 1, 1),       # two row blocks per image
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,p,s,launches', WGRAD)
def test_wgrad_halo_plan_matches_autograd(pkg, ce, N, H, W, Cin, Cout, k, p, s, launches):
    """jvae_conv_wgrad_emulate executes the halo weight-gradient plan(s) on the host (MN-major operands, taps stacked along M
    through LBO, vertically adjacent taps along N, the per-plane split of stride-2 layers) against torch's autograd"""
    nat = pkg._native
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(Cin, Cout, k, stride=s, padding=p, bias=False)
    x = torch.randn(N, Cin, H, W)
    y = conv(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    want = conv.weight.grad
    Hq, Wq = y.shape[-2:]
    taps = [(i - p, j - p) for i in range(k) for j in range(k)]
    dw = torch.zeros(Cout, Cin, k * k)
    n = nat.conv_wgrad_emulate(_nhwc(dy, ce.r8(Cout)), Cout, _nhwc(x, ce.r8(Cin)), Cin, nat.taps_arg(taps), s, dw, 1, Cin * k * k, k * k)
    assert n == launches, 'plan count'
    assert not torch.isnan(dw).any()
    got = dw.view(Cout, Cin, k, k)
    assert float((got - want).abs().max()) <= 2e-4 * float(want.abs().max())


def test_random_geometries_through_both_planners(pkg, ce):
    """seeded sweep: every plan the halo planners make for random small layers (Conv2d stride 1 / 2, forward and weight gradient;
    stride-2 ConvTranspose2d merged) emulates to torch's result; 400 more geometries were swept when the emulators were written"""
    import random
    nat = pkg._native
    rnd = random.Random(11)
    done = 0
    for it in range(14):
        N = rnd.randint(1, 5); H = rnd.choice([6, 8, 9, 12, 16, 20]); W = rnd.choice([6, 8, 10, 12, 16])
        Ci = rnd.choice([3, 8, 16, 24, 32, 40, 64]); Co = rnd.choice([3, 8, 16, 24, 32, 48, 64]); k = rnd.choice([1, 3, 5])
        s = rnd.choice([1, 1, 2])
        torch.manual_seed(it)
        conv = torch.nn.Conv2d(Ci, Co, k, stride=s, padding=k // 2)
        conv.weight.data = conv.weight.data.to(torch.bfloat16).float()
        x = torch.randn(N, Ci, H, W)
        y = conv(x)
        Hq, Wq = y.shape[-2:]
        if Wq < 6:
            continue
        taps = [(i - k // 2, j - k // 2) for i in range(k) for j in range(k)]
        g = conv.weight.detach().permute(0, 2, 3, 1).reshape(Co, k * k, Ci)
        wm = ce.pack_gather_weights(g, list(range(k * k)), Ci).float().contiguous()
        out = torch.full((N, Hq, Wq, ce.r8(Co)), float('nan'))
        info = nat.conv_halo_emulate(_nhwc(x, ce.r8(Ci)), Ci, wm, wm.shape[0], nat.taps_arg(taps), s, Hq, Wq, out, Co, (1, 1), (0, 0),
                                     conv.bias.detach().contiguous(), 0)
        if info is not None:
            assert not torch.isnan(out).any(), (N, H, W, Ci, Co, k, s, info)
            assert float((out[..., :Co].permute(0, 3, 1, 2) - y.detach()).abs().max()) < 1e-3, (N, H, W, Ci, Co, k, s, info)
            done += 1
        dy = torch.randn_like(y)
        y.backward(dy)
        dw = torch.zeros(Co, Ci, k * k)
        n = nat.conv_wgrad_emulate(_nhwc(dy, ce.r8(Co)), Co, _nhwc(x, ce.r8(Ci)), Ci, nat.taps_arg(taps), s, dw, 1, Ci * k * k, k * k)
        if n is not None:
            want = conv.weight.grad
            assert float((dw.view(Co, Ci, k, k) - want).abs().max()) <= 1e-3 * float(want.abs().max()), (N, H, W, Ci, Co, k, s, n)
            done += 1
    for it in range(8):
        N = rnd.randint(1, 4); H = rnd.choice([6, 8, 12, 16]); W = rnd.choice([6, 8, 12, 16])
        Ci = rnd.choice([8, 16, 32, 48, 64]); Co = rnd.choice([3, 16, 24, 32, 64]); k = rnd.choice([2, 3, 4, 5, 6])
        p = rnd.choice([q for q in range(k) if 0 <= 2 + 2 * q - k <= 1])
        torch.manual_seed(100 + it)
        conv = torch.nn.ConvTranspose2d(Ci, Co, k, stride=2, padding=p, output_padding=2 + 2 * p - k)
        conv.weight.data = conv.weight.data.to(torch.bfloat16).float()
        x = torch.randn(N, Ci, H, W)
        y = conv(x).detach()
        ops = ce.deconv_form(k, p, 2, 2 * H, 2 * W)
        g = conv.weight.detach().permute(1, 2, 3, 0).reshape(Co, k * k, Ci)
        wm = ce.pack_gather_weights(g, [i for op in ops for i in op['idx']], Ci).float().contiguous()
        out = torch.full((N, 2 * H, 2 * W, ce.r8(Co)), float('nan'))
        info = nat.conv_halo_emulate(_nhwc(x, ce.r8(Ci)), Ci, wm, wm.shape[0], None, 1, H, W, out, Co, (2, 2), (0, 0),
                                     conv.bias.detach().contiguous(), 0, phases=nat.phases_arg(ops))
        if info is not None:
            assert not torch.isnan(out).any(), (N, H, W, Ci, Co, k, p, info)
            assert float((out[..., :Co].permute(0, 3, 1, 2) - y).abs().max()) < 1e-3, (N, H, W, Ci, Co, k, p, info)
            done += 1
    assert done >= 20

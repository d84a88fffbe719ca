"""GPU parity of the input-batch kernel (csrc/elementwise.cu: jvae_batch_u8_to_f32, through the C ABI) and of
DeviceBatchLoader end to end against the reference's own input path (tests/batch_reference.py: torchvision transforms
per sample + DataLoader): bit-exact, resident and pinned-host datasets, several epochs, ragged last batch."""
import numpy as np
import pytest
import torch

from batch_reference import reference_batches
from emu_kernels import emu_batch_u8_to_f32

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _images(n, h, w, c, seed=0):
    g = np.random.default_rng(seed)
    shape = (n, h, w, c) if c else (n, h, w)
    return g.integers(0, 256, size=shape, dtype=np.uint8), g.integers(0, 10, size=n)


CASES = [
    (32, 32, 3, ['flip', 'crop'], 'simple', None),
    (32, 32, 3, ['crop', 'flip'], 'simple', None),
    (32, 32, 3, [], 'simple', None),
    (28, 28, 0, ['crop'], 'pad', None),
    (40, 36, 3, ['flip', 'crop'], 'crop', (3, 32, 32)),
    (7, 5, 3, ['flip', 'crop'], 'simple', None),
]


@pytest.mark.parametrize('resident', [True, False])
@pytest.mark.parametrize('H,W,C,aug,transformer,out_shape', CASES)
def test_loader_equals_reference_dataloader(pkg, H, W, C, aug, transformer, out_shape, resident):
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    data, targets = _images(150, H, W, C)
    torch.manual_seed(5)
    want = reference_batches(data, targets, 32, aug, transformer, out_shape, epochs=2)
    torch.manual_seed(5)
    loader = DeviceBatchLoader(data, targets, 32, device=DEV, data_augmentation=aug, transformer=transformer,
                               out_shape=out_shape, rng='torchvision', resident=resident)
    n0 = pkg._native.launch_count()
    got = [(x, y) for _ in range(2) for x, y in loader]
    assert pkg._native.launch_count() - n0 == len(want) == 10
    for (gx, gy), (wx, wy) in zip(got, want):
        assert gx.is_cuda and gx.dtype == torch.float32 and gy.dtype == torch.int64
        assert torch.equal(gy.cpu(), wy)
        assert torch.equal(gx.cpu(), wx)


def test_kernel_equals_index_map_at_full_batch(pkg):
    """B = 512 CIFAR-shaped batch (BASELINE configs[1] input): kernel vs the index map, every decision combination"""
    nat = pkg._native
    data, _ = _images(2048, 32, 32, 3, seed=1)
    g = torch.Generator().manual_seed(2)
    idx = torch.randint(0, 2048, (512,), generator=g)
    flip = (torch.rand(512, generator=g) < 0.5).to(torch.uint8)
    crop = torch.randint(0, 9, (512, 2), generator=g, dtype=torch.int32)
    d = torch.from_numpy(data)
    for flip_first in (0, 1):
        for use_flip, use_crop in ((1, 1), (1, 0), (0, 1), (0, 0)):
            cfg = nat.BatchCfg(H=32, W=32, C=3, out_H=32, out_W=32, crop_pad=4 if use_crop else 0, flip_first=flip_first,
                               post_off_y=0, post_off_x=0)
            f, c = (flip if use_flip else None), (crop if use_crop else None)
            out = torch.empty(512, 3, 32, 32, device=DEV)
            nat.batch_u8_to_f32(cfg, d.to(DEV), idx.to(DEV), f.to(DEV) if f is not None else None,
                                c.to(DEV) if c is not None else None, out)
            assert torch.equal(out.cpu(), emu_batch_u8_to_f32(cfg, d, idx, f, c))


def test_loader_feeds_a_train_step(pkg):
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    torch.manual_seed(0)
    data, targets = _images(96, 32, 32, 3)
    net = pkg.ClassificationVariationalNetwork((3, 32, 32), 10, type='cvae', features='conv32', upsampler='deconv32',
                                               batch_norm='both', encoder=[], decoder=[], classifier=[], latent_dim=32,
                                               latent_sampling=2, gamma=0, output_activation='linear',
                                               sigma={'value': 1.0, 'learned': True},
                                               prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar'}).to(DEV)
    net.train()
    loader = DeviceBatchLoader(data, targets, 32, device=DEV, data_augmentation=['flip', 'crop'], seed=1)
    seen = 0
    for x, y in loader:
        losses, _ = net.train_step(x, y)
        assert torch.isfinite(losses['total']).all()
        seen += x.shape[0]
    assert seen == 96


def test_bad_arguments_fail_loudly(pkg):
    nat = pkg._native
    cfg = nat.BatchCfg(H=8, W=8, C=3, out_H=8, out_W=8, crop_pad=0, flip_first=0, post_off_y=0, post_off_x=0)
    d = torch.zeros(4, 8, 8, 3, dtype=torch.uint8, device=DEV)
    idx = torch.zeros(2, dtype=torch.int64, device=DEV)
    out = torch.empty(2, 3, 8, 8, device=DEV)
    with pytest.raises(nat.NativeError):        # crop offsets without RandomCrop padding
        nat.batch_u8_to_f32(cfg, d, idx, None, torch.zeros(2, 2, dtype=torch.int32, device=DEV), out)

"""Host logic of joint-vae_b200/utils/batch_loader.py without a GPU: under the same torch.manual_seed the planner draws the
permutation, flip flags and crop offsets in the order the reference's DataLoader + torchvision transforms consume the
global generator, and the kernel's index map (emulated in tests/emu_kernels.py) then reproduces the reference's batches
bit for bit.  The CUDA kernel itself is checked in tests/test_gpu_batch_loader.py."""
import numpy as np
import pytest
import torch

from batch_reference import reference_batches
from emu_kernels import emu_batch_u8_to_f32


def _images(n, h, w, c, seed=0):
    g = np.random.default_rng(seed)
    shape = (n, h, w, c) if c else (n, h, w)
    return g.integers(0, 256, size=shape, dtype=np.uint8), g.integers(0, 10, size=n)


CASES = [
    # (H, W, C (0: grey (N, H, W)), data_augmentation, transformer, out_shape)
    (32, 32, 3, [], 'simple', None),
    (32, 32, 3, ['flip'], 'simple', None),
    (32, 32, 3, ['crop'], 'simple', None),
    (32, 32, 3, ['flip', 'crop'], 'simple', None),       # train.py's usual --data-augmentation flip crop
    (32, 32, 3, ['crop', 'flip'], 'simple', None),
    (28, 28, 0, ['crop'], 'pad', None),                  # mnist-like: grey, Pad(2) -> 32 x 32
    (28, 28, 0, [], 'pad', None),
    (40, 36, 3, ['flip', 'crop'], 'crop', (3, 32, 32)),  # CenterCrop after the augmentation
    (7, 5, 3, ['flip', 'crop'], 'simple', None),         # padding 7 // 8 = 0: RandomCrop draws nothing
]


@pytest.mark.parametrize('H,W,C,aug,transformer,out_shape', CASES)
def test_same_batches_as_the_reference_dataloader(pkg, H, W, C, aug, transformer, out_shape):
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    data, targets = _images(37, H, W, C)
    torch.manual_seed(11)
    want = reference_batches(data, targets, 8, aug, transformer, out_shape, epochs=2)
    torch.manual_seed(11)
    loader = DeviceBatchLoader(data, targets, 8, data_augmentation=aug, transformer=transformer, out_shape=out_shape,
                               rng='torchvision')
    assert len(loader) == 5
    got = []
    for _ in range(2):
        for idx, flip, crop in loader.plan_epoch():
            got.append((emu_batch_u8_to_f32(loader.cfg, loader.data, idx, flip, crop), loader.targets[idx]))
    assert len(got) == len(want) == 10
    for (gx, gy), (wx, wy) in zip(got, want):
        assert gx.shape == wx.shape and gx.dtype == wx.dtype
        assert torch.equal(gy, wy)
        assert torch.equal(gx, wx)
    assert tuple(got[0][0].shape[1:]) == loader.out_shape


def test_unshuffled_order_and_drop_last(pkg):
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    data, targets = _images(21, 8, 8, 3)
    loader = DeviceBatchLoader(data, targets, 4, shuffle=False, drop_last=True)
    plans = list(loader.plan_epoch())
    assert len(plans) == len(loader) == 5
    assert torch.equal(torch.cat([p[0] for p in plans]), torch.arange(20))
    assert all(p[1] is None and p[2] is None for p in plans)


def test_vectorised_draws_are_seeded_and_in_range(pkg):
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    data, targets = _images(64, 32, 32, 3)
    a = [p for p in DeviceBatchLoader(data, targets, 16, data_augmentation=['flip', 'crop'], seed=3).plan_epoch()]
    b = [p for p in DeviceBatchLoader(data, targets, 16, data_augmentation=['flip', 'crop'], seed=3).plan_epoch()]
    for (i1, f1, c1), (i2, f2, c2) in zip(a, b):
        assert torch.equal(i1, i2) and torch.equal(f1, f2) and torch.equal(c1, c2)
        assert c1.dtype == torch.int32 and int(c1.min()) >= 0 and int(c1.max()) <= 8 and set(f1.tolist()) <= {0, 1}
    assert sorted(torch.cat([p[0] for p in a]).tolist()) == list(range(64))


def test_argument_errors(pkg):
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    data, targets = _images(4, 8, 8, 3)
    with pytest.raises(TypeError):
        DeviceBatchLoader(data.astype(np.float32), targets, 2)
    with pytest.raises(ValueError):
        DeviceBatchLoader(data, targets[:3], 2)
    with pytest.raises(ValueError):
        DeviceBatchLoader(data, targets, 2, data_augmentation=['rotate'])
    with pytest.raises(ValueError):
        DeviceBatchLoader(data, targets, 2, transformer='crop')
    with pytest.raises(Exception):          # no CPU fallback
        next(iter(DeviceBatchLoader(data, targets, 2, device='cpu')))


def test_rank_sharding_partitions_every_epoch(pkg):
    """data-parallel use: all ranks walk the same permutation and keep every world_size-th sample"""
    from jointvae_b200.utils.batch_loader import DeviceBatchLoader
    data, targets = _images(50, 8, 8, 3)
    loaders = [DeviceBatchLoader(data, targets, 4, data_augmentation=['flip', 'crop'], seed=5, rank=r, world_size=3)
               for r in range(3)]
    assert [len(l) for l in loaders] == [5, 5, 4]
    for epoch in range(2):
        plans = [list(l.plan_epoch()) for l in loaders]
        idx = [torch.cat([p[0] for p in pl]) for pl in plans]
        assert sorted(torch.cat(idx).tolist()) == list(range(50))            # a partition of the dataset
        assert [i.numel() for i in idx] == [17, 17, 16]
        if epoch == 0:
            first = idx
        else:
            assert any(not torch.equal(a, b) for a, b in zip(first, idx))    # a new permutation every epoch
        # augmentation decisions differ between ranks
        assert not torch.equal(plans[0][0][2], plans[1][0][2])
    with pytest.raises(ValueError):
        DeviceBatchLoader(data, targets, 4, rank=0, world_size=2)                     # no shared seed
    with pytest.raises(ValueError):
        DeviceBatchLoader(data, targets, 4, seed=1, rng='torchvision', rank=0, world_size=2)

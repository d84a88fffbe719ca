"""Input batches on the device.

The reference feeds the step from `torch.utils.data.DataLoader(trainset, batch_size, shuffle=True, num_workers=0)`
(cvae.py:2245-2256) whose items go one by one through PIL transforms (utils/torch_load.py:405-426: RandomHorizontalFlip,
RandomCrop(size, padding=size // 8, padding_mode='edge'), then Pad(2) / CenterCrop and ToTensor), are collated on the
host and copied with `x.to(device)` (cvae.py:2427).  Here the dataset stays a uint8 (N, H, W, C) array (what torchvision's
CIFAR10 / SVHN / MNIST hold in `.data`), resident in HBM when it fits (CIFAR-10 train: 150 MB) or in pinned host memory
otherwise, and one kernel per batch does gather + augmentation + ToTensor (include/jvae_b200.h: jvae_batch_u8_to_f32).
Batch k+1 is produced on a side stream while the step of batch k runs.

rng='torchvision' draws the permutation and the per-sample flip / crop decisions from torch's global CPU generator in
exactly the order the reference's DataLoader + transforms do, so `torch.manual_seed(s)` gives the same batches, bit for
bit, as the reference (checked in tests/test_batch_loader_cpu.py against a real DataLoader).  rng='vectorised' (default)
draws them with a few tensor ops from a private generator: same distribution, no per-sample Python work.
"""
import numpy as np
import torch

from .. import _native as nat


def _as_u8_nhwc(data):
    t = torch.as_tensor(np.asarray(data) if not torch.is_tensor(data) else data)
    if t.dtype != torch.uint8:
        raise TypeError(f'images must be uint8 (got {t.dtype}): pass the dataset\'s raw `.data` array')
    if t.dim() == 3:            # (N, H, W) grey-level sets (MNIST-like)
        t = t.unsqueeze(-1)
    if t.dim() != 4:
        raise ValueError('images must be (N, H, W) or (N, H, W, C)')
    return t.contiguous()


class DeviceBatchLoader:
    """Iterable of (x, y): x float32 (B, C, H', W') in [0, 1] on `device`, y int64 (B,) on `device`.

    data_augmentation: sequence of 'flip' / 'crop' applied in that order (torch_load.py:405-414).
    transformer: 'simple' (ToTensor only) | 'pad' (Pad(2), torch_load.py:422-423) | 'crop' (CenterCrop(out_shape),
    torch_load.py:419-420).  crop_padding: RandomCrop padding, default H // 8 (torch_load.py:411).
    """

    def __init__(self, data, targets, batch_size, device='cuda', data_augmentation=(), transformer='simple', out_shape=None,
                 crop_padding=None, shuffle=True, drop_last=False, rng='vectorised', seed=None, resident=None,
                 max_resident_bytes=32 << 30, rank=None, world_size=None):
        self.data = _as_u8_nhwc(data)
        n, H, W, C = self.data.shape
        self.targets = torch.as_tensor(np.asarray(targets) if not torch.is_tensor(targets) else targets).long().reshape(-1)
        if self.targets.numel() != n:
            raise ValueError('one target per image is required')
        if batch_size < 1:
            raise ValueError('batch_size must be positive')
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.device = torch.device(device)
        aug = list(data_augmentation)
        for t in aug:
            if t not in ('flip', 'crop'):
                raise ValueError(f'unknown data augmentation {t!r} (flip | crop)')
        if len(set(aug)) != len(aug):
            raise ValueError('each augmentation may appear once')
        self.data_augmentation = aug
        self.crop_padding = (H // 8 if crop_padding is None else int(crop_padding)) if 'crop' in aug else 0
        # RandomCrop draws nothing when the padded image already has the target size (torchvision get_params)
        self._crop_draws = 'crop' in aug and self.crop_padding > 0
        if rng not in ('torchvision', 'vectorised'):
            raise ValueError("rng must be 'torchvision' or 'vectorised'")
        self.rng = rng
        # data-parallel training (distributed.py): every rank walks the SAME permutation (shared seed) and keeps the samples
        # rank, rank + world_size, ...; the augmentation decisions come from a per-rank generator
        self.rank, self.world_size = (0, 1) if world_size in (None, 1) else (int(rank), int(world_size))
        if not 0 <= self.rank < self.world_size:
            raise ValueError('rank must be in [0, world_size)')
        if self.world_size > 1 and (seed is None or rng != 'vectorised'):
            raise ValueError("sharding across ranks needs rng='vectorised' and the same explicit seed on every rank")
        self._gen = torch.Generator()          # permutation
        self._gen_aug = torch.Generator()      # flip / crop draws
        if seed is not None:
            self._gen.manual_seed(int(seed))
            self._gen_aug.manual_seed(int(seed) + 7919 * (self.rank + 1))
        else:
            self._gen.seed()
            self._gen_aug.seed()

        oH, oW, off_y, off_x = H, W, 0, 0
        if transformer in ('simple', 'tensor', None):
            pass
        elif transformer == 'pad':
            oH, oW, off_y, off_x = H + 4, W + 4, -2, -2
        elif transformer == 'crop':
            if out_shape is None:
                raise ValueError("transformer='crop' needs out_shape")
            oH, oW = (int(out_shape[-2]), int(out_shape[-1]))
            if oH > H or oW > W:
                raise ValueError('CenterCrop larger than the image is not supported')
            off_y, off_x = int(round((H - oH) / 2.0)), int(round((W - oW) / 2.0))     # torchvision center_crop
        else:
            raise ValueError(f'unknown transformer {transformer!r} (simple | pad | crop)')
        self.transformer = transformer
        self.out_shape = (C, oH, oW)
        self.cfg = nat.BatchCfg(H=H, W=W, C=C, out_H=oH, out_W=oW, crop_pad=self.crop_padding if self._crop_draws else 0,
                                flip_first=int('flip' in aug and ('crop' not in aug or aug.index('flip') < aug.index('crop'))),
                                post_off_y=off_y, post_off_x=off_x)
        self._resident = (self.data.numel() <= max_resident_bytes) if resident is None else bool(resident)
        self._dev_data = self._dev_targets = None
        self._side = None
        self._ring, self._ring_pos = [], 0      # pinned host buffers, reused round-robin (pinning a fresh buffer per batch
        #                                         costs a cudaHostAlloc, milliseconds)

    # ------------------------------------------------------------------ host side: which samples, which decisions
    def _n_local(self):
        n = self.data.shape[0]
        return len(range(self.rank, n, self.world_size))

    def __len__(self):
        n = self._n_local()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def _epoch_order(self):
        n = self.data.shape[0]
        if not self.shuffle:
            return torch.arange(n)
        if self.rng == 'torchvision':
            # DataLoader.__iter__ -> _BaseDataLoaderIter.__init__ draws `_base_seed`, then RandomSampler.__iter__ draws the
            # seed of its private generator; both from the global CPU generator, in this order
            torch.empty((), dtype=torch.int64).random_()
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            return torch.randperm(n, generator=g)
        return torch.randperm(n, generator=self._gen)

    def _draw(self, nb):
        """flip flags (nb) uint8 or None, crop offsets (nb, 2) int32 or None"""
        want_flip = 'flip' in self.data_augmentation
        if not want_flip and not self._crop_draws:
            return None, None
        flip = torch.zeros(nb, dtype=torch.uint8) if want_flip else None
        crop = torch.zeros(nb, 2, dtype=torch.int32) if self._crop_draws else None
        span = 2 * self.crop_padding + 1
        if self.rng == 'torchvision':
            # per sample, in Compose order: RandomHorizontalFlip.forward: torch.rand(1) < p;
            # RandomCrop.get_params: torch.randint(0, h - th + 1, (1,)) then torch.randint(0, w - tw + 1, (1,))
            for s in range(nb):
                for t in self.data_augmentation:
                    if t == 'flip':
                        flip[s] = int(torch.rand(1) < 0.5)
                    elif self._crop_draws:
                        crop[s, 0] = int(torch.randint(0, span, size=(1,)).item())
                        crop[s, 1] = int(torch.randint(0, span, size=(1,)).item())
        else:
            if want_flip:
                flip = (torch.rand(nb, generator=self._gen_aug) < 0.5).to(torch.uint8)
            if self._crop_draws:
                crop = torch.randint(0, span, (nb, 2), generator=self._gen_aug, dtype=torch.int32)
        return flip, crop

    def plan_epoch(self):
        """generator of (index int64 (nb), flip, crop) per batch: everything the kernel needs besides the images"""
        order = self._epoch_order()
        if self.world_size > 1:
            order = order[self.rank::self.world_size]
        n, bs = order.numel(), self.batch_size
        for lo in range(0, n, bs):
            idx = order[lo:lo + bs]
            if idx.numel() < bs and self.drop_last:
                return
            flip, crop = self._draw(idx.numel())
            yield idx, flip, crop

    # ------------------------------------------------------------------ device side
    def _setup_device(self):
        if self.device.type != 'cuda':
            raise nat.NativeError('DeviceBatchLoader produces batches on CUDA devices only (there is no CPU fallback)')
        if self._dev_targets is None:
            self._dev_targets = self.targets.to(self.device)
            if self._resident:
                self._dev_data = self.data.to(self.device)
            else:
                self.data = self.data.pin_memory()
            self._side = torch.cuda.Stream(self.device)
            self._side.wait_stream(torch.cuda.current_stream(self.device))      # the resident copy was enqueued there

    def _host_slot(self):
        """one of 3 pinned (control block, staging) pairs; waits for the copies that last read it"""
        if not self._ring:
            bs = self.batch_size
            for _ in range(3):
                slot = {'ctrl': torch.empty(25 * bs, dtype=torch.uint8, pin_memory=True), 'stage': None, 'busy': None}
                if not self._resident:
                    slot['stage'] = torch.empty((bs,) + tuple(self.data.shape[1:]), dtype=torch.uint8, pin_memory=True)
                self._ring.append(slot)
        slot = self._ring[self._ring_pos]
        self._ring_pos = (self._ring_pos + 1) % len(self._ring)
        if slot['busy'] is not None:
            slot['busy'].synchronize()
        return slot

    def _produce(self, idx, flip, crop):
        """enqueue one batch on the side stream; returns (x, y, event, host buffers kept alive until consumed)"""
        nb = idx.numel()
        C, oH, oW = self.out_shape
        # one pinned control block per batch: [index int64 nb][targets int64 nb][crop int32 2 nb][flip uint8 nb]
        slot = self._host_slot()
        ctrl = slot['ctrl'][:25 * nb]
        ctrl[:8 * nb].view(torch.int64).copy_(idx if self._resident else torch.arange(nb))
        torch.index_select(self.targets, 0, idx, out=ctrl[8 * nb:16 * nb].view(torch.int64))
        if crop is not None:
            ctrl[16 * nb:24 * nb].view(torch.int32).copy_(crop.reshape(-1))
        if flip is not None:
            ctrl[24 * nb:].copy_(flip)
        stage = None
        if not self._resident:      # host gather into pinned staging, then ONE uint8 copy (a quarter of the f32 batch)
            stage = slot['stage'][:nb]
            torch.index_select(self.data, 0, idx, out=stage)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._side):
            dctrl = ctrl.to(self.device, non_blocking=True)
            src = self._dev_data if self._resident else stage.to(self.device, non_blocking=True)
            x = torch.empty((nb, C, oH, oW), dtype=torch.float32, device=self.device)
            nat.batch_u8_to_f32(self.cfg, src, dctrl[:8 * nb].view(torch.int64),
                                dctrl[24 * nb:] if flip is not None else None,
                                dctrl[16 * nb:24 * nb].view(torch.int32) if crop is not None else None, x)
            y = dctrl[8 * nb:16 * nb].view(torch.int64)
            ev = torch.cuda.Event()
            ev.record(self._side)
            slot['busy'] = ev          # the host buffers may be rewritten once the copies queued before this event are done
        for t in (x, dctrl):        # allocated on the side stream, consumed on the caller's stream
            t.record_stream(cur)
        return x, y, ev, (ctrl, stage)

    def __iter__(self):
        self._setup_device()
        cur = torch.cuda.current_stream(self.device)
        pending = None
        for idx, flip, crop in self.plan_epoch():
            nxt = self._produce(idx, flip, crop)
            if pending is not None:
                x, y, ev, _ = pending
                cur.wait_event(ev)
                yield x, y
            pending = nxt
        if pending is not None:
            x, y, ev, _ = pending
            cur.wait_event(ev)
            yield x, y

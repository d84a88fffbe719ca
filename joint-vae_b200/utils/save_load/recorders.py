"""LossRecorder with the reference's interface and on-disk format (utils/save_load/recorders.py:13-370): the per-sample
loss / score / logit tensors of a test or OOD sweep, sample axis last, saved as `record-<set>.pth` =
`torch.save(recorder.__dict__)`.  It is the wire format between scoring and every result tool of the reference.

What differs is how batches land in the buffers: the reference builds an index tensor from a Python `range` for every
key of every batch and `scatter_`s (recorders.py:335-369), and `__getitem__` / `save` gather through another range
tensor.  Here a batch is one strided copy into a slice of the preallocated device buffer and reads are slices, so
recording costs one small copy kernel per key and never touches the host.  Files written by either implementation load
in the other (tests/test_recorder.py checks both directions against a file written by the unmodified reference).
"""
import logging
import os
import re

import numpy as np
import torch


class LossRecorder:

    _file_pattern = 'record-{w}.pth'
    _sample_dim = -1

    def __init__(self, batch_size, num_batch=1, device=None, **tensors):
        self.last_batch_size = None
        self._seed = None
        self._num_batch = 0
        self._samples = 0
        self.batch_size = batch_size
        self.reset()
        self._tensors = {}
        self.device = device
        if tensors:
            self._create_tensors(num_batch, device=device, **tensors)

    # ------------------------------------------------------------------ buffers
    def _create_tensors(self, num_batch, device=None, **tensors):
        assert not self._tensors
        self._num_batch = num_batch
        self._samples = num_batch * self.batch_size
        if not device and not self.device:
            device = next(iter(tensors.values())).device
        self.device = device or self.device
        for k, t in tensors.items():
            shape = list(t.shape)
            shape[self._sample_dim] = self._samples
            self._tensors[k] = torch.zeros(shape, dtype=t.dtype, device=self.device)
        self.last_batch_size = self.batch_size

    def to(self, device):
        for k in self._tensors:
            self._tensors[k] = self._tensors[k].to(device)

    def reset(self, seed=False):
        self._recorded_batches = 0
        if self._seed is None or seed:
            self._seed = np.random.randint(1, int(1e8))
        self.last_batch_size = self.batch_size

    def init_seed_for_dataloader(self):
        """same data order for every pass over a set (recorders.py:72-79)"""
        self._initial_seed = torch.seed()
        torch.manual_seed(self._seed)

    def restore_seed(self):
        torch.manual_seed(self._initial_seed)

    # ------------------------------------------------------------------ container protocol
    def keys(self):
        return self._tensors.keys()

    def __len__(self):
        return self._recorded_batches

    def __iter__(self):
        return iter(self._tensors)

    def __repr__(self):
        return 'Recorder for ' + ' '.join(str(k) for k in self.keys())

    @property
    def recorded_samples(self):
        return (len(self) - 1) * self.batch_size + self.last_batch_size

    def __getitem__(self, k):
        return self._tensors[k].narrow(self._sample_dim, 0, max(self.recorded_samples, 0)).clone()

    def pop(self, k):
        t = self[k]
        self._tensors.pop(k)
        return t

    # ------------------------------------------------------------------ capacity
    @property
    def num_batch(self):
        return self._num_batch

    @num_batch.setter
    def num_batch(self, n):
        if not self._tensors:
            return
        height = next(iter(self._tensors.values())).shape[self._sample_dim]
        want = n * self.batch_size
        if want > height:
            for k, t in self._tensors.items():
                shape = list(t.shape)
                shape[self._sample_dim] = want
                grown = torch.zeros(shape, dtype=t.dtype, device=t.device)
                grown.narrow(self._sample_dim, 0, height).copy_(t)
                self._tensors[k] = grown
        self._num_batch = n
        self._samples = n * self.batch_size
        self._recorded_batches = min(n, self._recorded_batches)

    def has_batch(self, number, only_full=False):
        """number starts at 0"""
        if number == len(self) - 1:
            return not only_full or self.last_batch_size == self.batch_size
        return number < self._recorded_batches

    # ------------------------------------------------------------------ batches
    def append_batch(self, extend=True, **tensors):
        if not self._tensors:
            self._create_tensors(1, **tensors)
        start = self._recorded_batches * self.batch_size
        if start + self.batch_size > self._samples:
            if not extend:
                raise IndexError
            self.num_batch *= 2
        sizes = {tensors[k].shape[self._sample_dim] for k in tensors}
        assert len(sizes) == 1, 'all batches have to be of same size'
        size = sizes.pop()
        assert size <= self.batch_size, 'appended batch to large'
        assert self.last_batch_size == self.batch_size      # only the last batch of a sweep may be short
        self.last_batch_size = size
        for k, t in tensors.items():
            if k not in self._tensors:
                raise KeyError(k)
            self._tensors[k].narrow(self._sample_dim, start, size).copy_(t)      # one strided device copy
        self._recorded_batches += 1

    def get_batch(self, i, *which, device=None, force_dict=False):
        if not which:
            if not self.keys():
                raise KeyError('empty recorder')
            return self.get_batch(i, *self.keys(), force_dict=True)
        if len(which) > 1 or force_dict:
            return {w: self.get_batch(i, w) for w in which}
        if not self.has_batch(i):
            raise IndexError(f'{i} >= {len(self)}')
        start = i * self.batch_size
        size = self.last_batch_size if i == len(self) - 1 else self.batch_size
        t = self._tensors[which[0]]
        if device:
            t = t.to(device)
        return t.narrow(self._sample_dim, start, size).clone()

    # ------------------------------------------------------------------ files
    def save(self, file_path, cut=True, append=False):
        if append:
            try:
                already = self.load(file_path)
                already.merge(self)
            except FileNotFoundError:
                already = self
            already.save(file_path, cut=cut, append=False)
            return
        if cut:
            self.num_batch = len(self)
            n = self.recorded_samples
            for k in self._tensors:
                self._tensors[k] = self._tensors[k].narrow(self._sample_dim, 0, n).contiguous()
        torch.save(self.__dict__, file_path)

    @classmethod
    def load(cls, file_path, device=None, **kw):
        if 'map_location' not in kw and not torch.cuda.is_available():
            kw['map_location'] = torch.device('cpu')
            device = 'cpu'
        kw.setdefault('weights_only', False)
        d = torch.load(file_path, **kw)
        r = cls(d['batch_size'], d['_num_batch'], **d['_tensors'])
        for k in ('_seed', '_tensors', '_recorded_batches', '_aux'):
            if k in d:
                setattr(r, k, d[k])
        for k in d:
            if not k.startswith('_'):
                setattr(r, k, d[k])
        if isinstance(r.last_batch_size, dict):       # files of older reference versions
            r.last_batch_size = next(iter(r.last_batch_size.values()))
        if device:
            r.device = device
            for k in r._tensors:
                if r._tensors[k].device != torch.device(device):
                    r._tensors[k] = r._tensors[k].to(device)
        return r

    @classmethod
    def loadall(cls, dir_path, *w, file_name=None, output='recorders', **kw):
        """every `record-<name>.pth` of a directory (or the named ones); output 'recorders' or 'paths'"""
        file_name = file_name or cls._file_pattern
        pick = (lambda p: cls.load(p, **kw)) if output.startswith('record') else (lambda p: p)
        found = {}
        if not w:
            pattern = file_name.replace('.', r'\.').replace('{w}', '(?P<name>.+)')
            for f in os.listdir(dir_path):
                m = re.match(pattern, f)
                if m:
                    found[m.group('name')] = pick(os.path.join(dir_path, f))
        for word in w:
            path = os.path.join(dir_path, file_name.format(w=word))
            if os.path.exists(path):
                found[word] = pick(path)
            else:
                logging.warning('%s not found', os.path.basename(path))
        return found

    # ------------------------------------------------------------------ algebra
    def copy(self, device=None):
        new = type(self)(self.batch_size)
        for i in range(len(self)):
            new.append_batch(**self.get_batch(i, device=device))
        return new

    def merge(self, other, axis='samples'):
        assert isinstance(other, type(self))
        assert axis in ('samples', 'keys'), 'axis has to be either sampels or keys '
        if axis == 'samples':
            total = self.recorded_samples + other.recorded_samples
            common = set(self) & set(other)
            joined = {k: torch.cat((self[k], other[k]), dim=self._sample_dim) for k in common}
            self.num_batch = len(self) + other.recorded_samples // self.batch_size + 1
            self._tensors = joined
            self.last_batch_size = (total - 1) % self.batch_size + 1
            self._recorded_batches = (total - 1) // self.batch_size + 1
        else:
            assert self.recorded_samples == other.recorded_samples
            common = set(self) & set(other)
            assert not common, 'can not merge recorder with common keys ({})'.format(', '.join(common))
            self._tensors.update(other._tensors)

    def split(self, *keys, keep=False):
        other = self.copy()
        for k in list(self):
            if k in keys:
                if not keep:
                    self.pop(k)
            else:
                other.pop(k)
        return other

"""TEST INFRASTRUCTURE ONLY: the input path of the reference re-assembled from its own building blocks, as the oracle for
joint-vae_b200/utils/batch_loader.py.  A torchvision-style dataset over a uint8 array (`Image.fromarray(self.data[i])`,
as torchvision's CIFAR10 / MNIST do), the transform list utils/torch_load.py:405-426, 472-473 composes
(train_transforms in data_augmentation order, then the post transforms, ToTensor last) and the DataLoader of
cvae.py:2245-2249 (shuffle=True, num_workers=0)."""
import numpy as np
import torch
from PIL import Image
from torchvision import transforms


class ArrayImages(torch.utils.data.Dataset):
    def __init__(self, data, targets, transform):
        self.data, self.targets, self.transform = data, targets, transform

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        a = self.data[i]
        img = Image.fromarray(a if a.ndim == 3 else a, mode=None if a.ndim == 3 else 'L')
        return self.transform(img), int(self.targets[i])


def reference_transform(shape, data_augmentation, transformer, out_shape=None, imagenet=False):
    """torch_load.py:405-426"""
    train = []
    for t in data_augmentation:
        if t == 'flip':
            t_ = transforms.RandomHorizontalFlip()
        if t == 'crop':
            size = shape[1:]
            padding = 0 if imagenet else size[0] // 8
            t_ = transforms.RandomCrop(size, padding=padding, padding_mode='edge')
        train.append(t_)
    post = []
    if transformer == 'crop':
        post.append(transforms.CenterCrop(out_shape[1:]))
    elif transformer == 'pad':
        post.append(transforms.Pad(2))
    post.append(transforms.ToTensor())
    return transforms.Compose(train + post)


def reference_batches(data, targets, batch_size, data_augmentation, transformer, out_shape=None, shuffle=True, epochs=1):
    data = np.asarray(data)
    shape = (data.shape[3] if data.ndim == 4 else 1,) + tuple(data.shape[1:3])
    ds = ArrayImages(data, targets, reference_transform(shape, data_augmentation, transformer, out_shape))
    loader = torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=0)
    out = []
    for _ in range(epochs):
        for x, y in loader:
            out.append((x, y))
    return out

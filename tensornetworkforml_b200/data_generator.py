"""Caller-side input contract: mirror of the reference's ``data_generator`` (DG:6-194).

Host-side and one-off: same function names, arguments, RNG call order and loader output
(``list[(ndarray(S,2), int)]`` per batch), so seeded scripts written against the reference produce the same data.
The feature map itself is also available on the device (``tnml_feature_map``) for raw-pixel inputs.
"""
from __future__ import annotations

import numpy as np
from torch.utils.data import DataLoader, Dataset, SubsetRandomSampler


def create_dataset(n_samples, linear_dim=5, sigma=0.5, prob_zero=0.5):
    """Two-diagonals toy images (DG:6-52): label 0 = anti-diagonal, 1 = main diagonal, mixed with uniform noise.

    RNG order is the reference's: ``np.random.choice`` for the labels, then ``np.random.rand`` for the noise."""
    one = np.eye(linear_dim)
    zero = one[::-1, :]
    labels = np.random.choice([0, 1], size=n_samples, p=[prob_zero, 1 - prob_zero])
    data = np.where((labels == 0)[:, None, None], zero[None], one[None]).astype(np.float64)
    noise = np.random.rand(n_samples, linear_dim, linear_dim) * sigma
    return data * (1 - sigma) + noise, labels


def stripe_templates(linear_dim=14, n_labels=10):
    """The ``n_labels`` fixed stripe images of create_multiclass_dataset (even labels: row stripes, odd: column)."""
    templates = np.zeros((n_labels, linear_dim, linear_dim))
    period = max(2, (n_labels + 1) // 2)
    idx = np.arange(linear_dim)
    for k in range(n_labels):
        stripe = ((idx + k // 2) % period == 0).astype(np.float64)
        if k % 2 == 0:
            templates[k] = stripe[:, None] * np.ones((1, linear_dim))
        else:
            templates[k] = np.ones((linear_dim, 1)) * stripe[None, :]
    return templates


def create_multiclass_dataset(n_samples, linear_dim=14, n_labels=10, sigma=0.7):
    """Additive extension: ``n_labels`` fixed stripe templates (even labels: row
    stripes, odd labels: column stripes), ``label ~ randint``, mixed with uniform noise exactly like DG:49-50."""
    templates = stripe_templates(linear_dim, n_labels)
    labels = np.random.randint(0, n_labels, n_samples)
    noise = np.random.rand(n_samples, linear_dim, linear_dim) * sigma
    return templates[labels] * (1 - sigma) + noise, labels


def _generate_on_device(templates, n_samples, sigma, prob_first, seed, device):
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("the device-side generators need a CUDA device (there is no CPU fallback)")
    dev = torch.device(device if device is not None else "cuda")
    n_labels, side = templates.shape[0], templates.shape[1]
    t = torch.from_numpy(np.ascontiguousarray(templates.reshape(n_labels, -1), dtype=np.float64)).to(dev)
    x = torch.empty((n_samples, side, side), dtype=torch.float64, device=dev)
    labels = torch.empty(n_samples, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("tnml_generate_dataset", t.data_ptr(), x.data_ptr(), labels.data_ptr(), n_samples, side * side,
                  n_labels, float(sigma), float(prob_first), int(seed) & (2 ** 64 - 1),
                  torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.current_stream(dev).synchronize()       # `t` is released when this function returns
    return x, labels


def create_dataset_device(n_samples, linear_dim=5, sigma=0.5, prob_zero=0.5, seed=0, device=None):
    """create_dataset (DG:6-52) generated ON the device: (x (n, d, d) float64, labels (n,) int32) as CUDA tensors, ready
    for ``Network.forward_raw`` -- nothing crosses PCIe.  Same construction (templates, label probabilities, noise
    mixing); the random numbers come from a counter-based generator, not NumPy's stream, so use the host function when
    a seeded reference run has to be reproduced."""
    one = np.eye(linear_dim)
    return _generate_on_device(np.stack([one[::-1, :], one]), n_samples, sigma, prob_zero, seed, device)


def create_multiclass_dataset_device(n_samples, linear_dim=14, n_labels=10, sigma=0.7, seed=0, device=None):
    """create_multiclass_dataset generated on the device (see create_dataset_device)."""
    return _generate_on_device(stripe_templates(linear_dim, n_labels), n_samples, sigma, -1.0, seed, device)


def get_MNIST_dataset(data_root_dir='./datasets', download=True):
    """MNIST as NumPy arrays (DG:55-87).  Needs torchvision and, with download=True, network access."""
    from torchvision.datasets import MNIST
    train = MNIST(data_root_dir, train=True, download=download)
    test = MNIST(data_root_dir, train=False, download=download)

    def unpack(ds):
        return np.array([np.array(s[0]) for s in ds]), np.array([np.array(s[1]) for s in ds])
    train_data, train_labels = unpack(train)
    test_data, test_labels = unpack(test)
    return train_data, train_labels, test_data, test_labels


class NumpyDataset(Dataset):
    """NumPy (data, label) pairs as a torch Dataset (DG:90-122)."""

    def __init__(self, data, label):
        self.data = data
        self.label = label

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return (self.data[index], self.label[index])


def psi(x):
    """phi(x) = [sin(pi x/2), cos(pi x/2)] on the last axis -- sin first, as in the code (DG:165-167)."""
    x = np.asarray(x)
    return np.stack((np.sin(np.pi * x / 2), np.cos(np.pi * x / 2)), axis=-1)


def _identity_collate(batch):
    return batch


def prepare_dataset(data, label, train_perc, val_perc, train_batch_size, val_batch_size, test_batch_size):
    """Feature-map, split and wrap in DataLoaders (DG:125-194).  Batches are ``list[(x:(S,2), y:int)]``."""
    x = psi(data.reshape(len(data), -1))
    m = int(len(x) * train_perc)
    train_set = NumpyDataset(x[:m], label[:m])
    test_set = NumpyDataset(x[m:], label[m:])
    train_len = int(m * (1 - val_perc))
    train_sampler = SubsetRandomSampler(np.arange(train_len))
    val_sampler = SubsetRandomSampler(np.arange(train_len, m))
    train_loader = DataLoader(train_set, train_batch_size, sampler=train_sampler, drop_last=True,
                              collate_fn=_identity_collate)
    val_loader = DataLoader(train_set, val_batch_size, sampler=val_sampler, drop_last=True,
                            collate_fn=_identity_collate)
    test_loader = DataLoader(test_set, test_batch_size, drop_last=False, collate_fn=_identity_collate)
    return train_loader, val_loader, test_loader

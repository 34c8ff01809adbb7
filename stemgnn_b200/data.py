"""Device-resident data path.

`DeviceWindowLoader` iterates a `ForecastDataset` exactly like
`torch.utils.data.DataLoader(dataset, batch_size, shuffle, drop_last, num_workers=0)` does in the
reference handler (handler.py:135-138): the SAME samplers produce the SAME index batches (so a seeded run
visits identical batches), but the normalised series lives in HBM once and every batch is built there by one
gather kernel (`stemgnn_gather_windows`) instead of B `__getitem__` calls, a collate and an H2D copy.
The host DataLoader tops out near 70 k windows/s single-threaded; the CUDA forward runs at
27 k windows/s per GPU, so the host path would cost a third of the inference time.
"""
import ctypes

import numpy as np
import torch
import torch.utils.data as torch_data

from . import _lib, runtime


class DeviceWindowLoader:
    def __init__(self, dataset, batch_size, shuffle=False, drop_last=False, device="cuda:0", generator=None):
        self.dataset = dataset
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceWindowLoader needs a CUDA device (there is no CPU fallback)")
        self.batch_size, self.drop_last = batch_size, drop_last
        # same float64 -> float32 conversion as ForecastDataset.__getitem__ (forecast_dataloader.py:61-62)
        series = torch.from_numpy(np.ascontiguousarray(dataset.data)).type(torch.float)
        self.series = series.to(self.device)
        self.T, self.N = self.series.shape
        self.W, self.H = dataset.window_size, dataset.horizon
        self.end_idx = torch.tensor(dataset.x_end_idx, dtype=torch.int32)
        base = torch_data.RandomSampler(dataset, generator=generator) if shuffle else torch_data.SequentialSampler(dataset)
        self.batch_sampler = torch_data.BatchSampler(base, batch_size, drop_last)

    def __len__(self):
        return len(self.batch_sampler)

    def __iter__(self):
        lib = _lib.load()
        # torch's DataLoader iterator draws its `_base_seed` from the global RNG before the sampler draws the
        # permutation seed; consume the same value so that a seeded run visits the same batches either way
        torch.empty((), dtype=torch.int64).random_()
        for indices in self.batch_sampler:
            ends = self.end_idx[indices].to(self.device, non_blocking=True)
            B = len(indices)
            x = torch.empty(B, self.W, self.N, dtype=torch.float32, device=self.device)
            y = torch.empty(B, self.H, self.N, dtype=torch.float32, device=self.device)
            rc = lib.stemgnn_gather_windows(self.series.data_ptr(), self.T, self.N, ends.data_ptr(), B, self.W,
                                            self.H, x.data_ptr(), y.data_ptr(), runtime._stream_ptr(self.device))
            _lib.check(rc, "stemgnn_gather_windows")
            yield x, y

"""Device-resident data path: `DeviceWindowLoader` yields bit-identical batches, in
the identical order, as the reference's host `DataLoader(ForecastDataset)` — only the place where the
windows are assembled changes."""
import time

import numpy as np
import pytest
import torch
import torch.utils.data as torch_data

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dataset(T=700, N=37, W=12, H=3, method="z_score", interval=1):
    from data_loader.forecast_dataloader import ForecastDataset
    rng = np.random.default_rng(4)
    data = rng.normal(size=(T, N)) * 3 + 7
    data[5, 3] = np.nan
    stat = {"mean": np.nanmean(data, 0).tolist(), "std": np.nanstd(data, 0).tolist()} if method else None
    return ForecastDataset(data, W, H, normalize_method=method, norm_statistic=stat, interval=interval)


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("bs,interval", [(32, 1), (29, 1), (64, 3)])
def test_device_loader_matches_host_dataloader_bit_exact(shuffle, bs, interval):
    from stemgnn_b200.data import DeviceWindowLoader
    ds = _dataset(interval=interval)
    host = torch_data.DataLoader(ds, batch_size=bs, shuffle=shuffle, drop_last=False, num_workers=0)
    dev = DeviceWindowLoader(ds, bs, shuffle=shuffle, drop_last=False, device=DEV)
    assert len(host) == len(dev)
    for epoch in range(2):
        torch.manual_seed(100 + epoch)
        hb = [(x.clone(), y.clone()) for x, y in host]
        torch.manual_seed(100 + epoch)
        db = [(x.cpu(), y.cpu()) for x, y in dev]
        assert len(hb) == len(db)
        for (hx, hy), (dx, dy) in zip(hb, db):
            assert torch.equal(hx, dx) and torch.equal(hy, dy)
    assert hb[-1][0].shape[0] == len(ds) - (len(hb) - 1) * bs          # ragged last batch kept


def test_device_loader_throughput_vs_host():
    from stemgnn_b200.data import DeviceWindowLoader
    ds = _dataset(T=6000, N=358)
    host = torch_data.DataLoader(ds, batch_size=32, shuffle=True, drop_last=False, num_workers=0)
    dev = DeviceWindowLoader(ds, 32, shuffle=True, device=DEV)
    t0 = time.perf_counter()
    n = sum(x.to(DEV, non_blocking=True).shape[0] for x, _ in host)
    torch.cuda.synchronize()
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    m = sum(x.shape[0] for x, _ in dev)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    assert n == m == len(ds)
    print(f"host DataLoader {n / t_host:.0f} windows/s, device loader {m / t_dev:.0f} windows/s")
    assert t_dev < t_host

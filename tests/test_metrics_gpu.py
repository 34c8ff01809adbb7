"""Device-side validation metrics (stemgnn_eval_metrics) against the reference's numpy path
(data_loader.forecast_dataloader.de_normalized + utils.math_utils.evaluate, handler.py:74-82)."""
import numpy as np
import pytest
import torch

from data_loader.forecast_dataloader import de_normalized
from utils.math_utils import evaluate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["z_score", "min_max", None])
@pytest.mark.parametrize("count,H,N", [(1, 1, 1), (37, 3, 140), (1000, 12, 33)])
def test_device_metrics_match_numpy(method, count, H, N):
    from stemgnn_b200.metrics import device_evaluate
    rng = np.random.default_rng(count * 31 + N)
    f_norm = rng.normal(size=(count, H, N))
    t_norm = rng.normal(size=(count, H, N)).astype(np.float32)
    stat = None
    if method == "z_score":
        std = rng.uniform(0.5, 20, size=N)
        std[0] = 0.0                                            # zero std is replaced by 1 (reference :17)
        stat = dict(mean=rng.normal(size=N) * 50, std=list(std))
    elif method == "min_max":
        lo = rng.normal(size=N) * 10
        stat = dict(min=lo, max=lo + rng.uniform(1, 100, size=N))
    f, t = (de_normalized(f_norm, method, stat), de_normalized(t_norm, method, stat)) if method else (f_norm, t_norm)
    want = evaluate(t, f), evaluate(t, f, by_node=True), evaluate(t_norm, f_norm)
    got = device_evaluate(torch.from_numpy(f_norm).cuda(), torch.from_numpy(t_norm).cuda(), method, stat)
    for g3, w3 in zip(got, want):
        for g, w in zip(g3, w3):
            np.testing.assert_allclose(g, w, rtol=1e-12, atol=0)


def test_device_metrics_nan_semantics():
    """0/0 in the MAPE ratio is NaN in numpy (np.minimum propagates it); x/0 is clipped to 5."""
    from stemgnn_b200.metrics import device_evaluate
    f = np.array([[[0.0, 1.0, 2.0]]]); t = np.array([[[0.0, 0.0, 2.0]]], dtype=np.float32)
    with np.errstate(all="ignore"):
        want = evaluate(t, f, by_node=True)
    got = device_evaluate(torch.from_numpy(f).cuda(), torch.from_numpy(t).cuda(), None, None)[1]
    assert np.isnan(got[0][0]) and np.isnan(want[0][0])
    assert got[0][1] == 5.0 == want[0][1]
    np.testing.assert_allclose(got[0][2], want[0][2])


def test_validate_uses_device_metrics(monkeypatch):
    """handler.validate on CUDA gives the same dict as its host-numpy path (STEMGNN_HOST_METRICS)."""
    import argparse
    from models import handler
    from models.base_model import Model
    from data_loader.forecast_dataloader import ForecastDataset
    rng = np.random.default_rng(0)
    data = rng.normal(size=(200, 12)).cumsum(axis=0) + 30.0
    stat = dict(mean=data.mean(0), std=list(data.std(0)))
    ds = ForecastDataset(data, window_size=12, horizon=3, normalize_method="z_score", norm_statistic=stat)
    torch.manual_seed(0)
    m = Model(12, 2, 12, 5, horizon=3).to("cuda:0")
    loader = handler._make_loader(ds, 32, False, "cuda:0")
    a = handler.validate(m, loader, "cuda:0", "z_score", stat, 12, 12, 3)
    monkeypatch.setenv("STEMGNN_HOST_METRICS", "1")
    loader = handler._make_loader(ds, 32, False, "cuda:0")
    b = handler.validate(m, loader, "cuda:0", "z_score", stat, 12, 12, 3)
    for k in ("mae", "mape", "rmse"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-10)
    for k in ("mae_node", "mape_node", "rmse_node"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-10)

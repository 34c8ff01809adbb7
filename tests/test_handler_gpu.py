"""End-to-end drop-in check on the GPU: the re-hosted `models.handler.train/test` (the entry points the
reference main.py calls) drive the CUDA model through an epoch of training, checkpointing (whole-module
pickle, as the reference does), validation and test on an ECG-like synthetic series."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(**kw):
    d = dict(train=True, evaluate=True, dataset="synthetic", window_size=12, horizon=3, train_length=7,
             valid_length=2, test_length=1, epoch=2, lr=1e-3, multi_layer=5, device="cuda:0",
             validate_freq=1, batch_size=32, norm_method="z_score", optimizer="RMSProp", early_stop=False,
             exponential_decay_step=5, decay_rate=0.5, dropout_rate=0.5, leakyrelu_rate=0.2)
    d.update(kw)
    return argparse.Namespace(**d)


def _series(T=420, N=20, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(T)[:, None]
    phase = rng.uniform(0, 6.28, size=(1, N))
    base = np.sin(2 * np.pi * t / 24.0 + phase) + 0.3 * np.sin(2 * np.pi * t / 7.0 + 2 * phase)
    mix = rng.normal(size=(N, N)) * 0.1 + np.eye(N)
    return (base @ mix + 0.05 * rng.normal(size=(T, N))) * 3.0 + 10.0


def test_train_validate_test_roundtrip(tmp_path, capsys):
    from models import handler
    data = _series()
    n = len(data)
    tr, va, te = data[: int(0.7 * n)], data[int(0.7 * n): int(0.9 * n)], data[int(0.9 * n):]
    out_train, out_test = str(tmp_path / "train"), str(tmp_path / "test")
    os.makedirs(out_train); os.makedirs(out_test)
    torch.manual_seed(0)
    metrics, stat = handler.train(tr, va, _args(), out_train)
    assert set(metrics) >= {"mae", "mape", "rmse"} and np.isfinite(metrics["mae"])
    assert os.path.exists(os.path.join(out_train, "norm_stat.json"))
    assert os.path.exists(os.path.join(out_train, "_stemgnn.pt")) and os.path.exists(os.path.join(out_train, "1_stemgnn.pt"))
    handler.test(te, _args(), out_train, out_test)
    assert os.path.exists(os.path.join(out_test, "predict.csv"))
    out = capsys.readouterr().out
    assert "Total Trainable Params" in out and "Performance on test set" in out
    # the checkpoint is the pickled drop-in Model and keeps working after reload
    m = handler.load_model(out_train)
    assert type(m).__module__ == "models.base_model"
    x = torch.randn(5, 12, 20, device="cuda:0")
    with torch.no_grad():
        f, a = m(x)
    assert f.shape == (5, 3, 20) and torch.isfinite(f).all()
    # training reduced the normalised error below the "predict the mean" level (= 1.0 for z-scored data)
    losses = [float(l.split("train_total_loss")[1]) for l in out.splitlines() if "train_total_loss" in l]
    assert losses[-1] < losses[0] and losses[-1] < 1.0


def test_adam_minmax_and_early_stop(tmp_path):
    from models import handler
    data = _series(T=300, N=12, seed=3)
    tr, va = data[:210], data[210:]
    out = str(tmp_path / "t"); os.makedirs(out)
    m, stat = handler.train(tr, va, _args(optimizer="Adam", norm_method="min_max", epoch=3, early_stop=True,
                                          early_stop_step=1, batch_size=16), out)
    assert "min" in stat and np.isfinite(m["rmse"])

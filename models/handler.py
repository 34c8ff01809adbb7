"""Train / evaluate driver with the entry points of the reference `models/handler.py`
(microsoft/StemGNN): `train(train_data, valid_data, args, result_file)`,
`test(test_data, args, result_train_file, result_test_file)`, `validate(...)`, `inference(...)`,
`save_model` / `load_model` — so the reference `main.py` runs unchanged against this package
(`python run_main.py --device cuda:0 ...`).  The model it drives is `models.base_model.Model`,
whose forward/backward are the stemgnn_b200 CUDA kernels; everything here is host orchestration.

Differences from the reference that are forced by the modern stack:
`np.float` -> `float`, `torch.load(..., weights_only=False)`.  Behaviour kept on purpose: the
checkpoint is the whole pickled module, epoch 0 is saved under the "best" name (reference :21),
LR decays every `exponential_decay_step` epochs, early stopping counts non-improving validations.
"""
import json
import os
import time
from datetime import datetime

import numpy as np
import torch
import torch.nn as nn
import torch.utils.data as torch_data

from data_loader.forecast_dataloader import ForecastDataset, de_normalized
from models.base_model import Model
from utils.math_utils import evaluate

_CKPT = '_stemgnn.pt'


def _ckpt_path(model_dir, epoch):
    return os.path.join(model_dir, (str(epoch) if epoch else '') + _CKPT)


def save_model(model, model_dir, epoch=None):
    if model_dir is None:
        return
    os.makedirs(model_dir, exist_ok=True)
    with open(_ckpt_path(model_dir, epoch), 'wb') as f:
        torch.save(model, f)


def load_model(model_dir, epoch=None):
    if not model_dir:
        return None
    os.makedirs(model_dir, exist_ok=True)
    path = _ckpt_path(model_dir, epoch)
    if not os.path.exists(path):
        return None
    with open(path, 'rb') as f:
        return torch.load(f, weights_only=False)


def _inference_tensors(model, dataloader, device, node_cnt, window_size, horizon):
    """Rolling forecast: the model emits `len_out` steps per call; the window is shifted by the
    prediction until `horizon` steps exist (one call per batch when the model emits the full
    horizon).  Returns (forecast, target) numpy arrays of shape (count, horizon, node).
    Forecasts and targets stay on the device until the end of the pass (one D2H copy instead of the
    reference's two per batch, handler.py:60-63); the values are the same."""
    forecasts, targets = [], []
    model.eval()
    with torch.no_grad():
        for inputs, target in dataloader:
            inputs, target = inputs.to(device), target.to(device)
            steps = torch.zeros(inputs.size(0), horizon, node_cnt, dtype=torch.float64, device=inputs.device)
            done = 0
            while done < horizon:
                out, _ = model(inputs)
                len_out = out.size(1)
                if len_out == 0:
                    raise Exception('Get blank inference result')
                inputs[:, :window_size - len_out, :] = inputs[:, len_out:window_size, :].clone()
                inputs[:, window_size - len_out:, :] = out.clone()
                take = min(horizon - done, len_out)
                steps[:, done:done + take, :] = out[:, :take, :].detach()
                done += take
            forecasts.append(steps)
            targets.append(target.detach())
    return torch.cat(forecasts, dim=0), torch.cat(targets, dim=0)


def inference(model, dataloader, device, node_cnt, window_size, horizon):
    """Reference signature (handler.py:41-71): (forecast, target) numpy arrays of shape (count, horizon, node)."""
    f, t = _inference_tensors(model, dataloader, device, node_cnt, window_size, horizon)
    return f.cpu().numpy(), t.cpu().numpy()


def validate(model, dataloader, device, normalize_method, statistic, node_cnt, window_size, horizon,
             result_file=None):
    f_dev, t_dev = _inference_tensors(model, dataloader, device, node_cnt, window_size, horizon)
    if f_dev.is_cuda and os.environ.get('STEMGNN_HOST_METRICS') is None:
        # de-normalisation + MAPE/MAE/RMSE reductions on the device: one D2H copy of (6, N) float64 per validation
        # (reference: de_normalized + 3 x evaluate in numpy on the host, handler.py:74-82)
        from stemgnn_b200.metrics import device_evaluate
        score, score_by_node, score_norm = device_evaluate(f_dev, t_dev, normalize_method, statistic)
        forecast = target = None
    else:
        forecast_norm, target_norm = f_dev.cpu().numpy(), t_dev.cpu().numpy()
        if normalize_method and statistic:
            forecast = de_normalized(forecast_norm, normalize_method, statistic)
            target = de_normalized(target_norm, normalize_method, statistic)
        else:
            forecast, target = forecast_norm, target_norm
        score = evaluate(target, forecast)
        score_by_node = evaluate(target, forecast, by_node=True)
        score_norm = evaluate(target_norm, forecast_norm)
    print(f'NORM: MAPE {score_norm[0]:7.9%}; MAE {score_norm[1]:7.9f}; RMSE {score_norm[2]:7.9f}.')
    print(f'RAW : MAPE {score[0]:7.9%}; MAE {score[1]:7.9f}; RMSE {score[2]:7.9f}.')
    if result_file:
        os.makedirs(result_file, exist_ok=True)
        if forecast is None:         # the CSV dumps need the first-step rows on the host (test() only)
            f0, t0 = f_dev[:, :1, :].cpu().numpy(), t_dev[:, :1, :].cpu().numpy()
            if normalize_method and statistic:
                f0, t0 = de_normalized(f0, normalize_method, statistic), de_normalized(t0, normalize_method, statistic)
            forecast, target = f0, t0
        pred, truth = forecast[:, 0, :], target[:, 0, :]
        np.savetxt(f'{result_file}/target.csv', truth, delimiter=",")
        np.savetxt(f'{result_file}/predict.csv', pred, delimiter=",")
        np.savetxt(f'{result_file}/predict_abs_error.csv', np.abs(pred - truth), delimiter=",")
        np.savetxt(f'{result_file}/predict_ape.csv', np.abs((pred - truth) / truth), delimiter=",")
    return dict(mae=score[1], mae_node=score_by_node[1], mape=score[0], mape_node=score_by_node[0],
                rmse=score[2], rmse_node=score_by_node[2])


def _make_loader(dataset, batch_size, shuffle, device):
    """Same batches as the reference's `DataLoader(dataset, batch_size, shuffle=..., drop_last=False,
    num_workers=0)` (handler.py:135-138, :200-201).  On a CUDA device the windows are gathered in HBM by
    `stemgnn_b200.data.DeviceWindowLoader` (identical sampler objects, so identical index order)."""
    if torch.device(device).type == 'cuda' and os.environ.get('STEMGNN_HOST_LOADER') is None:
        from stemgnn_b200.data import DeviceWindowLoader
        return DeviceWindowLoader(dataset, batch_size, shuffle=shuffle, drop_last=False, device=device)
    return torch_data.DataLoader(dataset, batch_size=batch_size, drop_last=False, shuffle=shuffle, num_workers=0)


def _norm_statistic(train_data, method):
    if method == 'z_score':
        return {"mean": np.mean(train_data, axis=0).tolist(), "std": np.std(train_data, axis=0).tolist()}
    if method == 'min_max':
        return {"min": np.min(train_data, axis=0).tolist(), "max": np.max(train_data, axis=0).tolist()}
    return None


def train(train_data, valid_data, args, result_file):
    node_cnt = train_data.shape[1]
    model = Model(node_cnt, 2, args.window_size, args.multi_layer, horizon=args.horizon)
    model.to(args.device)
    if len(train_data) == 0:
        raise Exception('Cannot organize enough training data')
    if len(valid_data) == 0:
        raise Exception('Cannot organize enough validation data')

    normalize_statistic = _norm_statistic(train_data, args.norm_method)
    if normalize_statistic is not None:
        with open(os.path.join(result_file, 'norm_stat.json'), 'w') as f:
            json.dump(normalize_statistic, f)

    # On a CUDA device the step (zero_grad .. optimizer.step, handler.py:160-166 of the reference) is one captured CUDA
    # graph with a fused optimiser kernel and a device-side loss accumulator (stemgnn_b200.trainer.FusedTrainer);
    # STEMGNN_EAGER_TRAIN=1 keeps the reference's eager torch loop.
    fused = None
    if torch.device(args.device).type == 'cuda' and os.environ.get('STEMGNN_EAGER_TRAIN') is None:
        from stemgnn_b200.trainer import FusedTrainer
        fused = FusedTrainer(model, optimizer=args.optimizer, lr=args.lr, eps=1e-08, betas=(0.9, 0.999))
        optim = scheduler = None
    else:
        if args.optimizer == 'RMSProp':
            optim = torch.optim.RMSprop(params=model.parameters(), lr=args.lr, eps=1e-08)
        else:
            optim = torch.optim.Adam(params=model.parameters(), lr=args.lr, betas=(0.9, 0.999))
        scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer=optim, gamma=args.decay_rate)

    ds_kw = dict(window_size=args.window_size, horizon=args.horizon, normalize_method=args.norm_method,
                 norm_statistic=normalize_statistic)
    train_loader = _make_loader(ForecastDataset(train_data, **ds_kw), args.batch_size, True, args.device)
    valid_loader = _make_loader(ForecastDataset(valid_data, **ds_kw), args.batch_size, False, args.device)
    criterion = nn.MSELoss(reduction='mean').to(args.device)
    print(f"Total Trainable Params: {sum(p.numel() for p in model.parameters() if p.requires_grad)}")

    best_mae, stale, metrics = np.inf, 0, {}
    for epoch in range(args.epoch):
        t0 = time.time()
        model.train()
        loss_total, cnt = 0.0, 0
        for inputs, target in train_loader:
            inputs, target = inputs.to(args.device), target.to(args.device)
            cnt += 1
            if fused is not None:
                fused.step(inputs, target)           # one graph replay; the loss stays on the device
                continue
            model.zero_grad()
            forecast, _ = model(inputs)
            loss = criterion(forecast, target)
            loss.backward()
            optim.step()
            loss_total += float(loss)
        if fused is not None:
            loss_total = fused.pop_loss()            # the epoch's only loss read-back
        print('| end of epoch {:3d} | time: {:5.2f}s | train_total_loss {:5.4f}'.format(
            epoch, time.time() - t0, loss_total / cnt))
        save_model(model, result_file, epoch)
        if (epoch + 1) % args.exponential_decay_step == 0:
            if fused is not None:
                fused.set_lr(fused.lr * args.decay_rate)     # ExponentialLR(gamma=decay_rate).step()
            else:
                scheduler.step()
        if (epoch + 1) % args.validate_freq == 0:
            print('------ validate on data: VALIDATE ------')
            metrics = validate(model, valid_loader, args.device, args.norm_method, normalize_statistic,
                               node_cnt, args.window_size, args.horizon, result_file=result_file)
            if best_mae > metrics['mae']:
                best_mae, stale = metrics['mae'], 0
                save_model(model, result_file)
            else:
                stale += 1
        if args.early_stop and stale >= args.early_stop_step:
            break
    return metrics, normalize_statistic


def test(test_data, args, result_train_file, result_test_file):
    with open(os.path.join(result_train_file, 'norm_stat.json'), 'r') as f:
        normalize_statistic = json.load(f)
    model = load_model(result_train_file)
    node_cnt = test_data.shape[1]
    test_set = ForecastDataset(test_data, window_size=args.window_size, horizon=args.horizon,
                               normalize_method=args.norm_method, norm_statistic=normalize_statistic)
    test_loader = _make_loader(test_set, args.batch_size, False, args.device)
    m = validate(model, test_loader, args.device, args.norm_method, normalize_statistic, node_cnt,
                 args.window_size, args.horizon, result_file=result_test_file)
    print('Performance on test set: MAPE: {:5.2f} | MAE: {:5.2f} | RMSE: {:5.4f}'.format(
        m['mape'], m['mae'], m['rmse']))

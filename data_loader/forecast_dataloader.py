"""Sliding-window forecasting dataset with the semantics of the reference
`data_loader/forecast_dataloader.py` (microsoft/StemGNN): optional z-score / min-max normalisation
with externally supplied statistics, forward/backward fill of missing values, and items
`(x: (window, N) float32, y: (horizon, N) float32)` whose window ends at every `interval`-th row."""
import numpy as np
import pandas as pd
import torch
import torch.utils.data as torch_data


def _safe_std(std):
    return np.asarray([1 if s == 0 else s for s in std])


def normalized(data, normalize_method, norm_statistic=None):
    """Returns (normalised data, statistics).  min-max uses scale = max-min+1e-5 and clips to [0,1];
    z-score replaces zero std by 1 (reference :6-22)."""
    if normalize_method == 'min_max':
        if not norm_statistic:
            norm_statistic = dict(max=np.max(data, axis=0), min=np.min(data, axis=0))
        lo = np.asarray(norm_statistic['min'])      # lists (norm_stat.json) are accepted too
        span = np.asarray(norm_statistic['max']) - lo + 1e-5
        data = np.clip((data - lo) / span, 0.0, 1.0)
    elif normalize_method == 'z_score':
        if not norm_statistic:
            norm_statistic = dict(mean=np.mean(data, axis=0), std=np.std(data, axis=0))
        std = [1 if s == 0 else s for s in norm_statistic['std']]
        data = (data - norm_statistic['mean']) / std
        norm_statistic['std'] = std
    return data, norm_statistic


def de_normalized(data, normalize_method, norm_statistic):
    """Inverse of `normalized` (min-max uses +1e-8 here, as the reference does at :29)."""
    if normalize_method == 'min_max':
        if not norm_statistic:
            norm_statistic = dict(max=np.max(data, axis=0), min=np.min(data, axis=0))
        lo = np.asarray(norm_statistic['min'])
        span = np.asarray(norm_statistic['max']) - lo + 1e-8
        return data * span + lo
    if normalize_method == 'z_score':
        if not norm_statistic:
            norm_statistic = dict(mean=np.mean(data, axis=0), std=np.std(data, axis=0))
        return data * _safe_std(norm_statistic['std']) + norm_statistic['mean']
    return data


class ForecastDataset(torch_data.Dataset):
    def __init__(self, df, window_size, horizon, normalize_method=None, norm_statistic=None, interval=1):
        self.window_size, self.horizon, self.interval = window_size, horizon, interval
        self.normalize_method, self.norm_statistic = normalize_method, norm_statistic
        frame = pd.DataFrame(df)
        self.data = frame.ffill(limit=len(frame)).bfill(limit=len(frame)).values
        self.df_length = len(self.data)
        self.x_end_idx = self.get_x_end_idx()
        if normalize_method:
            self.data, _ = normalized(self.data, normalize_method, norm_statistic)

    def get_x_end_idx(self):
        """Exclusive end row of every input window: window, window+interval, ... while a full
        horizon still fits behind it."""
        ends = range(self.window_size, self.df_length - self.horizon + 1)
        return [ends[j * self.interval] for j in range(len(ends) // self.interval)]

    def __len__(self):
        return len(self.x_end_idx)

    def __getitem__(self, index):
        hi = self.x_end_idx[index]
        x = torch.from_numpy(self.data[hi - self.window_size:hi]).type(torch.float)
        y = torch.from_numpy(self.data[hi:hi + self.horizon]).type(torch.float)
        return x, y

"""Latitude-weighted verification scores, one pass over prediction and target.

Same names, arguments and return shapes as the functions of the reference's `era5_data/score.py` that `test()` calls
(`models/pangu_sample.py:531-569`):

    weighted_rmse_torch_channels(pred, target, mask=None)   era5_data/score.py:126-161
    weighted_acc_torch_channels(pred, target)               era5_data/score.py:181-201

`pred` / `target`: `[n, c, h, w]` or `[c, h, w]` CUDA fp32 tensors; the result has the leading dimensions.  The reference
makes about ten full-tensor passes per call; `pangu_lat_weighted_score_sums` (csrc/bwd_kernels.cu) reads each field once and
returns the five weighted sums per plane, from which BOTH scores follow -- `scores()` gives them together, optionally
subtracting a per-plane climatology for the anomaly correlation (`pangu_sample.py:549-556`).  No CPU path.
"""
import math

import torch

from . import abi, ops
from .abi import PanguError


_LAT_W = {}


def latitude_weights(num_lat, device):
    """era5_data/score.py:99-106: num_lat * cos(3.1416/180 * lat(j)) / sum_j cos(...), lat(j) = 90 - j * 180 / (num_lat - 1),
    fp32 with the reference's own constant 3.1416.  A table of `num_lat` constants: evaluated once on the host with the same
    fp32 expression as the reference (device `cos` differs from the host's in the last bit) and kept on the device."""
    key = (int(num_lat), str(torch.device(device)))
    w = _LAT_W.get(key)
    if w is None:
        j = torch.arange(start=0, end=num_lat)
        lat = 90. - j * 180. / float(num_lat - 1)
        c = torch.cos(3.1416 / 180. * lat)
        w = _LAT_W[key] = (num_lat * c / torch.sum(c)).float().contiguous().to(device)
    return w


def _planes(t, name):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise PanguError(f"pangu_b200.score: '{name}' must be a CUDA tensor (no CPU fallback)")
    if t.dim() not in (3, 4):
        raise PanguError(f"pangu_b200.score: '{name}' must be [n, c, h, w] or [c, h, w], got {tuple(t.shape)}")
    return t.detach().float().contiguous()


def score_sums(pred, target, mask=None, clim=None):
    """(fp64 tensor [*lead, 5], h * w): sum w m (p-t)^2, sum w m, sum w a b, sum w a^2, sum w b^2 (include/pangu_b200.h)."""
    p, t = _planes(pred, "pred"), _planes(target, "target")
    if p.shape != t.shape:
        raise PanguError(f"pangu_b200.score: pred {tuple(p.shape)} and target {tuple(t.shape)} differ")
    lead, (H, W) = p.shape[:-2], p.shape[-2:]
    planes = math.prod(lead)
    if mask is not None:
        mask = mask.to(p.device).float().contiguous()
        if tuple(mask.shape) != (H, W):
            raise PanguError(f"pangu_b200.score: mask must be [{H}, {W}], got {tuple(mask.shape)}")
    if clim is not None:
        clim = clim.to(p.device).float().reshape(-1).contiguous()
        if clim.numel() != planes:
            raise PanguError(f"pangu_b200.score: clim must have {planes} entries, got {clim.numel()}")
    w = latitude_weights(H, p.device)
    sums = torch.empty(planes, 5, dtype=torch.float64, device=p.device)
    ops._chk(p, torch.float32, "pred")                    # also selects p's device / stream for the launch
    ops._call("lat_weighted_score_sums", "pangu_lat_weighted_score_sums",
              (p.data_ptr(), t.data_ptr(), mask.data_ptr() if mask is not None else None,
               clim.data_ptr() if clim is not None else None, w.data_ptr(), planes, H, W, sums.data_ptr(), ops._stream(),),
              nbytes=float(p.numel() * 8))
    return sums.reshape(*lead, 5), H * W


def scores(pred, target, mask=None, clim=None):
    """(rmse, acc) with the leading dimensions of `pred`, fp32; `clim` = per-plane climatology for the anomaly correlation."""
    s, hw = score_sums(pred, target, mask, clim)
    rmse = torch.sqrt(s[..., 0] / (s[..., 1] if mask is not None else hw))
    acc = s[..., 2] / torch.sqrt(s[..., 3] * s[..., 4])
    return rmse.float(), acc.float()


def weighted_rmse_torch_channels(pred, target, mask=None):
    return scores(pred, target, mask)[0]


def weighted_acc_torch_channels(pred, target):
    return scores(pred, target)[1]


def weighted_rmse_torch(pred, target):
    """era5_data/score.py:164-167: mean over the batch dimension."""
    return torch.mean(weighted_rmse_torch_channels(pred, target), dim=0)


def weighted_acc_torch(pred, target):
    return torch.mean(weighted_acc_torch_channels(pred, target), dim=0)

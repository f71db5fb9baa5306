"""Generate tests/golden/reference_score_goldens.npz from the UNMODIFIED reference's era5_data/score.py.

Run in the build container only (needs /root/reference):   python tests/golden/make_score_golden.py

score.py imports only numpy / torch, so it is loaded as-is (by file path: the `era5_data` package __init__ is not needed).
Stored: seeded inputs (small: 3 planes of 33 x 64, and a batch [2, 3, 33, 64]), a 0/1 mask, and what the reference's
weighted_rmse_torch_channels (with and without mask), weighted_acc_torch_channels and latitude weights return for them.
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_score", "/root/reference/era5_data/score.py")
score = importlib.util.module_from_spec(spec)
spec.loader.exec_module(score)

g = torch.Generator().manual_seed(2024)
H, W = 33, 64
pred3 = torch.randn(3, H, W, generator=g) * 3.0 + 1.0
targ3 = pred3 + 0.5 * torch.randn(3, H, W, generator=g)
pred4 = torch.randn(2, 3, H, W, generator=g)
targ4 = 0.7 * pred4 + 0.3 * torch.randn(2, 3, H, W, generator=g)
mask = (torch.rand(H, W, generator=g) > 0.4).float()
lat_t = torch.arange(start=0, end=H)
s = torch.sum(torch.cos(3.1416 / 180. * score.lat(lat_t, H)))
out = dict(
    pred3=pred3.numpy(), targ3=targ3.numpy(), pred4=pred4.numpy(), targ4=targ4.numpy(), mask=mask.numpy(),
    lat_weight=score.latitude_weighting_factor_torch(lat_t, H, s).numpy(),
    rmse3=score.weighted_rmse_torch_channels(pred3, targ3).numpy(),
    rmse3_masked=score.weighted_rmse_torch_channels(pred3, targ3, mask).numpy(),
    rmse4=score.weighted_rmse_torch_channels(pred4, targ4).numpy(),
    rmse4_masked=score.weighted_rmse_torch_channels(pred4, targ4, mask).numpy(),
    acc3=score.weighted_acc_torch_channels(pred3, targ3).numpy(),
    acc4=score.weighted_acc_torch_channels(pred4, targ4).numpy(),
    rmse4_mean=score.weighted_rmse_torch(pred4, targ4).numpy(),
)
np.savez_compressed(os.path.join(HERE, "reference_score_goldens.npz"), **out)
print({k: v.shape for k, v in out.items()})

"""The reference's training loss (models/pangu_sample.py:163-218, default branch) as ONE fused pass per output tensor:
target normalisation (era5_data/utils_data.py normData), per-variable weights, L1, mean, the two loss weights -- and the
gradient w.r.t. the model outputs, written by the same kernel."""
import torch

from . import abi, ops

# era5_data/config.py:52-55
UPPER_WEIGHTS = (3.00, 0.60, 1.50, 0.77, 0.54)
SURFACE_WEIGHTS = (1.50, 0.77, 0.66, 3.00)
UPPER_LOSS_WEIGHT, SURFACE_LOSS_WEIGHT = 1.0, 0.25


def _flat(t, n):
    return None if t is None else t.detach().reshape(-1).float().contiguous()[:n].contiguous()


class _WeightedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, output_surface, target, target_surface, stats, w_upper, w_surface, lw_upper, lw_surface):
        dev = output.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        need = output.requires_grad or output_surface.requires_grad
        grads = []
        surface_mean, surface_std, upper_mean, upper_std = stats if stats is not None else (None,) * 4
        for o, t, m, s, w, lw in ((output, target, upper_mean, upper_std, w_upper, lw_upper),
                                  (output_surface, target_surface, surface_mean, surface_std, w_surface, lw_surface)):
            o = ops._chk(o.detach().contiguous(), torch.float32, "output")
            t = ops._chk(t.detach().contiguous(), torch.float32, "target")
            if o.shape != t.shape:
                raise abi.PanguError("weighted_l1_loss: output and target shapes differ")
            nvar = w.numel()
            plane_elems = o.shape[-2] * o.shape[-1]
            planes = o.numel() // plane_elems
            if planes % nvar:
                raise abi.PanguError("weighted_l1_loss: planes are not a multiple of the variable count")
            d = torch.empty_like(o) if need else None
            m, s = _flat(m, planes), _flat(s, planes)
            ops._call("weighted_l1_loss", "pangu_weighted_l1_loss",
                      (ops._ptr(o), ops._ptr(t), ops._ptr(m), ops._ptr(s), ops._ptr(w), planes, planes // nvar, plane_elems,
                       float(lw) / o.numel(), ops._ptr(loss), ops._ptr(d), ops._stream(),),
                      nbytes=float(o.numel() * (8 + 4 * need)))
            grads.append(d)
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.have = need
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.have:
            return (None,) * 9
        d, ds = ctx.saved_tensors
        return d * g, ds * g, None, None, None, None, None, None, None


def weighted_l1_loss(output, output_surface, target, target_surface, statistics_last=None,
                     upper_weights=UPPER_WEIGHTS, surface_weights=SURFACE_WEIGHTS,
                     upper_loss_weight=UPPER_LOSS_WEIGHT, surface_loss_weight=SURFACE_LOSS_WEIGHT):
    """loss = mean(L1(output, norm(target)) * upper_weights) * upper_loss_weight
            + mean(L1(output_surface, norm(target_surface)) * surface_weights) * surface_loss_weight
    (models/pangu_sample.py:205-218).  statistics_last = (surface_mean [4], surface_std [4], upper_mean [5*13],
    upper_std [5*13]) in the order of era5_data.utils_data.weatherStatistics_output, or None when the targets are
    already normalised.  Batch 1 (like the model)."""
    dev = output.device
    if output.shape[0] != 1:
        raise abi.PanguError("weighted_l1_loss: batch 1 only (per-plane statistics are indexed without a batch axis)")
    wu = torch.as_tensor(upper_weights, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    ws = torch.as_tensor(surface_weights, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    return _WeightedL1.apply(output, output_surface, target, target_surface, statistics_last, wu, ws,
                             float(upper_loss_weight), float(surface_loss_weight))

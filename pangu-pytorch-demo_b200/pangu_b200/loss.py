"""The reference's training losses (models/pangu_sample.py:163-218) as ONE fused pass per output tensor: target
normalisation (era5_data/utils_data.py normData), per-variable weights, the optional custom mask, L1, the mean and the
loss weights -- and the gradient w.r.t. the model outputs, written by the same kernel.

    training_loss(output, output_surface, target, target_surface, statistics_last,
                  only_use_wind_speed_loss=False, custom_mask=None)        # the branch structure of train(), :183-204

      default            mean(|o - t| * w_var) * loss_weight, upper + surface           (:201-204)   weighted_l1_loss
      custom mask        sum(|o - t| * w_var * mask) / mask.sum() instead of the mean   (:196-199)   weighted_l1_loss(mask=)
      wind speed         mean(|ws(o) - ws(t)|), surface (u10, v10) + upper (u, v)       (:184-193)   wind_speed_l1_loss
      wind speed + mask  sum(|ws(o) - ws(t)| * mask) / mask.sum()                       (:187-190)   wind_speed_l1_loss(mask=)

with ws(u, v) = sqrt(u^2 + v^2) (get_wind_speed, :74-93) on the NORMALISED fields, as the reference computes it.
Batch 1 (like the model).  No CPU path."""
import torch

from . import abi, ops

# era5_data/config.py:52-55
UPPER_WEIGHTS = (3.00, 0.60, 1.50, 0.77, 0.54)
SURFACE_WEIGHTS = (1.50, 0.77, 0.66, 3.00)
UPPER_LOSS_WEIGHT, SURFACE_LOSS_WEIGHT = 1.0, 0.25
# channel positions used by get_wind_speed (models/pangu_sample.py:76-88): surface u10 / v10, upper u / v
SURFACE_UV, UPPER_UV = (1, 2), (3, 4)


def _flat(t, n):
    return None if t is None else t.detach().reshape(-1).float().contiguous()[:n].contiguous()


def _mask(mask, shape, dev):
    """custom mask [H, W] (any dtype; models/pangu_sample.py:122-127) -> (fp32 device tensor, valid_points)."""
    if mask is None:
        return None, None
    m = mask.detach().to(dev).float().contiguous()
    if tuple(m.shape) != tuple(shape):
        raise abi.PanguError(f"custom mask must be {tuple(shape)}, got {tuple(m.shape)}")
    return m, float(m.sum())


class _WeightedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, output_surface, target, target_surface, stats, w_upper, w_surface, lw_upper, lw_surface, mask,
                denoms):
        dev = output.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        need = output.requires_grad or output_surface.requires_grad
        grads = []
        surface_mean, surface_std, upper_mean, upper_std = stats if stats is not None else (None,) * 4
        for o, t, m, s, w, lw, den in ((output, target, upper_mean, upper_std, w_upper, lw_upper, denoms[0]),
                                       (output_surface, target_surface, surface_mean, surface_std, w_surface, lw_surface, denoms[1])):
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
            scale = float(lw) / (o.numel() if den is None else den)
            if mask is None:
                ops._call("weighted_l1_loss", "pangu_weighted_l1_loss",
                          (ops._ptr(o), ops._ptr(t), ops._ptr(m), ops._ptr(s), ops._ptr(w), planes, planes // nvar, plane_elems,
                           scale, ops._ptr(loss), ops._ptr(d), ops._stream(),),
                          nbytes=float(o.numel() * (8 + 4 * need)))
            else:
                ops._call("weighted_l1_loss", "pangu_weighted_l1_loss_masked",
                          (ops._ptr(o), ops._ptr(t), ops._ptr(m), ops._ptr(s), ops._ptr(w), ops._ptr(mask), planes, planes // nvar,
                           plane_elems, scale, ops._ptr(loss), ops._ptr(d), ops._stream(),),
                          nbytes=float(o.numel() * (8 + 4 * need)))
            grads.append(d)
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.have = need
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.have:
            return (None,) * 11
        d, ds = ctx.saved_tensors
        return (d * g, ds * g) + (None,) * 9


def weighted_l1_loss(output, output_surface, target, target_surface, statistics_last=None,
                     upper_weights=UPPER_WEIGHTS, surface_weights=SURFACE_WEIGHTS,
                     upper_loss_weight=UPPER_LOSS_WEIGHT, surface_loss_weight=SURFACE_LOSS_WEIGHT, mask=None,
                     surface_mask_denominator="valid_points"):
    """loss = mean(L1(output, norm(target)) * upper_weights) * upper_loss_weight
            + mean(L1(output_surface, norm(target_surface)) * surface_weights) * surface_loss_weight
    (models/pangu_sample.py:201-218).  statistics_last = (surface_mean [4], surface_std [4], upper_mean [5*13],
    upper_std [5*13]) in the order of era5_data.utils_data.weatherStatistics_output, or None when the targets are
    already normalised.  Batch 1 (like the model).

    mask: the reference's custom mask [H, W] (`use_custom_mask`, :120-127): the means become
    sum(L1 * weights * mask) / mask.sum() (train(), :196-199).  `surface_mask_denominator="valid_points*channels"`
    gives the surface term of the validation / test loops instead (:467, :337)."""
    dev = output.device
    if output.shape[0] != 1:
        raise abi.PanguError("weighted_l1_loss: batch 1 only (per-plane statistics are indexed without a batch axis)")
    wu = torch.as_tensor(upper_weights, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    ws = torch.as_tensor(surface_weights, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    m, valid = _mask(mask, output.shape[-2:], dev)
    denoms = (None, None)
    if m is not None:
        if surface_mask_denominator not in ("valid_points", "valid_points*channels"):
            raise abi.PanguError("surface_mask_denominator: 'valid_points' or 'valid_points*channels'")
        denoms = (valid, valid * (output_surface.shape[1] if surface_mask_denominator == "valid_points*channels" else 1))
    return _WeightedL1.apply(output, output_surface, target, target_surface, statistics_last, wu, ws,
                             float(upper_loss_weight), float(surface_loss_weight), m, denoms)


class _WindSpeedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, output_surface, target, target_surface, stats, mask, valid):
        dev = output.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        need = output.requires_grad or output_surface.requires_grad
        surface_mean, surface_std, upper_mean, upper_std = stats if stats is not None else (None,) * 4
        grads = []
        for o, t, m, s, (cu, cv) in ((output, target, upper_mean, upper_std, UPPER_UV),
                                     (output_surface, target_surface, surface_mean, surface_std, SURFACE_UV)):
            o = ops._chk(o.detach().contiguous(), torch.float32, "output")
            t = ops._chk(t.detach().contiguous(), torch.float32, "target")
            if o.shape != t.shape:
                raise abi.PanguError("wind_speed_l1_loss: output and target shapes differ")
            H, W = o.shape[-2:]
            nvar = o.shape[1]
            levels = o.numel() // (nvar * H * W)                        # 13 upper-air levels, 1 for the surface fields
            o3, t3 = o.reshape(nvar, levels, H * W), t.reshape(nvar, levels, H * W)
            d = torch.zeros_like(o) if need else None
            d3 = d.reshape(nvar, levels, H * W) if need else None
            st = [None] * 4
            if m is not None:
                mm, ss = _flat(m, nvar * levels).reshape(nvar, levels), _flat(s, nvar * levels).reshape(nvar, levels)
                st = [mm[cu].contiguous(), ss[cu].contiguous(), mm[cv].contiguous(), ss[cv].contiguous()]
            scale = 1.0 / (levels * H * W if valid is None else valid)
            ops._call("wind_speed_l1_loss", "pangu_wind_speed_l1_loss",
                      (ops._ptr(o3[cu]), ops._ptr(o3[cv]), ops._ptr(t3[cu]), ops._ptr(t3[cv]), ops._ptr(st[0]), ops._ptr(st[1]),
                       ops._ptr(st[2]), ops._ptr(st[3]), ops._ptr(mask), levels, H * W, scale, ops._ptr(loss),
                       ops._ptr(d3[cu]) if need else None, ops._ptr(d3[cv]) if need else None, ops._stream(),),
                      nbytes=float(levels * H * W * 4 * (4 + 2 * need)))
            grads.append(d)
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.have = need
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.have:
            return (None,) * 7
        d, ds = ctx.saved_tensors
        return (d * g, ds * g) + (None,) * 5


def wind_speed_l1_loss(output, output_surface, target, target_surface, statistics_last=None, mask=None):
    """mean(|ws(output_surface) - ws(norm(target_surface))|) + mean(|ws(output) - ws(norm(target))|) with
    ws = sqrt(u^2 + v^2) over (u10, v10) / (u, v) (models/pangu_sample.py:74-93, :184-193); with the custom mask the means
    become sum(. * mask) / mask.sum() (:187-190).  The other variables get zero gradient, as in the reference."""
    if output.shape[0] != 1:
        raise abi.PanguError("wind_speed_l1_loss: batch 1 only")
    m, valid = _mask(mask, output.shape[-2:], output.device)
    return _WindSpeedL1.apply(output, output_surface, target, target_surface, statistics_last, m, valid)


def training_loss(output, output_surface, target, target_surface, statistics_last=None, only_use_wind_speed_loss=False,
                  custom_mask=None, upper_weights=UPPER_WEIGHTS, surface_weights=SURFACE_WEIGHTS,
                  upper_loss_weight=UPPER_LOSS_WEIGHT, surface_loss_weight=SURFACE_LOSS_WEIGHT):
    """The loss of the reference's train() for one sample, all four branches of models/pangu_sample.py:183-204."""
    if only_use_wind_speed_loss:
        return wind_speed_l1_loss(output, output_surface, target, target_surface, statistics_last, custom_mask)
    return weighted_l1_loss(output, output_surface, target, target_surface, statistics_last, upper_weights, surface_weights,
                            upper_loss_weight, surface_loss_weight, custom_mask)

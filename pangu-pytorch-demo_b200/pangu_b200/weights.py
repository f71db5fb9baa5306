"""Weight import and the compact Earth-specific bias.

* `load_named_weights(model, weights, name_map)` -- the loop of the reference's `models/onnx2torch.py:124-161`: every model
  parameter is filled from the array its look-up row names (`keys_all.csv`: torch_name -> onnx_name); 2-D arrays are stored
  transposed in the ONNX graph (`:141-144`), every other rank is copied as it is; parameters are frozen like the reference does
  (`requires_grad = False`).  Extension: a 5-D `earth_specific_bias` may be given as the paper's COMPACT table
  `[3312, T, heads]`; it is expanded on the device by `expand_bias_table`.
* `expand_bias_table(table)` / `reduce_bias_grad(d_full)` -- the gather the reference keeps as commented code
  (`models/layers.py:355,442-449`) and its adjoint, as CUDA index kernels (`pangu_bias_table_expand` / `_reduce`): a fine-tune run that
  wants the compact parameterisation (40 MB instead of 1 GB of bias parameters / gradients to all-reduce) keeps the table as
  its parameter, expands it before the step and reduces the dense gradient after it.
"""
import numpy as np
import torch

from . import abi, ops
from .abi import PanguError

TABLE_ROWS = 3312            # (2*2) * (6*6) * (2*12-1), models/layers.py:355


def expand_bias_table(table):
    """table fp32 CUDA [3312, T, heads] -> [1, T, heads, 144, 144] (the shape of `EarthAttention3D.earth_specific_bias`)."""
    ops._chk(table, torch.float32, "table")
    if table.dim() != 3 or table.shape[0] != TABLE_ROWS:
        raise PanguError(f"expand_bias_table: table must be [{TABLE_ROWS}, T, heads], got {tuple(table.shape)}")
    T, heads = int(table.shape[1]), int(table.shape[2])
    full = torch.empty((1, T, heads, 144, 144), dtype=torch.float32, device=table.device)
    ops._call("bias_table_expand", "pangu_bias_table_expand", (ops._ptr(table), ops._ptr(full), T, heads, ops._stream(),),
              nbytes=float(full.numel() * 8))
    return full


def reduce_bias_grad(d_full, d_table=None):
    """Dense bias gradient [1, T, heads, 144, 144] (or without the 1) -> gradient of the compact table [3312, T, heads]
    (accumulated into `d_table` when given)."""
    ops._chk(d_full, torch.float32, "d_full")
    if d_full.dim() == 5:
        d_full = d_full[0]
    if d_full.dim() != 4 or tuple(d_full.shape[2:]) != (144, 144):
        raise PanguError(f"reduce_bias_grad: expected [T, heads, 144, 144], got {tuple(d_full.shape)}")
    T, heads = int(d_full.shape[0]), int(d_full.shape[1])
    if d_table is None:
        d_table = torch.zeros((TABLE_ROWS, T, heads), dtype=torch.float32, device=d_full.device)
    else:
        ops._chk(d_table, torch.float32, "d_table")
        if tuple(d_table.shape) != (TABLE_ROWS, T, heads):
            raise PanguError(f"reduce_bias_grad: d_table must be [{TABLE_ROWS}, {T}, {heads}]")
    ops._call("bias_table_reduce", "pangu_bias_table_reduce", (ops._ptr(d_full), ops._ptr(d_table), T, heads, ops._stream(),),
              nbytes=float(d_full.numel() * 4 + d_table.numel() * 8))
    return d_table


def load_named_weights(model, weights, name_map, freeze=True):
    """Fill `model`'s parameters from `weights` (dict: source name -> numpy array / tensor) through `name_map`
    (dict or iterable of (torch_name, source_name) pairs, the two columns of the reference's keys_all.csv).
    Returns the number of parameters loaded; parameters without a row (or whose row has no source name) are left alone,
    as in the reference.  Shape rules of models/onnx2torch.py:136-160."""
    name_map = dict(name_map)
    count = 0
    with torch.no_grad():
        for name, param in model.named_parameters():
            src = name_map.get(name)
            if not isinstance(src, str):
                continue
            if src not in weights:
                raise PanguError(f"load_named_weights: '{src}' (for parameter '{name}') is not in the weight dictionary")
            w = torch.as_tensor(np.asarray(weights[src])) if not torch.is_tensor(weights[src]) else weights[src]
            if param.dim() == 2:
                w = w.t()                                             # stored transposed (onnx2torch.py:141-144)
            elif param.dim() == 5 and w.dim() == 3 and w.shape[0] == TABLE_ROWS:
                if not torch.cuda.is_available():
                    raise PanguError("load_named_weights: a compact bias table is expanded on the GPU (no CPU fallback)")
                dev = param.device if param.is_cuda else torch.device("cuda", torch.cuda.current_device())
                w = expand_bias_table(w.to(dev, torch.float32).contiguous())
            if tuple(w.shape) != tuple(param.shape):
                raise PanguError(f"load_named_weights: '{name}' is {tuple(param.shape)}, '{src}' gives {tuple(w.shape)}")
            param.copy_(w.to(device=param.device, dtype=param.dtype))
            if freeze:
                param.requires_grad = False
            count += 1
    return count

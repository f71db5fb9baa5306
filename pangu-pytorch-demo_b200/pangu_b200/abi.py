"""ctypes binding of libpangu_b200.so -- one prototype per entry point of include/pangu_b200.h.

The library is built in-tree by `pangu_b200/build.py` (nvcc, sm_100a).  If it is missing the import of
any op raises: the product has no fallback path.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# $PANGU_B200_LIB: another build of the same C ABI (A/B measurements of kernel versions on one box, tools/ab_build.sh)
LIB_PATH = os.environ.get("PANGU_B200_LIB") or os.path.join(HERE, "libpangu_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU = 0, 1
ROLL_NONE, ROLL_SHIFT, WINDOWED = 0, 1, 2


class PanguError(RuntimeError):
    pass


class Geom(Structure):
    """pangu_geom: token grid of one stage."""
    _fields_ = [("Z", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32), ("heads", c_int32)]


class Band(Structure):
    """pangu_band: latitude band of one stage's window grid (include/pangu_b200.h)."""
    _fields_ = [("h0", c_int32), ("hrows", c_int32), ("hw0", c_int32), ("nhw", c_int32), ("wrap", c_int32),
                ("halo", c_int32), ("halo_lo", c_int32)]


_PROTOS = {
    "pangu_last_error": (c_char_p, []),
    "pangu_abi_version": (c_int, []),
    "pangu_has_tcgen05": (c_int, []),
    "pangu_window_partition": (c_int, [c_void_p, c_void_p, POINTER(Geom), c_int, c_int, c_void_p]),
    "pangu_window_reverse": (c_int, [c_void_p, c_void_p, POINTER(Geom), c_int, c_int, c_void_p]),
    "pangu_window_source_index": (c_int, [c_void_p, POINTER(Geom), c_int, c_void_p]),
    "pangu_shift_mask": (c_int, [c_void_p, POINTER(Geom), c_void_p]),
    "pangu_set_pdl": (c_int, [c_int]),
    "pangu_position_index": (c_int, [c_void_p, c_void_p]),
    "pangu_bias_table_expand": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "pangu_bias_table_reduce": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "pangu_linear": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                             c_int, c_int, c_int, c_void_p]),
    "pangu_linear_bf16_ex": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_int64, c_int32, c_int32, c_int, c_int, c_void_p]),
    "pangu_ln_residual": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int32, c_float, c_void_p]),
    "pangu_linear_ln_residual_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p]),
    "pangu_mlp_ln_residual_bf16": (c_int, [c_void_p] * 10 + [c_int64, c_int32, c_float, c_void_p]),
    "pangu_attn_proj_mlp_bf16": (c_int, [c_void_p] * 13 + [c_int64, c_void_p, c_void_p, c_int64, c_int32, c_float, c_float, c_void_p]),
    "pangu_debug_mlp_trace": (c_int, [c_void_p, c_int32]),
    "pangu_debug_attn_trace": (c_int, [c_void_p, c_int32]),
    "pangu_window_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(Geom), c_int, c_int,
                                       c_void_p]),
    "pangu_window_attention_band": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                            c_void_p, POINTER(Geom), POINTER(Band), c_int, c_int, c_void_p]),
    "pangu_patch_embed_gather": (c_int, [c_void_p] * 10 + [c_int, c_void_p]),
    "pangu_patch_embed_gather_rows": (c_int, [c_void_p] * 10 + [c_int, c_int32, c_int32, c_int32, c_void_p]),
    "pangu_patch_recover_scatter": (c_int, [c_void_p] * 5),
    "pangu_patch_recover_scatter_rows": (c_int, [c_void_p] * 4 + [c_int32, c_int32, c_void_p]),
    "pangu_patch_recover_scatter_denorm": (c_int, [c_void_p] * 4 + [c_int32, c_int32] + [c_void_p] * 5),
    "pangu_downsample_merge_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int32, c_int32, c_int32,
                                          c_int32, c_float, c_void_p]),
    "pangu_upsample_shuffle_ln": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int32, c_int32,
                                          c_int32, c_int32, c_int32, c_float, c_void_p]),
    "pangu_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pangu_concat_cast_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    # fine-tune backward
    "pangu_linear_bf16_add": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                      c_int32, c_void_p]),
    "pangu_linear_wgrad_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                        c_void_p]),
    "pangu_colsum": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int32, c_void_p, c_void_p]),
    "pangu_ln_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_int32, c_float, c_void_p]),
    "pangu_upsample_shuffle_ln_backward": (c_int, [c_void_p] * 6 + [c_int32] * 5 + [c_float, c_void_p]),
    "pangu_downsample_merge_ln_backward": (c_int, [c_void_p] * 6 + [c_int32] * 4 + [c_float, c_void_p]),
    "pangu_gelu_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pangu_gelu_backward_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "pangu_window_attention_train": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(Geom), c_int,
                                             c_void_p]),
    "pangu_window_attention_backward": (c_int, [c_void_p] * 9 + [POINTER(Geom), c_int, c_void_p]),
    "pangu_patch_recover_gather_backward": (c_int, [c_void_p] * 4 + [c_int32, c_int32, c_void_p]),
    "pangu_weighted_l1_loss": (c_int, [c_void_p] * 5 + [c_int32, c_int32, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "pangu_weighted_l1_loss_masked": (c_int, [c_void_p] * 6 + [c_int32, c_int32, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "pangu_wind_speed_l1_loss": (c_int, [c_void_p] * 9 + [c_int32, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pangu_linear_bf16_aux": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                      c_int, c_int, c_void_p, c_void_p]),
    "pangu_lat_weighted_score_sums": (c_int, [c_void_p] * 5 + [c_int32, c_int32, c_int32, c_void_p, c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_PROTOS)


def lib():
    """Load the shared library (once).  Raises PanguError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PanguError(
            f"{LIB_PATH} is missing: build it with `python {os.path.join(HERE, 'build.py')}` "
            "(nvcc, sm_100a). There is no fallback path.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        if not hasattr(handle, name):
            raise PanguError(f"{LIB_PATH} does not export {name}; rebuild it")
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().pangu_last_error()
        raise PanguError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")

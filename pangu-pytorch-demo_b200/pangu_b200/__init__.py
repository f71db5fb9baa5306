"""pangu_b200 -- host-side plumbing for the B200 (sm_100a) Pangu-Weather kernels.

    abi         ctypes binding of libpangu_b200.so (include/pangu_b200.h); fails loudly if the
                library is missing -- there is no CPU or eager-PyTorch fallback
    ops         torch.Tensor -> raw pointer wrappers around the C ABI
    functional  op sequences of the reference modules (models/layers.py) in fp32 / bf16
    dist        latitude-band sharding + halo exchange, data-parallel helpers
"""
from . import abi  # noqa: F401

__all__ = ["abi"]

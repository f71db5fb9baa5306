"""K7 index kernels on the GPU, through the C ABI, bit-exact against the oracle's closed forms (which
tests/test_oracle_golden.py pins to the reference) and against the committed goldens."""
import hashlib

import numpy as np
import pytest
import torch

import pangu_oracle as orc

pytestmark = pytest.mark.gpu

GEOMS = {"A24": (8, 181, 24), "B24": (8, 91, 24), "A": (8, 181, 360), "B": (8, 91, 180)}


@pytest.fixture(scope="module")
def ops():
    from pangu_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("tag", list(GEOMS))
@pytest.mark.parametrize("roll", [0, 1])
def test_source_index(ops, goldens, tag, roll):
    Z, H, W = GEOMS[tag]
    got = ops.window_source_index(Z, H, W, roll, "cuda").cpu().numpy()
    assert np.array_equal(got, orc.window_source_index(Z, H, W, bool(roll)))
    key = f"index.{tag}.roll{roll}.sha256"
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == goldens[key].tobytes()


@pytest.mark.parametrize("tag,C", [("A", 192), ("B", 384), ("A24", 192)])
@pytest.mark.parametrize("roll", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_partition_reverse_bit_exact(ops, tag, C, roll, dtype):
    Z, H, W = GEOMS[tag]
    N = Z * H * W
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, C, generator=g).to(dtype)
    src = torch.from_numpy(orc.window_source_index(Z, H, W, bool(roll)))
    want = torch.cat((x, torch.zeros(1, C, dtype=dtype)), 0)[src]          # -1 -> zero row
    xd = x.cuda()
    win = ops.window_partition(xd, Z, H, W, roll)
    assert torch.equal(win.cpu(), want)
    # pad rows are exactly zero and the reverse is the exact inverse on real tokens
    back = ops.window_reverse(win, Z, H, W, roll)
    assert torch.equal(back.cpu(), x)


@pytest.mark.parametrize("tag", list(GEOMS))
def test_shift_mask(ops, goldens, tag):
    Z, H, W = GEOMS[tag]
    m = ops.shift_mask(Z, H, W, "cuda").cpu().numpy()
    assert np.array_equal(m, orc.shift_mask(Z, H, W))
    assert np.array_equal(np.packbits((m != 0).reshape(-1)), goldens[f"mask.{tag}.bits"])


def test_position_index(ops, goldens):
    idx = ops.position_index("cuda").cpu().numpy()
    assert np.array_equal(idx, orc.position_index())
    assert np.array_equal(idx, goldens["position_index"].astype(np.int64))


def test_windowed_identity_mode(ops):
    idx = ops.window_source_index(2, 7, 24, 2, "cuda").cpu().numpy()      # nLon=2, T=2
    assert np.array_equal(idx.reshape(-1), np.arange(idx.size))


# geometries outside the model's two stages (ragged / minimal grids): the index kernels are closed forms of (Z, H, W), pinned on
# the CPU against the reference's tensor-op sequence by tests/test_oracle_properties.py
SMALL = [(2, 1, 12), (4, 7, 24), (8, 13, 36), (6, 19, 12), (8, 25, 48)]


@pytest.mark.parametrize("Z,H,W", SMALL)
@pytest.mark.parametrize("roll", [0, 1])
def test_index_kernels_on_small_and_minimal_grids(ops, Z, H, W, roll):
    assert np.array_equal(ops.window_source_index(Z, H, W, roll, "cuda").cpu().numpy(), orc.window_source_index(Z, H, W, bool(roll)))
    x = torch.randn(Z * H * W, 8, generator=torch.Generator().manual_seed(H))
    src = torch.from_numpy(orc.window_source_index(Z, H, W, bool(roll)))
    win = ops.window_partition(x.cuda(), Z, H, W, roll)
    assert torch.equal(win.cpu(), torch.cat((x, torch.zeros(1, 8)), 0)[src])
    assert torch.equal(ops.window_reverse(win, Z, H, W, roll).cpu(), x)
    if H >= 7:
        assert np.array_equal(ops.shift_mask(Z, H, W, "cuda").cpu().numpy(), orc.shift_mask(Z, H, W))

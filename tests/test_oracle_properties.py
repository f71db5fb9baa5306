"""Size-independent checks of the oracle's integer closed forms against the reference's own tensor-op sequence, replayed op by
op on small geometries the committed goldens do not cover (the goldens pin the two model geometries; these pin the FORMULA):
F.pad + torch.roll + view / permute partition (models/layers.py:224-262), its inverse (:269-293), gen_mask (:187-216) and
_construct_index (:371-411).  CPU only, bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pangu_oracle as orc

GEOMS = [(2, 1, 12), (4, 7, 24), (8, 13, 36), (6, 19, 12), (8, 25, 48)]         # (Z, H, W): Z even, W % 12 == 0, (H + 5) % 6 == 0
                                                                                 # (the reference pads a fixed 5 rows, models/layers.py:228)


def _partition(x, Z, H, W, roll):
    """models/layers.py:224-262 on a [1, Z*H*W, C] tensor -> [nLon, T, 144, C] (op-by-op restatement used as brute force)."""
    C = x.shape[-1]
    x = x.view(1, Z, H, W, C)
    pad = 5
    x = F.pad(x, (0, 0, 0, 0, 0, pad))                                          # :228 pads H by 5 at the end
    Hp = H + pad
    if roll:
        x = torch.roll(x, shifts=(-1, -3, -6), dims=(1, 2, 3))                  # :237-238
    x = x.view(1, Z // 2, 2, Hp // 6, 6, W // 12, 12, C).permute(0, 5, 1, 3, 2, 4, 6, 7)
    return x.reshape(W // 12, (Z // 2) * (Hp // 6), 144, C), Hp


@pytest.mark.parametrize("Z,H,W", GEOMS)
@pytest.mark.parametrize("roll", [False, True])
def test_source_index_is_pad_roll_partition(Z, H, W, roll):
    g = orc.window_geometry(Z, H, W)
    ids = torch.arange(1, Z * H * W + 1, dtype=torch.float64).view(1, -1, 1)    # token n carries the value n + 1; pad rows carry 0
    x = ids.view(1, Z, H, W, 1)
    x = F.pad(x, (0, 0, 0, 0, 0, g["Hp"] - H))
    if roll:
        x = torch.roll(x, shifts=(-1, -3, -6), dims=(1, 2, 3))
    win = x.view(1, Z // 2, 2, g["Hp"] // 6, 6, W // 12, 12, 1).permute(0, 5, 1, 3, 2, 4, 6, 7).reshape(g["nLon"], g["T"], 144)
    want = win.long().numpy() - 1                                               # -1 = zero pad row
    assert np.array_equal(orc.window_source_index(Z, H, W, roll), want)


@pytest.mark.parametrize("Z,H,W", [g for g in GEOMS if g[1] >= 7])       # Hp = 6 leaves gen_mask's first two h slices empty
def test_shift_mask_is_gen_mask(Z, H, W):
    """gen_mask, models/layers.py:187-216: region image over the padded grid with the reference's slices -- including the
    `+6` start of the second h slice (:197) --, partitioned like the data, -100 where the ids of a pair differ."""
    g = orc.window_geometry(Z, H, W)
    Hp = g["Hp"]
    img = torch.zeros(1, Z, Hp, W, 1)
    z_slices = (slice(0, -2), slice(-2, -1), slice(-1, None))
    h_slices = (slice(0, -6), slice(6, -3), slice(-3, None))
    cnt = 0
    for zs in z_slices:
        for hs in h_slices:
            img[:, zs, hs, :, :] = cnt
            cnt += 1
    win = img.view(1, Z // 2, 2, Hp // 6, 6, W // 12, 12, 1).permute(0, 5, 1, 3, 2, 4, 6, 7).reshape(g["nLon"], g["T"], 144)
    mask = (win[:, :, None, :] - win[:, :, :, None])
    mask = torch.where(mask != 0, torch.tensor(-100.0), torch.tensor(0.0))
    assert torch.equal(mask[0], mask[-1])                                       # identical for every longitude window
    got = orc.shift_mask(Z, H, W)
    assert np.array_equal(got, mask[0].numpy())
    # the compact group ids of the CUDA kernels induce the same partition
    gid = orc.shift_group_ids_closed_form(Z, H, W)
    assert np.array_equal(np.where(gid[:, None, :] != gid[:, :, None], np.float32(-100.0), np.float32(0.0)), got)


@pytest.mark.parametrize("Z,H,W", GEOMS)
def test_reverse_undoes_partition_and_crop_drops_the_pad(Z, H, W):
    """models/layers.py:269-293: reverse view/permute, roll back (+1, +3, +6), crop rows >= H  ==  scatter by source index."""
    C = 3
    x = torch.randn(1, Z * H * W, C, generator=torch.Generator().manual_seed(Z * H + W))
    for roll in (False, True):
        win, Hp = _partition(x.clone(), Z, H, W, roll)
        y = win.reshape(1, W // 12, Z // 2, Hp // 6, 2, 6, 12, C).permute(0, 2, 4, 3, 5, 1, 6, 7).reshape(1, Z, Hp, W, C)
        if roll:
            y = torch.roll(y, shifts=(1, 3, 6), dims=(1, 2, 3))
        y = y[:, :, :H].reshape(1, Z * H * W, C)
        assert torch.equal(y, x)
        src = torch.from_numpy(orc.window_source_index(Z, H, W, roll)).reshape(-1)
        real = src >= 0
        out = torch.zeros(Z * H * W, C)
        out[src[real]] = win.reshape(-1, C)[real]
        assert torch.equal(out, x[0])
        assert int(real.sum()) == Z * H * W and torch.equal(win.reshape(-1, C)[~real], torch.zeros(int((~real).sum()), C))


def test_position_index_is_construct_index():
    """EarthAttention3D._construct_index, models/layers.py:371-411, op by op (meshgrid / flatten / broadcast subtraction)."""
    wz, wh, ww = 2, 6, 12
    coords_zi, coords_zj = torch.arange(wz), -torch.arange(wz) * wz
    coords_hi, coords_hj = torch.arange(wh), -torch.arange(wh) * wh
    coords_w = torch.arange(ww)
    c1 = torch.stack(torch.meshgrid([coords_zi, coords_hi, coords_w], indexing="ij")).flatten(1)
    c2 = torch.stack(torch.meshgrid([coords_zj, coords_hj, coords_w], indexing="ij")).flatten(1)
    co = (c1[:, :, None] - c2[:, None, :]).permute(1, 2, 0).contiguous()
    co[:, :, 2] += ww - 1
    co[:, :, 1] *= 2 * ww - 1
    co[:, :, 0] *= (2 * ww - 1) * wh * wh
    idx = co.sum(-1).flatten()
    assert np.array_equal(orc.position_index(), idx.numpy().astype(np.int64))

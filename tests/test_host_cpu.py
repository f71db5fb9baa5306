"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/pangu_b200.h declares, the drop-in modules honour the reference's parameter / module contract,
and the product fails loudly (no fallback) when asked to run without a GPU."""
import copy
import ctypes
import io
import os
import re

import pytest
import torch

import pangu_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pangu_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pangu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pangu_b200 import abi
    assert os.path.exists(abi.LIB_PATH), "libpangu_b200.so not built: run __graft_entry__.build()"
    handle = ctypes.CDLL(abi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/pangu_b200.h but not exported"
    assert set(abi.exported_symbols()) == set(declared), "ctypes prototypes out of sync with the header"
    lib = abi.lib()
    assert lib.pangu_abi_version() == 1 and lib.pangu_has_tcgen05() == 1


def test_bad_arguments_are_reported_without_a_gpu():
    from pangu_b200 import abi
    lib = abi.lib()
    g = abi.Geom(8, 180, 360, 192, 6)            # (180+5) % 6 != 0
    assert lib.pangu_window_source_index(None, g, 0, None) == -1
    assert b"bad argument" in lib.pangu_last_error()
    with pytest.raises(abi.PanguError):
        abi.check(-1, "probe")


def test_model_contract_matches_reference():
    from models.pangu_model import PanguModel
    import models.layers as L
    for name in ("PatchEmbedding_pretrain", "PatchEmbedding", "EarthSpecificLayer", "EarthSpecificBlock", "Mlp",
                 "EarthAttention3D", "DownSample", "UpSample", "PatchRecovery_pretrain", "PatchRecovery"):
        assert hasattr(L, name)
    m = PanguModel(device="cpu")
    shapes = orc.param_shapes()
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())            # 223 keys of keys_all.csv, same order
    assert all(tuple(v.shape) == shapes[k] for k, v in sd.items())
    assert sum(p.numel() for p in m.parameters()) == 276659936
    assert not list(m.buffers())
    # Linear sub-modules stay nn.Linear (LoRA targets them by isinstance, finetune/lora_tune.py:170-172)
    blk = m.layers[0].blocks[0]
    for lin in (blk.attention.linear1, blk.attention.linear2, blk.linear.linear1, blk.linear.linear2,
                m.downsample.linear, m.upsample.linear1, m.upsample.linear2):
        assert type(lin) is torch.nn.Linear
    assert m.layers[0].use_checkpoint is True                 # evaluated while training=True
    assert blk.attention.position_index.tolist() == orc.position_index().tolist()
    # init follows models/pangu_model.py:52-59
    assert float(blk.norm1.weight.min()) == 1.0 and float(blk.attention.linear1.bias.abs().max()) == 0.0
    assert 0.015 < float(blk.attention.earth_specific_bias.std()) < 0.025
    # state_dict round trip with the DDP "module." prefix stripped (finetune/finetune_fully.py:212)
    m.load_state_dict({k: v for k, v in orc.synth_params(0).items()}, strict=True)
    # deepcopy and whole-module pickling (models/pangu_sample.py:370-372)
    m2 = copy.deepcopy(m.layers[1].blocks[0])
    assert torch.equal(m2.attention.earth_specific_bias, m.layers[1].blocks[0].attention.earth_specific_bias)
    buf = io.BytesIO()
    torch.save(m.downsample, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert torch.equal(m3.linear.weight, m.downsample.linear.weight)


def test_no_cpu_fallback():
    from models.pangu_model import PanguModel
    from pangu_b200.abi import PanguError
    m = PanguModel(device="cpu").eval()
    inp, inp_s = torch.zeros(1, 5, 13, 721, 1440), torch.zeros(1, 4, 721, 1440)
    with pytest.raises(PanguError, match="no CPU fallback"):
        m(inp, inp_s, None, None, None)


def test_ort_like_session_surface_and_no_cpu_path():
    """pangu_b200.session.InferenceSession mirrors the onnxruntime call of inference/inference_singleOutput.py:146-147;
    without a CUDA model it must refuse to construct (no CPU forecast path)."""
    from models.pangu_model import PanguModel
    from pangu_b200.abi import PanguError
    from pangu_b200.session import InferenceSession
    m = PanguModel(device="cpu").eval()
    z = torch.zeros(1)
    with pytest.raises(PanguError, match="CUDA"):
        InferenceSession(m, (z, z, z, z), (z, z, z, z), z, z)
    s = InferenceSession.__new__(InferenceSession)               # the metadata surface needs no device
    assert [(a.name, a.shape) for a in s.get_inputs()] == [("input", [5, 13, 721, 1440]), ("input_surface", [4, 721, 1440])]
    assert [a.name for a in s.get_outputs()] == ["output", "output_surface"]
    with pytest.raises(PanguError, match="float32"):
        InferenceSession._as_f32(__import__("numpy").zeros((4, 721, 1440)), (4, 721, 1440), "input_surface")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pangu-pytorch-demo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "pangu_oracle" not in src and "oracle/" not in src, f"{f} references the oracle"


def test_wrapped_or_hooked_linears_are_folded_or_refused_never_ignored():
    """ADVICE r1 / finetune/lora_tune.py:169-186: the kernels read weights off the sub-modules, so a peft LoRA wrapper is
    folded into the operand (W + scaling * B @ A, differentiable w.r.t. A and B), and everything the fold cannot express
    (active LoRA dropout, unknown wrappers, sub-classes, forward hooks) raises PanguError."""
    from fake_peft import LoraLinear, ModulesToSave
    from pangu_b200 import functional as PF
    from pangu_b200.abi import PanguError
    torch.manual_seed(0)
    base = torch.nn.Linear(24, 40)
    w0, b0 = PF.lin_wb(base)
    assert w0 is base.weight and b0 is base.bias
    lora = LoraLinear(torch.nn.Linear(24, 40), r=4, lora_alpha=8)
    w, b = PF.lin_wb(lora)
    x = torch.randn(5, 24)
    assert torch.allclose(torch.nn.functional.linear(x, w, b), lora(x), atol=1e-6)         # the fold IS the adapter's forward
    g = torch.autograd.grad((torch.nn.functional.linear(x, w, b) ** 2).sum(), [lora.lora_A["default"].weight,
                                                                              lora.lora_B["default"].weight])
    gr = torch.autograd.grad((lora(x) ** 2).sum(), [lora.lora_A["default"].weight, lora.lora_B["default"].weight])
    assert all(torch.allclose(a, c, atol=1e-5) for a, c in zip(g, gr))
    assert not isinstance(w, torch.nn.Parameter) and PF.WeightCache._transient(w)          # folded operands are never cached
    lora.merged = True
    assert PF.lin_wb(lora)[0] is lora.base_layer.weight
    lora.merged = False
    drop = LoraLinear(torch.nn.Linear(24, 40), r=4, lora_dropout=0.1).train()
    with pytest.raises(PanguError, match="lora_dropout"):
        PF.lin_wb(drop)
    PF.lin_wb(drop.eval())                                                                # dropout is the identity in eval()
    conv = ModulesToSave(torch.nn.Conv1d(8, 6, 1))
    assert PF.lin_wb(conv)[0] is conv.modules_to_save["default"].weight

    class MyLinear(torch.nn.Linear):
        def forward(self, x):
            return super().forward(x) * 2

    with pytest.raises(PanguError, match="expected nn.Linear"):
        PF.lin_wb(MyLinear(3, 3))
    hooked = torch.nn.Linear(3, 3)
    hooked.register_forward_hook(lambda m, i, o: o + 1)
    with pytest.raises(PanguError, match="hooks"):
        PF.lin_wb(hooked)
    with pytest.raises(PanguError, match="LayerNorm"):
        PF.norm_wb(torch.nn.GroupNorm(1, 4))


def test_wants_graph_follows_autograd_semantics():
    """eval() with grad mode on and trainable parameters builds the graph (nn.Module semantics); no_grad,
    frozen parameters or set_forward_only() keep the fused forward-only path."""
    import models.layers as L
    from pangu_b200 import autograd as AG
    m = L.Mlp(192, 0).eval()
    x = torch.zeros(4, 192)
    assert AG.wants_graph(m, x)
    with torch.no_grad():
        assert not AG.wants_graph(m, x)
    m.set_forward_only()
    assert not AG.wants_graph(m, x)
    m.set_forward_only(False)
    for p in m.parameters():
        p.requires_grad_(False)
    assert not AG.wants_graph(m, x) and AG.wants_graph(m, x.clone().requires_grad_(True))

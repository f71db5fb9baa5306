"""A stand-in with the attribute surface of peft's LoRA layers (peft is not installed in this image).

`LoraLinear` mirrors what `peft.tuners.lora.Linear` exposes and what `get_peft_model(model, LoraConfig(r, lora_alpha,
target_modules, lora_dropout, modules_to_save))` (finetune/lora_tune.py:173-186) leaves in the module tree:
`base_layer` (the frozen nn.Linear), ModuleDicts `lora_A` / `lora_B` / `lora_dropout`, dict `scaling`,
`active_adapters`, `merged`, `disable_adapters`, and a forward that computes
    base(x) + B(A(dropout(x))) * scaling                      (peft/tuners/lora/layer.py Linear.forward)
`ModulesToSave` mirrors peft.utils.other.ModulesToSaveWrapper (`original_module`, `modules_to_save[adapter]`).
"""
import copy
import math

import torch
from torch import nn


class LoraLinear(nn.Module):
    def __init__(self, base, r=16, lora_alpha=16, lora_dropout=0.0, adapter="default", seed=0):
        super().__init__()
        self.base_layer = base
        for p in base.parameters():
            p.requires_grad_(False)
        self.lora_A = nn.ModuleDict({adapter: nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter: nn.Linear(r, base.out_features, bias=False)})
        self.lora_dropout = nn.ModuleDict({adapter: nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.scaling = {adapter: lora_alpha / r}
        self.active_adapters = [adapter]
        self.merged = False
        self.disable_adapters = False
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():                                  # peft zero-initialises B; tests want a visible adapter
            self.lora_A[adapter].weight.copy_(torch.randn(r, base.in_features, generator=g) / math.sqrt(base.in_features))
            self.lora_B[adapter].weight.copy_(torch.randn(base.out_features, r, generator=g) * 0.05)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x):
        y = self.base_layer(x)
        if self.merged or self.disable_adapters:
            return y
        for a in self.active_adapters:
            y = y + self.lora_B[a](self.lora_A[a](self.lora_dropout[a](x))) * self.scaling[a]
        return y


class ModulesToSave(nn.Module):
    def __init__(self, module, adapter="default"):
        super().__init__()
        self.original_module = module
        self.modules_to_save = nn.ModuleDict({adapter: copy.deepcopy(module)})
        self.active_adapters = [adapter]
        self.disable_adapters = False
        for p in self.original_module.parameters():
            p.requires_grad_(False)

    def forward(self, *a, **k):
        return self.modules_to_save[self.active_adapters[0]](*a, **k)


def wrap_linears(model, **kw):
    """What get_peft_model does with target_modules = every nn.Linear (finetune/lora_tune.py:169-172)."""
    for name, mod in list(model.named_modules()):
        for child_name, child in list(mod.named_children()):
            if type(child) is nn.Linear:
                setattr(mod, child_name, LoraLinear(child, **kw))
    for p in model.parameters():
        p.requires_grad_(False)
    for m in model.modules():
        if isinstance(m, LoraLinear):
            for p in list(m.lora_A.parameters()) + list(m.lora_B.parameters()):
                p.requires_grad_(True)
    return model

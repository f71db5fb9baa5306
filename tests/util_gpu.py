"""Helpers shared by the GPU parity tests."""
import torch

import pangu_oracle as orc


def load_params(module, params, prefix, device="cuda"):
    sd = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
    module.load_state_dict(sd, strict=True)
    return module.to(device).eval()


def check_digest(goldens, name, t, tol):
    flat = t.detach().reshape(-1)
    assert flat.numel() == int(goldens[name + ".numel"]), name
    pos = torch.from_numpy(goldens[name + ".pos"]).to(flat.device)
    ref = torch.from_numpy(goldens[name + ".val"])
    got = flat[pos].float().cpu()
    err = orc.rel_l2(got, ref)
    assert err <= tol, f"{name}: rel-L2 {err:.3e} > {tol:g}"
    return err

"""TEST INFRASTRUCTURE ONLY -- fixtures for `pixart_sigma_b200.iddpm` from the UNMODIFIED reference sampler:
`IDDPM(str(steps)).p_sample_loop(model, z.shape, z, clip_denoised=..., model_kwargs=..., device='cpu')`
(scripts/inference.py:95-101) with a deterministic toy denoiser and a seeded torch RNG (the loop draws one `randn_like`
per step).  Stored: the initial noise, the final sample, the timestep map and the timesteps the model was called with.
Run in the build container only:   python oracle/gen_golden_iddpm.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refshim import install_reference_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {"iddpm_sample_s10": dict(steps="10", clip=False, seed=3), "iddpm_sample_s100": dict(steps="100", clip=False, seed=4),
         "iddpm_sample_s25_clip": dict(steps="25", clip=True, seed=5), "iddpm_sample_sections": dict(steps="4,6,10", clip=False, seed=6)}


def toy_model(calls=None):
    def model(x, timestep, scale=1.0, **kw):
        if calls is not None:
            calls.append(timestep.clone())
        tt = (timestep.float() / 1000.0).view(-1, 1, 1, 1)
        return torch.cat([scale * 0.6 * torch.tanh(x) * torch.cos(3.0 * tt) + 0.1 * tt, 0.8 * torch.sin(2.0 * x + tt)], dim=1)
    return model


def main():
    install_reference_shims()
    from diffusion import IDDPM
    for name, c in CASES.items():
        g = torch.Generator().manual_seed(c["seed"])
        z = torch.randn(3, 4, 8, 8, generator=g)
        calls = []
        torch.manual_seed(100 + c["seed"])
        d = IDDPM(c["steps"])
        out = d.p_sample_loop(toy_model(calls), z.shape, z, clip_denoised=c["clip"], model_kwargs=dict(scale=0.9), device="cpu")
        torch.save({"steps": c["steps"], "clip": c["clip"], "seed": c["seed"], "z": z, "out": out, "timestep_map": list(d.timestep_map),
                    "calls": torch.stack(calls), "generator": "oracle/gen_golden_iddpm.py", "reference_commit": "1ce521af"},
                   os.path.join(OUT, name + ".pt"))
        print(name, d.num_timesteps, "steps, |out| =", float(out.norm()))


if __name__ == "__main__":
    main()

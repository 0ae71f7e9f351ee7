"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/dpm_*.pt by running the UNMODIFIED reference sampler
(`diffusion.DPMS` from /root/reference, imported through oracle/refshim.py) on seeded inputs with the deterministic toy
denoiser of oracle/dpm_oracle.py.  Build container only:  python oracle/gen_golden_dpm.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dpm_oracle as do               # noqa: E402
from oracle.refshim import install_reference_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {"dpm_s2": dict(steps=2, n=1, hw=(8, 8), cfg=4.5), "dpm_s5": dict(steps=5, n=2, hw=(8, 12), cfg=4.5),
         "dpm_s20": dict(steps=20, n=2, hw=(16, 16), cfg=4.5), "dpm_s7_cfg7": dict(steps=7, n=1, hw=(8, 8), cfg=7.0)}


def inputs(case, seed=0):
    g = torch.Generator().manual_seed(seed)
    n, (h, w) = case["n"], case["hw"]
    z = torch.randn(n, 4, h, w, generator=g)
    cond = torch.randn(n, 1, 6, 8, generator=g) * 0.5
    uncond = (torch.randn(1, 1, 6, 8, generator=g) * 0.5).repeat(n, 1, 1, 1)
    return z, cond, uncond


def main():
    install_reference_shims()
    import tqdm as _tqdm  # noqa: F401  (the reference's loop wraps its range in tqdm)
    from diffusion import DPMS
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        z, cond, uncond = inputs(case)
        solver = DPMS(do.toy_model, condition=cond, uncondition=uncond, cfg_scale=case["cfg"], model_kwargs={})
        with torch.no_grad():
            out, inter = solver.sample(z.clone(), steps=case["steps"], order=2, skip_type="time_uniform",
                                       method="multistep", return_intermediate=True)
        # the model-input times the reference hands to the denoiser (survey appendix A: 999.0, 749.25, ... for 4 steps)
        seen = []
        def spy(x, t, c, **kw):
            seen.append(float(t[0]))
            return do.toy_model(x, t, c, **kw)
        DPMS(spy, condition=cond, uncondition=uncond, cfg_scale=case["cfg"], model_kwargs={}).sample(
            z.clone(), steps=case["steps"], order=2, skip_type="time_uniform", method="multistep")
        torch.save(dict(case=case, out=out, x_after_first_update=inter[1] if len(inter) > 1 else inter[0],
                        model_times=torch.tensor(seen)), os.path.join(OUT, name + ".pt"))
        print(name, tuple(out.shape), "times", [round(v, 3) for v in seen[:4]], "...")


if __name__ == "__main__":
    main()

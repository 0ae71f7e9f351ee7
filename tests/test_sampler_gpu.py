"""GPU tests (-m gpu) of the DPM-Solver++ row: the fused step kernel and the sampling loop of
pixart_sigma_b200/sampler.py against the oracle restatement (oracle/dpm_oracle.py) and the reference fixtures.

Tolerances: step kernel vs the fp32 formula 1e-6 (fp32 arithmetic, reciprocal instead of division, FMA contraction);
loop with the deterministic toy denoiser vs the reference fixture 1e-5 (20 steps of the above); loop around the sm_100a
PixArtMS vs the oracle loop around the oracle forward 2e-2 (4 denoiser evaluations at the model's ~3e-3 each, amplified by
the solver's 1/sigma factors)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dpm_oracle as do            # noqa: E402
from oracle import pixart_oracle as po         # noqa: E402
from oracle.gen_golden_dpm import CASES, inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"

if torch.cuda.is_available():
    from pixart_sigma_b200 import PixArtMS, lib, sampler


def _rel(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("second_order", [False, True])
def test_step_kernel_matches_formula(dtype, second_order):
    g = torch.Generator().manual_seed(3)
    n, h, w = 3, 24, 20
    full = torch.randn(2 * n, 8, h, w, generator=g).to(dtype).to(DEV)       # learn-sigma output: eps = first 4 channels
    out = full[:, :4]                                                       # the view forward_with_dpmsolver returns
    x = torch.randn(n, 4, h, w, generator=g).to(DEV)
    prev = torch.randn(n, 4, h, w, generator=g).to(DEV)
    k = dict(cfg_scale=4.5, sigma_s=0.83, alpha_s=0.557, a=0.91, b=-0.21, c=(-0.08 if second_order else 0.0))
    eu, ec = out.float().chunk(2)
    eps = eu + k["cfg_scale"] * (ec - eu)
    x0 = (x - k["sigma_s"] * eps) / k["alpha_s"]
    want = k["a"] * x - k["b"] * x0 - k["c"] * (x0 - prev)
    got_x, got_prev = x.clone(), prev.clone()
    lib.dpm_solver_pp_step(out, got_x, got_prev, **k)
    assert _rel(got_x, want) < 1e-6
    assert _rel(got_prev, x0) < 1e-6


@pytest.mark.parametrize("name", sorted(CASES))
def test_loop_with_toy_denoiser_matches_reference_fixture(name):
    case = CASES[name]
    gold = torch.load(os.path.join(GOLD, name + ".pt"))
    z, cond, uncond = inputs(case)
    solver = sampler.DPMS(do.toy_model, condition=cond.to(DEV), uncondition=uncond.to(DEV), cfg_scale=case["cfg"], model_kwargs={})
    out = solver.sample(z.to(DEV), steps=case["steps"], order=2, skip_type="time_uniform", method="multistep")
    assert out.shape == gold["out"].shape and _rel(out, gold["out"]) < 1e-5
    out2, inter = solver.sample(z.to(DEV), steps=case["steps"], return_intermediate=True)
    assert len(inter) == case["steps"] and _rel(inter[0], gold["x_after_first_update"]) < 1e-5
    graphed = solver.sample(z.to(DEV), steps=case["steps"], cuda_graph=True)
    again = solver.sample(z.to(DEV), steps=case["steps"], cuda_graph=True)      # replay of the cached graph
    assert torch.equal(graphed, again) and _rel(graphed, gold["out"]) < 1e-5


def test_unguided_sampling_matches_oracle():
    case = CASES["dpm_s5"]
    z, cond, uncond = inputs(case)
    solver = sampler.DPMS(do.toy_model, condition=cond.to(DEV), uncondition=None, cfg_scale=4.5)
    out = solver.sample(z.to(DEV), steps=5)
    # oracle with cfg_scale 1: eps = eu + 1 * (ec - eu) = ec, i.e. the conditional branch alone
    want = do.sample(do.toy_model, z, cond, uncond, 1.0, steps=5)
    assert _rel(out, want) < 1e-5


def test_loop_around_the_sm100_model_matches_oracle_loop():
    """4-step CFG sampling of a depth-2 PixArtMS at 256px (latent 32x32) on the sm_100a kernels vs the oracle loop around the
    fp32 oracle forward on the same bf16-rounded weights."""
    cfg = po.OracleConfig(depth=2, input_size=32, pe_interpolation=0.5)
    sd = {k: v.to(torch.bfloat16).float() for k, v in po.synthetic_state_dict(cfg, seed=0).items()}
    z, _, y, mask = po.synthetic_inputs(cfg, 1, (32, 32), seed=5, lens=[77])
    null_y = po.synthetic_inputs(cfg, 1, (32, 32), seed=6)[2]
    y, null_y = y.to(torch.bfloat16).float(), null_y.to(torch.bfloat16).float()
    with torch.device(DEV):
        m = PixArtMS(depth=2, input_size=32, pe_interpolation=0.5, model_max_length=300)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and missing == ["pos_embed"]
    m = m.to(torch.bfloat16).eval()
    kw = dict(data_info=None, mask=mask.to(DEV))
    solver = sampler.DPMS(m.forward_with_dpmsolver, condition=y.to(DEV), uncondition=null_y.to(DEV), cfg_scale=4.5, model_kwargs=kw)
    launches = lib.launch_count()
    out = solver.sample(z.to(DEV), steps=4, order=2, skip_type="time_uniform", method="multistep")
    assert lib.launch_count() - launches >= 4 * (2 * 11 + 1)           # 4 x (2 blocks x 11 kernels + the step kernel)

    def oracle_model(x, t, c, **kwargs):
        return po.forward_with_dpmsolver(sd, cfg, x, t, c, data_info=None, mask=mask)
    want = do.sample(oracle_model, z, y, null_y, 4.5, steps=4)
    err = _rel(out, want)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity.txt"), "a") as f:
        f.write(f"sampler 4-step loop depth2 256px: rel_err vs oracle loop = {err:.3e}\n")
    assert err < 2e-2
    graphed = solver.sample(z.to(DEV), steps=4, cuda_graph=True)
    assert _rel(graphed, out) < 1e-5

"""CPU tests of the DPM-Solver++ row: the oracle restatement against fixtures produced by the unmodified
reference sampler (oracle/gen_golden_dpm.py), and the product's host-side step plan against the oracle's."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dpm_oracle as do            # noqa: E402
from oracle.gen_golden_dpm import CASES, inputs  # noqa: E402
from pixart_sigma_b200 import sampler          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_sampler_matches_reference_fixture(name):
    """Bit-level agreement is expected (same float32 operations in the same order); 1e-6 leaves room for BLAS-free
    elementwise differences between torch builds."""
    case = CASES[name]
    g = torch.load(os.path.join(GOLD, name + ".pt"))
    z, cond, uncond = inputs(case)
    seen = []

    def spy(x, t, c, **kw):
        seen.append(float(t[0]))
        return do.toy_model(x, t, c, **kw)

    out, inter = do.sample(spy, z, cond, uncond, case["cfg"], steps=case["steps"], return_intermediate=True)
    assert len(seen) == case["steps"]                                   # one denoiser call per step, none at t_end
    assert torch.allclose(torch.tensor(seen), g["model_times"], atol=1e-3)
    assert float((out - g["out"]).norm() / g["out"].norm()) < 1e-6
    ref_first = g["x_after_first_update"]          # the reference's intermediates start with the initial x
    assert float((inter[0] - ref_first).norm() / ref_first.norm()) < 1e-6


def test_model_times_of_the_four_step_probe():
    """steps=4 => model times 999.0, 749.25, 499.5, 249.75."""
    sch = do.DiscreteSchedule()
    co = do.step_coefficients(sch, do.time_steps(sch, 4))
    assert [round(c["t_input"], 2) for c in co] == [999.0, 749.25, 499.5, 249.75]
    assert [c["order"] for c in co] == [1, 2, 2, 1]


@pytest.mark.parametrize("steps", [2, 5, 20, 33])
def test_product_step_plan_matches_oracle(steps):
    solver = sampler.DPMS(lambda *a, **k: None, condition=torch.zeros(1, 1, 2, 2), uncondition=torch.zeros(1, 1, 2, 2), cfg_scale=4.5)
    plan = solver.plan(steps)
    sch = do.DiscreteSchedule()
    want = do.step_coefficients(sch, do.time_steps(sch, steps))
    assert len(plan) == len(want) == steps
    for p, w in zip(plan, want):
        assert p["order"] == w["order"]
        for k in ("t_input", "sigma_s", "alpha_s", "a", "b", "c"):
            assert abs(p[k] - w[k]) <= 1e-6 * max(1.0, abs(w[k])), (k, p[k], w[k])


def test_sampler_rejects_cpu_tensors_and_unsupported_modes():
    solver = sampler.DPMS(lambda *a, **k: None, condition=torch.zeros(1, 1, 2, 2), uncondition=torch.zeros(1, 1, 2, 2), cfg_scale=4.5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        solver.sample(torch.zeros(1, 4, 8, 8), steps=4)
    with pytest.raises(NotImplementedError):
        solver.sample(torch.zeros(1, 4, 8, 8), steps=4, method="singlestep")
    with pytest.raises(NotImplementedError):
        sampler.DPMS(lambda *a, **k: None, None, None, 1.0, model_type="v")

"""CPU tests of the IDDPM sampling loop (`pixart_sigma_b200.iddpm`) against fixtures from the unmodified reference sampler
(oracle/gen_golden_iddpm.py): same respaced schedule, the model sees the original timestep indices, and -- with the same
torch RNG stream -- the same samples."""
import os

import pytest
import torch

from oracle.gen_golden_iddpm import toy_model
from pixart_sigma_b200.iddpm import IDDPM, space_timesteps

CASES = ["iddpm_sample_s10", "iddpm_sample_s100", "iddpm_sample_s25_clip", "iddpm_sample_sections"]


@pytest.mark.parametrize("name", CASES)
def test_p_sample_loop_matches_reference(golden_dir, name):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    d = IDDPM(fix["steps"])
    assert d.timestep_map == fix["timestep_map"] and d.num_timesteps == len(fix["timestep_map"])
    calls = []
    torch.manual_seed(100 + fix["seed"])
    out = d.p_sample_loop(toy_model(calls), fix["z"].shape, fix["z"], clip_denoised=fix["clip"], model_kwargs=dict(scale=0.9),
                          device="cpu")
    assert torch.equal(torch.stack(calls), fix["calls"])                     # original timestep indices, newest first
    assert torch.allclose(out, fix["out"], rtol=1e-5, atol=1e-6), float((out - fix["out"]).abs().max())


def test_space_timesteps_edges():
    assert space_timesteps(1000, "1000") == list(range(1000))
    assert space_timesteps(1000, "1") == [0]
    assert space_timesteps(100, "ddim10") == list(range(0, 100, 10))
    with pytest.raises(ValueError):
        space_timesteps(10, "20")


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        IDDPM("10", predict_xstart=True)
    with pytest.raises(NotImplementedError):
        IDDPM("10").p_sample_loop(toy_model(), (1, 4, 8, 8), torch.zeros(1, 4, 8, 8), cond_fn=lambda *a: None, device="cpu")

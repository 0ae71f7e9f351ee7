"""IDDPM ancestral sampling loop around the denoiser -- the third caller of the path in `scripts/inference.py`
(`--sampling_algo iddpm`, inference.py:89-102: `IDDPM(str(sample_steps)).p_sample_loop(model.forward_with_cfg, z.shape, z,
clip_denoised=False, model_kwargs=..., device=...)`), next to the DPM-Solver++ loop (`sampler.py`) and the training loss
(`training.py`).  SURVEY.md 8a row 18 keeps these outer loops in PyTorch: this is a restatement of

    diffusion/iddpm.py:10-53              IDDPM(...) factory (linear betas, learned-range variance, epsilon prediction)
    diffusion/model/respace.py:12-94      space_timesteps / SpacedDiffusion (re-derived betas, timestep map)
    diffusion/model/respace.py:121-134    _WrappedModel (the model sees the ORIGINAL timestep index, as `timestep=`)
    diffusion/model/gaussian_diffusion.py:280-361, 405-446, 493-540   p_mean_variance / p_sample / p_sample_loop_progressive

so that sampling runs without the reference package (the reference's own object also works unchanged on the model).
Pure elementwise torch on (n, 4, h, w) latents around one `forward_with_cfg` per step; pinned against the unmodified
reference sampler by `tests/golden/iddpm_sample_*.pt` (same torch RNG stream: one `randn_like` per step).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch


def space_timesteps(num_timesteps: int, section_counts) -> List[int]:
    """Timesteps of the original process kept by a respacing string such as "100" or "10,15,20" (respace.py:12-63)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == want:
                    return sorted(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return sorted(set(steps))


class IDDPMSampler:
    """`IDDPM(timestep_respacing)` restricted to what inference uses: linear schedule, epsilon prediction, learned-range
    variance (`learn_sigma=True, pred_sigma=True`)."""

    def __init__(self, timestep_respacing, diffusion_steps: int = 1000):
        scale = 1000 / diffusion_steps                                        # gaussian_diffusion.py:107-116
        base_betas = np.linspace(scale * 0.0001, scale * 0.02, diffusion_steps, dtype=np.float64)
        base_ac = np.cumprod(1.0 - base_betas, axis=0)
        if timestep_respacing is None or timestep_respacing == "":
            timestep_respacing = [diffusion_steps]
        keep = set(space_timesteps(diffusion_steps, timestep_respacing))
        self.timestep_map, betas, last = [], [], 1.0                          # respace.py:73-86
        for i, ac in enumerate(base_ac):
            if i in keep:
                betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        betas = np.array(betas, dtype=np.float64)
        self.num_timesteps = len(betas)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)                       # gaussian_diffusion.py:214-227
        self.tab = {
            "sqrt_recip_ac": np.sqrt(1.0 / ac), "sqrt_recipm1_ac": np.sqrt(1.0 / ac - 1),
            "post_logvar": np.log(np.append(post_var[1], post_var[1:])) if len(post_var) > 1 else np.array([]),
            "post_c1": betas * np.sqrt(ac_prev) / (1.0 - ac), "post_c2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
            "log_betas": np.log(betas),
        }
        self._dev: Dict[str, torch.Tensor] = {}

    def _x(self, name: str, t: torch.Tensor) -> torch.Tensor:                 # _extract_into_tensor, :1029-1041
        key = f"{name}@{t.device}"
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.tab[name]).to(t.device)
        return self._dev[key][t].float()[:, None, None, None]

    def p_mean_variance(self, model, x, t, clip_denoised: bool = True, model_kwargs: Optional[dict] = None):
        """:280-361 with ModelVarType.LEARNED_RANGE and ModelMeanType.EPSILON; the model is called like _WrappedModel does."""
        n, C = x.shape[:2]
        map_t = torch.tensor(self.timestep_map, device=t.device, dtype=t.dtype)[t]
        out = model(x, timestep=map_t, **(model_kwargs or {}))
        assert out.shape == (n, 2 * C, *x.shape[2:])
        eps, var_values = torch.split(out, C, dim=1)
        min_log, max_log = self._x("post_logvar", t), self._x("log_betas", t)
        frac = (var_values + 1) / 2
        log_variance = frac * max_log + (1 - frac) * min_log
        pred_xstart = self._x("sqrt_recip_ac", t) * x - self._x("sqrt_recipm1_ac", t) * eps
        if clip_denoised:
            pred_xstart = pred_xstart.clamp(-1, 1)
        mean = self._x("post_c1", t) * pred_xstart + self._x("post_c2", t) * x
        return {"mean": mean, "variance": torch.exp(log_variance), "log_variance": log_variance, "pred_xstart": pred_xstart}

    def p_sample(self, model, x, t, clip_denoised: bool = True, model_kwargs: Optional[dict] = None):
        """:405-446: x_{t-1} = mean + [t != 0] * exp(0.5 * log_variance) * noise."""
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
        noise = torch.randn_like(x)
        nonzero = (t != 0).float().view(-1, 1, 1, 1)
        return {"sample": out["mean"] + nonzero * torch.exp(0.5 * out["log_variance"]) * noise, "pred_xstart": out["pred_xstart"]}

    @torch.no_grad()
    def p_sample_loop_progressive(self, model, shape: Sequence[int], noise=None, clip_denoised: bool = True,
                                  model_kwargs: Optional[dict] = None, device=None, progress: bool = False) -> Iterator[dict]:
        if device is None:
            device = next(model.parameters()).device
        img = noise if noise is not None else torch.randn(*shape, device=device)
        for i in reversed(range(self.num_timesteps)):                          # :519-540
            t = torch.tensor([i] * shape[0], device=device)
            out = self.p_sample(model, img, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
            yield out
            img = out["sample"]

    def p_sample_loop(self, model, shape, noise=None, clip_denoised: bool = True, denoised_fn=None, cond_fn=None,
                      model_kwargs: Optional[dict] = None, device=None, progress: bool = False) -> torch.Tensor:
        """Same call signature as gaussian_diffusion.py:448-491 (denoised_fn / cond_fn are not used by inference.py and
        must be None)."""
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn are not supported by the B200 IDDPM sampling loop")
        final = None
        for final in self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                    model_kwargs=model_kwargs, device=device, progress=progress):
            pass
        return final["sample"]


def IDDPM(timestep_respacing, noise_schedule: str = "linear", use_kl: bool = False, sigma_small: bool = False,
          predict_xstart: bool = False, learn_sigma: bool = True, pred_sigma: bool = True, rescale_learned_sigmas: bool = False,
          diffusion_steps: int = 1000, snr: bool = False, return_startx: bool = False) -> IDDPMSampler:
    """Factory with the reference's signature (diffusion/iddpm.py:10-22); only the configuration inference uses exists."""
    if (noise_schedule != "linear" or use_kl or sigma_small or predict_xstart or not learn_sigma or not pred_sigma or
            rescale_learned_sigmas or snr or return_startx):
        raise NotImplementedError("only IDDPM(str(steps)) with its defaults (linear schedule, epsilon prediction, learned-range "
                                  "variance) is restated here")
    return IDDPMSampler(timestep_respacing, diffusion_steps)

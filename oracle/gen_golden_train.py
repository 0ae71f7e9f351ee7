"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/train_*.pt by running the UNMODIFIED reference training
computation (`/root/reference` through oracle/refshim.py): `IDDPM(str(1000), learn_sigma=True, pred_sigma=True,
snr=False).training_losses(model, x0, t, model_kwargs)` (train_scripts/train.py:189-197, 410) followed by
`loss.mean().backward()`, fp32 on CPU, seeded synthetic weights / inputs (regenerated from seeds, not stored).

Stored per case: the loss terms, and for EVERY parameter a gradient fingerprint -- its L2 norm and 64 entries at seeded
positions -- so the fixture stays small while pinning every gradient.  Also `train_loss_only.pt`: the loss terms of the
reference for a fixed synthetic model output (pins `pixart_sigma_b200.training.IDDPMLoss` without any model).

Run in the build container only:   python oracle/gen_golden_train.py
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pixart_oracle as po              # noqa: E402
from oracle.gen_golden import build_reference       # noqa: E402
from oracle.refshim import install_reference_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # depth-2 XL/2-width model, 256px latents (N = 256 tokens), ragged caption lengths, t spanning the schedule (incl. 0:
    # the decoder-NLL branch of the vb term)
    "train_d2_b2": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5), batch=2, hw=(32, 32), t=[0, 731], lens=[300, 41]),
    "train_d2_b3_512": dict(cfg=dict(depth=2, input_size=64, pe_interpolation=1.0), batch=3, hw=(64, 64), t=[999, 5, 250],
                            lens=[120, 300, 7]),
    # KV-compressed second block (conv sr=2 + LayerNorm, PixArt_blocks.py:84-89): pins the gradients of attn.sr / attn.norm
    "train_d2_kvconv": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5, kv_sampling="conv", kv_scale_factor=2,
                                     kv_compress_layer=[1]), batch=2, hw=(32, 32), t=[400, 90], lens=[64, 300]),
}


def fingerprint_positions(numel: int, seed: int = 1234, k: int = 64) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + numel % 9973)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def train_inputs(cfg: po.OracleConfig, batch, hw, t, lens, seed=0):
    x, _, y, mask = po.synthetic_inputs(cfg, batch, hw, seed=seed, lens=lens)
    g = torch.Generator().manual_seed(seed + 77)
    x0 = x * 0.5                                                   # latents ~ VAE scale
    noise = torch.randn(x.shape, generator=g)
    return x0, torch.tensor(t, dtype=torch.long), y, mask, noise


def main():
    install_reference_shims()
    from diffusion import IDDPM
    os.makedirs(OUT, exist_ok=True)
    diffusion = IDDPM(str(1000), learn_sigma=True, pred_sigma=True, snr=False)

    # (1) loss only: a fixed "model output"
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(4, 4, 16, 16, generator=g) * 0.5
    x0[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -0.9995, 0.9995])       # exercise the discretized-likelihood edge branches
    noise = torch.randn(4, 4, 16, 16, generator=g)
    fake = torch.randn(4, 8, 16, 16, generator=g)
    t = torch.tensor([0, 1, 500, 999])
    terms = diffusion.training_losses(lambda x, timestep, **kw: fake, x0, t, model_kwargs={}, noise=noise)
    torch.save({"x0": x0, "noise": noise, "fake": fake, "t": t, "terms": {k: v.clone() for k, v in terms.items()},
                "x_t": diffusion.q_sample(x0, t, noise=noise), "generator": "oracle/gen_golden_train.py",
                "reference_commit": "1ce521af"}, os.path.join(OUT, "train_loss_only.pt"))
    print("train_loss_only:", {k: v.tolist() for k, v in terms.items()})

    # (2) full training step through the reference model
    for name, c in CASES.items():
        t0 = time.time()
        cfg = po.OracleConfig(**c["cfg"])
        sd = po.synthetic_state_dict(cfg, seed=0)
        ref = build_reference(cfg, sd).train()
        ref.y_embedder.uncond_prob = 0.0                           # token_drop calls .cuda() (PixArt_blocks.py:394) and is random
        x0, t, y, mask, noise = train_inputs(cfg, c["batch"], c["hw"], c["t"], c["lens"])
        terms = diffusion.training_losses(ref, x0, t, model_kwargs=dict(y=y, mask=mask, data_info=None), noise=noise)
        terms["loss"].mean().backward()
        grads = {}
        for n, p in ref.named_parameters():
            gflat = p.grad.detach().flatten().float()
            grads[n] = {"norm": gflat.norm().clone(), "vals": gflat[fingerprint_positions(gflat.numel())].clone()}
        torch.save({"case": name, "cfg": c["cfg"], "batch": c["batch"], "hw": c["hw"], "t": c["t"], "lens": c["lens"],
                    "terms": {k: v.detach().clone() for k, v in terms.items()}, "grads": grads,
                    "generator": "oracle/gen_golden_train.py", "reference_commit": "1ce521af"},
                   os.path.join(OUT, name + ".pt"))
        print(f"{name}: loss {terms['loss'].tolist()} {len(grads)} grads ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()

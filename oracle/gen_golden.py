"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the UNMODIFIED reference
(`/root/reference`, imported through oracle/refshim.py) on seeded synthetic weights and inputs.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden.py
Weights and inputs are NOT stored: both are regenerated from seeds by
`oracle.pixart_oracle.synthetic_state_dict / synthetic_inputs`, so fixtures stay a few hundred KB.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pixart_oracle as po            # noqa: E402
from oracle.refshim import install_reference_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (ctor overrides, batch, latent (H, W), timesteps, prefix lens or None, extra)
CASES = {
    # full XL/2, BASELINE config c1 (256px CPU gate): t=500 with the 77-token mask of the survey probe
    "xl2_256_b1_mask77": dict(cfg=dict(depth=28, input_size=32, pe_interpolation=0.5), batch=1, hw=(32, 32),
                              t=[500.0], lens=[77]),
    # full XL/2, CFG-style batch 2, fractional DPM-Solver times, ragged lengths incl. full and tiny
    "xl2_256_b2_ragged": dict(cfg=dict(depth=28, input_size=32, pe_interpolation=0.5), batch=2, hw=(32, 32),
                              t=[749.25, 49.95], lens=[300, 8]),
    "d2_nomask": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5), batch=2, hw=(32, 32),
                      t=[999.0, 499.5], lens=None),
    "d2_nonsquare": dict(cfg=dict(depth=2, input_size=64, pe_interpolation=1.0), batch=1, hw=(48, 80),
                         t=[249.75], lens=[120]),
    "d2_kvconv": dict(cfg=dict(depth=2, input_size=64, pe_interpolation=1.0, kv_sampling="conv", kv_scale_factor=2,
                               kv_compress_layer=[1]), batch=1, hw=(64, 64), t=[949.05], lens=[33]),
    "d2_kvave": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5, kv_sampling="ave", kv_scale_factor=2,
                              kv_compress_layer=[0, 1]), batch=1, hw=(32, 32), t=[500.0], lens=[64]),
    "d2_kvuniform": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5, kv_sampling="uniform",
                                  kv_scale_factor=2, kv_compress_layer=[1]), batch=1, hw=(32, 32), t=[500.0], lens=[64]),
    "d2_kvuniform_every": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5, kv_sampling="uniform_every",
                                        kv_scale_factor=2, kv_compress_layer=[1]), batch=1, hw=(32, 32), t=[500.0],
                               lens=[64]),
    "d2_qknorm": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5, qk_norm=True), batch=1, hw=(32, 32),
                      t=[500.0], lens=[200]),
    "d2_micro": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5, micro_condition=True), batch=2,
                     hw=(32, 32), t=[500.0, 20.0], lens=[300, 150], micro=True),
    "d2_emptykeys": dict(cfg=dict(depth=2, input_size=32, pe_interpolation=0.5), batch=2, hw=(32, 32),
                         t=[500.0, 500.0], lens=[0, 5]),
}


def build_reference(cfg: po.OracleConfig, sd):
    from diffusion.model.nets.PixArtMS import PixArtMS
    kvc = None
    if cfg.kv_sampling is not None:
        kvc = dict(sampling=cfg.kv_sampling, scale_factor=cfg.kv_scale_factor, kv_compress_layer=list(cfg.kv_compress_layer))
    m = PixArtMS(input_size=cfg.input_size, patch_size=cfg.patch_size, hidden_size=cfg.hidden_size, depth=cfg.depth,
                 num_heads=cfg.num_heads, pe_interpolation=cfg.pe_interpolation, model_max_length=cfg.model_max_length,
                 micro_condition=cfg.micro_condition, qk_norm=cfg.qk_norm, kv_compress_config=kvc)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k == "pos_embed" for k in missing), missing       # recomputed per forward (checkpoint.py:54-60)
    return m.eval()


def main():
    install_reference_shims()
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        cfg = po.OracleConfig(**c["cfg"])
        sd = po.synthetic_state_dict(cfg, seed=0)
        x, t, y, mask = po.synthetic_inputs(cfg, c["batch"], c["hw"], seed=0, timesteps=c["t"], lens=c["lens"])
        data_info = None
        if c.get("micro"):
            data_info = {"img_hw": torch.tensor([[256.0, 256.0]]).repeat(c["batch"], 1),
                         "aspect_ratio": torch.tensor([[1.0]]).repeat(c["batch"], 1)}
        ref = build_reference(cfg, sd)
        taps = {}
        hooks = [ref.blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("block0", o.detach().clone())),
                 ref.blocks[-1].register_forward_hook(lambda m, i, o: taps.__setitem__("block_last", o.detach().clone()))]
        with torch.no_grad():
            out = ref(x, t, y, mask=mask, data_info=data_info)
        for hk in hooks:
            hk.remove()
        fix = {"case": name, "cfg": c["cfg"], "batch": c["batch"], "hw": c["hw"], "t": c["t"], "lens": c["lens"],
               "micro": bool(c.get("micro")), "out": out.float().clone(),
               # token-strided taps keep the fixtures small but still pin every block boundary value range
               "block0_tap": taps["block0"][:, ::37].float().clone(),
               "block_last_tap": taps["block_last"][:, ::37].float().clone(),
               "reference_commit": "1ce521af", "generator": "oracle/gen_golden.py"}
        torch.save(fix, os.path.join(OUT, name + ".pt"))
        print(f"{name}: out {tuple(out.shape)} |out|={out.norm():.4f} ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()

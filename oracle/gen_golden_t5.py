"""TEST INFRASTRUCTURE ONLY -- fixtures for `pixart_sigma_b200.t5.T5EncoderModel` from the reference's own dependency:
`transformers.T5EncoderModel(...)(input_ids, attention_mask)['last_hidden_state']` (diffusion/model/t5.py:10,107-110) in fp32 on
seed-defined synthetic weights (`synthetic_t5_state_dict`: no transformers initialiser involved, so the test can rebuild them without
transformers).  Stored: config, weight seed, token ids, mask, the fp32 output and the position-bias table of block 0.
Run in the build container only:   python oracle/gen_golden_t5.py   (transformers 5.5 here)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
CASES = {"t5_small_ragged": dict(cfg=dict(vocab_size=384, d_model=256, d_kv=64, d_ff=512, num_layers=3, num_heads=4), seed=11,
                                 B=3, L=40, lens=[40, 7, 23]),
         "t5_wide_300": dict(cfg=dict(vocab_size=256, d_model=512, d_kv=64, d_ff=1024, num_layers=2, num_heads=8), seed=12,
                             B=2, L=300, lens=[300, 77])}


def synthetic_t5_state_dict(cfg: dict, seed: int) -> dict:
    """Deterministic weights in transformers' key layout: matrices ~ N(0, 1 / fan_in) (q: N(0, 1 / (fan_in d_kv)), T5's own scale), RMS-norm weights 1 + 0.2 N(0, 1), the bias table
    8 N(0, 1) / sqrt(d_model), values rounded to bf16 so that the bf16 model and the fp32 oracle hold the same numbers."""
    g = torch.Generator().manual_seed(seed)
    D, inner, F_, H = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"], cfg["num_heads"]
    r = lambda *s, std=1.0: (torch.randn(*s, generator=g) * std).to(torch.bfloat16).float()
    sd = {"shared.weight": r(cfg["vocab_size"], D)}
    sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer."
        # q as in T5's Mesh-TensorFlow initialisation, (d_model d_kv)^-1/2: the UNSCALED logits q.k then have unit variance
        sd[p + "0.SelfAttention.q.weight"] = r(inner, D, std=(D * cfg["d_kv"]) ** -0.5)
        for n in "kv":
            sd[p + f"0.SelfAttention.{n}.weight"] = r(inner, D, std=D ** -0.5)
        sd[p + "0.SelfAttention.o.weight"] = r(D, inner, std=inner ** -0.5)
        if i == 0:
            sd[p + "0.SelfAttention.relative_attention_bias.weight"] = r(32, H, std=8.0 * D ** -0.5)
        sd[p + "0.layer_norm.weight"] = (1 + 0.2 * torch.randn(D, generator=g)).to(torch.bfloat16).float()
        sd[p + "1.DenseReluDense.wi_0.weight"] = r(F_, D, std=D ** -0.5)
        sd[p + "1.DenseReluDense.wi_1.weight"] = r(F_, D, std=D ** -0.5)
        sd[p + "1.DenseReluDense.wo.weight"] = r(D, F_, std=F_ ** -0.5)
        sd[p + "1.layer_norm.weight"] = (1 + 0.2 * torch.randn(D, generator=g)).to(torch.bfloat16).float()
    sd["encoder.final_layer_norm.weight"] = (1 + 0.2 * torch.randn(D, generator=g)).to(torch.bfloat16).float()
    return sd


def main():
    import transformers
    for name, c in CASES.items():
        cfg = c["cfg"]
        hf = transformers.T5EncoderModel(transformers.T5Config(feed_forward_proj="gated-gelu", dropout_rate=0.0, **cfg)).eval()
        missing, unexpected = hf.load_state_dict(synthetic_t5_state_dict(cfg, c["seed"]), strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        g = torch.Generator().manual_seed(100 + c["seed"])
        ids = torch.randint(0, cfg["vocab_size"], (c["B"], c["L"]), generator=g)
        mask = (torch.arange(c["L"])[None] < torch.tensor(c["lens"])[:, None]).long()
        with torch.no_grad():
            out = hf(input_ids=ids, attention_mask=mask)["last_hidden_state"]
            bias = hf.encoder.block[0].layer[0].SelfAttention.compute_bias(c["L"], c["L"])[0]
        torch.save({"cfg": cfg, "seed": c["seed"], "input_ids": ids, "attention_mask": mask, "last_hidden_state": out.half(),
                    "position_bias_h0": bias[0].clone() if c["L"] <= 64 else None, "generator": "oracle/gen_golden_t5.py",
                    "transformers": transformers.__version__}, os.path.join(OUT, name + ".pt"))
        print(name, tuple(out.shape), "|out| =", float(out.norm()))


if __name__ == "__main__":
    main()

"""GPU parity tests (-m gpu) of the T5 caption encoder: `pxa_rmsnorm_bf16` against torch, and `pixart_sigma_b200.t5.T5EncoderModel`
against transformers' T5EncoderModel (the reference's own dependency, diffusion/model/t5.py:10,107-110) run in fp32 on the host with
the same bf16-valued weights.  Bar 1e-2 normwise on `last_hidden_state`: bf16 GEMM operands, fp32 residual stream and statistics."""
import pytest
import torch

from oracle import pixart_oracle as po

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")
if torch.cuda.is_available():
    from pixart_sigma_b200 import lib
    from pixart_sigma_b200.t5 import T5EncoderModel


@pytest.mark.parametrize("M,C", [(300, 4096), (1201, 4096), (77, 256), (9, 10240)])
def test_rmsnorm_matches_torch(M, C):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(M, C, generator=g) * 3 + 0.2).cuda()
    w = (1 + 0.2 * torch.randn(C, generator=g)).to(torch.bfloat16).cuda()
    out = torch.full((M, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    lib.rmsnorm(x, w, out, eps=1e-6)
    want = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
    assert torch.isfinite(out.float()).all()
    assert po.rel_err(out.float(), want) < 4e-3
    big = torch.empty(M, C + 64, dtype=torch.bfloat16, device="cuda")            # strided output rows
    lib.rmsnorm(x, w, big[:, :C], eps=1e-6)
    assert torch.equal(big[:, :C], out)


@pytest.mark.parametrize("B,H,L,lens", [(1, 1, 128, None), (2, 4, 40, [40, 7]), (3, 8, 300, [300, 77, 129]), (2, 64, 300, [300, 8]),
                                        (1, 3, 384, [383]), (2, 2, 257, None), (1, 2, 16, [5])])
@pytest.mark.parametrize("toeplitz", [False, True])
def test_t5_attention_kernel_matches_torch(B, H, L, lens, toeplitz):
    """pxa_t5_attn_d64_bf16 on column slices of a fused (B*L, 3*H*64) qkv buffer against fp32 torch: unscaled logits + per-head bias
    [H, L, L] + additive key mask, softmax, P V.  1 / 2 / 3 key boxes, partial last boxes / chunks / 16-key steps, one sample
    running into the next one's rows inside a TMA box."""
    g = torch.Generator().manual_seed(5)
    inner = H * 64
    qkv = (torch.randn(B * L, 3 * inner, generator=g) * 0.5).to(torch.bfloat16).cuda()
    if toeplitz:       # T5's form: the bias depends on j - i only; the kernel takes the (H, 2L-1) vector and stages it in smem
        rel = (torch.randn(H, 2 * L - 1, generator=g) * 2).cuda()
        pos = torch.arange(L, device="cuda")
        bias = rel[:, pos[None, :] - pos[:, None] + L - 1].contiguous()
    else:
        rel = None
        bias = (torch.randn(H, L, L, generator=g) * 2).cuda()
    key_bias = None
    if lens is not None:
        keep = (torch.arange(L)[None] < torch.tensor(lens)[:, None]).float()
        key_bias = ((1.0 - keep) * torch.finfo(torch.float32).min).cuda().contiguous()
    out = torch.full((B * L, inner), float("nan"), dtype=torch.bfloat16, device="cuda")
    q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
    lib.t5_attn(q, k, v, out, None if toeplitz else bias, key_bias, B=B, H=H, L=L, scale=1.0, rel_bias=rel)
    q4, k4, v4 = (t.reshape(B, L, H, 64).transpose(1, 2).float() for t in (q, k, v))
    s = q4 @ k4.transpose(-1, -2) + bias[None]
    if key_bias is not None:
        s = s + key_bias.view(B, 1, 1, L)
    want = (torch.softmax(s, -1) @ v4).transpose(1, 2).reshape(B * L, inner)
    assert torch.isfinite(out.float()).all()
    assert po.rel_err(out.float(), want) < 6e-3
    with pytest.raises(lib.PxaError):
        big = torch.empty(385, 64, dtype=torch.bfloat16, device="cuda")
        lib.t5_attn(big, big, big, torch.empty_like(big), torch.zeros(1, 385, 385, device="cuda"), None, B=1, H=1, L=385)


def _pair(cfg, seed):
    torch.manual_seed(seed)
    hc = transformers.T5Config(feed_forward_proj="gated-gelu", dropout_rate=0.0, **cfg)
    hf = transformers.T5EncoderModel(hc).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "layer_norm" in n:
                p.add_(0.2 * torch.randn_like(p))
            if "relative_attention_bias" in n:
                p.mul_(8.0)
    m = T5EncoderModel(cfg)
    m.load_state_dict(hf.state_dict())
    m = m.to(torch.bfloat16).cuda()
    hf.load_state_dict({k: v.float().cpu() for k, v in m.state_dict().items()})
    return m, hf


@pytest.mark.parametrize("name,cfg,B,L,lens", [
    ("small", dict(vocab_size=384, d_model=256, d_kv=64, d_ff=512, num_layers=3, num_heads=4), 3, 40, [40, 7, 23]),
    ("xxl-width-2-layers", dict(vocab_size=512, d_model=4096, d_kv=64, d_ff=10240, num_layers=2, num_heads=64), 2, 300, [300, 77]),
    ("xxl-width-1-layer-b5", dict(vocab_size=512, d_model=4096, d_kv=64, d_ff=10240, num_layers=1, num_heads=64), 5, 300,
     [300, 8, 129, 300, 64]),          # M = 1500 rows: the CTA-pair GEMM path
])
@pytest.mark.parametrize("attn_impl", ["kernel", "torch"])
def test_t5_encoder_matches_transformers(name, cfg, B, L, lens, attn_impl):
    if attn_impl == "torch" and name == "xxl-width-2-layers":
        pytest.skip("the PyTorch attention core is covered by the small and the M = 1500 geometry (saves a 386 M parameter CPU model)")
    m, hf = _pair(cfg, seed=0)
    m.attn_impl = attn_impl
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, cfg["vocab_size"], (B, L), generator=g)
    mask = (torch.arange(L)[None] < torch.tensor(lens)[:, None]).long()
    n0 = lib.launch_count()
    got = m(input_ids=ids.cuda(), attention_mask=mask.cuda())["last_hidden_state"]
    torch.cuda.synchronize()
    launches = lib.launch_count() - n0
    # per block: 2 RMS norms, 5 GEMMs (q|k|v as one, o, wi_0, wi_1, wo), the attention kernel; + the final norm
    assert launches == cfg["num_layers"] * (8 if attn_impl == "kernel" else 7) + 1
    torch.set_num_threads(min(32, max(torch.get_num_threads(), 8)))
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    assert got.shape == want.shape and got.dtype == torch.bfloat16 and torch.isfinite(got.float()).all()
    err = po.rel_err(got.float().cpu(), want)
    valid = mask.bool()
    err_valid = po.rel_err(got.float().cpu()[valid], want[valid])
    print(f"T5 {name} attn={attn_impl}: last_hidden_state rel_err {err:.3e} (valid tokens {err_valid:.3e}), {launches} kernel launches")
    assert err < 1e-2 and err_valid < 1e-2

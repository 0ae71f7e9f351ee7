"""CPU tests of the T5 caption encoder (pixart_sigma_b200/t5.py): checkpoint key layout and bucketing against the installed
`transformers` (the reference's own dependency, diffusion/model/t5.py:10), and the host-side glue -- embedding, relative-position
bias, key mask, layer order, in-place residual updates, final norm -- against transformers' T5EncoderModel in fp32 with the two
kernel wrappers replaced by torch stand-ins of their contracts.  The kernels run in tests/test_t5_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pixart_oracle as po
from pixart_sigma_b200 import lib, t5

transformers = pytest.importorskip("transformers")

SMALL = dict(vocab_size=384, d_model=256, d_kv=64, d_ff=512, num_layers=3, num_heads=4)


def hf_model(cfg, seed=0):
    torch.manual_seed(seed)
    hc = transformers.T5Config(feed_forward_proj="gated-gelu", dropout_rate=0.0, **cfg)
    m = transformers.T5EncoderModel(hc).eval()
    with torch.no_grad():                                 # T5 inits layer norms to 1 and the bias table small: make them matter
        for n, p in m.named_parameters():
            if "layer_norm" in n:
                p.add_(0.2 * torch.randn_like(p))
            if "relative_attention_bias" in n:
                p.mul_(8.0)
    return m


def _standins(monkeypatch):
    def gemm(a, w, bias, out, *, epilogue=lib.EPI_BIAS, residual=None, block_n=0, **kw):
        assert not kw and bias is None and block_n in (0, 128, 192, 256)
        y = F.linear(a.float(), w.float())
        if epilogue == lib.EPI_BIAS_GELU:
            y = F.gelu(y, approximate="tanh")
        elif epilogue == lib.EPI_BIAS_RESIDUAL:
            assert residual is out and out.dtype == torch.float32
            y = residual + y
        else:
            assert epilogue == lib.EPI_BIAS
        out.copy_(y)
        return out

    def rmsnorm(x, weight, out, eps=1e-6):
        out.copy_(x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight.float())
        return out

    def t5_attn(q, k, v, out, bias, key_bias=None, *, B, H, L, scale=1.0, rel_bias=None):
        if rel_bias is not None:                      # Toeplitz form: bias[h, i, j] = rel_bias[h, j - i + L - 1]
            assert bias is None and rel_bias.shape == (H, 2 * L - 1)
            pos = torch.arange(L)
            bias = rel_bias[:, pos[None, :] - pos[:, None] + L - 1]
        q4, k4, v4 = (t.reshape(B, L, H, 64).transpose(1, 2).float() for t in (q, k, v))
        s = q4 @ k4.transpose(-1, -2) * scale + bias[None]
        if key_bias is not None:
            s = s + key_bias.view(B, 1, 1, L)
        out.view(B, L, H, 64).copy_((torch.softmax(s, -1) @ v4).transpose(1, 2))
        return out

    monkeypatch.setattr(lib, "gemm", gemm)
    monkeypatch.setattr(lib, "rmsnorm", rmsnorm)
    monkeypatch.setattr(lib, "t5_attn", t5_attn)
    monkeypatch.setattr(t5, "_need_kernels", lambda w: None)          # the product refuses CPU tensors (last test)


def test_state_dict_keys_and_buckets_equal_transformers():
    hf = hf_model(SMALL)
    m = t5.T5EncoderModel(SMALL)
    assert set(m.state_dict()) == set(hf.state_dict())
    assert all(m.state_dict()[k].shape == v.shape for k, v in hf.state_dict().items())
    missing, unexpected = m.load_state_dict(hf.state_dict())
    assert not missing and not unexpected
    assert m.encoder.embed_tokens.weight is m.shared.weight
    from transformers.models.t5.modeling_t5 import T5Attention
    pos = torch.arange(300)
    rel = pos[None] - pos[:, None]
    assert torch.equal(t5.relative_position_bucket(rel), T5Attention._relative_position_bucket(rel, True, 32, 128))
    assert t5.T5_V1_1_XXL["num_layers"] == 24 and t5.T5_V1_1_XXL["d_model"] == 4096 and t5.T5_V1_1_XXL["d_ff"] == 10240


@pytest.mark.parametrize("attn_impl", ["kernel", "torch"])
def test_forward_glue_equals_transformers_fp32(monkeypatch, attn_impl):
    _standins(monkeypatch)
    hf = hf_model(SMALL)
    m = t5.T5EncoderModel(SMALL)
    m.attn_impl = attn_impl
    m.load_state_dict(hf.state_dict())
    m = m.to(torch.bfloat16)
    hf.load_state_dict({k: v.float() for k, v in m.state_dict().items()})          # the same bf16-valued weights on both sides
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, SMALL["vocab_size"], (3, 40), generator=g)
    mask = (torch.arange(40)[None] < torch.tensor([40, 7, 23])[:, None]).long()
    got = m(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    assert got.shape == want.shape == (3, 40, 256) and got.dtype == torch.bfloat16
    assert po.rel_err(got.float(), want) < 1e-2                                    # bf16 activations between the stages
    valid = mask.bool()
    assert po.rel_err(got.float()[valid], want[valid]) < 1e-2
    got2 = m(ids, None).last_hidden_state                                           # no mask, positional call, attribute access
    with torch.no_grad():
        want2 = hf(input_ids=ids).last_hidden_state
    assert po.rel_err(got2.float(), want2) < 1e-2


def test_relative_bias_vector_is_the_toeplitz_form_of_the_position_bias():
    """`pxa_t5_attn_d64_bf16` takes the bias by offset: position_bias(L)[h, i, j] == relative_bias(L)[h, j - i + L - 1]."""
    m = t5.T5EncoderModel(SMALL)
    with torch.no_grad():
        m.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight.normal_()
    for L in (1, 7, 40, 300):
        pb, rel = m.position_bias(L), m.relative_bias(L)
        pos = torch.arange(L)
        assert rel.shape == (SMALL["num_heads"], 2 * L - 1)
        assert torch.equal(pb, rel[:, pos[None, :] - pos[:, None] + L - 1])


def test_replace_t5_model_swaps_the_model_inside_the_reference_embedder(monkeypatch):
    """diffusion/model/t5.py:12-111: the embedder keeps `.model`, `.device`; after the swap `get_text_embeddings`' call
    `self.model(input_ids=..., attention_mask=...)['last_hidden_state']` runs on this module with the checkpoint's weights."""
    _standins(monkeypatch)
    hf = hf_model(SMALL)

    class Embedder:                                   # the attributes of the reference's T5Embedder that matter here
        device = torch.device("cpu")
        torch_dtype = torch.bfloat16
        model = hf

    emb = Embedder()
    new = t5.replace_t5_model(emb)
    assert emb.model is new and isinstance(new, t5.T5EncoderModel) and new.dtype == torch.bfloat16
    assert new.cfg["d_model"] == SMALL["d_model"] and new.cfg["num_layers"] == SMALL["num_layers"]
    ids = torch.randint(0, SMALL["vocab_size"], (2, 16), generator=torch.Generator().manual_seed(9))
    mask = torch.ones_like(ids)
    got = emb.model(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask)["last_hidden_state"]
    assert po.rel_err(got.float(), want) < 2e-2        # bf16-rounded weights here vs the fp32 checkpoint there


def test_unsupported_arguments_raise_and_caches_follow_weight_updates(monkeypatch):
    _standins(monkeypatch)
    m = t5.T5EncoderModel(SMALL).to(torch.bfloat16)
    ids = torch.randint(0, SMALL["vocab_size"], (1, 8), generator=torch.Generator().manual_seed(2))
    with pytest.raises(NotImplementedError):
        m(ids, None, output_attentions=True)
    with pytest.raises(NotImplementedError):
        m(ids, None, inputs_embeds=torch.zeros(1, 8, SMALL["d_model"]))
    a = m(ids, None, return_dict=True).last_hidden_state.clone()
    sa = m.encoder.block[0].layer[0].SelfAttention
    with torch.no_grad():
        sa.q.weight.mul_(0.5)                         # an in-place update bumps the version: the stacked copy is rebuilt
    b = m(ids, None).last_hidden_state.clone()
    assert not torch.equal(a, b)
    sa.q.weight.data.mul_(2.0)                        # a `.data` write does not: the documented contract is clear_caches()
    m.clear_caches()
    c = m(ids, None).last_hidden_state
    assert torch.equal(a, c)


def test_product_refuses_to_run_without_the_kernels():
    m = t5.T5EncoderModel(SMALL)
    with pytest.raises(RuntimeError, match="sm_100a kernels only"):
        m(torch.zeros(1, 8, dtype=torch.long))


# ------------------------------------------------------------------------------------------------- committed golden vectors
@pytest.mark.parametrize("name", ["t5_small_ragged", "t5_wide_300"])
def test_forward_glue_matches_the_committed_transformers_fixture(monkeypatch, name):
    """tests/golden/t5_*.pt: outputs of transformers' T5EncoderModel (fp32) on seed-defined weights, written by
    oracle/gen_golden_t5.py in the build container -- the pin does not depend on the transformers version of the box that runs this."""
    import os
    from oracle.gen_golden_t5 import synthetic_t5_state_dict
    fix = torch.load(os.path.join(os.path.dirname(__file__), "golden", name + ".pt"))
    _standins(monkeypatch)
    m = t5.T5EncoderModel(fix["cfg"])
    m.load_state_dict(synthetic_t5_state_dict(fix["cfg"], fix["seed"]))
    m = m.to(torch.bfloat16)                                  # the synthetic weights are bf16-valued: nothing is lost
    if fix["position_bias_h0"] is not None:
        L = fix["input_ids"].shape[1]
        assert torch.allclose(m.position_bias(L)[0], fix["position_bias_h0"], atol=1e-6)
    got = m(input_ids=fix["input_ids"], attention_mask=fix["attention_mask"])["last_hidden_state"]
    assert po.rel_err(got.float(), fix["last_hidden_state"].float()) < 1e-2

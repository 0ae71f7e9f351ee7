"""CPU tests of the SDXL-VAE assembly (pixart_sigma_b200/vae.py): state-dict layout and parameter count of the public checkpoint,
and the host-side glue (block order, layouts, shortcut / residual wiring, quant convolutions, latent distribution) against
oracle/vae_oracle.py with the three kernel wrappers replaced by torch stand-ins of their documented contracts -- the kernels
themselves are checked on the GPU (tests/test_vae_gpu.py).  No CUDA needed."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pixart_oracle as po
from oracle import vae_oracle as vo
from pixart_sigma_b200 import lib, vae

SMALL = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(64, 128, 128, 128), layers_per_block=1,
             norm_num_groups=32, scaling_factor=0.13025)


def _standins(monkeypatch):
    def conv3x3_nhwc(x, w_packed, bias, out, residual=None):
        y = F.conv2d(x.float().permute(0, 3, 1, 2), w_packed.float().permute(0, 3, 1, 2), bias.float(), padding=1).permute(0, 2, 3, 1)
        out.copy_(y + residual.float() if residual is not None else y)
        return out

    def groupnorm_silu_nhwc(x, gamma, beta, out, *, groups=32, eps=1e-6, silu=True, stats_ws=None):
        y = F.group_norm(x.float().permute(0, 3, 1, 2), groups, gamma.float(), beta.float(), eps)
        out.copy_((F.silu(y) if silu else y).permute(0, 2, 3, 1))
        return out

    def gemm(a, w, bias, out, **kw):
        assert not kw
        out.copy_(F.linear(a.float(), w.float(), bias.float() if bias is not None else None))
        return out

    monkeypatch.setattr(lib, "conv3x3_nhwc", conv3x3_nhwc)
    monkeypatch.setattr(lib, "groupnorm_silu_nhwc", groupnorm_silu_nhwc)
    monkeypatch.setattr(lib, "gemm", gemm)
    monkeypatch.setattr(vae, "_need_kernels", lambda *a: None)


def _init(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return m


def test_state_dict_layout_is_the_public_sdxl_vae():
    m = vae.AutoencoderKL()
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 83_653_863          # parameter count of the public SDXL-VAE checkpoint
    for k, shape in {"decoder.conv_in.weight": (512, 4, 3, 3), "decoder.mid_block.attentions.0.to_q.weight": (512, 512),
                     "decoder.mid_block.attentions.0.to_out.0.bias": (512,), "decoder.mid_block.resnets.1.conv2.weight": (512, 512, 3, 3),
                     "decoder.up_blocks.2.resnets.0.conv_shortcut.weight": (256, 512, 1, 1),
                     "decoder.up_blocks.2.upsamplers.0.conv.weight": (256, 256, 3, 3), "decoder.up_blocks.3.resnets.2.conv1.weight": (128, 128, 3, 3),
                     "decoder.conv_norm_out.weight": (128,), "decoder.conv_out.weight": (3, 128, 3, 3),
                     "encoder.conv_in.weight": (128, 3, 3, 3), "encoder.down_blocks.1.resnets.0.conv_shortcut.weight": (256, 128, 1, 1),
                     "encoder.down_blocks.2.downsamplers.0.conv.weight": (512, 512, 3, 3), "encoder.conv_out.weight": (8, 512, 3, 3),
                     "quant_conv.weight": (8, 8, 1, 1), "post_quant_conv.weight": (4, 4, 1, 1)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd and "encoder.down_blocks.3.downsamplers.0.conv.weight" not in sd
    assert not any("_tail" in k or "_cache" in k for k in sd)
    assert m.config.scaling_factor == 0.13025


def test_decode_and_encode_glue_equal_the_oracle(monkeypatch):
    _standins(monkeypatch)
    m = _init(vae.AutoencoderKL(SMALL)).to(torch.bfloat16)
    sd = {k: v.float() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    z = torch.randn(2, 4, 8, 12, generator=g).to(torch.bfloat16)       # non-square, widths 12 .. 96: the multi-aspect buckets
    got = m.decode(z).sample
    want = vo.decode(sd, z.float())
    assert got.shape == want.shape == (2, 3, 64, 96)
    assert po.rel_err(got.float(), want) < 2e-2                        # bf16 activations between the (fp32 stand-in) stages
    img = torch.randn(1, 3, 64, 96, generator=g).to(torch.bfloat16)
    dist = m.encode(img).latent_dist
    want_m = vo.encode_moments(sd, img.float())
    assert dist.mean.shape == (1, 4, 8, 12)
    assert po.rel_err(torch.cat([dist.mean, dist.logvar], 1), torch.cat([want_m[:, :4], want_m[:, 4:].clamp(-30, 20)], 1)) < 2e-2
    s1 = dist.sample(generator=torch.Generator().manual_seed(3))
    s2 = dist.sample(generator=torch.Generator().manual_seed(3))
    assert torch.equal(s1, s2) and torch.equal(dist.mode(), dist.mean)
    assert po.rel_err((s1 - dist.mean) / dist.std, torch.randn(dist.mean.shape, generator=torch.Generator().manual_seed(3))) < 1e-5


def test_product_refuses_to_run_without_the_kernels():
    m = vae.AutoencoderKL(SMALL).to(torch.bfloat16)
    with pytest.raises(RuntimeError, match="sm_100a kernels only"):
        m.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(RuntimeError, match="sm_100a kernels only"):
        m.encode(torch.zeros(1, 3, 64, 64))

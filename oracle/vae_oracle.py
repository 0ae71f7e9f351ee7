"""TEST INFRASTRUCTURE -- fp32 CPU restatement of the SDXL-VAE (diffusers `AutoencoderKL`) around the reference's two call sites:
`vae.decode(samples / vae.config.scaling_factor).sample` (scripts/inference.py:136; loaded at :193-196) and
`vae.encode(img).latent_dist.sample()` (train_scripts/train.py:149; loaded at :85-88).

PARITY UNPINNED: diffusers is a third-party dependency of the reference (requirements.txt:2, git HEAD, un-vendored) and is neither
under /root/reference nor installed here, and the weights (`PixArt-alpha/pixart_sigma_sdxlvae_T5_diffusers/vae`) are not available
offline.  This file restates the published architecture of that checkpoint's `config.json` (block_out_channels 128/256/512/512,
layers_per_block 2, norm_num_groups 32, latent_channels 4, scaling_factor 0.13025; 83 653 863 parameters -- the count is checked in
tests/test_vae_cpu.py) with diffusers' module names, on plain functional torch ops:
  ResnetBlock2D   GN(32, eps 1e-6) -> SiLU -> conv3x3 -> GN -> SiLU -> conv3x3, + (1x1 conv_shortcut of) the input
  Attention       GN -> to_q / to_k / to_v (Linear, bias) -> ONE head of dim C: softmax(q k^T / sqrt(C)) v -> to_out.0 -> + input
  UNetMidBlock2D  resnets.0 -> attentions.0 -> resnets.1
  Upsample2D      nearest x2 -> conv3x3;    Downsample2D(padding=0)  pad (0,1,0,1) -> conv3x3 stride 2
  Decoder         conv_in -> mid_block -> up_blocks[0..3] (3 resnets each, upsampler on 0..2) -> conv_norm_out -> SiLU -> conv_out
  Encoder         conv_in -> down_blocks[0..3] (2 resnets each, downsampler on 0..2) -> mid_block -> conv_norm_out -> SiLU -> conv_out
  AutoencoderKL   decode = decoder(post_quant_conv(z));  encode = DiagonalGaussian(quant_conv(encoder(x)))
Only tests/ may import this module (the product path never does)."""
from typing import Dict

import torch
import torch.nn.functional as F

from .pixart_oracle import vae_resblock

Tensor = torch.Tensor
CONFIG = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, groups=32, latent_channels=4, eps=1e-6)


def _conv(sd, p, x, **kw):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], **kw)


def mid_attention(sd: Dict[str, Tensor], p: str, x: Tensor, groups: int = 32, eps: float = 1e-6) -> Tensor:
    B, C, H, W = x.shape
    h = F.group_norm(x, groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], eps)
    h = h.view(B, C, H * W).transpose(1, 2)                                   # (B, HW, C)
    q, k, v = (F.linear(h, sd[f"{p}.to_{n}.weight"], sd[f"{p}.to_{n}.bias"]) for n in "qkv")
    a = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1) @ v
    a = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return a.transpose(1, 2).reshape(B, C, H, W) + x


def mid_block(sd, p, x):
    x = vae_resblock(sd, p + ".resnets.0", x)
    x = mid_attention(sd, p + ".attentions.0", x)
    return vae_resblock(sd, p + ".resnets.1", x)


def _count(sd, fmt):
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def decoder(sd: Dict[str, Tensor], z: Tensor, p: str = "decoder") -> Tensor:
    x = _conv(sd, p + ".conv_in", z, padding=1)
    x = mid_block(sd, p + ".mid_block", x)
    for i in range(_count(sd, p + ".up_blocks.{}.resnets.0.conv1.weight")):         # the checkpoint decides: 4 blocks x 3 resnets
        for j in range(_count(sd, f"{p}.up_blocks.{i}.resnets." + "{}.conv1.weight")):
            x = vae_resblock(sd, f"{p}.up_blocks.{i}.resnets.{j}", x)
        if f"{p}.up_blocks.{i}.upsamplers.0.conv.weight" in sd:
            x = _conv(sd, f"{p}.up_blocks.{i}.upsamplers.0.conv", F.interpolate(x, scale_factor=2.0, mode="nearest"), padding=1)
    x = F.silu(F.group_norm(x, 32, sd[p + ".conv_norm_out.weight"], sd[p + ".conv_norm_out.bias"], 1e-6))
    return _conv(sd, p + ".conv_out", x, padding=1)


def encoder(sd: Dict[str, Tensor], img: Tensor, p: str = "encoder") -> Tensor:
    x = _conv(sd, p + ".conv_in", img, padding=1)
    for i in range(_count(sd, p + ".down_blocks.{}.resnets.0.conv1.weight")):       # 4 blocks x 2 resnets
        for j in range(_count(sd, f"{p}.down_blocks.{i}.resnets." + "{}.conv1.weight")):
            x = vae_resblock(sd, f"{p}.down_blocks.{i}.resnets.{j}", x)
        if f"{p}.down_blocks.{i}.downsamplers.0.conv.weight" in sd:
            x = _conv(sd, f"{p}.down_blocks.{i}.downsamplers.0.conv", F.pad(x, (0, 1, 0, 1)), stride=2)
    x = mid_block(sd, p + ".mid_block", x)
    x = F.silu(F.group_norm(x, 32, sd[p + ".conv_norm_out.weight"], sd[p + ".conv_norm_out.bias"], 1e-6))
    return _conv(sd, p + ".conv_out", x, padding=1)


def decode(sd, z):
    return decoder(sd, _conv(sd, "post_quant_conv", z))


def encode_moments(sd, img):
    return _conv(sd, "quant_conv", encoder(sd, img))

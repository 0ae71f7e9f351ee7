"""SDXL-VAE decoder ResBlock on the sm_100a implicit-GEMM 3x3 convolution.

The reference decodes latents with diffusers' `AutoencoderKL` (scripts/inference.py:136, 193-196); diffusers is not
vendored in the reference and not installed here, so this module mirrors the public `ResnetBlock2D` of the SDXL-VAE
decoder -- parameter names `norm1, conv1, norm2, conv2, conv_shortcut`, GroupNorm(32, eps 1e-6) -> SiLU -> Conv3x3 twice,
1x1 shortcut when the channel count changes -- with both 3x3 convolutions (68 % of the decoder FLOPs) on
`pxa_conv3x3_nhwc_bf16` and GroupNorm + SiLU in front of each as one NHWC kernel pair (`pxa_groupnorm_silu_nhwc_bf16`:
statistics, apply), so a block is 2 convolutions + 2 x 2 light passes with no layout copies in between.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib


class DecoderResBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: Optional[int] = None, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self._packed = {}

    def _pack(self, conv: nn.Conv2d) -> torch.Tensor:
        """[Cout, Cin, 3, 3] -> [Cout, 3, 3, Cin] (tap-major K), cached until the parameter changes."""
        key = (id(conv), conv.weight._version, conv.weight.data_ptr())
        hit = self._packed.get(id(conv))
        if hit is None or hit[0] != key:
            hit = (key, conv.weight.detach().permute(0, 2, 3, 1).contiguous())
            self._packed[id(conv)] = hit
        return hit[1]

    @torch.no_grad()
    def forward(self, x: torch.Tensor, temb=None) -> torch.Tensor:
        """x (B, C, H, W) bf16 on a B200 -> (B, Cout, H, W) bf16 (channels_last memory format)."""
        if not x.is_cuda or self.conv1.weight.dtype != torch.bfloat16:
            raise RuntimeError("DecoderResBlock runs on the sm_100a kernels only: CUDA tensors and bf16 weights required")
        B, Cc, H, W = x.shape
        x = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)     # no copy when the caller already is NHWC
        x_nhwc = x.permute(0, 2, 3, 1)                                          # contiguous (B, H, W, C) view
        h = lib.groupnorm_silu_nhwc(x_nhwc, self.norm1.weight, self.norm1.bias, torch.empty_like(x_nhwc),
                                    groups=self.norm1.num_groups, eps=self.norm1.eps)
        h1 = torch.empty(B, H, W, self.out_channels, dtype=torch.bfloat16, device=x.device)
        lib.conv3x3_nhwc(h, self._pack(self.conv1), self.conv1.bias, h1)
        h2 = lib.groupnorm_silu_nhwc(h1, self.norm2.weight, self.norm2.bias, torch.empty_like(h1),
                                     groups=self.norm2.num_groups, eps=self.norm2.eps)
        if self.conv_shortcut is not None:
            sc = torch.empty(B * H * W, self.out_channels, dtype=torch.bfloat16, device=x.device)
            lib.gemm(x_nhwc.reshape(B * H * W, Cc), self.conv_shortcut.weight.view(self.out_channels, Cc),
                     self.conv_shortcut.bias, sc)
            sc = sc.view(B, H, W, self.out_channels)
        else:
            sc = x_nhwc
        out = torch.empty(B, H, W, self.out_channels, dtype=torch.bfloat16, device=x.device)
        lib.conv3x3_nhwc(h2, self._pack(self.conv2), self.conv2.bias, out, residual=sc.contiguous())
        return out.permute(0, 3, 1, 2)


class UpsampleConv(nn.Module):
    """Nearest-neighbour x2 upsample followed by a 3x3 convolution (diffusers `Upsample2D` of the SDXL-VAE decoder up
    blocks; 2.8 TFLOP of the 10.5 TFLOP decode at 1024px).  The interpolation is data movement (PyTorch); the
    convolution runs on `pxa_conv3x3_nhwc_bf16`."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)
        self._packed = None

    @torch.no_grad()
    def forward(self, x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        if not x.is_cuda or self.conv.weight.dtype != torch.bfloat16:
            raise RuntimeError("UpsampleConv runs on the sm_100a kernels only: CUDA tensors and bf16 weights required")
        key = (self.conv.weight._version, self.conv.weight.data_ptr())
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, self.conv.weight.detach().permute(0, 2, 3, 1).contiguous())
        B, Cc, H, W = x.shape
        up = F.interpolate(x.to(torch.bfloat16), scale_factor=2.0, mode="nearest")
        up = up.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)           # (B, 2H, 2W, C)
        out = torch.empty(B, 2 * H, 2 * W, Cc, dtype=torch.bfloat16, device=x.device)
        lib.conv3x3_nhwc(up, self._packed[1], self.conv.bias, out)
        return out.permute(0, 3, 1, 2)


def patch_diffusers_decoder(vae) -> int:
    """Route every ResnetBlock2D of a diffusers AutoencoderKL *decoder* through `DecoderResBlock` (shares the
    parameters).  Returns the number of blocks patched.  (diffusers is not installed in the build container; this is
    the binding a user of scripts/inference.py adds after loading `vae`.)"""
    n = 0
    for mod in vae.decoder.modules():
        if type(mod).__name__ == "ResnetBlock2D" and getattr(mod, "time_emb_proj", None) is None:
            blk = DecoderResBlock(mod.in_channels, mod.out_channels, mod.norm1.num_groups, mod.norm1.eps)
            blk.norm1, blk.conv1, blk.norm2, blk.conv2 = mod.norm1, mod.conv1, mod.norm2, mod.conv2
            blk.conv_shortcut = getattr(mod, "conv_shortcut", None)
            scale = getattr(mod, "output_scale_factor", 1.0)
            mod.forward = (lambda b, s: (lambda x, temb=None, *a, **k: b(x) / s))(blk, scale)
            n += 1
        elif type(mod).__name__ == "Upsample2D" and getattr(mod, "conv", None) is not None and mod.conv.kernel_size == (3, 3):
            up = UpsampleConv(mod.conv.in_channels)
            up.conv = mod.conv
            mod.forward = (lambda u: (lambda x, *a, **k: u(x)))(up)
            n += 1
    return n

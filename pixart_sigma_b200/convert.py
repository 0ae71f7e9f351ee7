"""Checkpoint key mapping between the reference `.pth` layout (the layout `PixArtMS.state_dict()` keeps) and the
diffusers `Transformer2DModel` / `PixArtTransformer2DModel` layout.

`to_diffusers` follows `tools/convert_pixart_to_diffusers.py:29-155` key for key (fused qkv / kv projections are split
into to_q / to_k / to_v; `pos_embed`, `y_embedder.y_embedding` and EMA bookkeeping are not part of the diffusers
transformer); `from_diffusers` is its inverse, so HF-format PixArt-Sigma weights load into `pixart_sigma_b200.PixArtMS`
without the reference package.  Pure tensor re-keying (views / `torch.cat`), no kernels involved.
"""
from __future__ import annotations

from typing import Dict

import torch

_GLOBAL = [  # (diffusers key, reference key)
    ("pos_embed.proj.weight", "x_embedder.proj.weight"), ("pos_embed.proj.bias", "x_embedder.proj.bias"),
    ("caption_projection.linear_1.weight", "y_embedder.y_proj.fc1.weight"),
    ("caption_projection.linear_1.bias", "y_embedder.y_proj.fc1.bias"),
    ("caption_projection.linear_2.weight", "y_embedder.y_proj.fc2.weight"),
    ("caption_projection.linear_2.bias", "y_embedder.y_proj.fc2.bias"),
    ("adaln_single.emb.timestep_embedder.linear_1.weight", "t_embedder.mlp.0.weight"),
    ("adaln_single.emb.timestep_embedder.linear_1.bias", "t_embedder.mlp.0.bias"),
    ("adaln_single.emb.timestep_embedder.linear_2.weight", "t_embedder.mlp.2.weight"),
    ("adaln_single.emb.timestep_embedder.linear_2.bias", "t_embedder.mlp.2.bias"),
    ("adaln_single.linear.weight", "t_block.1.weight"), ("adaln_single.linear.bias", "t_block.1.bias"),
    ("proj_out.weight", "final_layer.linear.weight"), ("proj_out.bias", "final_layer.linear.bias"),
    ("scale_shift_table", "final_layer.scale_shift_table"),
]
_MICRO = [
    ("adaln_single.emb.resolution_embedder.linear_1.weight", "csize_embedder.mlp.0.weight"),
    ("adaln_single.emb.resolution_embedder.linear_1.bias", "csize_embedder.mlp.0.bias"),
    ("adaln_single.emb.resolution_embedder.linear_2.weight", "csize_embedder.mlp.2.weight"),
    ("adaln_single.emb.resolution_embedder.linear_2.bias", "csize_embedder.mlp.2.bias"),
    ("adaln_single.emb.aspect_ratio_embedder.linear_1.weight", "ar_embedder.mlp.0.weight"),
    ("adaln_single.emb.aspect_ratio_embedder.linear_1.bias", "ar_embedder.mlp.0.bias"),
    ("adaln_single.emb.aspect_ratio_embedder.linear_2.weight", "ar_embedder.mlp.2.weight"),
    ("adaln_single.emb.aspect_ratio_embedder.linear_2.bias", "ar_embedder.mlp.2.bias"),
]
_BLOCK = [  # per block, un-split entries: (diffusers suffix, reference suffix)
    ("scale_shift_table", "scale_shift_table"),
    ("attn1.to_out.0.weight", "attn.proj.weight"), ("attn1.to_out.0.bias", "attn.proj.bias"),
    ("ff.net.0.proj.weight", "mlp.fc1.weight"), ("ff.net.0.proj.bias", "mlp.fc1.bias"),
    ("ff.net.2.weight", "mlp.fc2.weight"), ("ff.net.2.bias", "mlp.fc2.bias"),
    ("attn2.to_q.weight", "cross_attn.q_linear.weight"), ("attn2.to_q.bias", "cross_attn.q_linear.bias"),
    ("attn2.to_out.0.weight", "cross_attn.proj.weight"), ("attn2.to_out.0.bias", "cross_attn.proj.bias"),
]
_QK_NORM = [("attn1.q_norm.weight", "attn.q_norm.weight"), ("attn1.q_norm.bias", "attn.q_norm.bias"),
            ("attn1.k_norm.weight", "attn.k_norm.weight"), ("attn1.k_norm.bias", "attn.k_norm.bias")]


def _depth_of(sd: Dict[str, torch.Tensor], prefix: str) -> int:
    idx = [int(k.split(".")[1]) for k in sd if k.startswith(prefix + ".")]
    return max(idx) + 1 if idx else 0


def to_diffusers(state_dict: Dict[str, torch.Tensor], micro_condition: bool = False, qk_norm: bool = False) -> Dict[str, torch.Tensor]:
    """Reference / pixart_sigma_b200 `state_dict` -> diffusers transformer keys (convert_pixart_to_diffusers.py:29-155)."""
    sd = state_dict.get("state_dict", state_dict)
    out = {d: sd[r] for d, r in _GLOBAL + (_MICRO if micro_condition else [])}
    for i in range(_depth_of(sd, "blocks")):
        src, dst = f"blocks.{i}.", f"transformer_blocks.{i}."
        for d, r in _BLOCK + (_QK_NORM if qk_norm else []):
            out[dst + d] = sd[src + r]
        for kind in ("weight", "bias"):
            q, k, v = torch.chunk(sd[src + f"attn.qkv.{kind}"], 3, dim=0)              # :92-99
            out[dst + f"attn1.to_q.{kind}"], out[dst + f"attn1.to_k.{kind}"], out[dst + f"attn1.to_v.{kind}"] = q, k, v
            k2, v2 = torch.chunk(sd[src + f"cross_attn.kv_linear.{kind}"], 2, dim=0)   # :137-138
            out[dst + f"attn2.to_k.{kind}"], out[dst + f"attn2.to_v.{kind}"] = k2, v2
    return out


def from_diffusers(diffusers_sd: Dict[str, torch.Tensor], micro_condition: bool = False, qk_norm: bool = False) -> Dict[str, torch.Tensor]:
    """Inverse of `to_diffusers`: diffusers transformer keys -> the `.pth` layout `PixArtMS.load_state_dict` takes
    (`pos_embed` and `y_embedder.y_embedding` are not in the diffusers file: load with strict=False)."""
    dsd = diffusers_sd
    out = {r: dsd[d] for d, r in _GLOBAL + (_MICRO if micro_condition else [])}
    for i in range(_depth_of(dsd, "transformer_blocks")):
        src, dst = f"transformer_blocks.{i}.", f"blocks.{i}."
        for d, r in _BLOCK + (_QK_NORM if qk_norm else []):
            out[dst + r] = dsd[src + d]
        for kind in ("weight", "bias"):
            out[dst + f"attn.qkv.{kind}"] = torch.cat([dsd[src + f"attn1.to_{n}.{kind}"] for n in "qkv"], dim=0)
            out[dst + f"cross_attn.kv_linear.{kind}"] = torch.cat([dsd[src + f"attn2.to_{n}.{kind}"] for n in "kv"], dim=0)
    return out

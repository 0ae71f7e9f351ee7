"""T5-v1.1-XXL caption encoder on the sm_100a kernels.

The reference encodes prompts with transformers' `T5EncoderModel` (diffusion/model/t5.py:12,107-110 `self.model(input_ids=...,
attention_mask=...)['last_hidden_state']`; scripts/inference.py:80, train_scripts/train.py:166): 24 layers, d_model 4096, 64 heads
of 64, gated-GELU feed-forward of 10240, 4.7 B parameters, 300 tokens per caption.  This module keeps that call and the
checkpoint's state-dict keys (`shared.weight`, `encoder.block.i.layer.0.SelfAttention.{q,k,v,o}.weight`,
`encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight`, `encoder.block.i.layer.{0,1}.layer_norm.weight`,
`encoder.block.i.layer.1.DenseReluDense.{wi_0,wi_1,wo}.weight`, `encoder.final_layer_norm.weight`) and runs

  * the seven Linear layers of a block -- 99 % of the FLOPs (2 x 4.7 G per token) -- on `pxa_gemm_bf16` (tcgen05): q / k / v, the
    attention output projection accumulated IN PLACE into the fp32 residual stream (fp32 residual epilogue), wi_0 with the
    GELU(tanh) epilogue (`gelu_new`), wi_1, and wo again in place into the stream;
  * both T5LayerNorms (RMS norm, no mean, no bias) on `pxa_rmsnorm_bf16`, reading the fp32 stream;
  * the attention core on `pxa_t5_attn_d64_bf16` (tcgen05; T5 adds a learned [heads, 300, 300] relative-position bias and the key
    mask to UNSCALED logits at head_dim 64, so the DiT's head_dim-72 flash kernel does not apply; with <= 384 keys the whole logit
    row sits in TMEM: no key loop, no online softmax), q | k | v coming from ONE GEMM against row-stacked weights;
  * what stays PyTorch: the embedding gather, the bias table look-up (once per forward) and the gate product gelu(wi_0 x) * (wi_1 x).

Deliberate deviation, towards the fp32 model: the residual stream between layers is fp32 (transformers keeps it in the checkpoint
dtype; the reference loads T5 in fp16 / bf16).  At 300 tokens per caption one forward is bound by streaming the 9.4 GB of bf16
weights once (1.45 ms at the measured 6.5 TB/s) against 2.8 TFLOP of GEMMs (2.1 ms at the sustained bf16 peak)."""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import lib

T5_V1_1_XXL = dict(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                   relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6)


class _RMSNorm(nn.Module):          # transformers T5LayerNorm
    def __init__(self, d: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _SelfAttention(nn.Module):    # transformers T5Attention (encoder: bidirectional, no cache)
    def __init__(self, cfg, has_bias: bool):
        super().__init__()
        inner = cfg["num_heads"] * cfg["d_kv"]
        self.q, self.k, self.v = (nn.Linear(cfg["d_model"], inner, bias=False) for _ in range(3))
        self.o = nn.Linear(inner, cfg["d_model"], bias=False)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(cfg["relative_attention_num_buckets"], cfg["num_heads"])


class _LayerSelfAttention(nn.Module):
    def __init__(self, cfg, has_bias: bool):
        super().__init__()
        self.SelfAttention = _SelfAttention(cfg, has_bias)
        self.layer_norm = _RMSNorm(cfg["d_model"], cfg["layer_norm_epsilon"])


class _DenseGated(nn.Module):       # transformers T5DenseGatedActDense, act = gelu_new
    def __init__(self, cfg):
        super().__init__()
        self.wi_0 = nn.Linear(cfg["d_model"], cfg["d_ff"], bias=False)
        self.wi_1 = nn.Linear(cfg["d_model"], cfg["d_ff"], bias=False)
        self.wo = nn.Linear(cfg["d_ff"], cfg["d_model"], bias=False)


class _LayerFF(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.DenseReluDense = _DenseGated(cfg)
        self.layer_norm = _RMSNorm(cfg["d_model"], cfg["layer_norm_epsilon"])


class _Block(nn.Module):
    def __init__(self, cfg, has_bias: bool):
        super().__init__()
        self.layer = nn.ModuleList([_LayerSelfAttention(cfg, has_bias), _LayerFF(cfg)])


class _Stack(nn.Module):
    def __init__(self, cfg, embed: nn.Embedding):
        super().__init__()
        self.embed_tokens = embed
        self.block = nn.ModuleList([_Block(cfg, i == 0) for i in range(cfg["num_layers"])])
        self.final_layer_norm = _RMSNorm(cfg["d_model"], cfg["layer_norm_epsilon"])


class EncoderOutput(dict):
    """`model(...)['last_hidden_state']` (diffusion/model/t5.py:110) and `.last_hidden_state` both work."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        return list(self.values())[k] if isinstance(k, int) else super().__getitem__(k)


def relative_position_bucket(rel: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """Bidirectional T5 bucketing of (key position - query position): half of the buckets per sign; within a sign the first half
    are exact offsets, the rest logarithmic up to max_distance (Mesh-TensorFlow `_relative_position_bucket`, as in transformers)."""
    nb = num_buckets // 2
    out = (rel > 0).to(torch.long) * nb
    n = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    return out + torch.where(n < max_exact, n, large)


def _res_bn(n: int) -> int:
    return 256 if n % 256 == 0 else (192 if n % 192 == 0 else 128)


def _need_kernels(w: torch.Tensor) -> None:
    if not w.is_cuda or w.dtype != torch.bfloat16:
        raise RuntimeError("T5EncoderModel runs on the sm_100a kernels only: move it to a B200 and cast it to bfloat16 "
                           "(model.to(torch.bfloat16); fp16 checkpoints: cast at load)")


class T5EncoderModel(nn.Module):
    def __init__(self, config: Optional[dict] = None):
        super().__init__()
        cfg = dict(T5_V1_1_XXL)
        cfg.update(config or {})
        self.cfg = cfg
        self.shared = nn.Embedding(cfg["vocab_size"], cfg["d_model"])
        self.encoder = _Stack(cfg, self.shared)                 # `encoder.embed_tokens.weight` is tied to `shared.weight`
        self._ws = {}
        self._qkv_w = {}
        # "torch": the attention core through PyTorch (fp32 matmuls); env PXA_T5_ATTN for A/B runs of bench.py
        self.attn_impl = os.environ.get("PXA_T5_ATTN", "kernel")

    @property
    def dtype(self):
        return self.shared.weight.dtype

    @property
    def device(self):
        return self.shared.weight.device

    def _buf(self, name, shape, dtype):
        t = self._ws.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype or t.device != self.device:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._ws[name] = t
        return t

    def _stacked_qkv(self, li: int, sa) -> torch.Tensor:
        """[q; k; v] weights of block li row-stacked (3 * inner, d_model) for one GEMM: a derived bf16 copy (2.4 GB for the XXL model,
        next to 9.4 GB of weights), rebuilt when a parameter's version or storage changes."""
        ps = (sa.q.weight, sa.k.weight, sa.v.weight)
        key = tuple((p._version, p.data_ptr()) for p in ps)
        hit = self._qkv_w.get(li)
        if hit is None or hit[0] != key:
            hit = (key, torch.cat([p.detach() for p in ps], 0).contiguous())
            self._qkv_w[li] = hit
        return hit[1]

    def clear_caches(self) -> None:
        """Drop the derived copies (row-stacked q | k | v weights) and the activation workspace.  The stacked weights are keyed on the
        parameters' version counters and storage, which in-place writes through `.data` bypass: call this after such an update."""
        self._qkv_w.clear()
        self._ws.clear()

    def position_bias(self, L: int) -> torch.Tensor:
        """(H, L, L) fp32: relative_attention_bias[bucket(j - i), h] -- layer 0 owns the table, every layer uses it."""
        cfg = self.cfg
        pos = torch.arange(L, device=self.device)
        bucket = relative_position_bucket(pos[None, :] - pos[:, None], cfg["relative_attention_num_buckets"],
                                          cfg["relative_attention_max_distance"])
        table = self.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight
        return table.float()[bucket].permute(2, 0, 1).contiguous()

    def relative_bias(self, L: int) -> torch.Tensor:
        """(H, 2L-1) fp32: the bias by key-minus-query offset d = j - i, entry d + L - 1 (position_bias(L)[h, i, j] = this[h, j - i + L - 1])."""
        cfg = self.cfg
        d = torch.arange(-(L - 1), L, device=self.device)
        bucket = relative_position_bucket(d, cfg["relative_attention_num_buckets"], cfg["relative_attention_max_distance"])
        table = self.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight
        return table.float()[bucket].t().contiguous()

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, **kwargs) -> EncoderOutput:
        _need_kernels(self.encoder.block[0].layer[0].SelfAttention.q.weight)
        unsupported = [k for k, v in kwargs.items() if k != "return_dict" and not (v is None or v is False)]
        if unsupported:         # transformers' other inputs / outputs (inputs_embeds, output_attentions, ...): the reference passes none
            raise NotImplementedError(f"T5EncoderModel.forward: unsupported arguments {unsupported}")
        cfg = self.cfg
        B, L = input_ids.shape
        D, H, dk, F_ = cfg["d_model"], cfg["num_heads"], cfg["d_kv"], cfg["d_ff"]
        inner, M, bf = H * dk, B * L, torch.bfloat16
        input_ids = input_ids.to(self.device)
        x32 = self._buf("x32", (M, D), torch.float32)
        x32.copy_(self.shared.weight[input_ids.reshape(-1)])
        key_bias = None
        if attention_mask is not None:                                                # additive key mask, as transformers' extended mask
            keep = attention_mask.to(self.device).to(torch.float32).view(B, L)
            key_bias = ((1.0 - keep) * torch.finfo(torch.float32).min).contiguous()
        native = self.attn_impl == "kernel"
        if native and L > lib.T5_ATTN_MAX_KEYS:
            raise RuntimeError(f"pxa_t5_attn_d64_bf16 holds the whole logit row in TMEM: at most {lib.T5_ATTN_MAX_KEYS} tokens per "
                               f"caption (got {L}; the reference uses 120 / 300).  model.attn_impl = 'torch' runs longer inputs.")
        if native:
            rel = self.relative_bias(L)                                               # (H, 2L-1) fp32: the bias is Toeplitz
        else:
            pos_bias = self.position_bias(L)                                          # (H, L, L) fp32
            bias = pos_bias[None] if key_bias is None else pos_bias[None] + key_bias.view(B, 1, 1, L)
        xn, qkv = self._buf("xn", (M, D), bf), self._buf("qkv", (M, 3 * inner), bf)
        q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
        ao, h0, h1 = self._buf("ao", (M, inner), bf), self._buf("h0", (M, F_), bf), self._buf("h1", (M, F_), bf)
        for li, blk in enumerate(self.encoder.block):
            at, ff = blk.layer[0], blk.layer[1]
            sa, dense = at.SelfAttention, ff.DenseReluDense
            # (1) x += o( softmax(q k^T + bias) v ),  q | k | v = ONE GEMM of rmsnorm(x) against the row-stacked weights
            #     (no 1/sqrt(d): T5 folds it into the initialisation)
            lib.rmsnorm(x32, at.layer_norm.weight, xn, eps=at.layer_norm.variance_epsilon)
            lib.gemm(xn, self._stacked_qkv(li, sa), None, qkv)
            if native:
                lib.t5_attn(q, k, v, ao, None, key_bias, B=B, H=H, L=L, scale=1.0, rel_bias=rel)
            else:                                                                     # attn_impl = "torch": A/B, and captions > 384 tokens
                q4, k4, v4 = (t.reshape(B, L, H, dk).transpose(1, 2).float() for t in (q, k, v))
                p = torch.softmax(q4 @ k4.transpose(-1, -2) + bias, dim=-1)
                ao.view(B, L, H, dk).copy_((p @ v4).transpose(1, 2))
            # block_n = 256 divides d_model 4096 (the auto choice for the in-place update, 192-column tiles on the CTA pair, is tuned
            # for the DiT's N = 1152 and would leave a partial last column tile here)
            lib.gemm(ao, sa.o.weight, None, x32, epilogue=lib.EPI_BIAS_RESIDUAL, residual=x32, block_n=_res_bn(D))
            # (2) x += wo( gelu_new(wi_0 n) * wi_1 n ),  n = rmsnorm(x)
            lib.rmsnorm(x32, ff.layer_norm.weight, xn, eps=ff.layer_norm.variance_epsilon)
            lib.gemm(xn, dense.wi_0.weight, None, h0, epilogue=lib.EPI_BIAS_GELU)
            lib.gemm(xn, dense.wi_1.weight, None, h1)
            h0.mul_(h1)
            lib.gemm(h0, dense.wo.weight, None, x32, epilogue=lib.EPI_BIAS_RESIDUAL, residual=x32, block_n=_res_bn(D))
        fin = self.encoder.final_layer_norm
        out = torch.empty(M, D, dtype=bf, device=self.device)
        lib.rmsnorm(x32, fin.weight, out, eps=fin.variance_epsilon)
        return EncoderOutput(last_hidden_state=out.view(B, L, D))


def replace_t5_model(embedder, config: Optional[dict] = None):
    """Swap the transformers model inside the reference's `T5Embedder` (diffusion/model/t5.py:12-88: `self.model`, `self.device`,
    `self.torch_dtype`) for this one: same weights (`state_dict` keys are identical), cast to bf16, on the embedder's device.
    `embedder.get_text_embeddings(texts)` (t5.py:90-111) then runs unchanged.  Returns the new model."""
    old = embedder.model
    cfg = dict(config or {})
    hf_cfg = getattr(old, "config", None)
    if hf_cfg is not None and not cfg:                      # take the geometry from the loaded checkpoint
        cfg = {k: getattr(hf_cfg, k) for k in T5_V1_1_XXL if hasattr(hf_cfg, k)}
    new = T5EncoderModel(cfg)
    missing, unexpected = new.load_state_dict(old.state_dict(), strict=False)
    if unexpected or [k for k in missing if k != "encoder.embed_tokens.weight"]:
        raise RuntimeError(f"T5 checkpoint layout mismatch: missing {missing}, unexpected {unexpected}")
    embedder.model = new.to(torch.bfloat16).to(embedder.device).eval()
    return embedder.model

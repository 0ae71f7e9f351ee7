"""Attention forward alone at the c3 shapes: self-attention (8 x 16 heads x 4096 x 4096), the packed var-len cross-attention
(8 x 16 x 4096 queries over <= 300 text tokens per sample) and the c4 KV-compressed shape (2 x 16 x 16384 x 4096).  CUDA events,
20 repetitions after 3 warm-ups.  PXA_ATTN_VARIANT selects the kernel variant (2 = one CTA per item, 4 / unset = persistent)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

dev, H, D = "cuda", 16, 72
C = H * D


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B, Nq, Nk in ((8, 4096, 4096), (2, 16384, 4096)):
    q = torch.randn(B * Nq, C, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * Nk, 2, C, device=dev).to(torch.bfloat16)
    o = torch.empty(B * Nq, C, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: lib.flash_attn(q, kv[:, 0], kv[:, 1], o, B=B, H=H, Nq=Nq, Nk=Nk, kv_rows=B * Nk, q_strides=(C, D),
                                       k_strides=(2 * C, D), v_strides=(2 * C, D)))
    print(f"self  B={B} Nq={Nq} Nk={Nk}: {ms * 1e3:8.1f} us ({4.0 * B * H * Nq * Nk * D / ms / 1e9:6.0f} TFLOP/s)", flush=True)
    # fp32_attention grade P V (P as bf16 hi + lo, PxaAttnArgs.p_precision = 1): 1.5 x the tensor work, same exp2 work
    ms = timeit(lambda: lib.flash_attn(q, kv[:, 0], kv[:, 1], o, B=B, H=H, Nq=Nq, Nk=Nk, kv_rows=B * Nk, q_strides=(C, D),
                                       k_strides=(2 * C, D), v_strides=(2 * C, D), fp32_p=True))
    print(f"self  B={B} Nq={Nq} Nk={Nk} fp32_p: {ms * 1e3:8.1f} us ({4.0 * B * H * Nq * Nk * D / ms / 1e9:6.0f} model TFLOP/s)", flush=True)

B, Nq = 8, 4096
for lens in ([300] * 8, [300, 77, 120, 256, 33, 300, 180, 64]):
    kv_len = torch.tensor(lens, dtype=torch.int32, device=dev)
    kv_off = torch.tensor([sum(lens[:i]) for i in range(B)], dtype=torch.int32, device=dev)
    rows = sum(lens)
    q = torch.randn(B * Nq, C, device=dev).to(torch.bfloat16)
    kv = torch.randn(rows, 2, C, device=dev).to(torch.bfloat16)
    o = torch.empty(B * Nq, C, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: lib.flash_attn(q, kv[:, 0], kv[:, 1], o, B=B, H=H, Nq=Nq, Nk=max(lens), kv_rows=rows, kv_len=kv_len, kv_off=kv_off,
                                       q_strides=(C, D), k_strides=(2 * C, D), v_strides=(2 * C, D)))
    fl = 4.0 * H * Nq * D * sum(lens)
    print(f"cross B={B} Nq={Nq} lens={lens}: {ms * 1e3:8.1f} us ({fl / ms / 1e9:6.0f} TFLOP/s)", flush=True)

"""The two attention kernels added at the end of round 2, alone: the fp32_attention mode of the head_dim-72 flash kernel
(P as bf16 hi + lo, c3 shape 8 x 16 x 4096 x 4096) and the T5 attention kernel (4 captions x 64 heads x 300 tokens; dense bias table
vs the Toeplitz vector in smem).  CUDA events, 20 repetitions after 3 warm-ups; `--ncu`: exactly one launch of each (for
`ncu --set full -c 3`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

dev = "cuda"
once = "--ncu" in sys.argv


def timeit(fn, reps=20):
    if once:
        fn()
        torch.cuda.synchronize()
        return float("nan")
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


B, H, D, N = 8, 16, 72, 4096
C = H * D
q = torch.randn(B * N, C, device=dev).to(torch.bfloat16)
kv = torch.randn(B * N, 2, C, device=dev).to(torch.bfloat16)
o = torch.empty(B * N, C, dtype=torch.bfloat16, device=dev)
ms = timeit(lambda: lib.flash_attn(q, kv[:, 0], kv[:, 1], o, B=B, H=H, Nq=N, Nk=N, kv_rows=B * N, q_strides=(C, D), k_strides=(2 * C, D),
                                   v_strides=(2 * C, D), fp32_p=True))
print(f"flash_attn_d72 fp32_p  B={B} N={N}: {ms * 1e3:8.1f} us ({4.0 * B * H * N * N * D / ms / 1e9:6.0f} model TFLOP/s)", flush=True)

B, H, L = 4, 64, 300
inner = H * 64
qkv = (torch.randn(B * L, 3 * inner, device=dev) * 0.5).to(torch.bfloat16)
out = torch.empty(B * L, inner, dtype=torch.bfloat16, device=dev)
rel = torch.randn(H, 2 * L - 1, device=dev)
pos = torch.arange(L, device=dev)
dense = rel[:, pos[None, :] - pos[:, None] + L - 1].contiguous()
lens = torch.tensor([300, 77, 150, 240], device=dev)
kb = ((torch.arange(L, device=dev)[None] >= lens[:, None]).float() * torch.finfo(torch.float32).min).contiguous()
qq, kk, vv = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
fl = 4.0 * B * H * L * L * 64
ms = timeit(lambda: lib.t5_attn(qq, kk, vv, out, None, kb, B=B, H=H, L=L, rel_bias=rel))
print(f"t5_attn_d64 toeplitz   B={B} H={H} L={L}: {ms * 1e3:8.1f} us ({fl / ms / 1e9:6.1f} TFLOP/s)", flush=True)
if not once:
    ms = timeit(lambda: lib.t5_attn(qq, kk, vv, out, dense, kb, B=B, H=H, L=L))
    print(f"t5_attn_d64 dense bias B={B} H={H} L={L}: {ms * 1e3:8.1f} us ({fl / ms / 1e9:6.1f} TFLOP/s)", flush=True)

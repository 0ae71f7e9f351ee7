"""Item-level cycle trace of CTA 0 of the PERSISTENT attention forward (variant 4 + debug_trace, PXA_ITRACE in attn_sm100.cu):
where the time goes at the boundary between two work items.  usage: python tools/attn_itrace.py [Nk [variant 4|5]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

B, H, N = 8, 16, 4096
Nk = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
VARIANT = 5          # item-level stamps are recorded by variant 5 (persistent, no stagger) + debug_trace
g = torch.Generator().manual_seed(0)
q = torch.randn(B * N, H * 72, generator=g).to(torch.bfloat16).cuda()
kv = torch.randn(B * Nk, 2, H * 72, generator=g).to(torch.bfloat16).cuda()
out = torch.empty(B * N, H * 72, dtype=torch.bfloat16, device="cuda")
trace = torch.zeros(18, 512, dtype=torch.int64, device="cuda")
kw = dict(B=B, H=H, Nq=N, Nk=Nk, kv_rows=B * Nk, q_strides=(H * 72, 72), k_strides=(2 * H * 72, 72), v_strides=(2 * H * 72, 72), variant=VARIANT)
for _ in range(2):
    lib.flash_attn(q, kv[:, 0], kv[:, 1], out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.flash_attn(q, kv[:, 0], kv[:, 1], out, debug_trace=trace, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"Nk={Nk} variant={VARIANT}: kernel {ms:.3f} ms, {4 * B * H * N * Nk * 72 / ms / 1e9:.1f} TFLOP/s")
tr = trace.cpu()
t0 = int(tr[0, 0])
n_items = int((tr[0, 0::8] > 0).sum())
print(f"CTA 0 walked {n_items} items; cycles relative to its first item start")
print("softmax warp 0: [item start | first S in regs | last P published | o_full | output stored]   MMA: [item start | Q,K ready | prologue issued | o_full committed]   TMA: [item start | q_empty | loads issued]")
for i in range(n_items):
    sm = [int(v) - t0 for v in tr[0, 8 * i:8 * i + 5]]
    mm = [int(v) - t0 for v in tr[16, 8 * i:8 * i + 4]]
    tm = [int(v) - t0 for v in tr[17, 8 * i:8 * i + 3]]
    if i < 4:
        print(f"  item {i:2d}: softmax {sm}  mma {mm}  tma {tm}")
per = [int(tr[0, 8 * (i + 1)]) - int(tr[0, 8 * i]) for i in range(n_items - 1)]
loop = [int(tr[0, 8 * i + 2]) - int(tr[0, 8 * i + 1]) for i in range(n_items)]
epi = [int(tr[0, 8 * i + 4]) - int(tr[0, 8 * i + 2]) for i in range(n_items)]
gap = [int(tr[0, 8 * (i + 1) + 1]) - int(tr[0, 8 * i + 4]) for i in range(n_items - 1)]
print("item period:", per)
print("softmax loop (first S -> last P):", loop)
print("epilogue (last P -> stored):", epi)
print("next item's first S after the stores:", gap)

"""Mlp branch at the c3 shape (M = 8 x 4096): fc1+GELU and fc2+gated-residual as two launches vs pxa_mlp_fused_bf16 (one
persistent kernel, hidden ring in L2) for a few (group, lag, ring, fc2 K-split) choices.  CUDA events, L2 flush (256 MB write) before every
repetition.  usage: python tools/mlp_fused_bench.py [M [group,lag,ring,ksplit ...]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
C, HID, dev = 1152, 4608, "cuda"
xn = torch.randn(M, C, device=dev).to(torch.bfloat16)
w1 = (torch.randn(HID, C, device=dev) * C ** -0.5).to(torch.bfloat16)
b1 = torch.randn(HID, device=dev).to(torch.bfloat16)
w2 = (torch.randn(C, HID, device=dev) * HID ** -0.5).to(torch.bfloat16)
b2 = torch.randn(C, device=dev).to(torch.bfloat16)
x32 = torch.randn(M, C, device=dev)
gate = torch.randn(8, 6, C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hid = torch.empty(M, HID, dtype=torch.bfloat16, device=dev)
kw = dict(gate=gate[:, 5], gate_batch_stride=6 * C, rows_per_batch=4096)


def timeit(step):
    for _ in range(3):
        step()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def two():
    lib.gemm(xn, w1, b1, hid, epilogue=lib.EPI_BIAS_GELU)
    lib.gemm(hid, w2, b2, x32, epilogue=lib.EPI_BIAS_RESIDUAL, residual=x32, reverse_tiles=True, **kw)


ms = timeit(two)
print(f"two launches (fc1 + GELU, fc2 + gated residual, L2 chaining): {ms * 1e3:8.1f} us ({4.0 * M * C * HID / ms / 1e9:6.0f} TFLOP/s)", flush=True)
cfgs = ((8, 1, 3, 1), (8, 1, 3, 3), (4, 1, 3, 3), (4, 2, 4, 3), (2, 4, 6, 3), (1, 8, 10, 3), (1, 4, 6, 3), (2, 2, 4, 3), (1, 12, 14, 3), (2, 4, 6, 2))
if len(sys.argv) > 2:
    cfgs = tuple(tuple(int(v) for v in c.split(",")) for c in sys.argv[2:])
for group, lag, ring, ks in cfgs:
    hws, fws = lib.mlp_fused_workspace(M, HID, dev, group=group, ring=ring)
    msf = timeit(lambda: lib.mlp_fused(xn, w1, b1, w2, b2, x32, hidden_ws=hws, flags_ws=fws, group=group, ring=ring, lag=lag, k_splits=ks, **kw))
    print(f"one persistent kernel, group {group} lag {lag} ring {ring} fc2 k-splits {ks} ({hws.numel() * 2 / 2 ** 20:5.1f} MiB hidden ring): "
          f"{msf * 1e3:8.1f} us ({4.0 * M * C * HID / msf / 1e9:6.0f} TFLOP/s)", flush=True)

"""MLP (fc1 + GELU -> fc2 + gated residual) at the c3 shape (M = 8 x 4096 rows) as ONE pair of launches versus G row-groups
that reuse one small hidden buffer, so that the hidden activations of a group are produced and consumed inside the 126 MB
L2 and never reach HBM (the L2-resident variant).  CUDA events, an L2 flush (256 MB write) before every
timed repetition so that each variant starts cold like it does inside a model step.
usage: python tools/mlp_group_bench.py [M]"""
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


def run(groups: int):
    rows = M // groups
    hid = torch.empty(rows, HID, dtype=torch.bfloat16, device=dev)

    def step():
        for g in range(groups):
            sl = slice(g * rows, (g + 1) * rows)
            lib.gemm(xn[sl], w1, b1, hid, epilogue=lib.EPI_BIAS_GELU)
            # the gate view must start at the group's first sample: rows_per_batch rows share a gate row
            b0 = (g * rows) // 4096
            lib.gemm(hid, w2, b2, x32[sl], epilogue=lib.EPI_BIAS_RESIDUAL, residual=x32[sl], gate=gate[b0:, 5],
                     gate_batch_stride=6 * C, rows_per_batch=4096 if rows >= 4096 else rows)
    for _ in range(3):
        step()
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2]


base = None
for G in (1, 2, 4, 8, 16):
    if M % G or (M // G) % 256:
        continue
    ms = run(G)
    base = base or ms
    tf = 4.0 * M * C * HID / ms / 1e9
    print(f"M={M} groups={G:2d} rows/group={M // G:6d} hidden buffer {M // G * HID * 2 / 2**20:6.1f} MiB: {ms * 1e3:8.1f} us "
          f"({tf:6.0f} TFLOP/s, {ms / base:5.3f} x the 2-launch time)", flush=True)

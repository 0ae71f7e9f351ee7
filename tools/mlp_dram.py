"""DRAM traffic of the Mlp branch at the c3 shape, two launches vs pxa_mlp_fused_bf16: run under
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv
Every variant is launched twice behind a 256 MB L2 flush; the second launch is the one to read.
usage: python tools/mlp_dram.py [group lag ring ksplit]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

M, C, HID, dev = 32768, 1152, 4608, "cuda"
group, lag, ring, ks = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (lib.MLP_GROUP, lib.MLP_LAG, lib.MLP_RING, 0)
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
hws, fws = lib.mlp_fused_workspace(M, HID, dev, group=group, ring=ring)
for _ in range(2):
    flush.zero_()
    lib.gemm(xn, w1, b1, hid, epilogue=lib.EPI_BIAS_GELU)
    lib.gemm(hid, w2, b2, x32, epilogue=lib.EPI_BIAS_RESIDUAL, residual=x32, reverse_tiles=True, **kw)
for _ in range(2):
    flush.zero_()
    lib.mlp_fused(xn, w1, b1, w2, b2, x32, hidden_ws=hws, flags_ws=fws, group=group, ring=ring, lag=lag, k_splits=ks, **kw)
torch.cuda.synchronize()
print("ok", group, lag, ring, ks)

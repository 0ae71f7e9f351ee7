"""GEMM micro-benchmark at the model's shapes (M = 8 x 4096 tokens): TFLOP/s per (N, K, epilogue, block_n)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib

M = 32768
dev = "cuda"
def run(N, K, epi, bn, pair=1, iters=20):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    kw = {}
    if epi == lib.EPI_BIAS_RESIDUAL:
        out = torch.randn(M, N, device=dev)
        kw = dict(residual=out, gate=torch.randn(8, 6, N, device=dev)[:, 2], gate_batch_stride=6 * N, rows_per_batch=4096)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.gemm(a, w, b, out, epilogue=epi, block_n=bn, cta_pair=pair, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.gemm(a, w, b, out, epilogue=epi, block_n=bn, cta_pair=pair, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2 * M * N * K / ms / 1e9

names = {0: "bias", 1: "gelu", 2: "resid"}
for (N, K, epi) in [(3456, 1152, 0), (1152, 1152, 0), (1152, 1152, 2), (4608, 1152, 1), (1152, 4608, 2), (2304, 1152, 0)]:
    for bn in (128, 192, 256):
        if N % bn and bn != 192:
            pass
        for pair in (1, 2):
            ms, tf = run(N, K, epi, bn, pair)
            print(f"N={N:5d} K={K:5d} {names[epi]:5s} BN={bn:3d} {'pair' if pair == 2 else '1cta'}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)

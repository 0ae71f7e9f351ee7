"""Round-2 GEMM micro-benchmark at M = 8 x 4096 rows: the residual epilogue variants (register-staged vs TMA-streamed, single
CTA vs CTA pair, with / without the fused-LayerNorm by-products) and the LN consumer epilogues next to the plain ones.
CUDA events, 20 repetitions after 3 warm-ups (back to back: L2-warm weights, activations >> L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

M, dev = 32768, "cuda"
RPB = 4096


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def resid(K, pair, res_epi, aux, byprod, bn=0, gate=True):
    N = 1152
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    x = torch.randn(M, N, device=dev)
    g = torch.randn(8, 6, N, device=dev)
    xa = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if (aux or byprod) else None
    st = torch.empty(M, 8, 2, device=dev) if byprod else None
    sc = (1 + 0.1 * torch.randn(8, 2, N, device=dev)) if byprod else None
    kw = dict(epilogue=lib.EPI_BIAS_RESIDUAL, residual=x, rows_per_batch=RPB, cta_pair=pair, res_epilogue=res_epi, out_aux=xa,
              block_n=bn)
    if gate:
        kw.update(gate=g[:, 2], gate_batch_stride=6 * N)
    if byprod:
        kw.update(aux_scale=sc[:, 0], aux_scale_batch_stride=2 * N, row_stats_out=st)
    ms = timeit(lambda: lib.gemm(a, w, b, x, **kw))
    return ms, 2.0 * M * N * K / ms / 1e9


def plain(N, K, epi, ln, epi_warps=0):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if ln:
        st = torch.rand(M, 8, 2, device=dev)
        st[:, :, 1] += 1000.0
        u, v = torch.randn(8, N, device=dev), torch.randn(8, N, device=dev)
        fn = lambda: lib.gemm(a, w, None, out, epilogue=epi, rows_per_batch=RPB, ln_stats=st, ln_u=u, ln_v=v,
                              ln_uv_batch_stride=N, ln_dim=K, epi_warps=epi_warps)
    else:
        fn = lambda: lib.gemm(a, w, b, out, epilogue=epi, epi_warps=epi_warps)
    ms = timeit(fn)
    return ms, 2.0 * M * N * K / ms / 1e9


print("== residual GEMMs, N = 1152 (us, TFLOP/s)")
for K in (1152, 4608):
    for name, kw in [("1cta tma  reduce (no aux)      ", dict(pair=1, res_epi=0, aux=False, byprod=False)),
                     ("1cta tma  rmw + aux            ", dict(pair=1, res_epi=0, aux=True, byprod=False)),
                     ("1cta tma  rmw + aux*s + stats  ", dict(pair=1, res_epi=0, aux=True, byprod=True)),
                     ("pair regs (round 1)            ", dict(pair=2, res_epi=1, aux=False, byprod=False)),
                     ("pair regs + aux                ", dict(pair=2, res_epi=1, aux=True, byprod=False)),
                     ("pair tma  reduce (no aux)      ", dict(pair=2, res_epi=2, aux=False, byprod=False)),
                     ("pair tma  rmw + aux            ", dict(pair=2, res_epi=2, aux=True, byprod=False)),
                     ("pair tma  rmw + aux*s + stats  ", dict(pair=2, res_epi=2, aux=True, byprod=True))]:
        for bn in ((192, 256) if kw["pair"] == 2 else (256,)):
            try:
                ms, tf = resid(K, bn=bn, **kw)
                print(f"K={K:5d} {name} BN={bn}: {ms * 1e3:8.1f} us {tf:7.1f} TF", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"K={K:5d} {name} BN={bn}: FAILED {e}", flush=True)
print("== bias / GELU GEMMs vs their fused-LayerNorm forms, K = 1152")
for N, e0, e1 in ((3456, lib.EPI_BIAS, lib.EPI_LN_BIAS), (4608, lib.EPI_BIAS_GELU, lib.EPI_LN_BIAS_GELU)):
    for ln, epi in ((False, e0), (True, e1)):
        for ew in (4, 8):
            ms, tf = plain(N, 1152, epi, ln, ew)
            print(f"N={N} {'LN-fused epilogue' if ln else 'plain epilogue   '} {ew} epilogue warps: {ms * 1e3:8.1f} us {tf:7.1f} TF", flush=True)
for N in (1152, 2304):
    for ew in (4, 8):
        ms, tf = plain(N, 1152, lib.EPI_BIAS, False, ew)
        print(f"N={N} plain epilogue    {ew} epilogue warps: {ms * 1e3:8.1f} us {tf:7.1f} TF", flush=True)

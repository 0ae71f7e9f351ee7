"""Cycle trace of CTA 0's residual-epilogue issuer thread (debug aid, see PXA_GTRACE in gemm_sm100.cu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib
M, N, K = 32768, 1152, int(sys.argv[1]) if len(sys.argv) > 1 else 1152
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
b = torch.randn(N, device="cuda").to(torch.bfloat16)
x = torch.randn(M, N, device="cuda")
aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
gate = torch.randn(8, 6, N, device="cuda")
tr = torch.zeros(4096, dtype=torch.int64, device="cuda")
kw = dict(epilogue=lib.EPI_BIAS_RESIDUAL, residual=x, gate=gate[:, 2], gate_batch_stride=6 * N, rows_per_batch=4096, out_aux=aux, cta_pair=1)
for _ in range(3):
    lib.gemm(a, w, b, x, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lib.gemm(a, w, b, x, debug_trace=tr, **kw); e1.record(); torch.cuda.synchronize()
print(f"K={K}: {e0.elapsed_time(e1)*1e3:.1f} us")
t = tr.cpu().tolist(); t0 = t[0]
# layout per tile: 2 stamps + 6 chunks x 5 stamps = 32
for tile in range(0, 6):
    base = tile * 32
    print(f"tile {tile}: wait_acc={t[base+1]-t[base]}  start={t[base]-t0}")
    for c in range(6):
        s = t[base + 2 + 5 * c: base + 2 + 5 * c + 5]
        prev = t[base + 1] if c == 0 else t[base + 2 + 5 * (c - 1) + 4]
        print(f"   chunk {c}: ldtm={s[0]-prev} wait_res={s[1]-s[0]} compute={s[2]-s[1]} drain={s[3]-s[2]} barrier={s[4]-s[3]}")

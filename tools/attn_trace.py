"""Cycle trace of one CTA of the attention kernel under a full-grid launch (debug aid, see PXA_TRACE in attn_sm100.cu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib

B, H, N = 8, 16, 4096
VARIANT = int(sys.argv[1]) if len(sys.argv) > 1 else 2     # 2: one CTA per item (item 0 of CTA 0); 4: persistent (CTA 0's third item)
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B * N, 3, H, 72, generator=g)).to(torch.bfloat16).cuda()
out = torch.empty(B * N, H * 72, dtype=torch.bfloat16, device="cuda")
trace = torch.zeros(18, 512, dtype=torch.int64, device="cuda")
kw = dict(B=B, H=H, Nq=N, Nk=N, kv_rows=B * N, q_strides=(3 * H * 72, 72), k_strides=(3 * H * 72, 72), v_strides=(3 * H * 72, 72),
          variant=VARIANT)
for _ in range(2):
    lib.flash_attn(qkv[:, 0], qkv[:, 1], qkv[:, 2], out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.flash_attn(qkv[:, 0], qkv[:, 1], qkv[:, 2], out, debug_trace=trace, **kw)
e1.record()
torch.cuda.synchronize()
print(f"variant {VARIANT}: kernel {e0.elapsed_time(e1):.3f} ms, {4*B*H*N*N*72/e0.elapsed_time(e1)/1e9:.1f} TFLOP/s")
tr = trace.cpu()
t0 = int(tr[tr > 0].min())
names = {0: "A w0", 4: "B w0"}
print("softmax stamps per 64-key sub-block n: [wait S | S ready | S in regs | row max | exp done | P in TMEM | P published]")
for w, nm in names.items():
    for n in list(range(0, 8)) + [20, 21, 40, 41]:
        row = [int(v) - t0 for v in tr[w, 7 * n:7 * n + 7]]
        d = [row[i + 1] - row[i] for i in range(6)]
        print(f"  {nm} n={n}: {row}  d(waitS,ld,max,exp,st,arrive)={d}")
per = [int(tr[0, 7 * (n + 1)]) - int(tr[0, 7 * n]) for n in range(8, 56)]
print("cycles per sub-block (softmax warp A w0 loop period), n=8..55:", per)
print("mean cycles per 128 keys:", 2 * sum(per) / len(per))
lag = [int(tr[4, 7 * n + 3]) - int(tr[0, 7 * n + 3]) for n in (4, 12, 20, 28, 36, 44, 52)]
print("tile B behind tile A at 'row max known' (cycles), n=4,12,..,52:", lag)

names6 = ["waitS", "ld", "max", "exp", "st", "arrive"]
for w, nm in names.items():
    acc = [0] * 6
    for n in range(8, 56):
        row = [int(v) for v in tr[w, 7 * n:7 * n + 7]]
        for i in range(6):
            acc[i] += row[i + 1] - row[i]
    print(f"  {nm} mean section cycles n=8..55:", {k: round(v / 48) for k, v in zip(names6, acc)})
mm = tr[16]
pw = [int(mm[4 * n + 1]) - int(mm[4 * n]) for n in range(8, 56)] + [int(mm[4 * n + 3]) - int(mm[4 * n + 2]) for n in range(8, 56)]
print("MMA thread: mean wait for P per tile visit (cycles):", round(sum(pw) / len(pw)))

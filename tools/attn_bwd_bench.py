"""Time the attention forward (with lse) and backward kernels alone at the training shapes (CUDA events, 10 reps after
3 warm-ups).  usage: python tools/attn_bwd_bench.py [B H N]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from pixart_sigma_b200 import lib  # noqa: E402

B, H, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 16, 4096)
D, C = 72, H * 72
dev = "cuda"
qkv = torch.randn(B * N, 3, H, D, device=dev).to(torch.bfloat16)
d_o = torch.randn(B * N, C, device=dev).to(torch.bfloat16)
o = torch.empty(B * N, C, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B, H, N, dtype=torch.float32, device=dev)
dqkv = torch.empty_like(qkv)
st = (3 * C, D)
kw = dict(B=B, H=H, Nq=N, Nk=N, kv_rows=B * N, q_strides=st, k_strides=st, v_strides=st, scale=D ** -0.5)


def fwd():
    lib.flash_attn(qkv[:, 0], qkv[:, 1], qkv[:, 2], o, lse=lse, **kw)


def bwd():
    lib.flash_attn_bwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], o, d_o, lse, dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], dq_strides=st,
                       dk_strides=st, dv_strides=st, **kw)


def timeit(fn, reps=10):
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


f = 4.0 * B * H * N * N * D
tf, tb = timeit(fwd), timeit(bwd)
print(f"B={B} H={H} N={N}: fwd {tf:.3f} ms ({f / tf / 1e9:.0f} TFLOP/s)  bwd {tb:.3f} ms "
      f"(model 2.5x fwd FLOPs: {2.5 * f / tb / 1e9:.0f} TFLOP/s; executed 3.5x: {3.5 * f / tb / 1e9:.0f} TFLOP/s)")

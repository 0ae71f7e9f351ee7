"""pxa_ln_modulate at the c3 shape (M = 32768 rows of 1152 fp32 -> bf16): grid-stride launch with row prefetch (auto) vs one row
per warp (max_ctas = M / 8) and other grid sizes.  CUDA events, 256 MB L2 flush before every repetition."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixart_sigma_b200 import lib  # noqa: E402

M, C, dev = 32768, 1152, "cuda"
x = torch.randn(M, C, device=dev)
mod = torch.randn(8, 6, C, device=dev)
out = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for ctas in (M // 8, 0, 148, 296, 444, 592, 1184):
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.ln_modulate(x, mod[:, 3], mod[:, 4], out, mod_batch_stride=6 * C, rows_per_batch=4096, max_ctas=ctas)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(f"max_ctas {ctas:5d}{' (auto)' if ctas == 0 else '       '}: {ms * 1e3:7.1f} us  ({M * C * 6 / ms / 1e6:6.0f} GB/s algorithmic)", flush=True)

"""Profiling target for ncu: a depth-D PixArt-Sigma-XL/2-width model at the c5 training shape (1024px, 4 images,
4096 tokens, fp32 master weights, activation checkpointing), `--warm` untimed training steps then ONE step bracketed by
cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  usage: python tools/train_profile.py [--depth 4] [--warm 1]"""
import argparse
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from pixart_sigma_b200 import build_model  # noqa: E402
from pixart_sigma_b200.parallel import GradBucketReducer  # noqa: E402
from pixart_sigma_b200.training import IDDPMLoss, train_step  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--imgs", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with torch.device(dev):
    m = build_model(dict(type="PixArtMS", depth=a.depth, input_size=128, pe_interpolation=2.0, model_max_length=300),
                    use_grad_checkpoint=True)
    for blk in m.blocks:
        torch.nn.init.normal_(blk.cross_attn.proj.weight, std=0.02)
    torch.nn.init.normal_(m.final_layer.linear.weight, std=0.02)
m = m.float().train()
red = GradBucketReducer(m)
g = torch.Generator().manual_seed(1)
x = (torch.randn(a.imgs, 4, 128, 128, generator=g) * 0.5).to(dev)
y = torch.randn(a.imgs, 1, 300, 4096, generator=g).to(torch.bfloat16).to(dev)
mask = (torch.arange(300)[None] < torch.randint(8, 301, (a.imgs, 1), generator=g)).to(torch.int16).view(a.imgs, 1, 1, 300).to(dev)
t = torch.randint(0, 1000, (a.imgs,), generator=g).to(dev)
noise = torch.randn(a.imgs, 4, 128, 128, generator=g).to(dev)
loss = IDDPMLoss()
for _ in range(a.warm):
    red.zero_grad()
    train_step(m, loss, x, t, y, mask, noise=noise, reducer=red)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
red.zero_grad()
lv = train_step(m, loss, x, t, y, mask, noise=noise, reducer=red)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(lv))

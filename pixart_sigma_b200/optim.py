"""Optimizer step on the flat gradient buckets.

The reference trains with `AdamW(lr=2e-5, weight_decay=3e-2, eps=1e-10)` built by `diffusion/utils/optimizer.py:236-245` from
`configs/PixArt_xl2_internal.py:48` (CAME is the other registered choice).  torch's optimizers walk 437 parameter tensors;
here the parameters of a bucket of `parallel.GradBucketReducer` are views of ONE flat fp32 buffer, like their gradients, and a
step is one `pxa_adamw_flat` launch per bucket (29 launches, 28 B of HBM traffic per parameter): torch.optim.AdamW's update
rule exactly (decoupled weight decay, bias-corrected first / second moments), optionally with the gradient un-scaling of a
loss scaler folded in.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import lib
from .parallel import GradBucketReducer


class FlatAdamW:
    """AdamW over the buckets of a `GradBucketReducer` whose parameters have been flattened (`reducer.flatten_params()`)."""

    def __init__(self, reducer: GradBucketReducer, lr: float = 2e-5, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-10,
                 weight_decay: float = 3e-2):
        self.reducer = reducer
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        reducer.flatten_params()
        self.state = [{"exp_avg": torch.zeros_like(b["flat_param"]), "exp_avg_sq": torch.zeros_like(b["flat_param"])}
                      for b in reducer.buckets]

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.reducer.zero_grad()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        """One update of every parameter from the (already all-reduced) flat gradients.  grad_scale: multiplied into the
        gradients first (1 / loss scale)."""
        self.step_count += 1
        self.reducer.check_views()
        for b, st in zip(self.reducer.buckets, self.state):
            if b["flat_param"].dtype != torch.float32:
                raise RuntimeError("FlatAdamW updates fp32 master parameters (use model.float())")
            lib.adamw_flat(b["flat_param"], b["flat"], st["exp_avg"], st["exp_avg_sq"], step=self.step_count, lr=self.lr,
                           betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, grad_scale=grad_scale)

    def state_dict(self) -> dict:
        return {"step": self.step_count, "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                "state": [{k: v.clone() for k, v in st.items()} for st in self.state]}

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd["step"])
        for st, src in zip(self.state, sd["state"]):
            for k in st:
                st[k].copy_(src[k])

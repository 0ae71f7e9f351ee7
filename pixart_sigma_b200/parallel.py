"""Multi-GPU host logic: batch-sharded inference (independent replicas) and the data-parallel
gradient all-reduce of training (`GradBucketReducer`).

Images are independent, so rank r denoises images [lo, hi) with its own full weight replica and the whole sampling
loop runs with NO collective; the only communication is an optional gather of the final latents.  CFG pairs stay on
one rank (the sampler concatenates [uncond, cond] per image).  One process per GPU, `torch.distributed` for plumbing
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: Optional[int] = None, world: Optional[int] = None):
    """Slice every (n_items, ...) tensor to this rank's images; None entries pass through."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    n = next(t.shape[0] for t in tensors if t is not None)
    lo, hi = shard_bounds(n, rank, world)
    return [None if t is None else t[lo:hi] for t in tensors]


def gather_batch(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Inverse of shard_batch for results: all ranks receive the full (n_items, ...) tensor in image order.
    Shards may be ragged (n_items % world != 0): they are padded to the largest shard for the all_gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((cap,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def rank_seed(base_seed: int, image_index: int) -> int:
    """Per-image noise seed that does not depend on how images are sharded, so an N-GPU run reproduces the 1-GPU run."""
    return (base_seed * 1_000_003 + image_index) % (2 ** 31 - 1)


# ------------------------------------------------------------------------------------------------- training: DDP all-reduce
class GradBucketReducer:
    """Bucketed, backward-overlapped gradient all-reduce (mean over ranks) -- what accelerate -> torch DDP does for the
    reference (`train_scripts/train.py:180,486`), laid out for this model: ONE bucket per PixArtMSBlock (21.3 M
    parameters, 85 MB fp32) plus one for everything else, each a flat fp32 buffer that the parameters' `.grad` are views
    of.  Blocks finish their backward in reverse order, so bucket i's all-reduce (NCCL over NVLink / NVSwitch, issued
    asynchronously the moment its last gradient has been accumulated) runs behind the backward of blocks i-1 .. 0;
    `finish()` waits for the stragglers.  No copy in or out of the buckets, no per-parameter collectives.

    The collective is the only data-path exchange of the training step: nothing else is sharded.
    """

    def __init__(self, model: torch.nn.Module, group=None, bucket_of=None, compress: Optional[str] = None):
        """compress: None = all-reduce the fp32 buckets as they are (what torch DDP does for the reference, 2.44 GB per step
        at XL/2); "bf16" = each bucket travels as a bf16 copy (1.22 GB; torch DDP's bf16_compress_hook semantics: cast,
        sum in bf16 on the wire, cast back) -- SURVEY.md 8e / 8f.3."""
        if compress not in (None, "bf16"):
            raise ValueError("compress must be None or 'bf16'")
        self.compress = compress
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if bucket_of is None:
            def bucket_of(name: str) -> str:
                parts = name.split(".")
                return ".".join(parts[:2]) if parts[0] == "blocks" and len(parts) > 2 else "rest"
        groups = {}
        for n, p in params:
            groups.setdefault(bucket_of(n), []).append(p)
        self.buckets = []
        for key, ps in groups.items():
            total = (sum(p.numel() for p in ps) + 3) // 4 * 4               # padded to 16 bytes (flat optimizer kernels)
            flat = torch.zeros(total, dtype=ps[0].dtype, device=ps[0].device)
            off = 0
            for p in ps:
                assert p.dtype == flat.dtype and p.device == flat.device, "one dtype / device per bucket"
                p.grad = flat[off: off + p.numel()].view_as(p)            # gradient accumulates straight into the bucket
                off += p.numel()
            wire = torch.empty(total, dtype=torch.bfloat16, device=flat.device) if compress == "bf16" else None
            self.buckets.append({"key": key, "flat": flat, "params": ps, "pending": 0, "work": None, "wire": wire})
        self._hooks = []
        self.overlap = True        # False: no collective from the hooks (CUDA-graph capture); finish() reduces every bucket
        for b in self.buckets:
            for p in b["params"]:
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, b=b: self._ready(b, _p)))
                # backward kernels that accumulate straight into p.grad (autograd.py) hand autograd no gradient; whether its
                # post-accumulate hook still fires for such a parameter depends on the torch version, so they signal as well
                # and `_ready` counts every parameter once per step
                p._pxa_grad_ready = (lambda b=b, p=p: self._ready(b, p))
        self.start()

    def flatten_params(self) -> None:
        """Make every parameter's storage a view of one flat buffer per bucket (same layout as the gradient buckets), so that
        an optimizer can update a whole bucket in one launch (`optim.FlatAdamW`).  Values are preserved; idempotent."""
        for b in self.buckets:
            if b.get("flat_param") is not None:
                continue
            fp = torch.zeros_like(b["flat"])
            off = 0
            for p in b["params"]:
                view = fp[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
            b["flat_param"] = fp

    def check_views(self) -> None:
        """Every parameter's `.grad` must still be its view into the flat bucket.  `optimizer.zero_grad()` /
        `model.zero_grad()` with the torch default set_to_none=True (or any code that rebinds `.grad`) silently detaches
        them; the buckets would then be reduced stale while the optimizer sees other tensors.  Detached gradients are
        re-attached when empty (None); a foreign tensor is an error."""
        for b in self.buckets:
            lo = b["flat"].data_ptr()
            hi = lo + b["flat"].numel() * b["flat"].element_size()
            off = 0
            for p in b["params"]:
                if p.grad is None:
                    p.grad = b["flat"][off: off + p.numel()].view_as(p)
                    p.grad.zero_()
                elif not (lo <= p.grad.data_ptr() < hi):
                    raise RuntimeError("GradBucketReducer: a parameter's .grad no longer aliases its gradient bucket (was .grad "
                                       "re-assigned?); zero gradients with reducer.zero_grad() or zero_grad(set_to_none=False)")
                off += p.numel()

    def start(self) -> None:
        """Arm the buckets for one backward pass (gradients are NOT zeroed: call `zero_grad()` between optimizer steps)."""
        self.check_views()
        for b in self.buckets:
            b["pending"], b["work"], b["seen"] = len(b["params"]), None, set()

    def zero_grad(self) -> None:
        for b in self.buckets:
            b["flat"].zero_()

    def _launch(self, b) -> None:
        b["flat"].div_(self.world)
        buf = b["flat"]
        if b["wire"] is not None:
            buf = b["wire"].copy_(b["flat"])
        b["work"] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _ready(self, b, p) -> None:
        if id(p) in b["seen"]:
            return
        b["seen"].add(id(p))
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1 and self.overlap:
            self._launch(b)

    def finish(self) -> None:
        """Block until every bucket has been reduced; buckets whose parameters got no gradient this step (unused
        parameters) are reduced here so that all ranks issue the same collectives."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b["work"] is None:
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
            if b["wire"] is not None:
                b["flat"].copy_(b["wire"])

    def grad_bytes(self) -> int:
        """Bytes handed to the collective per step (per rank)."""
        return sum((b["wire"] if b["wire"] is not None else b["flat"]).numel() *
                   (b["wire"] if b["wire"] is not None else b["flat"]).element_size() for b in self.buckets)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for b in self.buckets:
            for p in b["params"]:
                if hasattr(p, "_pxa_grad_ready"):
                    del p._pxa_grad_ready

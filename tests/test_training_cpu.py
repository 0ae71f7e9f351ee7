"""CPU tests (-m "not gpu") of the training host logic:
  * `IDDPMLoss` against the loss terms of the unmodified reference (tests/golden/train_loss_only.pt);
  * the oracle's autograd (oracle.forward_grad + the loss) against the reference's gradients (train_d2_b2.pt) -- this is
    what pins the gradient oracle the GPU tests compare the backward kernels with;
  * `GradBucketReducer`: world-size-2 gloo run reproduces the single-process gradient of the concatenated batch.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pixart_oracle as po
from oracle.gen_golden_train import fingerprint_positions, train_inputs
from pixart_sigma_b200.parallel import GradBucketReducer
from pixart_sigma_b200.training import IDDPMLoss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_iddpm_loss_matches_reference_terms(golden_dir):
    fix = torch.load(os.path.join(golden_dir, "train_loss_only.pt"))
    loss = IDDPMLoss()
    assert torch.allclose(loss.q_sample(fix["x0"], fix["t"], fix["noise"]), fix["x_t"], rtol=1e-6, atol=1e-6)
    terms = loss.training_losses(lambda x, timestep, **kw: fix["fake"], fix["x0"], fix["t"], {}, noise=fix["noise"])
    for k in ("mse", "vb", "loss"):
        assert torch.allclose(terms[k], fix["terms"][k], rtol=1e-5, atol=1e-6), (k, terms[k], fix["terms"][k])


def test_iddpm_loss_gradient_does_not_reach_the_mean_through_vb():
    """The vb term sees a detached epsilon (gaussian_diffusion.py:808): d vb / d eps == 0."""
    loss = IDDPMLoss()
    g = torch.Generator().manual_seed(0)
    out = torch.randn(2, 8, 8, 8, generator=g, requires_grad=True)
    x0, noise = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    terms = loss.training_losses(lambda x, timestep, **kw: out, x0, torch.tensor([3, 700]), {}, noise=noise)
    terms["vb"].sum().backward()
    assert float(out.grad[:, :4].abs().max()) == 0.0 and float(out.grad[:, 4:].abs().max()) > 0.0


@pytest.mark.slow
@pytest.mark.parametrize("name", ["train_d2_b2", "train_d2_kvconv"])
def test_oracle_autograd_matches_reference_gradients(golden_dir, name):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    cfg = po.OracleConfig(**fix["cfg"])
    sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in po.synthetic_state_dict(cfg, seed=0).items()}
    x0, t, y, mask, noise = train_inputs(cfg, fix["batch"], tuple(fix["hw"]), fix["t"], fix["lens"])
    model = lambda x, timestep, **kw: po.forward_grad(sd, cfg, x, timestep.float(), kw["y"], mask=kw["mask"])
    terms = IDDPMLoss().training_losses(model, x0, t, dict(y=y, mask=mask, data_info=None), noise=noise)
    for k in ("mse", "vb", "loss"):
        assert torch.allclose(terms[k], fix["terms"][k], rtol=1e-4, atol=1e-5), k
    terms["loss"].mean().backward()
    worst = 0.0
    for name, ref in fix["grads"].items():
        g = sd[name].grad.flatten()
        vals = g[fingerprint_positions(g.numel())]
        worst = max(worst, abs(float(g.norm()) - float(ref["norm"])) / float(ref["norm"]), po.rel_err(vals, ref["vals"]))
    assert worst < 2e-3, worst            # fp32 vs fp32, different summation orders


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Tiny(torch.nn.Module):
    """Same naming scheme as the model (blocks.<i>.* + others) so the default bucketing rule is exercised."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.blocks = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Tanh(), torch.nn.Linear(8, 8))
                                           for _ in range(3)])
        self.head = torch.nn.Linear(8, 2)
        self.unused = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x):
        for b in self.blocks:
            x = x + b(x)
        return self.head(x)


def _ddp_worker(rank, world, port, out_dir, compress=None):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _Tiny()
        red = GradBucketReducer(m, compress=compress)
        assert len(red.buckets) == 4 and {b["key"] for b in red.buckets} == {"blocks.0", "blocks.1", "blocks.2", "rest"}
        x = torch.randn(8, 8, generator=torch.Generator().manual_seed(9))
        lo, hi = rank * 4, rank * 4 + 4
        for step in range(2):                                    # two steps: buckets re-arm, zero_grad keeps the views
            red.zero_grad()
            red.start()
            m(x[lo:hi]).pow(2).mean().backward()
            red.finish()
        torch.save({n: p.grad.clone() for n, p in m.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
        assert all(p.grad.data_ptr() >= b["flat"].data_ptr() for b in red.buckets for p in b["params"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("compress", [None, "bf16"])
def test_grad_bucket_reducer_gloo_world2(tmp_path, compress):
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path), compress), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    m = _Tiny()
    x = torch.randn(8, 8, generator=torch.Generator().manual_seed(9))
    m(x).pow(2).mean().backward()                                # mean over the global batch == mean of the rank means
    for n, p in m.named_parameters():
        want = p.grad if p.grad is not None else torch.zeros_like(p)
        tol = dict(rtol=1e-5, atol=1e-7) if compress is None else dict(rtol=2e-2, atol=1e-4)    # bf16 on the wire
        assert torch.allclose(g0[n], want, **tol), n
        assert torch.equal(g0[n], g1[n]), n


def test_grad_bucket_reducer_single_process_is_a_noop_reduce():
    m = _Tiny()
    red = GradBucketReducer(m)
    m(torch.ones(2, 8)).sum().backward()
    red.finish()
    # one flat fp32 buffer per bucket, each padded to a multiple of 4 elements (16-byte granules of the flat optimizer kernel)
    assert red.grad_bytes() == sum((sum(p.numel() for p in b["params"]) + 3) // 4 * 4 for b in red.buckets) * 4
    assert sum(p.numel() for p in m.parameters()) * 4 <= red.grad_bytes() < sum(p.numel() for p in m.parameters()) * 4 + 16 * len(red.buckets)
    assert float(m.head.weight.grad.abs().sum()) > 0


def test_grad_bucket_reducer_survives_zero_grad_set_to_none_and_rejects_foreign_grads():
    """torch's default zero_grad(set_to_none=True) detaches the bucket views; start() re-attaches empty gradients
    and refuses a gradient tensor that lives outside its bucket."""
    m = _Tiny()
    red = GradBucketReducer(m)
    m(torch.ones(2, 8)).sum().backward()
    m.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in m.parameters())
    red.start()
    for b in red.buckets:
        lo, hi = b["flat"].data_ptr(), b["flat"].data_ptr() + b["flat"].numel() * 4
        assert all(p.grad is not None and lo <= p.grad.data_ptr() < hi and float(p.grad.abs().sum()) == 0 for p in b["params"])
    m(torch.ones(2, 8)).sum().backward()
    red.finish()
    assert float(sum(b["flat"].abs().sum() for b in red.buckets)) > 0          # the backward landed in the buckets again
    m.head.weight.grad = torch.zeros_like(m.head.weight)                         # a foreign tensor
    with pytest.raises(RuntimeError, match="no longer aliases"):
        red.start()


def test_bf16_shadow_cache_semantics():
    """`autograd._shadow`: bf16 copies live on the owning module, are reused while the parameter is unchanged, refreshed IN
    PLACE after an in-place update (same buffer: captured CUDA graphs keep pointing at it), rebuilt when the storage moves,
    and a bf16 parameter is its own shadow."""
    from pixart_sigma_b200 import autograd as ag
    lin = torch.nn.Linear(16, 8)
    w1 = ag._shadow(lin, "w")
    assert w1.dtype == torch.bfloat16 and torch.equal(w1, lin.weight.detach().to(torch.bfloat16))
    assert ag._shadow(lin, "w") is w1 and ag._shadow(lin, "b").shape == (8,)
    with torch.no_grad():
        lin.weight.add_(1.0)                                     # optimizer-style in-place update bumps _version
    w2 = ag._shadow(lin, "w")
    assert w2.data_ptr() == w1.data_ptr() and torch.equal(w2, lin.weight.detach().to(torch.bfloat16))
    lin.weight.data = lin.weight.data.clone()                    # storage moved (e.g. module.to()): new shadow values
    assert torch.equal(ag._shadow(lin, "w"), lin.weight.detach().to(torch.bfloat16))
    lin16 = torch.nn.Linear(16, 8).to(torch.bfloat16)
    assert ag._shadow(lin16, "w").data_ptr() == lin16.weight.data_ptr()
    assert ag._shadow(torch.nn.Linear(4, 4, bias=False), "b") is None
    ag.clear_shadow_cache(lin)
    assert "_pxa_shadow" not in lin.__dict__


class _ToyDenoiser(torch.nn.Module):
    """CPU stand-in with the denoiser's call signature (x, timestep, y, mask, data_info) -> (n, 8, h, w) and the model's
    parameter naming (blocks.<i>.*), so `train_step` + `IDDPMLoss` + `GradBucketReducer` run end to end without a GPU."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.inp = torch.nn.Conv2d(4, 8, 1)
        self.blocks = torch.nn.ModuleList([torch.nn.Conv2d(8, 8, 3, padding=1) for _ in range(2)])
        self.t_proj = torch.nn.Linear(1, 8)
        self.y_proj = torch.nn.Linear(16, 8)

    def forward(self, x, timestep, y, mask=None, data_info=None):
        h = self.inp(x) + self.t_proj(timestep.float()[:, None] / 1000.0)[:, :, None, None]
        cond = (self.y_proj(y.squeeze(1)) * mask.reshape(x.shape[0], -1, 1).float()).sum(1)
        for b in self.blocks:
            h = h + torch.tanh(b(h + cond[:, :, None, None]))
        return h


def _train_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pixart_sigma_b200.training import train_step
        m = _ToyDenoiser()
        red = GradBucketReducer(m)
        g = torch.Generator().manual_seed(21)
        x0, noise = torch.randn(4, 4, 8, 8, generator=g) * 0.5, torch.randn(4, 4, 8, 8, generator=g)
        y, t = torch.randn(4, 1, 6, 16, generator=g), torch.tensor([0, 250, 600, 999])
        mask = (torch.arange(6)[None] < torch.tensor([6, 2, 4, 1])[:, None]).long().view(4, 1, 1, 6)
        lo, hi = rank * 2, rank * 2 + 2                                        # this rank's images
        red.zero_grad()
        loss = train_step(m, IDDPMLoss(), x0[lo:hi], t[lo:hi], y[lo:hi], mask[lo:hi], noise=noise[lo:hi], reducer=red)
        torch.save({"loss": loss, "grads": {n: p.grad.clone() for n, p in m.named_parameters()}}, os.path.join(out_dir, f"t{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_data_parallel_train_step_gloo_world2(tmp_path):
    """N = 2 training step on CPU (gloo): per-rank IDDPM loss + backward + bucketed all-reduce equals the single-process
    step on the whole batch (the reducer averages; equal per-rank batch sizes)."""
    from pixart_sigma_b200.training import train_step
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "t0.pt"), torch.load(tmp_path / "t1.pt")
    m = _ToyDenoiser()
    g = torch.Generator().manual_seed(21)
    x0, noise = torch.randn(4, 4, 8, 8, generator=g) * 0.5, torch.randn(4, 4, 8, 8, generator=g)
    y, t = torch.randn(4, 1, 6, 16, generator=g), torch.tensor([0, 250, 600, 999])
    mask = (torch.arange(6)[None] < torch.tensor([6, 2, 4, 1])[:, None]).long().view(4, 1, 1, 6)
    loss = train_step(m, IDDPMLoss(), x0, t, y, mask, noise=noise)
    assert torch.allclose((r0["loss"] + r1["loss"]) / 2, loss, rtol=1e-5)
    for n, p in m.named_parameters():
        assert torch.allclose(r0["grads"][n], p.grad, rtol=1e-4, atol=1e-6), n
        assert torch.equal(r0["grads"][n], r1["grads"][n]), n

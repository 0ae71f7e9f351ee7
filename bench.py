#!/usr/bin/env python
"""bench.py -- denoise-steps/sec of the PixArt-Sigma-XL/2 denoiser hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one batch: one `PixArtMS.forward` of the CFG-batched latents of the
images a GPU holds (workload c3: 4 images/GPU at 1024px -> forward batch 8, 4096 tokens/sample), i.e. one DPM-Solver
model evaluation.  `value` = image-denoise-steps per second over all ranks (an image-step = one forward at batch 2,
reference diffusion/model/dpm_solver.py:328-331), inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e` is the same metric through the public API with pinned HOST inputs copied H2D and the eps result copied D2H
inside every timed step.  Weak scaling: per-GPU work is fixed, ranks are independent replicas (no collective in the
data path).  The working set per step (1.2 GB of weights + >1 GB of activations) is far larger than
the 126 MB L2, so no explicit L2 flush is needed between iterations.

`--workload c5` (BASELINE configs[4]) times the TRAINING step instead: IDDPM loss forward + backward through the
forward/backward kernels with per-block activation checkpointing + the bucketed gradient all-reduce (NCCL) overlapped
with the backward, 1024px, 4 images per GPU, fp32 master weights; `value` = trained images per second over all ranks.

One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C, L, HEADS, DEPTH = 1152, 300, 16, 28
WORKLOADS = {
    # name: (description, latent side, images per GPU, pe_interpolation, kv-compress)
    "c2": ("PixArt-Sigma-XL/2 512px, 4 images/GPU (forward batch 8, CFG pairs), 1024 tokens", 64, 4, 1.0, False),
    "c3": ("PixArt-Sigma-XL/2 1024-MS, 4 images/GPU (forward batch 8, CFG pairs), 4096 tokens "
           "[BASELINE configs[2]: batch 32 sharded over 8 GPUs]", 128, 4, 2.0, False),
    "c4": ("PixArt-Sigma-XL/2 2K-MS kv-compress sr=2 layers 14-27, 1 image/GPU (forward batch 2), 16384 tokens",
           256, 1, 4.0, True),
}


TRAIN_WORKLOAD = ("PixArt-Sigma-XL/2 1024-MS training step: IDDPM loss fwd + bwd (per-block activation checkpointing) + "
                  "bucketed gradient all-reduce, 4 images/GPU, 4096 tokens, fp32 master weights / bf16 kernels "
                  "[BASELINE configs[4]]", 128, 4, 2.0, False)


def flops_per_forward(n_tok: int, batch: int, kv_compress: bool):
    """Algorithmic FLOPs of one forward: per block 28NC^2 + 4LC^2 + 4N Nk C + 4NLC; returns
    (total, gemm part, attention part)."""
    gemm_blk = 28 * n_tok * C * C + 4 * L * C * C
    tot_gemm = tot_attn = 0
    for i in range(DEPTH):
        nk = n_tok // 4 if (kv_compress and i >= 14) else n_tok
        tot_gemm += gemm_blk
        tot_attn += 4 * n_tok * nk * C + 4 * n_tok * L * C
    emb = 2 * L * 4096 * C + 2 * L * C * C + 2 * n_tok * 16 * C + 2 * n_tok * C * 32 + (2 * 256 * C + 14 * C * C)
    return batch * (tot_gemm + tot_attn + emb), batch * (tot_gemm + 2 * L * 4096 * C + 2 * L * C * C), batch * tot_attn


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        # "under load" = samples drawing more than half the observed max power
        load = [s for s, pw in zip(sm, power) if power and pw > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(rows), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- CPU arms (oracle)
def cpu_reference_sample(side: int, pe: float, kv: bool, threads: int):
    """Time the oracle port of the reference forward (fp32, CPU) on a bounded sample of the workload: forward batch 2
    (= one image with CFG) through a 2-block and a 6-block slice of the 28-block model; per-block time = (t6-t2)/4,
    fixed (embedders + final layer) = t2 - 2*per-block, full forward = fixed + 28*per-block.  Returns
    (image-steps/s, seconds of CPU work measured)."""
    from oracle import pixart_oracle as po            # bench.py may use the oracle ONLY here (cpu baseline arm)
    torch.set_num_threads(threads)

    def run(depth):
        layers = list(range(depth // 2, depth)) if kv else []
        cfg = po.OracleConfig(depth=depth, input_size=side, pe_interpolation=pe, kv_sampling="conv" if kv else None,
                              kv_scale_factor=2 if kv else 1, kv_compress_layer=layers)
        sd = po.synthetic_state_dict(cfg, seed=0)
        x, t, y, mask = po.synthetic_inputs(cfg, 2, (side, side), lens=[300, 300])
        po.forward(sd, po.OracleConfig(depth=0, input_size=side, pe_interpolation=pe), x, t, y, mask=mask)  # warm-up
        t0 = time.perf_counter()
        po.forward(sd, cfg, x, t, y, mask=mask)
        return time.perf_counter() - t0

    t2, t6 = run(2), run(6)
    per_block = max((t6 - t2) / 4.0, 1e-9)
    fixed = max(t2 - 2.0 * per_block, 0.0)
    return 1.0 / (fixed + DEPTH * per_block), t2 + t6


def cpu_reference_train_sample(side: int, pe: float, threads: int):
    """Training counterpart of `cpu_reference_sample` (workload c5): the oracle port of the reference forward with autograd
    + the IDDPM loss + backward (fp32, CPU), ONE image, through a 1-block and a 3-block slice of the 28-block model;
    per-block time = (t3 - t1) / 2, fixed = t1 - per-block, full step = fixed + 28 x per-block.  Returns (images/s, seconds)."""
    from oracle import pixart_oracle as po            # bench.py may use the oracle ONLY in the cpu baseline / reference arms
    from pixart_sigma_b200.training import IDDPMLoss
    torch.set_num_threads(threads)

    def run(depth):
        cfg = po.OracleConfig(depth=depth, input_size=side, pe_interpolation=pe)
        sd = {k: v.requires_grad_(v.is_floating_point()) for k, v in po.synthetic_state_dict(cfg, seed=0).items()}
        x, _, y, mask = po.synthetic_inputs(cfg, 1, (side, side), lens=[300])
        model = lambda xx, timestep, **kw: po.forward_grad(sd, cfg, xx, timestep.float(), kw["y"], mask=kw["mask"])
        t0 = time.perf_counter()
        terms = IDDPMLoss().training_losses(model, x * 0.5, torch.tensor([500]), dict(y=y, mask=mask, data_info=None),
                                            noise=torch.randn_like(x))
        terms["loss"].mean().backward()
        return time.perf_counter() - t0

    run(0)                                             # warm-up (embedders, thread pool)
    t1, t3 = run(1), run(3)
    per_block = max((t3 - t1) / 2.0, 1e-9)
    fixed = max(t1 - per_block, 0.0)
    return 1.0 / (fixed + DEPTH * per_block), t1 + t3


def pick_cpu_threads() -> int:
    """torch CPU GEMMs of this size stop scaling (and regress) long before 128 threads: calibrate on a 1024x1152x4608
    fp32 matmul and keep the fastest of {all cores, 64, 32, 16}."""
    cores = os.cpu_count() or 1
    a, b = torch.randn(1024, 1152), torch.randn(1152, 4608)
    best, best_t = cores, None
    for n in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(n)
        a @ b
        t0 = time.perf_counter()
        for _ in range(5):
            a @ b
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    return best


def run_reference_arm(args, wl):
    desc, side, imgs, pe, kv = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = pick_cpu_threads()
    train = args.workload == "c5"
    sample_fn = (lambda: cpu_reference_train_sample(side, pe, threads)) if train else (lambda: cpu_reference_sample(side, pe, kv, threads))
    for _ in range(max(args.warmup, 0) and 1):
        sample_fn()
    vals, t_all = [], 0.0
    for _ in range(max(1, min(args.steps, 3))):
        v, dt = sample_fn()
        vals.append(v); t_all += dt
    value = statistics.median(vals)
    if train:
        sample = (f"oracle port of the reference training step (PixArtMS.forward with autograd + IDDPM loss + backward, fp32, torch "
                  f"CPU, {threads} threads) at {side * 8}px, ONE image: 1-block and 3-block slices timed, full 28-block step = "
                  f"fixed + 28 x per-block; median of {len(vals)} samples, {t_all:.1f} s of CPU work")
    else:
        sample = (f"oracle port of the reference PixArtMS.forward (fp32, torch CPU, {threads} threads) at {side * 8}px, forward "
                  f"batch 2 (one image with CFG): 2-block and 6-block slices timed, full 28-block forward = fixed + 28 x "
                  f"per-block; median of {len(vals)} samples, {t_all:.1f} s of CPU work")
    line = {"impl": "reference", "metric": "train-images/sec" if train else "denoise-steps/sec", "value": value,
            "unit": "images/s" if train else "image-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "note": "reference is pure Python; its CPU path timed via the oracle port "
                       "(the reference itself cannot travel to the GPU box)"},
            "cpu_baseline": {"value": value, "unit": "images/s" if train else "image-steps/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "images/s" if train else "image-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# --------------------------------------------------------------------------------------------- our arm
class KernelTimer:
    """CUDA-event timing of every GEMM / attention launch on the launching stream, inside the timed region."""

    def __init__(self, lib):
        self.lib, self.recs, self.on = lib, {"gemm": [], "attn": []}, False
        self._g, self._a = lib.gemm, lib.flash_attn

    def install(self):
        def g(*a, **k):
            return self._wrap("gemm", self._g, a, k)

        def f(*a, **k):
            return self._wrap("attn", self._a, a, k)
        self.lib.gemm, self.lib.flash_attn = g, f

    def _wrap(self, kind, fn, a, k):
        if not self.on:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        self.recs[kind].append((e0, e1))
        return r

    def totals_ms(self):
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.recs.items()}


def run_ours(args, wl, wl_name):
    import torch.distributed as dist
    from pixart_sigma_b200 import PixArtMS_XL_2, lib

    desc, side, imgs, pe, kv = wl
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    torch.manual_seed(1234 + rank)
    kvc = dict(sampling="conv", scale_factor=2, kv_compress_layer=list(range(14, 28))) if kv else None
    with torch.device(dev):
        model = PixArtMS_XL_2(input_size=side, pe_interpolation=pe, model_max_length=L, kv_compress_config=kvc)
        for blk in model.blocks:                      # un-zero the zero-initialised layers (SURVEY.md 8d weights)
            torch.nn.init.normal_(blk.cross_attn.proj.weight, std=0.02)
        torch.nn.init.normal_(model.final_layer.linear.weight, std=0.02)
    model = model.to(torch.bfloat16).eval()

    B = 2 * imgs                                       # CFG: [uncond, cond] per image
    n_tok = (side // 2) ** 2
    # DPM-Solver time grid of a 20-step run: (linspace(1, 1e-3, 21)[:-1] - 1e-3) * 1000 = 999.0, 949.05, ...
    tgrid = ((torch.linspace(1.0, 1e-3, 21)[:-1] - 1e-3) * 1000.0).tolist()
    g = torch.Generator().manual_seed(99 + rank)
    h_x = torch.randn(B, 4, side, side, generator=g).pin_memory()
    h_y = torch.randn(B, 1, L, 4096, generator=g).to(torch.bfloat16).pin_memory()
    lens = torch.randint(8, L + 1, (imgs,), generator=g)
    h_mask = (torch.arange(L)[None] < lens[:, None]).long().pin_memory()          # (imgs, L), repeated over CFG pairs
    h_t = torch.empty(B, dtype=torch.float32).pin_memory()
    h_out = torch.empty(B, 4, side, side, dtype=torch.bfloat16).pin_memory()
    d_x, d_y, d_mask = h_x.to(dev), h_y.to(dev), h_mask.to(dev)
    d_ts = [torch.full((B,), tv, device=dev) for tv in tgrid]

    graphed = None
    if args.cuda_graph:
        from pixart_sigma_b200.graph import GraphedForward
        graphed = GraphedForward(model)

    def step_resident(i):
        if graphed is not None:
            return graphed(d_x, d_ts[i % 20], d_y, d_mask)
        return model.forward_with_dpmsolver(d_x, d_ts[i % 20], d_y, None, mask=d_mask)

    def step_e2e(i):
        h_t.fill_(tgrid[i % 20])
        x = h_x.to(dev, non_blocking=True); t = h_t.to(dev, non_blocking=True)
        y = h_y.to(dev, non_blocking=True); mk = h_mask.to(dev, non_blocking=True)
        eps = model.forward_with_dpmsolver(x, t, y, None, mask=mk)
        h_out.copy_(eps, non_blocking=True)
        return eps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    timer = KernelTimer(lib)
    timer.install()
    with torch.no_grad():
        n_pre = lib.launch_count()
        model.forward_with_dpmsolver(d_x, d_ts[0], d_y, None, mask=d_mask)          # eager: kernel launches of one forward
        launches_per_forward = lib.launch_count() - n_pre
        for i in range(max(args.warmup, 3)):
            step_resident(i)
        step_e2e(0)
        sampler = ClockSampler(local) if rank == 0 else None
        n0 = lib.launch_count()
        timer.on = True
        ms = timed(step_resident, args.steps)
        timer.on = False
        launches = lib.launch_count() - n0
        if graphed is not None:
            launches = launches_per_forward * args.steps       # graph replays bypass the library's host-side counter
        ms_e2e = timed(step_e2e, args.steps)
        clocks = sampler.stop() if sampler else None
        loop = None
        if args.sampling_loop:
            # the caller of the path: a full 20-step DPM-Solver++ CFG sampling run of `imgs` images
            # through pixart_sigma_b200.sampler (fused step kernel; whole loop as one CUDA graph with --cuda-graph)
            from pixart_sigma_b200.sampler import DPMS
            cond, null = d_y[imgs:], d_y[:imgs]
            solver = DPMS(model.forward_with_dpmsolver, condition=cond, uncondition=null, cfg_scale=4.5,
                          model_kwargs=dict(data_info=None, mask=d_mask))
            z = d_x[:imgs].float()
            solver.sample(z, steps=20, cuda_graph=args.cuda_graph)                 # warm-up (and graph capture)
            ms_loop = timed(lambda i: solver.sample(z, steps=20, cuda_graph=args.cuda_graph), 1)
            loop = {"steps": 20, "images": imgs * world, "ms": ms_loop, "images_per_s": imgs * world / (ms_loop / 1000.0),
                    "denoise_steps_per_s": imgs * world * 20 / (ms_loop / 1000.0), "cuda_graph": bool(args.cuda_graph)}

    if rank == 0:
        tot, f_gemm, f_attn = flops_per_forward(n_tok, B, kv)
        ms_step = ms / args.steps
        value = imgs * world * args.steps / (ms / 1000.0)
        e2e_v = imgs * world * args.steps / (ms_e2e / 1000.0)
        sus, burst, hbm, src = measured_peaks()
        tt = timer.totals_ms()
        gemm_ms, gemm_n = tt["gemm"]
        attn_ms, attn_n = tt["attn"]
        gemm_tf = f_gemm * args.steps / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else None
        attn_tf = f_attn * args.steps / (attn_ms / 1000.0) / 1e12 if attn_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "latest_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("gemm_dram_bytes_per_launch")
        roof = {"bound": "tensor", "kernel": "pxa::gemm_bf16_kernel (all 7 GEMMs of the block; %.0f%% of the FLOPs)"
                % (100.0 * f_gemm / tot),
                "achieved": gemm_tf, "peak": sus, "unit": "TFLOP/s", "frac": (gemm_tf / sus) if gemm_tf else None,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({src})", "traffic": traffic,
                "launches_timed": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                "flops_per_launch_avg": f_gemm * args.steps / max(gemm_n, 1),
                "step_share": gemm_ms / ms if ms > 0 else None,
                "attention": {"kernel": "pxa::flash_attn_d72_kernel", "achieved": attn_tf, "frac": (attn_tf / sus) if attn_tf else None,
                              "launches_timed": attn_n, "step_share": attn_ms / ms if ms > 0 else None},
                "whole_step": {"achieved": tot / (ms_step / 1000.0) / 1e12, "frac": tot / (ms_step / 1000.0) / 1e12 / sus}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = pick_cpu_threads()
            v, dt = cpu_reference_sample(side, pe, kv, threads)
            cpu = {"value": v, "unit": "image-steps/s", "cores": threads, "kind": "port",
                   "sample": f"oracle port of the reference forward, fp32 torch CPU, {side * 8}px, forward batch 2: 2- and "
                             f"6-block slices timed ({dt:.1f} s), full forward = fixed + 28 x per-block"}
        h2d = h_x.numel() * 4 + h_y.numel() * 2 + h_mask.numel() * 8 + h_t.numel() * 4
        line = {"metric": "denoise-steps/sec", "value": value, "unit": "image-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{wl_name}: {desc}", "images_per_gpu": imgs, "forward_batch_per_gpu": B,
                           "tokens_per_sample": n_tok, "text_tokens": L, "parallelism": f"dp{world} (batch-sharded replicas)",
                           "l2": "inputs larger than L2 (weights 1.2 GB + activations per step); no flush needed",
                           "cuda_graph": bool(args.cuda_graph),
                           "tflop_per_step_per_gpu": tot / 1e12},
                "e2e": {"value": e2e_v, "unit": "image-steps/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h_out.numel() * 2},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        if loop is not None:
            line["sampling_loop"] = loop
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_train(args, wl):
    """BASELINE configs[4]: one training step = zero_grad + IDDPM loss forward + backward + gradient all-reduce."""
    import torch.distributed as dist
    from pixart_sigma_b200 import build_model, lib
    from pixart_sigma_b200.parallel import GradBucketReducer
    from pixart_sigma_b200.training import IDDPMLoss, train_step

    desc, side, imgs, pe, _ = wl
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    torch.manual_seed(1234)                                   # same initial weights on every rank (DDP replicas)
    with torch.device(dev):
        model = build_model(dict(type="PixArtMS_XL_2", input_size=side, pe_interpolation=pe, model_max_length=L),
                            use_grad_checkpoint=not args.no_checkpoint)
        for blk in model.blocks:
            torch.nn.init.normal_(blk.cross_attn.proj.weight, std=0.02)
        torch.nn.init.normal_(model.final_layer.linear.weight, std=0.02)
    model = model.float().train()
    reducer = GradBucketReducer(model)
    loss_fn = IDDPMLoss()
    n_tok = (side // 2) ** 2
    g = torch.Generator().manual_seed(7 + rank)
    h_x = (torch.randn(imgs, 4, side, side, generator=g) * 0.5).pin_memory()
    h_y = torch.randn(imgs, 1, L, 4096, generator=g).to(torch.bfloat16).pin_memory()
    lens = torch.randint(8, L + 1, (imgs,), generator=g)
    h_mask = (torch.arange(L)[None] < lens[:, None]).to(torch.int16).view(imgs, 1, 1, L).pin_memory()   # train.py:192 layout
    h_t = torch.randint(0, 1000, (imgs,), generator=g).pin_memory()
    h_noise = torch.randn(imgs, 4, side, side, generator=g).pin_memory()
    d_x, d_y, d_mask, d_t, d_noise = (v.to(dev) for v in (h_x, h_y, h_mask, h_t, h_noise))

    graphed = None
    if args.cuda_graph:
        from pixart_sigma_b200.training import GraphedTrainStep
        graphed = GraphedTrainStep(model, loss_fn, reducer, (d_x, d_t, d_y, d_mask, d_noise))

    def step_resident(i):
        if graphed is not None:
            return graphed(d_x, d_t, d_y, d_mask, d_noise)
        reducer.zero_grad()
        return train_step(model, loss_fn, d_x, d_t, d_y, d_mask, noise=d_noise, reducer=reducer)

    def step_e2e(i):
        x, y, mk, t, nz = (v.to(dev, non_blocking=True) for v in (h_x, h_y, h_mask, h_t, h_noise))
        if graphed is not None:
            return float(graphed(x, t, y, mk, nz))
        reducer.zero_grad()
        return float(train_step(model, loss_fn, x, t, y, mk, noise=nz, reducer=reducer))     # loss read back: D2H + sync

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for i in range(steps):
            fn(i)
        host_ms[fn.__name__] = (time.perf_counter() - t_host) * 1000.0 / steps     # host time to ENQUEUE a step (no sync)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    timer = KernelTimer(lib)
    timer.recs["attn_bwd"] = []
    timer.install()
    _ab, _wg = lib.flash_attn_bwd, lib.gemm_wgrad
    lib.flash_attn_bwd = lambda *a, **k: timer._wrap("attn_bwd", _ab, a, k)
    lib.gemm_wgrad = lambda *a, **k: timer._wrap("gemm", _wg, a, k)
    n_pre = lib.launch_count()
    reducer.zero_grad()
    train_step(model, loss_fn, d_x, d_t, d_y, d_mask, noise=d_noise, reducer=reducer)     # eager: kernel launches of one step
    launches_eager = lib.launch_count() - n_pre
    for i in range(max(args.warmup, 3)):
        loss0 = step_resident(i)
    step_e2e(0)
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.launch_count()
    timer.on = graphed is None                      # per-kernel events are not available inside a graph replay
    ms = timed(step_resident, args.steps)
    timer.on = False
    launches = lib.launch_count() - n0
    if graphed is not None:
        launches = launches_eager * args.steps       # graph replays bypass the library's host-side counter
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    if rank == 0:
        fwd, f_gemm, f_attn = flops_per_forward(n_tok, imgs, False)
        sus, burst, hbm, src = measured_peaks()
        ms_step = ms / args.steps
        tt = timer.totals_ms()
        recompute = 0 if args.no_checkpoint else 1
        gemm_flops = (3 + recompute) * f_gemm * args.steps            # fwd (+ recompute) + dgrad + wgrad
        gemm_ms, gemm_n = tt["gemm"]
        gemm_tf = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else None
        afw_ms, afw_n = tt["attn"]
        abw_ms, abw_n = tt["attn_bwd"]
        afw_tf = f_attn * args.steps / (afw_ms / 1000.0) / 1e12 if afw_ms > 0 else None   # recomputation reuses (o, lse)
        abw_tf = 3.5 * f_attn * args.steps / (abw_ms / 1000.0) / 1e12 if abw_ms > 0 else None     # 14 N Nk d executed (10 model)
        model_tf = 3 * fwd / (ms_step / 1000.0) / 1e12
        roof = {"bound": "tensor", "kernel": "pxa::gemm_bf16_kernel (forward, recompute, dgrad and wgrad GEMMs)",
                "achieved": gemm_tf, "peak": sus, "unit": "TFLOP/s", "frac": (gemm_tf / sus) if gemm_tf else None,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({src})", "traffic": None,
                "launches_timed": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                "flops_per_launch_avg": gemm_flops / max(gemm_n, 1), "step_share": gemm_ms / ms,
                "attention_fwd": {"kernel": "pxa::flash_attn_d72_kernel", "achieved": afw_tf, "frac": afw_tf / sus if afw_tf else None,
                                  "launches_timed": afw_n, "step_share": afw_ms / ms},
                "attention_bwd": {"kernel": "pxa::flash_attn_d72_bwd_kernel (dKV + dQ passes + delta)", "achieved": abw_tf,
                                  "frac": abw_tf / sus if abw_tf else None, "launches_timed": abw_n, "step_share": abw_ms / ms,
                                  "note": "executed FLOPs (14 N Nk d incl. the dQ-pass recomputation); model FLOPs are 10 N Nk d"},
                "whole_step": {"model_tflops": model_tf, "frac": model_tf / sus,
                               "note": "model FLOPs = 3 x forward (no recomputation counted)"}}
        h2d = h_x.numel() * 4 + h_y.numel() * 2 + h_mask.numel() * 2 + h_t.numel() * 8 + h_noise.numel() * 4
        line = {"metric": "train-images/sec", "value": imgs * world * args.steps / (ms / 1000.0), "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"c5: {desc}", "images_per_gpu": imgs, "tokens_per_sample": n_tok, "text_tokens": L,
                           "parallelism": f"ddp{world} (bucketed NCCL all-reduce of {reducer.grad_bytes() / 1e9:.2f} GB fp32 grads, "
                                          f"{len(reducer.buckets)} buckets, overlapped with backward)",
                           "grad_checkpointing": not args.no_checkpoint, "cuda_graph": bool(args.cuda_graph), "optimizer_step": "not included (fwd + bwd + all-reduce, as configs[4] states)",
                           "l2": "working set (2.4 GB fp32 weights + activations) larger than L2; no flush needed",
                           "tflop_model_per_step_per_gpu": 3 * fwd / 1e12, "peak_mem_gib": peak_mem, "loss": float(loss0),
                           "host_enqueue_ms_per_step": host_ms.get("step_resident")},
                "e2e": {"value": imgs * world * args.steps / (ms_e2e / 1000.0), "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": None}
        if world == 1 and not args.no_cpu_baseline:
            threads = pick_cpu_threads()
            v, dt = cpu_reference_train_sample(side, pe, threads)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle port of the reference training step (forward with autograd + IDDPM loss + "
                                              f"backward, fp32 torch CPU), {side * 8}px, one image: 1- and 3-block slices timed "
                                              f"({dt:.1f} s), full step = fixed + 28 x per-block"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--no-checkpoint", action="store_true", help="c5: train without per-block activation checkpointing")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sampling-loop", action="store_true",
                    help="also time one full 20-step DPM-Solver++ sampling run through pixart_sigma_b200.sampler "
                         "(extra key `sampling_loop`; not part of the timed region of `value` / `e2e`)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="replay the resident-input forward as one CUDA graph (pixart_sigma_b200.graph.GraphedForward); "
                         "per-kernel event timing (roofline) is unavailable in this mode")
    args = ap.parse_args()
    wl = TRAIN_WORKLOAD if args.workload == "c5" else WORKLOADS[args.workload]
    # The contract is ONE JSON line on stdout.  Libraries write banners there too (NCCL prints its version line at the
    # first communicator init), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved fd.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.workload == "c5" and args.impl != "reference":
        run_train(args, wl)
    elif args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()

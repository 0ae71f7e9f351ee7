"""Summarise gpurun_out ncu artefacts into profiles/<tag>_*.{md,csv} (tracked).  Usage: summarize_profile.py <tag>"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
if tag.startswith("c5"):
    lines = [f"# ncu summary {tag} (training step)", "", "Command: `tools/train_profile.py --depth 4` (launch list of ONE training step: IDDPM loss "
             "fwd + bwd, activation checkpointing, 1024px, 4 images, 4096 tokens) / `--depth 1` (`--set full` on the backward "
             "kernels), one B200, `--clock-control none`, `--profile-from-start off`. Per-launch times under ncu are cold-cache "
             "and serialised: compare SHARES.", ""]
else:
    lines = [f"# ncu summary {tag}", "", "Command: `bench.py --no-cuda-graph --no-extras --no-parity --steps 2 --warmup 3` (launch list) / `--steps 1` (`--set full`, GEMM + attention + norm kernels), "
             "workload c3 (1024px, forward batch 8), one B200, `--clock-control none`. Per-launch times under ncu are cold-cache and "
             "serialised: compare SHARES.", ""]
lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(rows):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        k = r["Kernel Name"].split("(")[0][:70]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    lines += ["## Launch list (gpu__time_duration.sum)", "", "| share | launches | avg us | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24 if tag.startswith('c5') else 12]:
        lines.append(f"| {100 * v[1] / tot:.2f}% | {v[0]} | {v[1] / v[0]:.1f} | `{k}` |")
    lines.append(f"\nTotal {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches.\n")
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.writelines(rows)
rp = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
rcsv = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_raw.csv")      # already converted on the GPU box (run_gpu_ncu.sh)
if os.path.exists(rp) or os.path.exists(rcsv):
    raw = open(rcsv).read() if os.path.exists(rcsv) else \
        subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) %"),
            ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
    lines += ["## `--set full` captures", "", "| kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|" + "---|" * len(cols)]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:60]
        vals = [f"{r[idx[c]]} {units[idx[c]]}".strip() if c in idx else "-" for c, _ in cols]
        lines.append(f"| `{name}` | " + " | ".join(vals) + " |")
    with open(os.path.join(out_dir, f"{tag}_raw.csv"), "w") as f:
        keep = ["Kernel Name"] + [c for c, _ in cols] + ["sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "lts__t_bytes.sum"]
        w = csv.writer(f)
        w.writerow(keep)
        for r in rows[2:]:
            w.writerow([r[idx[k]] if k in idx else "" for k in keep])
    # average DRAM bytes per GEMM launch (the bench's roofline.traffic for the GEMM family)
    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    gem = [r for r in rows[2:] if "gemm" in r[idx["Kernel Name"]]]
    if gem and not tag.startswith("c5"):
        import json
        tot = [to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
               to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in gem]
        json.dump({"tag": tag, "gemm_launches_captured": len(gem), "gemm_dram_bytes_per_launch": sum(tot) / len(tot),
                   "source": f"profiles/{tag}_raw.csv (ncu --set full of bench.py c3, GEMM launches of one PixArtMSBlock)",
                   "note": "mean of dram__bytes_read.sum + dram__bytes_write.sum over the GEMM launches of one ncu --set full capture"},
                  open(os.path.join(out_dir, "latest_traffic.json"), "w"), indent=1)
bp = os.path.join(ROOT, "gpurun_out", f"bench_{tag}.json")
if os.path.exists(bp):
    lines += ["", "## bench line of the same build", "", "```json", open(bp).read().strip(), "```"]
open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))

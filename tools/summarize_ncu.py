#!/usr/bin/env python
"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).  Usage: tools/summarize_ncu.py TAG [N]"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = [f"# ncu summary `{tag}` (N={N})", ""]

# ---- launch list
lc = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 10 and r[0].isdigit()]
    per = {}
    for r in rows:
        per.setdefault(r[4].split("(")[0], []).append(float(r[-1]))
    tot = sum(sum(v) for v in per.values())
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, our kernels, cold-cache & serialised)", "",
              "| kernel | launches | avg us | share of our kernels' time |", "|---|---|---|---|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| {k} | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {100*sum(v)/tot:.1f}% |")
    lines.append("")
    keep = os.path.join(out_dir, f"launches_{tag}.csv")
    with open(keep, "w") as f:
        f.write(open(lc).read())

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
    lines += [f"## Full capture of `{data[0][hdr.index('Kernel Name')] if 'Kernel Name' in hdr else 'kernel'}` (`ncu --set full --clock-control none`, {len(data)} launches)", "",
              "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
    traffic = None
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    try:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        def tobytes(v, u):
            m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v) * m
        traffic = sum(tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw]) for r in data) / len(data)
        lines += ["", f"DRAM traffic per launch (read+write): **{traffic/1e9:.4f} GB** vs algorithmic 8*N^2 = {8*N*N/1e9:.4f} GB "
                  f"(ratio {traffic/(8.0*N*N):.4f})."]
        tj = os.path.join(out_dir, "traffic.json")
        d = json.load(open(tj)) if os.path.exists(tj) else {}
        d[str(N)] = round(traffic)
        json.dump(d, open(tj, "w"), indent=1)
    except Exception as e:
        lines.append(f"(traffic not extracted: {e})")
    # stall reasons
    st = [(hdr[i], [float(r[i]) for r in data]) for i in range(len(hdr)) if "warp_issue_stalled" in hdr[i] and hdr[i].endswith("_per_warp_active.pct")]
    if st:
        lines += ["", "Top warp stall reasons (% of warp-active samples):", ""]
        for k, v in sorted(st, key=lambda kv: -kv[1][0])[:8]:
            lines.append(f"- {k.split('issue_stalled_')[1].replace('_per_warp_active.pct','')}: {v[0]:.1f}%")
open(os.path.join(out_dir, f"ncu_{tag}.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

#!/usr/bin/env python
"""Turn the ncu outputs of tools/collect_evidence.sh (gpurun_out/) into the tracked evidence files under profiles/:
  <tag>_launches_step_B256.csv     the raw per-launch list (gpu__time_duration, dram bytes) of ONE denoise step
  <tag>_launches_summary.txt       per-kernel launches / time share / DRAM bytes per launch
  roofline_traffic.json            dram__bytes_read+write per launch of each conv kernel (read by bench.py -> roofline.traffic)
  <tag>_ncu_<name>_raw.txt         selected raw metrics of a `--set full` capture
  <tag>_ncu_<name>_stalls.txt      top stall-sampled SASS lines of that capture
usage: summarise_ncu.py <tag>   (e.g. r01s4)"""
import csv, json, os, re, shutil, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

src = os.path.join(OUT, "launches_step_B256.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(PROF, f"{tag}_launches_step_B256.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(float))
    ids = defaultdict(set)
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"^void (ddif::)?", "", name)
        metric, val = r[ix["Metric Name"]], float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        if metric.startswith("dram__bytes"):
            val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if metric == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)  # -> us
        per[name][metric] += val
        ids[name].add(r[ix["ID"]])
    tot = sum(v["gpu__time_duration.sum"] for v in per.values())
    lines = [f"one denoise step, B=256 WV3 64x64 (ncu, serialised, cold-ish caches): {sum(len(v) for v in ids.values())} launches, {tot / 1e3:.3f} ms",
             f"{'kernel':<34}{'launches':>9}{'us total':>10}{'share %':>9}{'us/launch':>10}{'DRAM MB/launch':>16}"]
    traffic = {}
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[name])
        dram = (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / n
        lines.append(f"{name:<34}{n:>9}{v['gpu__time_duration.sum']:>10.1f}{100 * v['gpu__time_duration.sum'] / tot:>9.1f}{v['gpu__time_duration.sum'] / n:>10.1f}{dram / 1e6:>16.2f}")
    base = defaultdict(lambda: [0.0, 0])
    for name, v in per.items():
        b = base[re.sub(r"<.*", "", name)]
        b[0] += v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
        b[1] += len(ids[name])
    for name, (byt, n) in base.items():
        if "conv" in name:
            traffic[name + "_dram_bytes_per_launch"] = byt / n
    open(os.path.join(PROF, f"{tag}_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    traffic["source"] = f"profiles/{tag}_launches_step_B256.csv (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one denoise step)"
    json.dump(traffic, open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)
    print("\n".join(lines))

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct", "lts__throughput.avg.pct", "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "smsp__average_warps_issue_stalled", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max")
for f in sorted(os.listdir(OUT)):
    if not f.endswith(".ncu-rep"):
        continue
    name = f[:-8]
    rep = os.path.join(OUT, f)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) >= 3:
        with open(os.path.join(PROF, f"{tag}_ncu_{name}_raw.txt"), "w") as fo:
            kn = rr[2][rr[0].index("Kernel Name")] if "Kernel Name" in rr[0] else ""
            fo.write(f"# ncu --set full --clock-control none, {f}: {kn}\n")
            for i, k in enumerate(rr[0]):
                if any(k.startswith(x) for x in KEYS):
                    fo.write(f"{k} [{rr[1][i]}] = {rr[2][i]}\n")
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = f"/tmp/{name}_src.csv"
    open(tmp, "w").write(srcp)
    top = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_top.py"), tmp, "40"], capture_output=True, text=True).stdout
    open(os.path.join(PROF, f"{tag}_ncu_{name}_stalls.txt"), "w").write(top)
    print(name, "->", f"{tag}_ncu_{name}_raw.txt / _stalls.txt")

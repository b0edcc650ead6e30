#!/usr/bin/env python3
"""Turn the raw captures in gpurun_out/ into the tracked summaries under profiles/ (run locally after scripts/make_profiles.sh)."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

def short(name):
    m = re.search(r"(x_kernel|col_kernel)<mvd::Plan<\(int\)(\d+).*?>, \(int\)(\d)>", name)
    if m:
        kind = {"x_kernel": ["X_FWD", "X_RATIO", "X_UPDATE", "X_INV"], "col_kernel": ["COL_FWD", "COL_INV", "COL_CONV"]}[m.group(1)][int(m.group(3))]
        return f"{m.group(1)}<N={m.group(2)},{kind}>"
    return re.sub(r"\(.*", "", name)[:60]

# ---- launch list: per-kernel share of the step -------------------------------------------------------------------
rows = list(csv.reader(l for l in open(os.path.join(G, "launches.csv")) if not l.startswith("==")))
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(short(r[ik]), [0, 0.0])
    a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"launches_{R}.md"), "w") as f:
    f.write(f"# ncu launch list ({R}) -- `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu`\n\n")
    f.write("Cold-cache, serialised launch times: compare SHARES, not absolutes. First 600 launches (setup convolutions, spectra, warm-up, timed steps).\n\n")
    f.write("| kernel | launches | total ns | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.0f} | {t / tot:.3f} |\n")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"launches_{R}.csv"))

# ---- full captures -------------------------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size"]
STALL = "smsp__average_warps_issue_stalled_"
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic, out = {}, [f"# ncu --set full captures ({R}), c3 tile 1080x540x540 (2 tiles), B200\n"]
pass_of = {("x_kernel", "0"): "P1", ("x_kernel", "1"): "P5", ("x_kernel", "2"): "P9", ("col_kernel", "0"): "P2", ("col_kernel", "2"): "P3", ("col_kernel", "1"): "P4"}
for rep in ("prof_x_c3", "prof_col_c3"):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        m = re.search(r"(x_kernel|col_kernel).*?, (\d)>\(", name)
        pas = pass_of.get((m.group(1), m.group(2)), "?") if m else "?"
        out.append(f"\n## {pas}  `{short(name.replace('Plan<', 'mvd::Plan<(int)').replace(', ', ', (int)')) if False else name[:100]}`\n")
        out.append("| metric | value |\n|---|---|")
        rd = wr = 0.0
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {r[i]} {units[i]} |")
                if w == "dram__bytes_read.sum": rd = float(r[i]) * UNIT.get(units[i], 1.0)
                if w == "dram__bytes_write.sum": wr = float(r[i]) * UNIT.get(units[i], 1.0)
        st = sorted(((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr)
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")), reverse=True)
        out.append("| stall cycles per issue (top) | " + ", ".join(f"{n}={v:.2f}" for v, n in st[:6]) + " |")
        traffic[pas] = rd + wr
        if pas == "P2": traffic["P6"] = rd + wr
        if pas == "P3": traffic["P7"] = rd + wr
        if pas == "P4": traffic["P8"] = rd + wr
open(os.path.join(P, f"ncu_{R}.md"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
for fn in ("passes_c3.json",):
    if os.path.exists(os.path.join(G, fn)):
        shutil.copy(os.path.join(G, fn), os.path.join(P, fn.replace(".json", f"_{R}.json")))
print("profiles written:", sorted(os.listdir(P)))

#!/usr/bin/env python3
"""gpurun_out/{ncu_final.ncu-rep, final_launches.csv, final_*.json} -> tracked summaries under profiles/ (round 2)."""
import collections, csv, json, os, re, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def short(name):
    name = name.replace("(int)", "")
    m = re.search(r"(x_kernel_w|x_kernel_p|x_kernel|col_kernel)<.*?Plan<(\d+), (\d+), (\d+), (\d+).*?>, (\d)", name)
    if m:
        kinds = {"x": ["X_FWD", "X_RATIO", "X_UPDATE", "X_INV"], "c": ["COL_FWD", "COL_INV", "COL_CONV"]}[m.group(1)[0]]
        return f"{m.group(1)}<N={m.group(2)} ({m.group(3)}x{m.group(4)}{'x' + m.group(5) if m.group(5) != '1' else ''}),{kinds[int(m.group(6))]}>"
    return re.sub(r"\(.*", "", name)[:70]

# ---- launch list of the bench command ---------------------------------------------------------------------------------
src = os.path.join(G, "final_launches.csv")
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(short(r[ik]), [0, 0.0])
    a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, "launches_r02.md"), "w") as f:
    f.write("# ncu launch list (round 2, final build)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-parity`\n\n")
    f.write("Cold-cache, serialised launch times: compare SHARES, not absolutes.  The launches cover the set-up (image synthesis convolutions, "
            "PSF derivation, kernel spectra, weight masks, PsiInit), the warm-up and the timed steps.\n\n")
    f.write("| kernel | launches | total ns | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.0f} | {t / tot:.3f} |\n")
    # share of one view update on one tile: mean launch time of each c3 tile kernel x its launches per view update, next to the
    # CUDA-event shares of the committed bench line (all launches of a kernel name do the same work: same tile)
    per_update = {"x_kernel_w<N=540 (18x30),X_FWD>": ("P1", 1), "col_kernel<N=540 (27x20),COL_FWD>": ("P2 + P6", 2),
                  "col_kernel<N=540 (27x20),COL_CONV>": ("P3 + P7", 2), "col_kernel<N=540 (27x20),COL_INV>": ("P4 + P8", 2),
                  "x_kernel_w<N=540 (18x30),X_RATIO>": ("P5", 1), "x_kernel<N=540 (18x30),X_UPDATE>": ("P9", 1)}
    # only the launches of the two timed steps: the last 2 steps x 4 views x 2 tiles x 9 passes of the list (the same kernel names also
    # serve the set-up -- spectra, PsiInit blur -- with other extension modes and cache states)
    seq = [(short(r[ik]), float(r[iv].replace(",", ""))) for r in rows[1:] if len(r) > iv and r[im] == "gpu__time_duration.sum"]
    tail = [x for x in seq if x[0] in per_update][-144:]
    agg = collections.OrderedDict()
    for k, t in tail:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
    tb = sum(agg[k][1] / agg[k][0] * m for k, (_, m) in per_update.items() if k in agg)
    ev = None
    try:
        line = json.loads([l for l in open(os.path.join(P, "bench_c3_n1_r02.json")).read().splitlines() if l.startswith("{")][-1])
        pp = line["roofline"]["all_passes_ms_per_launch"]
        ev = {"P1": pp[0], "P2 + P6": pp[1] + pp[5], "P3 + P7": pp[2] + pp[6], "P4 + P8": pp[3] + pp[7], "P5": pp[4], "P9": pp[8]}
    except Exception:
        pass
    f.write("\nShare of one view update (c3 tile kernels, FFT length 540): ncu launch list vs the CUDA events of `bench_c3_n1_r02.json`:\n\n"
            "| kernel | passes | launches in the two timed steps | ncu mean ns / launch | ncu share | CUDA-event share |\n|---|---|---:|---:|---:|---:|\n")
    for k, (pas, m) in per_update.items():
        if k in agg:
            mean = agg[k][1] / agg[k][0]
            es = f"{ev[pas] / sum(ev.values()):.3f}" if ev else "-"
            f.write(f"| `{k}` | {pas} | {agg[k][0]} | {mean:.0f} | {mean * m / tb:.3f} | {es} |\n")
shutil.copy(src, os.path.join(P, "launches_r02.csv"))

# ---- ncu --set full of the nine passes ------------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
STALL = "smsp__average_warps_issue_stalled_"
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
txt = subprocess.run(["ncu", "-i", os.path.join(G, "ncu_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
names = ["P1", "P2", "P3", "P4", "P5", "P6", "P7", "P8", "P9"]
traffic = {}
out = ["# ncu --set full captures (round 2, final build), c3 tile 1080x540x540, one B200\n",
       "`ncu --set full --clock-control none --import-source on -k regex:Plan<(int)540 -s 96 -c 9 python scripts/prof_passes.py c3 1` "
       "(the nine passes of one view update on one tile, in launch order).  Times under the profiler are cold-cache and serialised; "
       "the bench line's CUDA-event times are the ones to quote.\n"]
for pas, r in zip(names, rows[2:11]):
    out.append(f"\n## {pas}  `{short(r[hdr.index('Kernel Name')])}`\n")
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
    out.append(f"| DRAM bytes per launch | {(rd + wr) / 1e9:.3f} GB |")
    traffic[pas] = rd + wr
open(os.path.join(P, "ncu_r02.md"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
for a, b in (("last_bench_c3.json", "bench_c3_n1_r02.json"), ("final_bench_c2.json", "bench_c2_n1_r02.json"), ("prof_passes_c3.json", "passes_c3_r02.json"),
             ("final_pytest.log", "pytest_gpu_r02.log"), ("final_smoke.log", "smoke_r02.log")):
    if os.path.exists(os.path.join(G, a)):
        lines = [l for l in open(os.path.join(G, a)).read().splitlines() if l.strip()]
        keep = [l for l in lines if l.startswith("{")] if a.endswith(".json") and "prof_passes" not in a else lines
        open(os.path.join(P, b), "w").write("\n".join(keep if keep else lines) + "\n")
print(json.dumps(traffic, indent=1))

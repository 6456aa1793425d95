#!/usr/bin/env python
"""Turn the ncu outputs of scripts/gpu_round.sh (gpurun_out/r01_top.ncu-rep, gpurun_out/r01_launches.csv) into the tracked
summaries under profiles/: r01_top_kernels.csv (per-kernel metrics), r01_traffic.json (DRAM bytes per launch, read by bench.py),
r01_launches_tc.csv (the launch list) and a share-of-step table on stdout."""
import csv, io, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic"]
# eager forward order of the kernels matched by the capture's -k regex (batch 32 x 4096)
STAGES = ["fps0", "knn0", "fps1", "knn1", "fps2", "three_nn0", "sa0", "sa1", "sa2", "fp2", "fp1", "fp0", "vlad0", "vlad1", "vlad2"]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[h], rows[h + 1], rows[h + 2:]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mult


def main():
    rep = os.path.join(OUT, "r01_top.ncu-rep")
    head, units, rows = raw_page(rep)
    col = {name: i for i, name in enumerate(head)}
    traffic = {}
    with open(os.path.join(PROF, "r01_top_kernels.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["stage", "Kernel Name", "Grid Size", "Block Size"] + [f"{m} [{units[col[m]]}]" for m in METRICS if m in col])
        for stage, r in zip(STAGES, rows):
            name = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0]
            w.writerow([stage, name, r[col["Grid Size"]], r[col["Block Size"]]] + [r[col[m]] for m in METRICS if m in col])
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            traffic[stage] = int(rd + wr)
    old = {}
    try:
        old = json.load(open(os.path.join(PROF, "r01_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    kept = [k for k in old if k not in traffic]
    for k in kept:
        traffic[k] = old[k]
    json.dump({"source": "profiles/r01_top_kernels.csv (ncu --set full --clock-control none, one launch each, batch 32 x 4096"
                         + (f"; {', '.join(kept)} from the previous capture of the same kernel)" if kept else ")"),
               "dram_bytes_per_launch": traffic}, open(os.path.join(PROF, "r01_traffic.json"), "w"), indent=1)
    # launch list -> share of the step
    src = os.path.join(OUT, "r01_launches.csv")
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    head = rows[h]
    kn, mv = head.index("Kernel Name"), head.index("Metric Value")
    with open(os.path.join(PROF, "r01_launches_tc.csv"), "w", newline="") as f:
        csv.writer(f).writerows(rows[h:])
    per = collections.OrderedDict()
    data = rows[h + 1:]
    # the last forward of the run: launches after the last occurrence of the first fps kernel
    starts = [i for i, r in enumerate(data) if "fps_kernel<8" in r[kn]]
    fwd = data[starts[-2]:starts[-1]] if len(starts) >= 2 else data     # the last COMPLETE forward (-c may cut the final one)
    for r in fwd:
        name = r[kn].replace("<unnamed>::", "").replace("void ", "")[:70]
        per[name] = per.get(name, 0.0) + float(r[mv].replace(",", "")) / 1000.0
    tot = sum(per.values())
    print("| kernel | µs per forward | share |\n|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        print(f"| `{k}` | {v:.1f} | {100 * v / tot:.1f} % |")
    print(f"| total ({len(fwd)} launches) | {tot:.1f} | 100 % |")


if __name__ == "__main__":
    main()

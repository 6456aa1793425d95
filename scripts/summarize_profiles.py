#!/usr/bin/env python
"""Turn the ncu outputs of scripts/gpu_profile.sh (gpurun_out/r02_step_metrics.csv = one whole eager forward with a metric set,
r02_top.ncu-rep = `--set full` of the top kernels, r02_extra_metrics.csv = kernels outside the bench step) into the tracked
summaries under profiles/: r02_top_kernels.csv, r02_extra_kernels.csv, r02_traffic.json (read by bench.py) and a share-of-step
table (profiles/r02_share_of_step.md)."""
import collections, csv, io, json, os, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
COLS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]
ORDER = {"fps_kernel": ["fps0", "fps1", "fps2"], "gather_rows_kernel": ["gather0", "gather1", "gather2"],
         "knn_index_kernel": ["index0", "index1"], "knn_pruned32_kernel": ["knn0", "knn1"], "knn_kernel": ["knn2"],
         "three_nn_kernel": ["three_nn2", "three_nn1"], "three_nn_pruned_kernel": ["three_nn0"],
         "sa_narrow_tc_kernel": ["sa0"], "mlp_tc_kernel": ["sa1", "sa2", "fp2", "fp1", "fp0"], "vlad_tc_kernel": ["vlad0", "vlad1", "vlad2"],
         "vlad_finalize_kernel": ["vlad0_fin", "vlad1_fin", "vlad2_fin"], "afa_att_kernel": ["afa_att"], "afa_att_tc_kernel": ["afa_att"], "afa_fc_tc_kernel": ["afa_fc"],
         "afa_softmax_kernel": ["afa_softmax"], "afa_fc_kernel": ["afa_fc"], "afa_finalize_kernel": ["afa_finalize"]}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {n: i for i, n in enumerate(rows[h])}
    out = collections.OrderedDict()
    for r in rows[h + 1:]:
        if not r or not r[0].isdigit():
            continue
        d = out.setdefault(int(r[0]), {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        try:
            d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", "")) * UNIT.get(r[col["Metric Unit"]], 1)
        except ValueError:
            pass
    return list(out.values())


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    return n.split("<")[0], n


def write_table(rows, path, stage_of=None):
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["stage", "kernel", "grid", "block"] + COLS)
        for i, r in enumerate(rows):
            w.writerow([stage_of(i, r) if stage_of else "", short(r["kernel"])[1], r["grid"], r["block"]] + [r.get(c, "") for c in COLS])


def main():
    step = launches(os.path.join(OUT, "r02_step_metrics.csv"))
    counters = collections.Counter()
    stages = []
    for r in step:
        base = short(r["kernel"])[0]
        names = ORDER.get(base, [base])
        stages.append(names[min(counters[base], len(names) - 1)])
        counters[base] += 1
    write_table(step, os.path.join(PROF, "r02_top_kernels.csv"), lambda i, r: stages[i])
    total = sum(r["gpu__time_duration.sum"] for r in step)
    traffic = {}
    lines = ["# One eager forward (batch 32 x 4096) under ncu: share of the step per kernel, tensor-pipe activity, DRAM and L2 traffic", "",
             "`ncu --metrics ... --clock-control none -k regex:<all hand-written kernels> -s 58 -c 29 python bench.py --mode eager` "
             "(scripts/gpu_profile.sh).  Times under ncu are serialised and cold-cache: the SHARE column is what compares with the "
             "live per-stage events of bench.py (`stage_ms`).", "",
             "| stage | kernel | grid | us | share | tensor pipe active % | issue active % | DRAM MB | L2 MB | regs |", "|---|---|---|---|---|---|---|---|---|---|"]
    merged = collections.OrderedDict()
    for st, r in zip(stages, step):
        key = "afa" if st.startswith("afa_") else st.replace("_fin", "")
        traffic[key] = traffic.get(key, 0) + int(r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0))
        lines.append("| %s | %s | %s | %.1f | %.1f %% | %.1f | %.1f | %.1f | %.1f | %d |" % (
            st, short(r["kernel"])[1][:34], r["grid"], r["gpu__time_duration.sum"], 100 * r["gpu__time_duration.sum"] / total,
            r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0), r.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
            (r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)) / 1e6, r.get("lts__t_bytes.sum", 0) / 1e6,
            int(r.get("launch__registers_per_thread", 0))))
    lines.append("")
    lines.append("Sum of kernel times under ncu: %.1f us." % total)
    open(os.path.join(PROF, "r02_share_of_step.md"), "w").write("\n".join(lines) + "\n")
    json.dump({"source": "profiles/r02_top_kernels.csv (ncu metric capture of one eager forward, --clock-control none, batch 32 x 4096): "
                         "dram__bytes_read.sum + dram__bytes_write.sum per launch (afa = its three launches)",
               "dram_bytes_per_launch": traffic}, open(os.path.join(PROF, "r02_traffic.json"), "w"), indent=1)
    extra = os.path.join(OUT, "r02_extra_metrics.csv")
    if os.path.exists(extra):
        write_table(launches(extra), os.path.join(PROF, "r02_extra_kernels.csv"))
    rep = os.path.join(OUT, "r02_top.ncu-rep")
    if os.path.exists(rep):
        keys = "gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio"
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", keys], capture_output=True, text=True).stdout
        open(os.path.join(PROF, "r02_top_full_set_raw.csv"), "w").write(txt)
    print("\n".join(lines))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns the ncu CSVs of profiles/capture.sh into the committed summaries:
    python profiles/summarize.py r1 4096
  profiles/<tag>_launches_summary.txt   per-kernel mean duration and share of a tracker step (launch list)
  profiles/<tag>_full_summary.txt       key `ncu --set full` metrics of the hot kernels
  profiles/traffic.json                 dram bytes (read + write) per launch per sequence, read by bench.py
"""
import csv
import json
import os
import sys
from collections import OrderedDict

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "gpurun_out")
STEP_KERNELS = ["pyramid_fused_kernel", "features_prepare_kernel", "init_pose_kernel", "sparse_align_kernel", "compose_poses_kernel",
                "reproject_prepare_kernel", "match_geom_kernel", "match_prepare_kernel", "match_direct_kernel", "seed_pose_table_kernel", "seeds_geom_kernel",
                "seeds_search_kernel", "epi_search_kernel", "lk_refine_kernel", "seeds_finish_kernel", "step_stats_kernel"]


def short(name):
    for k in STEP_KERNELS + ["synth_render_kernel", "fast_init_keys_kernel", "fast_kernel", "fast_finalize_kernel"]:
        if k in name:
            return k
    return name[:40]


def read_csv(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    return rows


# ---------------------------------------------------------------- launch list
rows = read_csv(os.path.join(SRC, tag + "_launches.csv"))
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
per = OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
    per.setdefault(short(r[ik]), []).append(v)
# steady state: drop the first launch of each step kernel when there are several (cold start)
mean = {k: (sum(v[1:]) / len(v[1:]) if len(v) > 2 else sum(v) / len(v)) for k, v in per.items()}
# lk_refine_kernel runs twice per step (map features, then seeds): its mean counts twice in the step total and its share is of both
PER_STEP = {"lk_refine_kernel": 2}
step_total = sum(mean[k] * PER_STEP.get(k, 1) for k in STEP_KERNELS if k in mean)
with open(os.path.join(HERE, tag + "_launches_summary.txt"), "w") as f:
    f.write("# %s ncu launch list (gpu__time_duration.sum, --clock-control none), command:\n" % tag)
    f.write("#   python bench.py --seqs %d --steps 2 --warmup 1 --no-latency --no-cpu-baseline --no-e2e --no-widen\n" % seqs)
    f.write("# cold-cache, serialised launches: compare SHARES, not absolutes.  One tracker step = one launch of each kernel below (lk_refine_kernel: two, share of both).\n")
    f.write("%-30s %6s %12s %s\n" % ("kernel", "n", "mean_us", "share_of_step"))
    for k in STEP_KERNELS:
        if k in mean:
            f.write("%-30s %6d %12.1f %8.3f\n" % (k, len(per[k]), mean[k], mean[k] * PER_STEP.get(k, 1) / step_total))
    f.write("# setup-only kernels (outside the timed region)\n")
    for k in per:
        if k not in STEP_KERNELS:
            f.write("%-30s %6d %12.1f\n" % (k, len(per[k]), mean[k]))

# ---------------------------------------------------------------- full capture
raw = os.path.join(SRC, tag + "_full_raw.csv")
if os.path.exists(raw):
    rows = read_csv(raw)
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
    last = OrderedDict()
    for r in rows[2:]:
        if len(r) == len(hdr):
            last[short(r[ix["Kernel Name"]])] = r      # keep the LAST captured instance of each kernel (steady state)

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    traffic = {}
    with open(os.path.join(HERE, tag + "_full_summary.txt"), "w") as f:
        f.write("# %s `ncu --set full --clock-control none --import-source on`, last captured launch of each hot kernel, %d sequences per launch\n" % (tag, seqs))
        for k, r in last.items():
            f.write("\n[%s]\n" % k)
            for w in want:
                if w in ix:
                    f.write("  %-82s %s %s\n" % (w, r[ix[w]], units[ix[w]]))
            if "dram__bytes_read.sum" in ix:
                rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
                wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
                traffic[k] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_per_sequence": (rd + wr) / seqs, "sequences_in_capture": seqs,
                              "source": "profiles/%s_full_summary.txt" % tag}
                def num(name):
                    return float(r[ix[name]].replace(",", "")) if name in ix else None
                traffic[k]["ncu"] = {"issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active") or num("sm__inst_issued.avg.pct_of_peak_sustained_active"),
                                     "alu_pipe_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                                     "fma_pipe_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                                     "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                                     "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                                     "warp_instructions_per_sequence": (num("smsp__inst_executed.sum") or 0) / seqs,
                                     "registers_per_thread": num("launch__registers_per_thread"),
                                     "stall_long_scoreboard": num("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                                     "stall_barrier": num("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")}
                f.write("  %-82s %.0f byte\n" % ("=> dram read+write per launch", rd + wr))
    json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
print("summaries written for", tag)

"""Turns the raw ncu outputs of `tools/gpu.sh launches convmetrics prof:<regex>` (gpurun_out/) into the small tracked summaries
under profiles/. Run in the build container after a profiling gpurun call:  python tools/summarize_profiles.py r02"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs(PROF, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("ay2::", "")
    return re.sub(r"\(.*", "", name)


def rows_of(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


# 1. launch list of the last eager step
p = os.path.join(OUT, "launches.csv")
if os.path.exists(p):
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in rows_of(p) if r.get("Metric Name") == "gpu__time_duration.sum"]
    starts = [i for i, r in enumerate(rows) if "space_to_depth" in r[0]]
    step = rows[starts[-1]:]
    agg = collections.OrderedDict()
    for n, v in step:
        agg.setdefault(short(n), [0, 0.0])
        agg[short(n)][0] += 1
        agg[short(n)][1] += v
    tot = sum(v for _, v in step)
    with open(os.path.join(PROF, f"{tag}_launches_step.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `python tools/profile_step.py 3` (eager replay of the\n")
        f.write("# benchmarked launch sequence: bs64 yolov5s 640x640, uint8 in, fused head -> NMS). Cold-cache serialised times: compare SHARES.\n")
        f.write(f"# launches in one step: {len(step)}; sum of durations {tot / 1e3:.1f} us\n")
        f.write("kernel,launches,total_us,share_pct\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{v / 1e3:.1f},{100 * v / tot:.1f}\n")
    with open(os.path.join(PROF, f"{tag}_launch_list.csv"), "w") as f:
        f.write("idx,kernel,duration_us\n")
        for i, (n, v) in enumerate(step):
            f.write(f"{i},{short(n)},{v / 1e3:.2f}\n")
    print(open(os.path.join(PROF, f"{tag}_launches_step.csv")).read())

# 2. per-launch metrics of the 52 conv launches of one step (DRAM traffic for roofline.traffic)
p = os.path.join(OUT, "conv_metrics.csv")
if not os.path.exists(p):
    p = os.path.join(OUT, "conv52.csv")
if os.path.exists(p):
    per = collections.OrderedDict()
    for r in rows_of(p):
        per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"])})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    def to_bytes(v):
        val, unit = v
        return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    def to_us(v):
        val, unit = v
        return val * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
    tot_rd = tot_wr = tot_us = tot_l2 = 0.0
    with open(os.path.join(PROF, f"{tag}_conv_launches_metrics.csv"), "w") as f:
        f.write(f"# ncu metrics (clock-control none) for the {len(per)} conv_tc_kernel / conv_chain_kernel launches of ONE bs64 yolov5s step, in launch order\n")
        f.write("idx,kernel,dur_us,dram_read_MB,dram_write_MB,l2_MB,tensor_pipe_pct,xu_pipe_pct,dram_pct\n")
        for i, (k, m) in enumerate(per.items()):
            rd, wr = to_bytes(m["dram__bytes_read.sum"]), to_bytes(m["dram__bytes_write.sum"])
            l2 = to_bytes(m["lts__t_bytes.sum"])
            us = to_us(m["gpu__time_duration.sum"])
            tot_rd += rd; tot_wr += wr; tot_us += us; tot_l2 += l2
            f.write(f"{i},{m['kernel']},{us:.2f},{rd / 1e6:.1f},{wr / 1e6:.1f},{l2 / 1e6:.1f},"
                    f"{m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0]:.1f},"
                    f"{m['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'][0]:.1f},"
                    f"{m['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]:.1f}\n")
    summary = {"launches": len(per), "dram_read_bytes": tot_rd, "dram_write_bytes": tot_wr, "dram_bytes": tot_rd + tot_wr,
               "l2_bytes": tot_l2, "sum_duration_us_under_ncu": tot_us,
               "algorithmic_bytes": 121.6e6 * 64, "note": "one bs64 step; traffic for bench.py roofline.traffic"}
    json.dump(summary, open(os.path.join(PROF, f"{tag}_conv_traffic.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1))

# 3. full captures -> raw metric CSV (small) for whatever .ncu-rep files exist
for rp in sorted(__import__("glob").glob(os.path.join(OUT, "prof_*.ncu-rep"))):
    rep = re.sub(r"[^A-Za-z0-9_]+", "_", os.path.basename(rp)[:-len(".ncu-rep")]).strip("_")
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if h in ("Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem") or re.match(
            r"(gpu__time_duration.sum|dram__bytes_(read|write).sum|lts__t_bytes.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
            r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct_of_peak_sustained_active|"
            r"sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
            r"smsp__cycles_active.avg|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|smsp__inst_executed.sum)$", h)]
    with open(os.path.join(PROF, f"{tag}_{rep}_full.csv"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); selected raw metrics, one row per launch\n")
        f.write(",".join(f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in keep) + "\n")
        for r in rows[2:]:
            f.write(",".join(short(r[i]).replace(",", " ") if hdr[i] == "Kernel Name" else r[i].replace(",", "") for i in keep) + "\n")
    print(open(os.path.join(PROF, f"{tag}_{rep}_full.csv")).read()[:1500])

"""Dev tool: condense ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py <tag> <launches.csv> <rep1.ncu-rep> [<rep2.ncu-rep> ...]

Writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launches_summary.csv (per kernel: launches, total / average
duration, DRAM bytes per launch), profiles/<tag>_ncu_full.csv (selected metrics of every `--set full` capture) and
profiles/traffic.json ({kernel: dram bytes per launch}; read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def short(name):
    return name.split("(")[0].replace("void ", "").replace("m3g::", "").strip()


def launches(tag, path):
    shutil.copy(path, os.path.join(PROF, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[ix["ID"]], {"name": short(r[ix["Kernel Name"]])})
        d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        t, u = d["gpu__time_duration.sum"]
        a[0] += 1
        a[1] += t * TIME.get(u, 1.0)
        for k, slot in (("dram__bytes_read.sum", 2), ("dram__bytes_write.sum", 3)):
            if k in d:
                a[slot] += d[k][0] * UNIT.get(d[k][1], 1.0)
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share", "avg_us", "dram_read_MB_per_launch",
                    "dram_write_MB_per_launch"])
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([n, a[0], f"{a[1]:.1f}", f"{a[1] / total:.4f}", f"{a[1] / a[0]:.1f}",
                        f"{a[2] / a[0] / 1e6:.1f}", f"{a[3] / a[0] / 1e6:.1f}"])
    return {n: (a[2] + a[3]) / a[0] for n, a in agg.items()}


def full(tag, reps):
    out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            rec = {"capture": os.path.basename(rep), "kernel": short(d.get("Kernel Name", "?"))}
            for m in METRICS:
                if m in d:
                    rec[m] = d[m] + (" " + u[m] if u.get(m) else "")
            out.append(rec)
    with open(os.path.join(PROF, f"{tag}_ncu_full.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["capture", "kernel"] + METRICS)
        for rec in out:
            w.writerow([rec.get("capture"), rec.get("kernel")] + [rec.get(m, "") for m in METRICS])
    return out


def main():
    tag, launch_csv, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(PROF, exist_ok=True)
    traffic = launches(tag, launch_csv)
    full(tag, reps)
    sys.path.insert(0, ROOT)
    import datetime

    import bench  # source_hash(): ties the capture to the CUDA sources it was taken on

    json.dump({"source": f"{tag}_launches.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch)",
               "source_hash": bench.source_hash(), "captured": datetime.date.today().isoformat(),
               "bytes_per_launch": {k: round(v) for k, v in traffic.items()}},
              open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns the .ncu-rep files a gpurun call brought back (gpurun_out/) into the small, tracked
summaries under profiles/: one CSV row per profiled launch with the metrics DESIGN.md quotes, and
roofline_traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py).

    python profiles/export_ncu.py r01 gpurun_out/r01_prof.ncu-rep gpurun_out/r01_cell.ncu-rep ...
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}          # bytes; durations in microseconds


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    start = out.index('"ID"')
    rows = list(csv.reader(io.StringIO(out[start:])))
    return rows[0], rows[1], rows[2:]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    recs = []
    for rep in reps:
        hdr, units, rows = raw_rows(rep)
        for r in rows:
            full = r[hdr.index("Kernel Name")]
            rec = {"report": os.path.basename(rep), "kernel": full.split("(const")[0].split("(hipr::")[0].strip()}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    v, u = r[i], units[i]
                    try:
                        v = float(v.replace(",", ""))
                        if u in SCALE:
                            v *= SCALE[u]
                            u = "us" if u in ("ns", "us", "ms", "s") else "byte"
                        elif u.endswith("/s") and u.split("/")[0] in SCALE:
                            v *= SCALE[u.split("/")[0]]
                            u = "byte/s"
                    except ValueError:
                        pass
                    rec[m + (" [%s]" % u if u else "")] = v
            recs.append(rec)
    keys = []
    for rec in recs:
        for k in rec:
            if k not in keys:
                keys.append(k)
    path = os.path.join(HERE, "%s_ncu_full_summary.csv" % tag)
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in recs:
            w.writerow(rec)
    print("wrote", path, len(recs), "launches")
    # dominant kernel's DRAM traffic per launch -> bench.py's roofline.traffic
    # (the plain float32 variant only: not the flat-field <.., true, ..> instantiation)
    k1 = [r for r in recs if "chansum_bulk_kernel" in r["kernel"]
          and not r["kernel"].split("chansum_bulk_kernel<")[1].split(",")[1:2] in (["1"], [" 1"], [" 1>"], [" (bool)1"])]
    if k1:
        tr = [r["dram__bytes_read.sum [byte]"] + r["dram__bytes_write.sum [byte]"] for r in k1]
        import hashlib
        src = os.path.join(os.path.dirname(HERE), "hiprfish-image-analysis_b200", "csrc", "chansum.cu")
        js = {"chansum_bytes_per_launch": sum(tr) / len(tr), "launches": len(tr), "source": k1[0]["report"],
              "workload": "2048x2048x95 float32 cube (1,593,835,520 B) -> float64 sum image", "round": tag,
              "chansum_cu_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
        with open(os.path.join(HERE, "roofline_traffic.json"), "w") as f:
            json.dump(js, f, indent=1)
        print("roofline_traffic.json", js)


if __name__ == "__main__":
    main()

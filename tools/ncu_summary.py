#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics the design decisions rest on.
usage: ncu_summary.py report.ncu-rep [rays-per-launch ...]   (with ray counts: L1 / L2 / DRAM bytes per ray, the
"L2/HBM bytes per ray" evidence BASELINE.json's north_star asks for; the n-th count belongs to the n-th kernel)"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "SM_B.TriageCompute.l1tex__t_sectors.sum", "lts__t_sectors.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    rays = [float(a) for a in sys.argv[2:]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print("==", r[hdr.index("Kernel Name")][:90])
        if n < len(rays) and rays[n] > 0:
            def per_ray(key, scale=1.0):
                if key not in hdr:
                    return float("nan")
                u = units[hdr.index(key)]
                v = to_bytes(r[hdr.index(key)], u) if "byte" in u else float(r[hdr.index(key)].replace(",", "")) * scale
                return v / rays[n]
            # --set full reports L1 and L2 traffic as 32-byte sectors (l1tex__t_sectors, lts__t_sectors)
            l1 = per_ray("SM_B.TriageCompute.l1tex__t_sectors.sum", 32.0) if "SM_B.TriageCompute.l1tex__t_sectors.sum" in hdr else per_ray("l1tex__t_sectors.sum", 32.0)
            print("  per ray (%.0f rays): L1 %.0f B, L2 %.0f B, DRAM read %.1f B + write %.1f B" % (
                rays[n], l1, per_ray("lts__t_sectors.sum", 32.0), per_ray("dram__bytes_read.sum"), per_ray("dram__bytes_write.sum")))
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "nan", "-nan")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("  stall %-40s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))


if __name__ == "__main__":
    main()

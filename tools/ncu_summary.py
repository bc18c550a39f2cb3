"""Key counters of an .ncu-rep (first matching launch) as plain text.
    python tools/ncu_summary.py report.ncu-rep [kernel substring]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors.sum", "l1tex__t_sectors_lookup_hit.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub not in name:
            continue
        print("kernel:", name.split("(")[0])
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k or h.endswith("." + k):
                    print("  %-75s %s %s" % (k, r[i], units[i]))
                    break
        print("  stall reasons (warps per issue-active cycle):")
        st = [(float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
              for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        for v, h in sorted(st, reverse=True)[:8]:
            print("    %-28s %.2f" % (h, v))
        break


main()

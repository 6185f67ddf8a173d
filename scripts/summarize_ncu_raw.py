"""Turn `ncu -i X.ncu-rep --page raw --csv` output into a short markdown summary for profiles/.

    python scripts/summarize_ncu_raw.py raw.csv "title" [algorithmic_bytes] > profiles/<name>.md
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"), ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
]


def main(path, title, alg_bytes=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nSource: `ncu --set full --clock-control none`, exported with `--page raw --csv` (`{path}`).\n")
    for vals in rows[2:]:
        print(f"## `{vals[hdr.index('Kernel Name')][:120]}`\n\n| metric | value |\n|---|---|")
        got = {}
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                got[key] = (vals[i], units[i])
                print(f"| {label} | {vals[i]} {units[i]} |")
        if alg_bytes and "dram__bytes_read.sum" in got:
            def mb(x):
                v, u = x
                v = float(v.replace(",", ""))
                return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            traffic = mb(got["dram__bytes_read.sum"]) + mb(got["dram__bytes_write.sum"])
            print(f"| **DRAM traffic (read+write)** | {traffic:.1f} MB |\n| **algorithmic bytes** | {float(alg_bytes)/1e6:.1f} MB |")
        print()


if __name__ == "__main__":
    main(*sys.argv[1:4])

"""Summarise an .ncu-rep (development tool): one block of key metrics per captured launch.

    python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed.sum", "thread instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for vals in rows[2:]:
            print(f"== {path}: {vals[col['Kernel Name']]}  (launch id {vals[col['ID']]})")
            for key, label in KEYS:
                if key in col:
                    print(f"   {label:24s} {vals[col[key]]:>16s} {units[col[key]]}")
            rd, wr = to_float(vals[col["dram__bytes_read.sum"]]), to_float(vals[col["dram__bytes_write.sum"]])
            dur = to_float(vals[col["gpu__time_duration.sum"]])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale[units[col["dram__bytes_read.sum"]]] + wr * scale[units[col["dram__bytes_write.sum"]]]
            tsc = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}[units[col["gpu__time_duration.sum"]]]
            print(f"   {'dram total':24s} {tot / 1e9:16.3f} GB   -> {tot / (dur * tsc) / 1e9:8.1f} GB/s under ncu")
            stalls = [(to_float(vals[i]), h) for h, i in col.items()
                      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
            stalls = sorted((s for s in stalls if s[0]), reverse=True)[:6]
            print("   top stalls (warps per issue): " + ", ".join(
                f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in stalls))


if __name__ == "__main__":
    main()

"""Condense an `ncu --set full` report into the CSV kept under profiles/ (one row per captured launch).

    python profiles/ncu_summary.py gpurun_out/r01_prof.ncu-rep > profiles/r01_ncu_full_summary.csv
"""
import csv
import subprocess
import sys

KEEP = [
    ("Kernel Name", "kernel"),
    ("Grid Size", "grid"),
    ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "l1_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = [(head.index(k), name) for k, name in KEEP if k in head]
    w = csv.writer(sys.stdout)
    w.writerow([name + (f" [{units[i]}]" if units[i] else "") for i, name in idx])
    for r in data:
        w.writerow([r[i] for i, _ in idx])


if __name__ == "__main__":
    main(sys.argv[1])

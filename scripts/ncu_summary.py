#!/usr/bin/env python
"""Extract the metrics the roofline discussion uses from an .ncu-rep into a small CSV (committed under profiles/).

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_x_summary.csv
"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum",
        "smsp__inst_executed_op_ldgsts.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__inst_executed.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    cols = [i for i, h in enumerate(hdr) if h in KEYS]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel"] + [f"{hdr[i]} [{units[i]}]" for i in cols])
        for r in rows[2:]:
            w.writerow([r[name_i][:90]] + [r[i] for i in cols])
    print("wrote", out, len(rows) - 2, "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

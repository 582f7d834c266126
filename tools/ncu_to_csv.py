#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the transposed CSV kept under profiles/ (metric, unit, one column per launch).

    python tools/ncu_to_csv.py gpurun_out/X.ncu-rep profiles/rN_name_ncu_raw.csv

Runs in the build container (ncu reads reports without a GPU)."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__cycles_elapsed.max",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    keep = list(KEEP) + [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(launches))])
        i = hdr.index("Kernel Name")
        w.writerow(["Kernel Name", ""] + [r[i] for r in launches])
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in launches])
    print("wrote", out, "launches:", len(launches))


if __name__ == "__main__":
    main()

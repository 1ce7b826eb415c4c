"""Pick the columns that matter out of `ncu -i X.ncu-rep --page raw --csv` dumps (one kernel per file) -> one small CSV.

Usage: python tools/summarise_ncu_full.py out.csv "label=raw.csv" ..."""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__cluster_size"]
out = csv.writer(open(sys.argv[1], "w", newline=""))
out.writerow(["capture"] + COLS)
for arg in sys.argv[2:]:
    label, path = arg.rsplit("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = [label]
    for c in COLS:
        if c in hdr:
            i = hdr.index(c)
            rec.append((vals[i] + " " + units[i]).strip())
        else:
            rec.append("")
    out.writerow(rec)

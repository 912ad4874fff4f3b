"""Reduce `ncu --page raw --csv` to our kernels and the columns the roofline argument uses.
python scripts/ncu_raw_pick.py raw.csv > profiles/xxx.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["ID", "Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [hdr.index(w) for w in want if w in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
for r in rows[2:]:
    kn = r[hdr.index("Kernel Name")]
    if "at::" not in kn and "elementwise" not in kn and "cub::" not in kn:
        r[hdr.index("Kernel Name")] = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        w.writerow([r[i] for i in idx])

"""Summarise an .ncu-rep (needs `ncu` on PATH, no GPU): per kernel duration, DRAM bytes, tensor-pipe %, L2/L1 throughput."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
idx = [(h, hdr.index(h)) for h in want if h in hdr]
for r in rows[2:]:
    print("-" * 100)
    for h, i in idx:
        print(f"{h:70s} {r[i]} {units[i]}")

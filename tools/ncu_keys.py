"""Key metrics of every kernel in an ncu report: python tools/ncu_keys.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg"]
tens = [h for h in hdr if "tensor" in h and ("pct" in h)]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90])
    for w in want + [t for t in tens if t not in want][:6]:
        if w in hdr:
            print("   %-85s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))

"""print the handful of raw metrics that decide "tail, latency or throughput" from an .ncu-rep
usage: python tools/ncu_keys.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, unit = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_bytes.sum',
        'launch__grid_size', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for val in rows[2:]:
    print("==", val[hdr.index("Kernel Name")][:60])
    for h, u, v in zip(hdr, unit, val):
        if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and float(v or 0) > 0.5:
            print(f"  {h} [{u}] {v}")

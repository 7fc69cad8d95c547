"""Post-process the ncu outputs of one profiling call into profiles/ (run here, on the CPU box).

On the GPU box (one gpurun call, ~2.5 GPU-minutes):

  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,\\
  sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,\\
  smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/rNN_metrics.csv --metrics $M \\
      python tools/profile_step.py --n 16000000 --fast 1
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/rNN_launches.csv \\
      --metrics gpu__time_duration.sum python tools/profile_step.py --n 16000000 --fast 1
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bh_group_bodies -c 1 \\
      -o gpurun_out/rNN_bh python tools/profile_step.py --n 16000000 --fast 1
  python bench.py > gpurun_out/bench_n1.json

Then:  python tools/refresh_profiles.py rNN     (writes profiles/rNN_*; needs ncu on PATH to read the .ncu-rep)
Source-level view of a capture:  ncu -i gpurun_out/rNN_bh.ncu-rep --page source --csv --print-source cuda,sass
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir, src = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")
T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}

# per-kernel counter table
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_table.py"), os.path.join(src, f"{tag}_metrics.csv")],
                       capture_output=True, text=True, check=True).stdout
with open(os.path.join(out_dir, f"{tag}_kernel_table_16M.txt"), "w") as f:
    f.write("# per-kernel ncu counters of ONE hot-path step (psim_step: polar pass and strict node centres included from\n"
            "# round 2 on), 16 M-body electrolyte, theta 1.0, parity_mode 0\n"
            "# (tools/refresh_profiles.py has the command line; table by tools/ncu_table.py)\n" + table)

# launch list
shutil.copy(os.path.join(src, f"{tag}_launches.csv"), os.path.join(out_dir, f"{tag}_launches_16M.csv"))
text = open(os.path.join(src, f"{tag}_launches.csv")).read()
rows = list(csv.DictReader(io.StringIO(text[text.index('"ID"'):])))
agg = defaultdict(lambda: [0.0, 0])
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("psim::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    agg[name][0] += float(r["Metric Value"].replace(",", "")) * T.get(r["Metric Unit"], 1e-6)
    agg[name][1] += 1
tot, n = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
bench = {}
try:
    bench_file = os.path.join(src, f"{tag}_bench_n1.json")
    bench = json.load(open(bench_file if os.path.exists(bench_file) else os.path.join(src, "bench_n1.json")))
except Exception:
    pass
lines = ["# ncu launch list of ONE hot-path step (psim_step), 16 M-body electrolyte, theta = 1.0, parity_mode 0",
         "# (cold-cache, serialised launches: compare SHARES with bench.py's phase_ms, not absolutes)", "",
         f"total {tot:.3f} ms in {n} launches" + (f"  (bench.py device-timed step: {bench['ms_per_step']:.2f} ms; phases "
                                                   f"{json.dumps({k: round(v, 2) for k, v in bench['phase_ms'].items()})})" if bench else ""),
         "", f"{'kernel':58s} {'ms':>8s} {'share':>7s} {'launches':>9s}"]
for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    lines.append(f"{k[:58]:58s} {ms:8.3f} {100 * ms / tot:6.1f}% {c:9d}")
open(os.path.join(out_dir, f"{tag}_launches_16M.txt"), "w").write("\n".join(lines) + "\n")

# full capture of the dominant kernel: details page, raw counters, DRAM traffic for bench.py's roofline.traffic
rep = os.path.join(src, f"{tag}_bh.ncu-rep")
if os.path.exists(rep) and shutil.which("ncu"):
    for page, name in (("details", f"{tag}_bh_group_bodies_16M_ncu.txt"), ("raw", f"{tag}_bh_group_bodies_16M_raw.csv")):
        res = subprocess.run(["ncu", "-i", rep, "--page", page] + (["--csv"] if page == "raw" else []),
                             capture_output=True, text=True)
        open(os.path.join(out_dir, name), "w").write(res.stdout)
    r = list(csv.reader(open(os.path.join(out_dir, f"{tag}_bh_group_bodies_16M_raw.csv"))))
    h, u, v = r[0], r[1], r[2]
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    g = lambda k: float(v[h.index(k)].replace(",", "")) * sc[u[h.index(k)]]
    json.dump({"kernel": "bh_group_bodies_kernel<false> (parity_mode 0)", "workload": "16M electrolyte theta=1.0",
               "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
               "source": f"ncu --set full --clock-control none --import-source on, profiles/{tag}_bh_group_bodies_16M_ncu.txt / _raw.csv",
               "n_bodies": 16000000}, open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)
if bench:
    json.dump(bench, open(os.path.join(out_dir, f"{tag}_bench_n1.json"), "w"))
print("profiles/ refreshed for", tag)

"""ncu --csv metric log -> per-kernel table (time-weighted averages over the launches of each kernel).
usage: python tools/ncu_table.py gpurun_out/metrics.csv [hbm_peak_gbs]"""
import csv
import io
import re
import sys
from collections import defaultdict

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6538.3
text = open(path).read()
start = text.index('"ID"')
rows = list(csv.DictReader(io.StringIO(text[start:])))
by_id = defaultdict(dict)
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("psim::", "").replace("(anonymous namespace)::", "")
    by_id[r["ID"]]["kernel"] = name
    by_id[r["ID"]][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])


def val(d, key, unit_scale=None):
    v, u = d.get(key, (0.0, ""))
    if unit_scale:
        v *= unit_scale.get(u, 1.0)
    return v


T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = defaultdict(lambda: defaultdict(float))
for d in by_id.values():
    a = agg[d["kernel"]]
    ms = val(d, "gpu__time_duration.sum", T)
    a["n"] += 1
    a["ms"] += ms
    a["dram"] += val(d, "dram__bytes_read.sum", B) + val(d, "dram__bytes_write.sum", B)
    a["l2"] += val(d, "lts__t_bytes.sum", B)
    for k, m in (("hit", "lts__t_sector_hit_rate.pct"), ("fma", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                 ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 ("tpi", "smsp__thread_inst_executed_per_inst_executed.ratio"),
                 ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        a[k] += val(d, m) * ms
    a["regs"] = max(a["regs"], val(d, "launch__registers_per_thread"))
total = sum(a["ms"] for a in agg.values())
print(f"# peaks: HBM {peak:.0f} GB/s (MEASURED_PEAKS.json); time-weighted averages over the launches of each kernel; total {total:.2f} ms")
print(f"{'kernel':46s} {'n':>3s} {'ms':>7s} {'share':>6s} {'DRAM GB/s':>9s} {'%HBM':>5s} {'L2 GB/s':>8s} {'L2hit%':>6s} {'FMA%':>5s} {'issue%':>6s} {'thr/inst':>8s} {'occ%':>5s} {'regs':>4s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    ms = a["ms"]
    if ms <= 0:
        continue
    gbs = a["dram"] / (ms * 1e-3) / 1e9
    print(f"{k[:46]:46s} {int(a['n']):3d} {ms:7.3f} {100 * ms / total:5.1f}% {gbs:9.0f} {100 * gbs / peak:5.1f} "
          f"{a['l2'] / (ms * 1e-3) / 1e9:8.0f} {a['hit'] / ms:6.1f} {a['fma'] / ms:5.1f} {a['issue'] / ms:6.1f} "
          f"{a['tpi'] / ms:8.1f} {a['occ'] / ms:5.1f} {int(a['regs']):4d}")

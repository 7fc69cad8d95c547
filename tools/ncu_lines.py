"""Per-source-line totals (stall samples, warp instructions, threads per instruction) of one ncu capture.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_lines.py x.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, out = None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit():
        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        out.append((cur, int(r[0]), r[1].strip(), num(r[4]), num(r[7]), num(r[8])))
ts, ti = sum(o[3] for o in out), sum(o[4] for o in out)
print(f"total stall samples {ts:.0f}, warp instructions {ti:.0f}")
print(f"{'file:line':28s} {'stall%':>7s} {'inst%':>7s} {'thr/inst':>8s}  source")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0] + ':' + str(o[1]):28s} {100 * o[3] / ts:7.2f} {100 * o[4] / ti:7.2f} {o[5] / max(o[4], 1):8.1f}  {o[2][:100]}")

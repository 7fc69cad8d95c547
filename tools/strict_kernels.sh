#!/bin/bash
# per-kernel device times of the strict_centres stage at 16 M bodies (ncu launch list)
for g in ${1:-electrolyte uniform}; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:strict -c 40 --csv --log-file gpurun_out/strict_$g.csv python tools/phase_times.py --n ${N:-16000000} --gen $g --strict 1 --reps 1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/strict_$g.csv")) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows[:10]:
    print("$g", r[4][:32], r[-1]); tot+=float(r[-1])
print("$g total us", tot/1e3)
PY
done

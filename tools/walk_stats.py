"""CPU statistics of the group walk's interaction lists (tests/emu emu_group_stats): how many nodes per walk are sure /
undecided / opened, with the 32 targets classified against 1, 2, 4 or 8 sub-boxes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Emu, electrolyte, uniform_pm1  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
gen = sys.argv[2] if len(sys.argv) > 2 else "electrolyte"
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
bd = dict(electrolyte=electrolyte, uniform=uniform_pm1)[gen](n)
emu = Emu()
emu.build(bd, 0)
sb = emu.sorted_bodies()
pts, rad = np.ascontiguousarray(sb[:, :2]), np.ascontiguousarray(sb[:, 3])
names = ["walks", "rounds", "visited", "sure", "undecided", "split", "opened_all", "leaves", "accepted", "reach_sure", "reach_und",
         "direct", "und_all_accept", "und_none_accept"]
for nsub in (1, 2, 4, 8):
    st = emu.group_stats(pts, rad, theta=theta, nsub=nsub).astype(np.float64)
    w = st[0]
    print(f"nsub={nsub}: " + "  ".join(f"{k}={v / w:.1f}" for k, v in zip(names[1:], st[1:])))

# bounding boxes of the 32-target groups (ideal: 32 bodies / rho = 512 A^2 -> 22.6 A square)
g = pts[: len(pts) // 32 * 32].reshape(-1, 32, 2)
ext = g.max(axis=1) - g.min(axis=1)
side = ext.max(axis=1)
print("group box longest side, quantiles 10/50/90/99/99.9 %:", np.round(np.quantile(side, [0.1, 0.5, 0.9, 0.99, 0.999]), 1))
print("mean area", float((ext[:, 0] * ext[:, 1]).mean()), " mean longest side", float(side.mean()))

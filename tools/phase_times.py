"""Per-phase device timings of the hot path (CUDA events on the library's stream)."""
import argparse
import sys
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import KE, electrolyte, uniform_pm1, clustered  # noqa: E402
from particlesim_b200 import Bodies, Simulation, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--gen", default="electrolyte")
ap.add_argument("--theta", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--fast", type=int, default=0)
ap.add_argument("--strict", type=int, default=0)
ap.add_argument("--cells", default="11.88", help="comma-separated cell sizes to time the short-range pass on")
args = ap.parse_args()

gen = dict(electrolyte=electrolyte, uniform=uniform_pm1, clustered=clustered)[args.gen]
t0 = time.time()
bd = gen(args.n)
print(f"generated {args.n} bodies in {time.time()-t0:.1f}s", flush=True)
b = Bodies(bd["pos"], vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
           species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
stream = torch.cuda.current_stream().cuda_stream
sim = Simulation(b, bd["hw"], bd["hh"], theta=args.theta, parity_mode=not args.fast, stream=stream, strict_centres=bool(args.strict))
sim.config.coulomb_constant = float(KE)
hw, hh = bd["hw"], bd["hh"]


def timed(name, fn, reps=args.reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:28s} min {min(ts):9.3f} ms  median {np.median(ts):9.3f} ms", flush=True)
    return min(ts)


C = sim._call
timed("build(CONTAINING)", lambda: C("psim_build", 0, 0.0, 0.0))
print(sim.stats())
timed("build(DOMAIN)", lambda: C("psim_build", 1, hw, hh))
for cs in [float(v) for v in args.cells.split(",")]:
    timed(f"cell_build {cs}", lambda: C("psim_cell_build", hw, hh, cs))
    timed(f"short_range(LJ) @ {cs}", lambda: C("psim_short_range", 7))
C("psim_reset_counters")
tf = timed("field+attract", lambda: C("psim_field", float(KE), 0.0, 0.0, 1, None, None))
st = sim.stats()
print("warp steps per group:", st["traversal_warp_steps"] / args.reps / ((args.n + 31) // 32))
C("psim_cell_build", hw, hh, 11.88)
timed("polar (ConjugatePair)", lambda: C("psim_apply_polar_forces", float(KE), 1))
if len(b.ebody):
    timed("update_electrons", lambda: C("psim_update_electrons", 0.0, 0.0, 5.0, float(KE)))
p = sim.step_params(do_polar=False)
timed("psim_step (no polar)", lambda: sim.step_device(p))
p = sim.step_params()
ts = timed("psim_step (full)", lambda: C("psim_step", p.__class__.from_buffer_copy(p)) if False else sim.step_device(p))
print(f"N={args.n} step {ts:.3f} ms -> {args.n/ts/1e3:.1f} Mparticles/s", flush=True)

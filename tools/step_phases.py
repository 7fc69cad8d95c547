"""Device-timed step and per-phase times of the bench workload (quick A/B runs: no oracle, no CPU baseline).

  python tools/step_phases.py [--n 16000000] [--config 4] [--steps 10] [--tag label]
Environment switches read by the library (e.g. PSIM_ONE_MUFU=0) apply as usual."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import KE  # noqa: E402
from particlesim_b200 import Bodies, Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--ieee", type=int, default=0)
ap.add_argument("--strict", type=int, default=1)
ap.add_argument("--tag", default="")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
n = args.n or cfg["n"]
bd = bench.make_workload(n, gen=cfg["gen"])
b = Bodies(bd["pos"], z=bd.get("z"), vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
           species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
stream = torch.cuda.current_stream().cuda_stream
sim = Simulation(b, bd["hw"], bd["hh"], domain_depth=float(bd.get("hd", 1.0)), theta=cfg["theta"], parity_mode=args.ieee,
                 stream=stream, strict_centres=bool(args.strict))
sim.config.coulomb_constant = float(KE)
sim.config.enable_out_of_plane = bool(bd.get("enable_out_of_plane", False))
p = sim.step_params(do_short_range=cfg["short"], do_electrons=cfg["electrons"], do_iterate=cfg["iterate"], do_polar=cfg["short"])
for _ in range(args.warmup):
    sim.step_device(p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    sim.step_device(p)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
ph, acc = np.zeros(8, np.float32), np.zeros(8, np.float64)
for _ in range(5):
    sim.step_device(p)
    sim._call("psim_phase_times", ph.ctypes.data)
    acc += ph
print(json.dumps({"tag": args.tag, "n": n, "config": args.config, "ms_per_step": round(ms, 3),
                  "Mparticles_s": round(n / ms / 1e3, 1),
                  "phase_ms": {k: round(float(v / 5), 3) for k, v in zip(bench.PHASES, acc)}}), flush=True)

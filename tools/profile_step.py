"""One hot-path step between cudaProfilerStart/Stop, for ncu (--profile-from-start off)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import KE, electrolyte, uniform_pm1, clustered  # noqa: E402
from particlesim_b200 import Bodies, Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16_000_000)
ap.add_argument("--gen", default="electrolyte")
ap.add_argument("--theta", type=float, default=1.0)
ap.add_argument("--fast", type=int, default=0)
ap.add_argument("--warm", type=int, default=1)
args = ap.parse_args()
bd = dict(electrolyte=electrolyte, uniform=uniform_pm1, clustered=clustered)[args.gen](args.n)
b = Bodies(bd["pos"], vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
           species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
sim = Simulation(b, bd["hw"], bd["hh"], theta=args.theta, parity_mode=not args.fast,
                 stream=torch.cuda.current_stream().cuda_stream)
sim.config.coulomb_constant = float(KE)
p = sim.step_params()
for _ in range(args.warm):
    sim.step_device(p)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sim.step_device(p)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", sim.stats())

"""Where the wall time of psim_step_host goes: wall clock per call, device phase times, and the same
call with parts of the traffic removed."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import KE, electrolyte  # noqa: E402
from particlesim_b200 import Bodies, Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16_000_000)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--flush", type=int, default=0, help="MB of host scratch written before each call (evicts the CPU caches)")
ap.add_argument("--fast", type=int, default=0)
args = ap.parse_args()
n = args.n
bd = electrolyte(n)
b = Bodies(bd["pos"], vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
           species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
sim = Simulation(b, bd["hw"], bd["hh"], theta=1.0, parity_mode=0 if args.fast else 1, stream=torch.cuda.current_stream().cuda_stream)
sim.config.coulomb_constant = float(KE)
params = sim.step_params()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_pos, h_vel, h_q = pin(bd["pos"]), pin(np.zeros((n, 2), np.float32)), pin(bd["charge"])
o_pos, o_vel, o_ef = (torch.empty(n, 2).pin_memory() for _ in range(3))
o_orig = torch.empty(n, dtype=torch.int32).pin_memory()
charge0 = torch.from_numpy(bd["charge"])
scratch = torch.empty(max(1, args.flush) * (1 << 18), dtype=torch.float32)


def run(name, vel=True, q=True, outs=(1, 1, 1, 1)):
    ts, ph = [], None
    for k in range(args.reps + 1):
        if args.flush:
            scratch.add_(1.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim._call("psim_step_host", C.byref(params), n, h_pos.data_ptr(), h_vel.data_ptr() if vel else None,
                  h_q.data_ptr() if q else None, o_pos.data_ptr() if outs[0] else None,
                  o_vel.data_ptr() if outs[1] else None, o_ef.data_ptr() if outs[2] else None,
                  o_orig.data_ptr())
        t1 = time.perf_counter()
        if k:
            ts.append((t1 - t0) * 1e3)
        ms = (C.c_float * 8)()
        sim._call("psim_phase_times", ms)
        ph = [round(v, 2) for v in ms]
        if outs[0]:
            h_pos.copy_(o_pos)
        else:
            sim._call("psim_download_bodies", o_pos.data_ptr(), *([None] * 10), o_orig.data_ptr())
            h_pos.copy_(o_pos)
        if outs[1]:
            h_vel.copy_(o_vel)
        h_q.copy_(charge0[o_orig.long()])
    print(f"{name:34s} wall {np.mean(ts):7.2f} ms   phases {ph}", flush=True)


run("all traffic")
run("no vel/charge in", vel=False, q=False)
run("charge in only", vel=False, q=True)
run("vel in only", vel=True, q=False)
run("no pos/vel/efield out", outs=(0, 0, 0, 1))
run("no vel/charge in, nothing out", vel=False, q=False, outs=(0, 0, 0, 1))

import sys, os, time
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import KE, electrolyte
from particlesim_b200 import Bodies, Simulation
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
polar = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bd = electrolyte(n)
b = Bodies(bd["pos"], vel=bd["vel"], mass=bd["mass"], radius=bd["radius"], charge=bd["charge"], species=bd["species"], ebody=bd["ebody"], erel=bd["erel"])
strict = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sim = Simulation(b, bd["hw"], bd["hh"], parity_mode=0, strict_centres=bool(strict), stream=torch.cuda.current_stream().cuda_stream)
sim.config.coulomb_constant = float(KE)
p = sim.step_params(do_polar=bool(polar))
ph = np.zeros(8, np.float32)
for k in range(16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.step_device(p)
    sim._call("psim_phase_times", ph.ctypes.data)
    dt = time.perf_counter() - t0
    if n <= 4_000_000 or k % 5 == 4:
        sim.download(("pos", "vel", "acc"))
        sim.download_electrons()
    st = sim.stats()
    print(k, f"{dt*1e3:8.1f} ms", "phases", np.round(ph, 2), "max|v|", np.abs(b.vel).max(), "max|acc|", np.abs(b.acc).max(), "nan pos", int(np.isnan(b.pos).sum()),
          "nan erel", int(np.isnan(b.erel).sum()), "max|erel|", np.nanmax(np.abs(b.erel)), "depth", st["max_depth"], "zero", st["zero_leaves"], "cap", st["cap_leaves"], flush=True)

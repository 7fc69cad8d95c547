"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import KE, clustered, electrolyte  # noqa: E402
from particlesim_b200 import Bodies, Simulation  # noqa: E402

for gen, n in ((electrolyte, 20_011), (clustered, 12_003)):
    bd = gen(n)
    if gen is electrolyte:
        bd["species"][:2000] = 1
    b = Bodies(bd["pos"], vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
               species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
    for mode, strict in ((1, True), (2, True), (0, True), (0, False)):
        sim = Simulation(b, bd["hw"], bd["hh"], parity_mode=mode, strict_centres=strict)
        sim.config.coulomb_constant = float(KE)
        for _ in range(2):
            sim.step_device()
        sim.sync()
        sim.collide(passes=2)
        sim.quadtree.build(sim.bodies)
        don = np.arange(0, n, 501).astype(np.uint32)
        sim.hop_alignment(don, [np.array([(d + 1) % n, (d + 7) % n], np.uint32) for d in don])
        nodes = sim.quadtree.nodes
        sim._call("psim_cell_build", bd["hw"], bd["hh"], 11.88)
        sim._neighbors(np.arange(0, n, 97), 3.96, False)
        print(gen.__name__, mode, strict, len(nodes), sim.stats()["max_depth"], flush=True)
        sim.close()
print("sanitize run done")

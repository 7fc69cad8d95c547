"""Generates tests/golden/golden.{json,npz}: small seeded inputs with the oracle's outputs.

The reference (Rust) cannot be built in this image and ships no golden vectors with inputs for this
path, so these fixtures pin the ORACLE (the reference restatement), after it passed
the reference's own known-answer tests in tests/test_oracle.py.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import KE, clustered, electrolyte, oracle_for, uniform_pm1  # noqa: E402

CASES = [
    dict(name="uniform_2k_containing", gen="uniform_pm1", n=2000, mode=0, theta=0.5, leaf=1, thread=1024),
    dict(name="electrolyte_3k_domain", gen="electrolyte", n=3000, mode=1, theta=1.0, leaf=1, thread=1024),
    dict(name="clustered_3k_leaf8", gen="clustered", n=3000, mode=0, theta=1.0, leaf=8, thread=32),
]

if __name__ == "__main__":
    gens = dict(uniform_pm1=uniform_pm1, electrolyte=electrolyte, clustered=clustered)
    arrays, meta = {}, {"seed": "0xC0FFEE", "k_e": float(KE), "cases": []}
    for case in CASES:
        b = gens[case["gen"]](case["n"])
        o = oracle_for(b, theta=case["theta"], leaf=case["leaf"], thread=case["thread"])
        o.build() if case["mode"] == 0 else o.build_with_domain(b["hw"], b["hh"])
        name = case["name"]
        for k in ("pos", "charge", "radius", "mass", "species"):
            arrays[f"{name}/{k}"] = b[k]
        arrays[f"{name}/perm"] = o.permutation()
        c = o.canonical()
        for f in ("depth", "is_leaf", "start", "end", "path_lo", "charge", "pos"):
            arrays[f"{name}/canon_{f}"] = c[f]
        e, _ = o.field(KE)
        arrays[f"{name}/e_field"] = e
        meta["cases"].append(dict(case, hw=b["hw"], hh=b["hh"]))
    # neighbour-count consumers on the clustered LithiumMetal set: the surround flags
    # after frames 10 and 13 (bodies nudged in between), and the z clamp against metal neighbours
    b = clustered(3000)
    rng = np.random.default_rng(11)
    o = oracle_for(b)
    o.update_surrounded_flags(b["hw"], b["hh"], frame=10)
    nudged = np.clip(b["pos"] + (rng.uniform(-1, 1, b["pos"].shape) * b["radius"][:, None]).astype(np.float32),
                     -b["hw"], b["hw"]).astype(np.float32)
    o.set_positions(nudged)
    o.update_surrounded_flags(b["hw"], b["hh"], frame=13)
    flags, last_pos, last_frame = o.surrounded()
    z0 = rng.uniform(-3, 3, len(nudged)).astype(np.float32)
    vz0 = rng.uniform(-1, 1, len(nudged)).astype(np.float32)
    z0[b["species"] == 1] = 0.0
    o2 = oracle_for(b)
    o2.set_bodies(b["pos"], z=z0, vz=vz0, mass=b["mass"], radius=b["radius"], charge=b["charge"], species=b["species"])
    o2.enforce_metal_z_boundaries(2.5, b["hw"], b["hh"])
    ob = o2.get_bodies()
    for k, v in dict(pos=b["pos"], radius=b["radius"], species=b["species"], mass=b["mass"], charge=b["charge"],
                     nudged=nudged, flags=flags, last_pos=last_pos, last_frame=last_frame, z0=z0, vz0=vz0,
                     z=ob["z"], vz=ob["vz"]).items():
        arrays[f"consumers/{k}"] = v
    meta["consumers"] = dict(hw=b["hw"], hh=b["hh"], max_z=2.5, frames=[10, 13])
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", len(arrays), "arrays")

"""GPU, BASELINE.json sizes: psim_field against the STRICT oracle (the reference restatement with its own
serial f32 node-centre sums, variant "") on the same bodies.

Bar (BASELINE.json north_star): relative L2 <= 1e-5 at the same opening angle, bit-exact permutation.
  * strict_centres = 1 (the benchmarked mode): asserted <= 1e-5, for the IEEE walk (parity_mode 1) and
    the fast-math walk (parity_mode 0); with parity_mode 2 the fields are the oracle's bit for bit.
  * strict_centres = 0 (f64 centre sums): reported; it differs from the reference by the reference's own
    summation noise (1e-4 class, DESIGN.md "node centres") and is only asserted <= 2e-3.
"""
import time

import numpy as np
import pytest

from helpers import KE, electrolyte, oracle_for, rel_l2, uniform_pm1
from test_gpu_tree import make_sim

pytestmark = pytest.mark.gpu

TOL = 1e-5


def device_field(bodies, mode, theta, **kw):
    sim = make_sim(bodies, theta=theta, **kw)
    if mode == 0:
        sim.quadtree.build(sim.bodies)
    else:
        sim.quadtree.build_with_domain(sim.bodies, bodies["hw"], bodies["hh"])
    sim.quadtree.field(sim.bodies, KE)
    e, ids = sim.bodies.e_field.copy(), sim.bodies.id.astype(np.int64)
    sim.close()
    return e, ids


@pytest.mark.parametrize("name,gen,n,mode,theta", [
    ("uniform_1M_theta0.5", uniform_pm1, 1_000_000, 0, 0.5),
    ("electrolyte_4M", electrolyte, 4_000_000, 1, 1.0),
    ("electrolyte_16M", electrolyte, 16_000_000, 0, 1.0),
])
def test_field_vs_strict_oracle_at_scale(cuda_device, name, gen, n, mode, theta):
    bodies = gen(n)
    t0 = time.time()
    o = oracle_for(bodies, theta=theta)
    o.build() if mode == 0 else o.build_with_domain(bodies["hw"], bodies["hh"])
    e_ref, _ = o.field(KE)
    perm = o.permutation()
    t_oracle = time.time() - t0
    del o
    out = {}
    for label, kw in (("strict+ieee", dict(strict_centres=True, parity_mode=1)),
                      ("strict+fast", dict(strict_centres=True, parity_mode=0)),
                      ("f64centres+ieee", dict(strict_centres=False, parity_mode=1))):
        e, ids = device_field(bodies, mode, theta, **kw)
        assert np.array_equal(ids, perm), f"{label}: body permutation differs from the reference's"
        assert np.all(np.isfinite(e))
        out[label] = rel_l2(e, e_ref)
    print(f"\n{name}: rel-L2 vs strict oracle {out} (oracle {t_oracle:.0f} s)")
    assert out["strict+ieee"] <= TOL, out
    assert out["strict+fast"] <= TOL, out
    assert out["f64centres+ieee"] <= 2e-3, out
    if n <= 1_000_000:  # reference-order walk on the reference's centres: the oracle's bits
        e, _ = device_field(bodies, mode, theta, strict_centres=True, parity_mode=2)
        assert np.array_equal(e, e_ref)

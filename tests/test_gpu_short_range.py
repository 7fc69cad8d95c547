"""GPU parity: cell list, neighbour queries, LJ / repulsion / stack pressure, integrator, electron
update and the fused step — through the C ABI against the oracle.  Index work is bit-exact; float
work within 1e-5 relative L2 (gather order differs from the reference's serial pair loop)."""
import numpy as np
import pytest

from helpers import KE, clustered, electrolyte, oracle_for, rel_l2, uniform_pm1
from test_gpu_tree import make_sim

pytestmark = pytest.mark.gpu
TOL = 1e-5


def both_after_build(bodies, variant="", **kw):
    sim = make_sim(bodies, **kw)
    o = oracle_for(bodies, variant=variant)
    sim.quadtree.build(sim.bodies)
    o.build()
    assert np.array_equal(o.permutation(), sim.bodies.id.astype(np.int64))
    return sim, o


@pytest.mark.parametrize("cell_size", [11.88, 3.96, 50.0])
def test_cell_list_contents(cuda_device, cell_size):
    bodies = clustered(30_000)
    # a few bodies outside the domain: coord() clamps them into the border cells
    bodies["pos"][:5] *= 1.5
    sim, o = both_after_build(bodies)
    hw, hh = bodies["hw"], bodies["hh"]
    sim.cell_list.update_domain_size(hw, hh)
    sim.cell_list.cell_size = cell_size
    sim.cell_list.rebuild(sim.bodies)
    o.cell_set_domain(hw, hh)
    o.cell_rebuild(cell_size)
    gx, gy, off, idx = sim.cell_list.cells()
    assert (gx, gy) == o.cell_dims()
    rng = np.random.default_rng(3)
    nonempty = np.nonzero(np.diff(off))[0]
    for c in np.concatenate([rng.choice(nonempty, 300), rng.integers(0, gx * gy, 100)]):
        assert np.array_equal(idx[off[c]:off[c + 1]].astype(np.int64), o.cell_contents(int(c)))
    assert off[-1] == len(bodies["pos"])
    sim.close()


def test_neighbor_queries(cuda_device):
    bodies = clustered(30_000)
    sim, o = both_after_build(bodies)
    hw, hh = bodies["hw"], bodies["hh"]
    sim.cell_list.update_domain_size(hw, hh)
    sim.cell_list.cell_size = 11.88
    sim.cell_list.rebuild(sim.bodies)
    o.cell_set_domain(hw, hh)
    o.cell_rebuild(11.88)
    rng = np.random.default_rng(5)
    q = rng.choice(len(bodies["pos"]), 400, replace=False)
    for cutoff in (3.96, 7.5, 30.0):
        got = sim._neighbors(q, cutoff, False)
        for i, g in zip(q, got):
            assert np.array_equal(g, o.cell_neighbors(int(i), cutoff))  # same order as the reference
            assert set(g) == set(o.tree_neighbors(int(i), cutoff))        # the tree query gives the same set
        metal = sim._neighbors(q[:100], cutoff, True)
        for i, g in zip(q[:100], metal):
            assert len(g) == o.cell_metal_neighbor_count(int(i), cutoff)
    assert sim.cell_list.metal_neighbor_count(sim.bodies, int(q[0]), 3.96) == o.cell_metal_neighbor_count(int(q[0]), 3.96)
    sim.close()


def test_lj_forces(cuda_device):
    from particlesim_b200 import forces
    bodies = clustered(40_000)
    sim = make_sim(bodies)
    o = oracle_for(bodies)
    hw, hh = bodies["hw"], bodies["hh"]
    sim.reset_acc()
    forces.prepare_spatial_structures(sim)
    forces.attract(sim)
    coul = sim.bodies.acc.copy()
    forces.apply_lj_forces(sim)
    o.reset_acc()
    o.prepare_spatial_structures(hw, hh)
    o.attract(KE)
    o.apply_lj_forces(True)
    ob = o.get_bodies()
    assert np.array_equal(ob["id"], sim.bodies.id)
    lj_dev = sim.bodies.acc - coul
    assert np.abs(lj_dev).max() > 0, "the clustered set must exercise LJ pairs"
    assert rel_l2(sim.bodies.acc, ob["acc"]) <= TOL
    sim.close()


def test_repulsion_and_stack_pressure(cuda_device):
    from particlesim_b200 import SimConfig, default_species_table, forces
    bodies = electrolyte(30_000)
    table = default_species_table()
    for sp in (0, 3, 4, 5):
        table["repulsion_enabled"][sp] = 1
    cfg = SimConfig(stack_pressure_enabled=True, stack_pressure=0.5, stack_pressure_decay=30.0)
    sim = make_sim(bodies, config=cfg, species_table=table)
    sim.config.coulomb_constant = float(KE)
    o = oracle_for(bodies)
    o.set_species_table(table)
    hw, hh = bodies["hw"], bodies["hh"]
    sim.reset_acc()
    forces.prepare_spatial_structures(sim)
    forces.attract(sim)
    forces.apply_lj_forces(sim)
    forces.apply_repulsive_forces(sim)
    forces.apply_stack_pressure(sim)
    o.reset_acc()
    o.prepare_spatial_structures(hw, hh)
    o.attract(KE)
    base = o.get_bodies()["acc"].copy()
    o.apply_lj_forces(True)
    o.apply_repulsive_forces(True)
    o.apply_stack_pressure(True, 0.5, 30.0, hw)
    ob = o.get_bodies()
    assert np.abs(ob["acc"] - base).max() > 0
    assert rel_l2(sim.bodies.acc, ob["acc"]) <= TOL
    sim.close()


@pytest.mark.parametrize("enable_z", [False, True])
def test_iterate(cuda_device, enable_z):
    from particlesim_b200 import SimConfig, forces
    bodies = electrolyte(30_000)
    rng = np.random.default_rng(11)
    bodies["vel"] = rng.normal(0, 2.0, bodies["pos"].shape).astype(np.float32)  # fast enough to hit the walls
    z = rng.uniform(-1, 1, len(bodies["pos"])).astype(np.float32)
    vz = rng.normal(0, 0.2, len(bodies["pos"])).astype(np.float32)
    cfg = SimConfig(damping_base=0.98, enable_out_of_plane=enable_z)
    from particlesim_b200 import Bodies, Simulation
    b = Bodies(bodies["pos"], z=z, vel=bodies["vel"], vz=vz, mass=bodies["mass"], radius=bodies["radius"],
               charge=bodies["charge"], species=bodies["species"])
    sim = Simulation(b, bodies["hw"], bodies["hh"], domain_depth=1.0, dt=5.0, config=cfg)
    sim.config.coulomb_constant = float(KE)
    o = oracle_for(bodies)
    o.set_bodies(bodies["pos"], z=z, vel=bodies["vel"], vz=vz, mass=bodies["mass"], radius=bodies["radius"],
                 charge=bodies["charge"], species=bodies["species"])
    hw, hh = bodies["hw"], bodies["hh"]
    sim.reset_acc()
    forces.prepare_spatial_structures(sim)
    forces.attract(sim)
    sim.iterate()
    o.reset_acc()
    o.prepare_spatial_structures(hw, hh)
    o.attract(KE)
    o.iterate(5.0, 0.98, hw, hh, 1.0, enable_z)
    ob = o.get_bodies()
    sim.download(("pos", "vel", "z", "vz"))
    assert np.abs(sim.bodies.pos[:, 0]).max() <= hw and np.abs(sim.bodies.pos[:, 1]).max() <= hh
    assert rel_l2(sim.bodies.pos, ob["pos"]) <= 1e-6
    assert rel_l2(sim.bodies.vel, ob["vel"]) <= TOL
    if enable_z:
        assert rel_l2(sim.bodies.z, ob["z"]) <= TOL and rel_l2(sim.bodies.vz, ob["vz"]) <= TOL
    else:
        assert np.array_equal(sim.bodies.z, ob["z"])
    sim.close()


def test_update_electrons(cuda_device):
    bodies = electrolyte(30_000)
    hw, hh = bodies["hw"], bodies["hh"]
    sim = make_sim(bodies)
    sim.background_e_field = (0.003, 0.001)
    sim.quadtree.build_with_domain(sim.bodies, hw, hh)
    sim.update_electrons()
    o = oracle_for(bodies)
    o.build_with_domain(hw, hh)
    o.update_electrons((0.003, 0.001), 5.0, KE, threads=0)
    ebody, erel, evel = o.get_electrons()
    assert np.array_equal(sim.bodies.ebody, ebody)
    assert rel_l2(sim.bodies.erel, erel) <= TOL
    assert rel_l2(sim.bodies.evel, evel) <= TOL
    sim.close()


def test_fused_step_matches_the_call_sequence(cuda_device):
    """psim_step == reset_acc, prepare_spatial_structures, attract, polar, LJ, repulsion, stack pressure,
    iterate, build_with_domain, update_electrons (simulation.rs:1000-1196) on the oracle"""
    bodies = electrolyte(40_000)
    # make some bodies LJ species so the short-range pass has work
    sel = np.arange(0, 4000)
    bodies["species"][sel] = 1
    bodies["charge"][sel] = 0.0
    bodies["radius"][sel] = 1.52
    bodies["mass"][sel] = 6.94
    keep = ~np.isin(bodies["ebody"], sel)
    bodies["ebody"], bodies["erel"] = bodies["ebody"][keep], bodies["erel"][keep]
    hw, hh = bodies["hw"], bodies["hh"]
    sim = make_sim(bodies)
    sim.step_device()
    sim.download(("pos", "vel", "acc", "e_field"))
    sim.download_electrons()
    orig = np.zeros(len(sim.bodies), np.uint32)
    sim._call("psim_download_bodies", *([None] * 11), orig.ctypes.data)
    o = oracle_for(bodies)
    o.reset_acc()
    o.prepare_spatial_structures(hw, hh)
    o.attract(KE)
    o.apply_polar_forces(KE, True, 1)  # Simulation::step runs it every step (simulation.rs:1007); do_polar defaults to 1
    o.apply_lj_forces(True)
    o.apply_repulsive_forces(True)
    o.iterate(5.0, 1.0, hw, hh, 1.0, False)
    o.build_with_domain(hw, hh)
    o.update_electrons((0.0, 0.0), 5.0, KE, threads=0)
    ob = o.get_bodies()
    assert np.array_equal(ob["id"], orig.astype(np.uint64))
    assert rel_l2(sim.bodies.pos, ob["pos"]) <= 1e-6
    assert rel_l2(sim.bodies.vel, ob["vel"]) <= TOL
    assert rel_l2(sim.bodies.acc, ob["acc"]) <= TOL
    ebody, erel, evel = o.get_electrons()
    sb = np.zeros(len(ebody), np.uint32)
    sr = np.zeros((len(ebody), 2), np.float32)
    sv = np.zeros((len(ebody), 2), np.float32)
    sim._call("psim_download_electrons", sb.ctypes.data, sr.ctypes.data, sv.ctypes.data)
    assert np.array_equal(sb, ebody)
    assert rel_l2(sr, erel) <= TOL and rel_l2(sv, evel) <= TOL
    sim.close()


@pytest.mark.parametrize("dipole_model,parity_mode", [(1, 1), (0, 1), (1, 0)])
def test_polar_forces(cuda_device, dipole_model, parity_mode):
    """forces.rs:52-175 (the pass between attract and LJ in Simulation::step), gather form vs the serial loop;
    parity_mode 0 = the MUFU rsqrt / rcp arithmetic of the benchmarked mode on the same pair terms"""
    from particlesim_b200 import forces
    bodies = electrolyte(30_000)
    hw, hh = bodies["hw"], bodies["hh"]
    sim = make_sim(bodies, parity_mode=parity_mode)
    o = oracle_for(bodies)
    sim.reset_acc()
    forces.prepare_spatial_structures(sim)
    forces.attract(sim)
    before = sim.bodies.acc.copy()
    forces.apply_polar_forces(sim, dipole_model)
    o.reset_acc()
    o.prepare_spatial_structures(hw, hh)
    o.attract(KE)
    base = o.get_bodies()["acc"].copy()
    o.apply_polar_forces(KE, True, dipole_model)
    ob = o.get_bodies()
    assert np.array_equal(ob["id"], sim.bodies.id)
    d_dev, d_ref = sim.bodies.acc - before, ob["acc"] - base
    assert np.abs(d_ref).max() > 0
    assert rel_l2(d_dev, d_ref) <= 1e-4  # the increment alone (differences of nearly equal fields)
    assert rel_l2(sim.bodies.acc, ob["acc"]) <= TOL
    sim.close()


def test_fused_step_with_polar_forces(cuda_device):
    bodies = electrolyte(30_000)
    hw, hh = bodies["hw"], bodies["hh"]
    sim = make_sim(bodies)
    sim.step_device(sim.step_params(do_polar=True))
    sim.download(("pos", "vel", "acc"))
    orig = np.zeros(len(sim.bodies), np.uint32)
    sim._call("psim_download_bodies", *([None] * 11), orig.ctypes.data)
    o = oracle_for(bodies)
    o.reset_acc()
    o.prepare_spatial_structures(hw, hh)
    o.attract(KE)
    o.apply_polar_forces(KE, True, 1)
    o.apply_lj_forces(True)
    o.apply_repulsive_forces(True)
    o.iterate(5.0, 1.0, hw, hh, 1.0, False)
    o.build_with_domain(hw, hh)
    o.update_electrons((0.0, 0.0), 5.0, KE, threads=0)
    ob = o.get_bodies()
    assert np.array_equal(ob["id"], orig.astype(np.uint64))
    assert rel_l2(sim.bodies.acc, ob["acc"]) <= TOL
    assert rel_l2(sim.bodies.vel, ob["vel"]) <= TOL
    assert rel_l2(sim.bodies.pos, ob["pos"]) <= 1e-6
    sim.close()


@pytest.mark.parametrize("do_electrons", [True, False])
def test_hosted_step_equals_update_step_download(cuda_device, do_electrons):
    """psim_step_host (pipelined copies) == psim_update_state + psim_step + psim_download_bodies, bit for
    bit per body, over consecutive steps fed with their own outputs.  The hosted step returns rows in the
    order of the step's first build, the sequence in the order the step leaves on the device: compare by
    original index."""
    bodies = electrolyte(60_000)
    n = len(bodies["pos"])
    a, b = make_sim(bodies), make_sim(bodies)
    rng = np.random.default_rng(5)
    # per-body state, by original index
    pos0 = bodies["pos"].copy()
    vel0 = rng.normal(0, 0.02, (n, 2)).astype(np.float32)
    q0 = bodies["charge"].copy()
    a_rows = np.arange(n)  # original index of each row, as each side's host sees it
    b_rows = np.arange(n)
    for step in range(4):
        # sequence on context a (rows in a's device order)
        pa, va, qa = (np.ascontiguousarray(v[a_rows]) for v in (pos0, vel0, q0))  # keep the arrays alive
        a._call("psim_update_state", n, pa.ctypes.data, va.ctypes.data, qa.ctypes.data)
        a.step_device(a.step_params(do_electrons=do_electrons))
        o_pos, o_vel, o_e = (np.zeros((n, 2), np.float32) for _ in range(3))
        o_orig = np.zeros(n, np.uint32)
        a._call("psim_download_bodies", o_pos.ctypes.data, None, o_vel.ctypes.data, *([None] * 7), o_e.ctypes.data,
                o_orig.ctypes.data)
        # hosted step on context b (rows in the order of b's previous outputs)
        out = b.step_host(np.ascontiguousarray(pos0[b_rows]), np.ascontiguousarray(vel0[b_rows]),
                          np.ascontiguousarray(q0[b_rows]), params=b.step_params(do_electrons=do_electrons))
        assert sorted(out["orig"].tolist()) == list(range(n))
        ia, ib = np.argsort(o_orig), np.argsort(out["orig"])
        assert np.array_equal(out["pos"][ib], o_pos[ia])
        assert np.array_equal(out["vel"][ib], o_vel[ia])
        assert np.array_equal(out["e_field"][ib], o_e[ia])
        if not do_electrons:
            assert np.array_equal(out["orig"], o_orig)  # no second sort: the same row order
        # the host changes the state between steps (collisions / foils would): by original index
        pos0[o_orig] = o_pos
        vel0[o_orig] = o_vel * np.float32(0.5)
        flip = rng.integers(0, n, 50)
        q0[flip] = -q0[flip]
        a_rows, b_rows = o_orig.astype(np.int64), out["orig"].astype(np.int64)
    if do_electrons:
        # both contexts end in the same device order with the same electrons
        for sim in (a, b):
            sim.download_electrons()
        assert np.array_equal(a.bodies.erel, b.bodies.erel) and np.array_equal(a.bodies.evel, b.bodies.evel)
        d_orig = np.zeros(n, np.uint32)
        b._call("psim_download_bodies", *([None] * 11), d_orig.ctypes.data)
        assert np.array_equal(d_orig, o_orig)
    a.close(), b.close()


def test_hand_derived_cases_on_the_device(cuda_device):
    """the hand-derived polar (ion next to a dipole, two facing dipoles) and first-five-metals cases of
    tests/test_oracle.py, straight through the C ABI: the device against the numbers derived by hand"""
    from particlesim_b200 import Bodies, Simulation
    k = float(KE)

    def polar(pos, charge, radius, mass, species, ebody, erel):
        b = Bodies(np.array(pos, np.float32), mass=mass, radius=radius, charge=charge, species=np.array(species, np.uint8),
                   ebody=np.array(ebody, np.uint32), erel=np.array(erel, np.float32))
        sim = Simulation(b, 50.0, 50.0)
        sim.config.coulomb_constant = k
        sim.reset_acc()
        sim._call("psim_cell_build", 50.0, 50.0, 7.5)
        sim._call("psim_apply_polar_forces", k, 1)
        sim.download(("acc",))
        acc = sim.bodies.acc.copy()
        sim.close()
        return acc

    acc = polar([[0, 0], [4, 0]], [0, 1], [2.5, 0.76], [88.06, 6.94], [4, 0], [0], [[0.5, 0.0]])
    force = (-4.0 * k / (20.0 * 4.0) - (-3.5 * k / (16.25 * 3.5))) * 0.8
    assert np.allclose(acc[0], [force / 88.06, 0.0], rtol=2e-6, atol=1e-12)
    assert np.allclose(acc[1], [-force / 6.94, 0.0], rtol=2e-6, atol=1e-12)
    acc = polar([[0, 0], [4, 0]], [0, 0], [2.5, 2.5], [88.06, 88.06], [4, 4], [0, 1], [[0.5, 0.0], [-0.5, 0.0]])
    q = 0.8
    nn, ne = -4.0 * q * k / (29.0 * 5.0), -3.5 * q * k / (16.25 * 3.5)
    en, ee = ne, -3.0 * q * k / (13.0 * 3.0)
    total_a = 2.0 * ((nn - ne) - (en - ee)) * q / 88.06
    assert np.allclose(acc[0], [total_a, 0.0], rtol=5e-6, atol=1e-12)
    assert np.allclose(acc[1], [-total_a, 0.0], rtol=5e-6, atol=1e-12)
    # first five metal neighbours (out_of_plane.rs:196-200)
    pos = np.array([[3, 3], [0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1], [0, 2]], np.float32)
    b = Bodies(pos, z=[3.0, 0, 0.5, -0.5, 1.0, 0.2, -3, -3], vz=[0.7] + [0] * 7, mass=[88.06] + [6.94] * 7,
               radius=[2.5] + [1.52] * 7, charge=np.zeros(8), species=np.array([4] + [1] * 7, np.uint8))
    sim = Simulation(b, 50.0, 50.0)
    sim.enforce_metal_z_boundaries(5.0)
    assert sim.bodies.z[0] == np.float32(np.float32(np.float32(-0.5) + np.float32(1.52)) + np.float32(0.01))
    assert sim.bodies.vz[0] == 0.0
    sim.close()


def test_hosted_step_after_a_reupload_uses_upload_order(cuda_device):
    """psim_step_host keeps a row map after a step with electrons (the device order then differs from the order of its
    outputs).  psim_upload_bodies starts over: the next hosted call takes its inputs in upload order, like a fresh
    context (a stale map would scatter positions, charges and velocities onto the wrong rows without any error)."""
    bodies = electrolyte(30_000)
    n = len(bodies["pos"])
    rng = np.random.default_rng(9)
    vel = rng.normal(0, 0.02, (n, 2)).astype(np.float32)
    a, b = make_sim(bodies), make_sim(bodies)
    p = a.step_params(do_electrons=True)
    a.step_host(bodies["pos"].copy(), vel.copy(), bodies["charge"].copy(), params=p)  # leaves a row map behind
    a.upload()                                                                          # same bodies, upload order again
    out_a = a.step_host(bodies["pos"].copy(), vel.copy(), bodies["charge"].copy(), params=p)
    out_b = b.step_host(bodies["pos"].copy(), vel.copy(), bodies["charge"].copy(), params=b.step_params(do_electrons=True))
    for k in ("orig", "pos", "vel", "e_field"):
        assert np.array_equal(out_a[k], out_b[k]), k
    a.close(), b.close()

"""GPU parity: keys, sort permutation, node topology, leaf membership, aggregates — through the C ABI,
against the oracle on the same seeded inputs.  Bit-exact for everything integer; charge and mass
bit-exact; node centres to rounding (see DESIGN.md "node centres")."""
import os

import numpy as np
import pytest

from helpers import (KE, Emu, assert_same_topology, canonical_from_nodes, check_next_pointers, clustered,
                     big_integer, electrolyte, fractional, oracle_for, rel_l2, uniform_pm1)

pytestmark = pytest.mark.gpu


def make_sim(bodies, **kw):
    from particlesim_b200 import Bodies, Simulation
    b = Bodies(bodies["pos"], vel=bodies.get("vel"), mass=bodies.get("mass"), radius=bodies.get("radius"),
               charge=bodies.get("charge"), species=bodies.get("species"), ebody=bodies.get("ebody"),
               erel=bodies.get("erel"))
    sim = Simulation(b, bodies["hw"], bodies["hh"], **kw)
    sim.config.coulomb_constant = float(KE)
    return sim


def build_both(bodies, mode, leaf=1, thread=1024, variant=""):
    sim = make_sim(bodies, leaf_capacity=leaf, thread_capacity=thread, strict_centres=(variant != "hp"))
    o = oracle_for(bodies, leaf=leaf, thread=thread, variant=variant)
    if mode == 0:
        sim.quadtree.build(sim.bodies)
        o.build()
    else:
        sim.quadtree.build_with_domain(sim.bodies, bodies["hw"], bodies["hh"])
        o.build_with_domain(bodies["hw"], bodies["hh"])
    return sim, o


CASES = [
    ("uniform_1", lambda: uniform_pm1(1)), ("uniform_2", lambda: uniform_pm1(2)),
    ("uniform_3", lambda: uniform_pm1(3)), ("uniform_33", lambda: uniform_pm1(33)),
    ("uniform_4097", lambda: uniform_pm1(4097)), ("uniform_100k", lambda: uniform_pm1(100_000)),
    ("electrolyte_50k", lambda: electrolyte(50_000)), ("clustered_60k", lambda: clustered(60_000)),
    ("fractional_70k", lambda: fractional(70_000)), ("bigint_4097", lambda: big_integer(4097)),
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name,gen", CASES)
def test_topology_permutation_aggregates(cuda_device, name, gen, mode):
    bodies = gen()
    sim, o = build_both(bodies, mode)
    n = len(bodies["pos"])
    # keys: sorted, and equal to the fp32-recurrence replay of the emulation
    keys = sim.quadtree.keys()
    assert np.all(keys[:-1] <= keys[1:])
    emu = Emu()
    emu.build(bodies, mode)
    assert np.array_equal(keys, emu.keys)
    # permutation: element for element (C = 1, duplicate-free)
    assert np.array_equal(sim.bodies.id.astype(np.int64), o.permutation())
    assert np.array_equal(sim.bodies.pos, o.get_bodies()["pos"])
    # topology
    nodes = sim.quadtree.nodes
    check_next_pointers(nodes)
    dc, oc = canonical_from_nodes(nodes), o.canonical()
    assert_same_topology(dc, oc)
    st = sim.stats()
    assert st["reference_nodes"] == len(oc) and st["max_depth"] == o.max_depth()
    # both ways of making the node charges are exercised: exact integer prefix differences for the reference's
    # integer charges, bottom-up level sweeps for the fractional set
    info = sim.build_info()
    assert info["integer_charges"] == (not name.startswith(("fractional", "bigint")) and
                                       os.environ.get("PSIM_INTEGER_CHARGES") != "0")
    assert info["charged_bodies"] == int(np.count_nonzero(bodies["charge"]))
    # aggregates
    assert np.array_equal(dc["charge"], oc["charge"])
    assert np.array_equal(dc["mass"], oc["mass"])
    leaf = oc["is_leaf"] == 1
    assert np.array_equal(dc["pos"][leaf], oc["pos"][leaf])
    scale = max(1.0, float(np.abs(bodies["pos"]).max()))
    assert np.abs(dc["pos"].astype(np.float64) - oc["pos"]).max() <= 2e-6 * scale * max(1.0, np.sqrt(n) / 30)
    sim.close()


@pytest.mark.parametrize("name,gen", [("uniform_100k", lambda: uniform_pm1(100_000)),
                                      ("clustered_60k", lambda: clustered(60_000))])
def test_centres_equal_f64_oracle(cuda_device, name, gen):
    """against the oracle variant that sums node centres in f64, centres agree to 1 ulp"""
    bodies = gen()
    sim, o = build_both(bodies, 0, variant="hp")
    dc, oc = canonical_from_nodes(sim.quadtree.nodes), o.canonical()
    assert_same_topology(dc, oc)
    a, b = dc["pos"].astype(np.float64), oc["pos"].astype(np.float64)
    ulp = np.spacing(np.abs(oc["pos"]).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(a - b) <= ulp)
    sim.close()


@pytest.mark.parametrize("leaf,thread", [(8, 32), (4, 1024), (1, 1), (3, 2), (16, 5), (2, 1024)])
@pytest.mark.parametrize("mode", [0, 1])
def test_leaf_and_thread_capacity(cuda_device, leaf, thread, mode):
    bodies = clustered(20_000)
    sim, o = build_both(bodies, mode, leaf, thread)
    dc, oc = canonical_from_nodes(sim.quadtree.nodes), o.canonical()
    assert_same_topology(dc, oc)
    # leaf membership as sets (order inside a multi-body leaf is the unstable partition's)
    operm, dperm = o.permutation(), sim.bodies.id.astype(np.int64)
    lm = oc["is_leaf"] == 1
    for a, b in zip(oc["start"][lm], oc["end"][lm]):
        assert set(operm[a:b]) == set(dperm[a:b])
    assert np.array_equal(dc["charge"], oc["charge"])
    assert np.allclose(dc["mass"], oc["mass"], rtol=1e-6, atol=0)
    sim.close()


def test_coincident_bodies_make_refused_leaves(cuda_device):
    """coincident bodies beyond the leaf capacity stay one zero-aggregate leaf"""
    b = uniform_pm1(1000)
    b["pos"][10:15] = b["pos"][10]
    b["pos"][500] = b["pos"][501]
    for leaf, thread in [(1, 1024), (4, 1024), (8, 4)]:
        sim, o = build_both(b, 0, leaf, thread)
        dc, oc = canonical_from_nodes(sim.quadtree.nodes), o.canonical()
        assert_same_topology(dc, oc)
        assert np.array_equal(dc["charge"], oc["charge"])
        assert o.flags() & 2
        assert sim.stats()["zero_leaves"] >= 1
        sim.close()


def test_all_bodies_identical_and_empty(cuda_device):
    b = uniform_pm1(64)
    b["pos"][:] = b["pos"][0]
    for mode in (0, 1):
        sim, o = build_both(b, mode)
        dc, oc = canonical_from_nodes(sim.quadtree.nodes), o.canonical()
        assert_same_topology(dc, oc)
        assert len(dc) == 1 and dc["charge"][0] == 0.0 and dc["mass"][0] == 0.0
        sim.close()
    # empty input: build is a no-op (quadtree.rs:155-158)
    from particlesim_b200 import Bodies, Simulation
    sim = Simulation(Bodies(np.zeros((0, 2), np.float32)), 10.0, 10.0)
    sim.quadtree.build(sim.bodies)
    assert sim.stats()["compact_nodes"] == 0 and len(sim.quadtree.nodes) == 0
    sim.quadtree.field(sim.bodies, KE)
    sim.close()


def test_adjacent_floats_deep_chain(cuda_device):
    """two bodies one ulp apart: the fp32 centre recurrence decides the depth, not a fixed grid"""
    b = uniform_pm1(3)
    b["pos"][:] = [[100, 100], [np.nextafter(np.float32(100), np.float32(200)), 100], [5, 5]]
    for half in (300.0, 8000.0):
        b["hw"] = b["hh"] = half
        sim, o = build_both(b, 1)
        assert_same_topology(canonical_from_nodes(sim.quadtree.nodes), o.canonical())
        assert sim.stats()["max_depth"] == o.max_depth() >= 25
        sim.close()


def test_single_body_degenerate_parameters(cuda_device):
    """src/body/tests/anion.rs:48-49: Quadtree::new(0.5, 0.01, 1, 1) builds on one body"""
    b = uniform_pm1(1)
    sim, o = build_both(b, 0, 1, 1)
    dc, oc = canonical_from_nodes(sim.quadtree.nodes), o.canonical()
    assert_same_topology(dc, oc)
    assert dc["charge"][0] == oc["charge"][0] == 0.0  # thread_capacity 1: the leaf is never aggregated
    sim.close()


def test_node_arena_overflow_is_an_error(cuda_device):
    from particlesim_b200 import PsimError
    bodies = clustered(5000)
    with pytest.raises(PsimError) as e:
        sim = make_sim(bodies, node_factor=1.0)
        sim._cfg  # noqa: B018
        # node_factor 1.0 => arena of n + 1024 nodes < 1.7 n
        sim.quadtree.build(sim.bodies)
    assert e.value.code == -4


def test_fused_step_survives_arena_overflow(cuda_device):
    """psim_step does not synchronise: on an arena overflow the field passes must walk nothing (not stale
    records) and the error must be reported by the next status call"""
    from particlesim_b200 import PsimError
    bodies = clustered(5000)
    for mode in (1, 2):
        sim = make_sim(bodies, node_factor=1.0, parity_mode=mode)
        sim.step_device(sim.step_params(do_electrons=False))
        sim.sync()
        with pytest.raises(PsimError) as e:
            sim._call("psim_build_status")
        assert e.value.code == -4
        sim.close()


@pytest.mark.parametrize("clump", [40, 3000])
def test_dense_subcell_clumps_sort_by_the_full_key(cuda_device, clump):
    """bodies closer than root / 65536 share the upper key word: the radix passes leave them in input order
    and the run fix-up (insertion sort for short runs, CTA radix sort for long ones) must finish the job"""
    rng = np.random.default_rng(5)
    b = uniform_pm1(6000 + clump)
    n = len(b["pos"])
    # a clump of distinct positions inside one 0.1 A cell of a 16 000 A domain (cell of the upper word: 0.24 A)
    base = np.array([1000.3, -2000.7], np.float32)
    grid = np.stack(np.meshgrid(np.arange(64), np.arange(64)), -1).reshape(-1, 2)[:clump]
    b["pos"][:clump] = base + (grid * np.float32(0.1 / 64)).astype(np.float32)
    sel = rng.permutation(n)
    for k in ("pos", "charge", "radius", "mass", "species"):
        if k in b and b[k] is not None:
            b[k] = np.ascontiguousarray(b[k][sel])
    b["hw"] = b["hh"] = 8000.0
    assert len(np.unique(b["pos"], axis=0)) == n
    for mode in (0, 1):
        sim, o = build_both(b, mode)
        keys = sim.quadtree.keys()
        assert np.all(keys[:-1] <= keys[1:])
        assert np.array_equal(sim.bodies.id.astype(np.int64), o.permutation())
        assert_same_topology(canonical_from_nodes(sim.quadtree.nodes), o.canonical())
        sim.close()


def test_c_client(cuda_device):
    """the plain-C client (tests/c/smoke.c, gcc -std=c99) gives the numbers of the Python path: same permutation,
    same sorted positions, same field bits"""
    import subprocess
    from test_abi import build_c_smoke
    exe = build_c_smoke()
    n = 300
    res = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.strip().splitlines()
    assert lines[-1] == "ok"
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[1:1 + n]])
    perm = rows[:, 0].astype(np.int64)
    assert np.array_equal(np.sort(perm), np.arange(n))
    # the same bodies through the Python mirror: positions regenerated from the C client's output
    pos = np.zeros((n, 2), np.float32)
    pos[perm] = rows[:, 1:3].astype(np.float32)
    idx = np.arange(n)
    q = np.where(idx % 2 == 1, -1.0, 1.0).astype(np.float32)
    bodies = dict(pos=pos, charge=q, radius=np.where(idx % 2 == 1, 2.0, 0.76).astype(np.float32),
                  mass=np.where(idx % 2 == 1, 145.0, 6.94).astype(np.float32), hw=50.0, hh=50.0)
    sim = make_sim(bodies, theta=0.5)
    sim.quadtree.build(sim.bodies)
    sim.quadtree.field(sim.bodies, KE)
    assert np.array_equal(sim.bodies.id.astype(np.int64), perm)
    assert np.array_equal(sim.bodies.e_field, rows[:, 3:5].astype(np.float32))
    o = oracle_for(bodies, theta=0.5)
    o.build()
    e, _ = o.field(KE)
    assert rel_l2(sim.bodies.e_field, e) <= 1e-5
    sim.close()

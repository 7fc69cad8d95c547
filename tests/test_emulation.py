"""CPU: the device construction logic (psim_core.cuh + tree_logic.cuh, compiled as plain C++ by
tests/emu) and the warp-lockstep traversal scheme, against the oracle.  This checks the ALGORITHM
before GPU time is spent; the GPU tests check the kernels themselves through the C ABI."""
import numpy as np
import pytest

from helpers import (KE, Emu, assert_same_topology, canonical_from_nodes, check_next_pointers, clustered,
                     electrolyte, oracle_for, rel_l2, uniform_pm1)


def run(bodies, mode, leaf=1, thread=1024, theta=1.0, variant="hp", slab=0):
    emu = Emu()
    o = oracle_for(bodies, theta=theta, leaf=leaf, thread=thread, variant=variant)
    o.build() if mode == 0 else o.build_with_domain(bodies["hw"], bodies["hh"])
    emu.build(bodies, mode, leaf, thread, slab)
    nodes = emu.nodes()
    check_next_pointers(nodes)
    ec, oc = canonical_from_nodes(nodes), o.canonical()
    assert_same_topology(ec, oc)
    assert np.array_equal(ec["charge"], oc["charge"])
    return emu, o, ec, oc


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 5000])
@pytest.mark.parametrize("mode", [0, 1])
def test_topology_small(n, mode):
    emu, o, ec, oc = run(uniform_pm1(n), mode)
    assert np.array_equal(emu.perm.astype(np.int64), o.permutation())
    assert np.array_equal(ec["mass"], oc["mass"])


@pytest.mark.parametrize("gen,n,mode,theta", [(uniform_pm1, 20000, 0, 1.0), (electrolyte, 20000, 1, 0.5),
                                              (clustered, 30000, 0, 1.0)])
def test_lockstep_traversal_is_bit_exact_given_equal_centres(gen, n, mode, theta):
    bodies = gen(n)
    emu, o, ec, oc = run(bodies, mode, theta=theta)
    assert np.array_equal(emu.perm.astype(np.int64), o.permutation())
    assert np.array_equal(ec["pos"], oc["pos"])  # f64 sums on both sides round to the same f32
    sb = emu.sorted_bodies()
    fe, steps, pairs = emu.walk(sb[:, :2], radius=sb[:, 3], theta=theta)
    fo, counters = o.acc_points(sb[:, :2], radius=sb[:, 3], k_e=KE)
    assert np.array_equal(fe, fo)
    assert pairs == counters[2]
    # and with the device centres injected into the strict reference-order tree
    s = oracle_for(bodies, theta=theta)
    s.build() if mode == 0 else s.build_with_domain(bodies["hw"], bodies["hh"])
    s.set_canonical_pos(ec["pos"])
    fi, _ = s.acc_points(sb[:, :2], radius=sb[:, 3], k_e=KE)
    assert np.array_equal(fe, fi)


@pytest.mark.parametrize("leaf,thread", [(8, 32), (4, 1024), (1, 1), (3, 2), (16, 5), (2, 1024)])
def test_capacities(leaf, thread):
    for bodies, mode in ((uniform_pm1(5000), 0), (clustered(5000), 1)):
        emu, o, ec, oc = run(bodies, mode, leaf, thread)
        operm = o.permutation()
        lm = oc["is_leaf"] == 1
        for a, b in zip(oc["start"][lm], oc["end"][lm]):
            assert set(operm[a:b]) == set(emu.perm[a:b].astype(np.int64))


def test_degenerate_inputs():
    b = uniform_pm1(1000)
    b["pos"][10:15] = b["pos"][10]
    b["pos"][500] = b["pos"][501]
    for leaf, thread in [(1, 1024), (4, 1024), (8, 4)]:
        emu, o, ec, oc = run(b, 0, leaf, thread)
        assert o.flags() & 2 and emu.meta()["zero_leaves"] >= 1
    b = uniform_pm1(64)
    b["pos"][:] = b["pos"][0]
    for mode in (0, 1):
        emu, o, ec, oc = run(b, mode)
        assert len(ec) == 1
    b = uniform_pm1(3)
    b["pos"][:] = [[100, 100], [np.nextafter(np.float32(100), np.float32(200)), 100], [5, 5]]
    for half in (300.0, 8000.0):
        b["hw"] = b["hh"] = half
        emu, o, ec, oc = run(b, 1)
        assert o.max_depth() >= 25


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("gen,n,mode,leaf,thread", [
    (uniform_pm1, 30000, 0, 1, 1024), (electrolyte, 20000, 1, 1, 1024), (clustered, 30000, 0, 1, 1024),
    (clustered, 20000, 1, 8, 32), (clustered, 20000, 0, 16, 5), (uniform_pm1, 33, 0, 1, 1024), (uniform_pm1, 2, 1, 1, 1024)])
def test_sharded_build_replay_equals_single_build(gen, n, mode, leaf, thread, world):
    """The algorithm of csrc/shard.cuh replayed serially with the shared host+device logic: key ranges
    aligned to the 65 536 top-level cells, virtual halo keys, per-bin tables, top heap, per-rank compaction
    must concatenate into exactly the single build's arrays (the GPU tests check the kernels themselves)."""
    emu = Emu()
    emu.build(gen(n), mode, leaf, thread)
    total, bad = emu.shard_check(world, leaf, thread)
    assert total == emu.M
    assert not bad.any(), dict(zip(["nodeA", "nodeB", "rec", "ndepth", "sentinel", "multi_writer", "travA", "travB"], bad.tolist()))


@pytest.mark.parametrize("slab", [7, 512, 4096])
@pytest.mark.parametrize("gen,n,mode,leaf,thread", [(uniform_pm1, 20000, 0, 1, 1024), (electrolyte, 20000, 1, 1, 1024),
                                                    (clustered, 30000, 0, 1, 1024), (clustered, 20000, 1, 8, 32),
                                                    (clustered, 5000, 0, 16, 5), (uniform_pm1, 3, 0, 1, 1024)])
def test_in_slab_sums_equal_the_level_sweeps(gen, n, mode, leaf, thread, slab):
    """tree_emit_kernel's scheme replayed serially: each slab of bodies sums the cells that lie inside it
    right after emitting them (skip pointers and body counts of internal cells from the keys), the level
    sweeps only visit the cells that straddle a slab boundary.  Topology, charge, mass and centres must equal
    the oracle's exactly as with one global sweep."""
    bodies = gen(n)
    emu, o, ec, oc = run(bodies, mode, leaf, thread, slab=slab)
    ref, _, rc, _ = run(bodies, mode, leaf, thread, slab=0)
    assert np.array_equal(ec["pos"], rc["pos"]) and np.array_equal(ec["mass"], rc["mass"])
    a, b = emu.meta(), ref.meta()
    assert all(a[k] == b[k] for k in a if k != "root") and np.array_equal(a["root"], b["root"])


@pytest.mark.parametrize("gen,n,mode,theta,leaf,thread", [
    (uniform_pm1, 20000, 0, 1.0, 1, 1024), (uniform_pm1, 20000, 0, 0.5, 1, 1024), (electrolyte, 20000, 1, 1.0, 1, 1024),
    (clustered, 30000, 0, 1.0, 1, 1024), (clustered, 20000, 1, 0.7, 8, 32)])
def test_group_walk_sums_exactly_the_reference_interaction_sets(gen, n, mode, theta, leaf, thread):
    """The device's default traversal shares one walk among 32 targets and classifies nodes against the
    group's bounding box.  Its serial port must give every target EXACTLY the interaction set of the
    reference-order walk (hash signatures of the accepted nodes and of the direct-term bodies, number of
    opened nodes) - the box tests only ever decide cases that are not borderline - and the same field up
    to the order of the additions."""
    bodies = gen(n)
    emu = Emu()
    emu.build(bodies, mode, leaf, thread)
    sb = emu.sorted_bodies()
    pts, rad = sb[:, :2], sb[:, 3]
    ref_sig = emu.walk_signatures(pts, rad, theta=theta)
    ref_f, _, _ = emu.walk(pts, radius=rad, theta=theta)
    f, sig, visited, lifo = emu.group_walk(pts, radius=rad, theta=theta)
    assert np.array_equal(sig, ref_sig)
    assert rel_l2(f, ref_f) <= 2e-6
    assert visited > 0
    # electron-like sample points (radius 0, off the bodies): the self-skip rule must include the own body
    rng = np.random.default_rng(1)
    pts2 = (pts + rng.normal(0, 0.4, pts.shape)).astype(np.float32)
    assert np.array_equal(emu.group_walk(pts2, theta=theta)[1], emu.walk_signatures(pts2, theta=theta))


def _strict_cases():
    rng = np.random.default_rng(1)
    m = 300_000
    return {
        "ones": np.ones(m),
        "signed_x": rng.uniform(-8000, 8000, m),
        "positive_x": rng.uniform(0, 8000, m),
        "masses": rng.choice([6.94, 145.0, 88.06, 90.08], m),
        "lognormal": rng.lognormal(0, 4, m) * rng.choice([-1, 1], m),
        "morton_like": np.concatenate([rng.uniform(-8000, 0, m // 4), rng.uniform(0, 8000, m // 4),
                                       rng.uniform(-8000, 0, m // 4), rng.uniform(0, 8000, m // 4)]),
        "tiny": rng.uniform(-1e-30, 1e-30, 50_000),
        "halves": rng.choice([0.5, 1.5, -0.5, 2.5, 0.25], m),
        "with_inf": np.concatenate([rng.uniform(0, 1, 5000), [np.inf], rng.uniform(0, 1, 5000)]),
        "zeros": np.zeros(10_000),
        "cancelling": np.tile([1e6, -1e6, 3.0, -3.0], 20_000),
    }


@pytest.mark.parametrize("name", list(_strict_cases()))
def test_strict_block_functions_equal_the_serial_f32_sum(name):
    """strict_logic.cuh (what strict_blockfn_kernel / strict_compose_kernel run per block): the running f32 sum
    of quadtree.rs:114-139 evaluated block-wise under a speculated binade, with the serial fall-back, gives
    the bits of the plain loop for every input, block size and starting value."""
    emu = Emu()
    a = _strict_cases()[name]
    for block in (512, 64, 7):
        for s0 in (0.0, 12345.678, -1e9):
            out, blocks, fallbacks = emu.strict_sum(a, block, s0)
            bits = out.view(np.uint32)
            assert bits[0] == bits[1] == bits[2], (name, block, s0, out)
    if name in ("positive_x", "masses", "ones"):  # a monotone sum leaves its binade ~24 times, whatever its length
        _, blocks, fallbacks = emu.strict_sum(a, 512, 0.0)
        assert fallbacks <= 40 and blocks > 500

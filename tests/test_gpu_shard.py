"""GPU: the Morton-range sharded build (shard.cuh) must produce exactly the single-GPU tree.  All ranks
live on one GPU here (one context each) and the exchanges are device copies (LoopbackComm), so the kernels
and the exchange protocol of a G-rank build are checked bit for bit without G GPUs; the NCCL plumbing itself
is covered by tools/check_sharded.py (torchrun) and the gloo test of tests/test_parallel_cpu.py."""
import numpy as np
import pytest

from helpers import KE, clustered, electrolyte, uniform_pm1

pytestmark = pytest.mark.gpu


def make_sim(bodies, **kw):
    import torch
    from particlesim_b200 import Bodies, Simulation
    b = Bodies(bodies["pos"], vel=bodies.get("vel"), mass=bodies.get("mass"), radius=bodies.get("radius"),
               charge=bodies.get("charge"), species=bodies.get("species"), ebody=bodies.get("ebody"),
               erel=bodies.get("erel"))
    sim = Simulation(b, bodies["hw"], bodies["hh"], stream=torch.cuda.current_stream().cuda_stream, **kw)
    sim.config.coulomb_constant = float(KE)
    return sim


def make_ranks(bodies, world, **kw):
    sims = []
    for r in range(world):
        s = make_sim(bodies, **kw)
        s.rank, s.world, s._shard_nb = r, world, max(len(bodies["pos"]), 1)
        s._call("psim_shard_init", r, world)
        sims.append(s)
    return sims


def trav(sim, count):
    import torch
    from particlesim_b200.parallel import shard_views
    v = shard_views(sim, torch)
    torch.cuda.synchronize()
    return v["travA"][:count].cpu().numpy().copy(), v["travB"][:count].cpu().numpy().copy()


def field_of(sim):
    n = len(sim.bodies)
    e = np.zeros((n, 2), np.float32)
    sim._call("psim_field", np.float32(KE), np.float32(0.0), np.float32(0.0), 0, e.ctypes.data, None)
    return e


def perm_of(sim):
    p = np.zeros(len(sim.bodies), np.uint32)
    sim._call("psim_get_permutation", p.ctypes.data)
    return p


CASES = [("uniform_100k", lambda: uniform_pm1(100_000), {}), ("electrolyte_50k", lambda: electrolyte(50_000), {}),
         ("clustered_60k", lambda: clustered(60_000), {}), ("uniform_33", lambda: uniform_pm1(33), {}),
         ("uniform_2", lambda: uniform_pm1(2), {}),
         ("clustered_cap8", lambda: clustered(20_000), dict(leaf_capacity=8, thread_capacity=32)),
         ("clustered_cap16_5", lambda: clustered(20_000), dict(leaf_capacity=16, thread_capacity=5))]


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name,gen,kw", CASES)
def test_sharded_build_equals_single_gpu_build(cuda_device, name, gen, kw, mode, world):
    import torch
    from particlesim_b200.parallel import LoopbackComm, sharded_build
    bodies = gen()
    hw, hh = np.float32(bodies["hw"]), np.float32(bodies["hh"])
    one = make_sim(bodies, **kw)
    one._call("psim_shard_init", 0, 1)
    one.rank, one.world, one._shard_nb = 0, 1, max(len(bodies["pos"]), 1)
    one._call("psim_build", mode, hw, hh)
    ref_perm, ref_e = perm_of(one), field_of(one)
    nodes = one.stats()["compact_nodes"]
    sims = make_ranks(bodies, world, **kw)
    body_lo, trav_lo = sharded_build(sims, mode, hw, hh, LoopbackComm(), torch)
    assert body_lo[0] == 0 and body_lo[-1] == len(bodies["pos"]) and np.all(np.diff(body_lo.astype(np.int64)) >= 0)
    T = int(trav_lo[-1])
    assert 0 < T <= nodes
    refA, refB = trav(one, T)
    for s in sims:
        assert np.array_equal(perm_of(s), ref_perm)
        a, b = trav(s, T)
        assert np.array_equal(a, refA)
        assert np.array_equal(b, refB)
        assert np.array_equal(field_of(s).view(np.uint32), ref_e.view(np.uint32))
        s.close()
    one.close()


def test_sharded_steps_equal_single_gpu_steps(cuda_device):
    """three full hot-path steps (both builds, field, LJ, integrator, electrons) with every build sharded
    four ways: the state must stay bit-identical to the single-GPU run"""
    import torch
    from particlesim_b200 import _lib
    from particlesim_b200.parallel import LoopbackComm, sharded_build
    bodies = electrolyte(40_000)
    bodies["species"][:4000] = 1
    one = make_sim(bodies)
    sims = make_ranks(bodies, 4)
    comm = LoopbackComm()
    # psim_step bins the LJ pass at the largest LJ cutoff (the pair sets do not depend on the cell size, the
    # order of the additions does)
    t = one.species_table
    cell = float(max(np.float32(r["lj_cutoff"]) * np.float32(r["lj_sigma"]) for r in t if r["lj_enabled"]))
    for _ in range(3):
        one.step_device(one.step_params(do_polar=False))
        p = sims[0].step_params(do_polar=False)
        for s in sims:
            s._call("psim_reset_acc")
        sharded_build(sims, _lib.BUILD_CONTAINING, 0.0, 0.0, comm, torch)
        for s in sims:
            s._call("psim_cell_build", p.hw, p.hh, cell)
            s._call("psim_field", p.k_e, p.bg_x, p.bg_y, 1, None, None)
            s._call("psim_short_range", _lib.SR_LJ | _lib.SR_REPULSION | _lib.SR_STACK_PRESSURE)
            s._call("psim_iterate", p.dt, p.damping_base, p.hw, p.hh, p.hd, int(p.enable_out_of_plane))
        sharded_build(sims, _lib.BUILD_DOMAIN, p.hw, p.hh, comm, torch)
        for s in sims:
            s._call("psim_update_electrons", p.bg_x, p.bg_y, p.dt, p.k_e)
    one.download(("pos", "vel", "e_field"))
    one.download_electrons()
    for s in sims:
        s.download(("pos", "vel", "e_field"))
        s.download_electrons()
        for k in ("pos", "vel", "e_field", "erel", "evel"):
            assert np.array_equal(getattr(s.bodies, k), getattr(one.bodies, k)), k
        s.close()
    one.close()


def test_sharded_build_at_scale(cuda_device):
    """BASELINE config 3 size (4 M, clustered: deep unbalanced tree, uneven bins), 5 ranks on one GPU: the
    pieces must concatenate into the single-GPU traversal tree and give the same field, bit for bit"""
    import torch
    from particlesim_b200.parallel import LoopbackComm, sharded_build
    bodies = clustered(4_000_000)
    hw, hh = np.float32(bodies["hw"]), np.float32(bodies["hh"])
    one = make_sim(bodies)
    one._call("psim_shard_init", 0, 1)
    one.rank, one.world, one._shard_nb = 0, 1, len(bodies["pos"])
    one._call("psim_build", 0, hw, hh)
    ref_perm, ref_e = perm_of(one), field_of(one)
    sims = make_ranks(bodies, 5)
    body_lo, trav_lo = sharded_build(sims, 0, hw, hh, LoopbackComm(), torch)
    counts = np.diff(body_lo.astype(np.int64))
    assert counts.sum() == len(bodies["pos"]) and counts.max() <= 1.3 * counts.mean(), counts
    T = int(trav_lo[-1])
    refA, refB = trav(one, T)
    for s in sims[::2]:
        assert np.array_equal(perm_of(s), ref_perm)
        a, b = trav(s, T)
        assert np.array_equal(a, refA) and np.array_equal(b, refB)
        assert np.array_equal(field_of(s).view(np.uint32), ref_e.view(np.uint32))
    for s in sims:
        s.close()
    one.close()


def test_sharded_tree_refuses_whole_array_calls(cuda_device):
    """after a sharded build a context holds its own piece of the node array only: the export, the key
    download and the reference-order walk (parity_mode 2) must say so instead of reading garbage"""
    import torch
    from particlesim_b200 import PsimError
    from particlesim_b200.parallel import LoopbackComm, sharded_build
    bodies = uniform_pm1(5000)
    sims = make_ranks(bodies, 2)
    sharded_build(sims, 0, np.float32(0), np.float32(0), LoopbackComm(), torch)
    s = sims[0]
    keys = np.zeros(5000, np.uint64)
    with pytest.raises(PsimError) as e:
        s._call("psim_get_keys", keys.ctypes.data)
    assert e.value.code == -5
    with pytest.raises(PsimError) as e:
        _ = s.quadtree.nodes
    assert e.value.code == -5
    s._cfg.parity_mode = 2
    s._call("psim_set_config", __import__("ctypes").byref(s._cfg))
    with pytest.raises(PsimError) as e:
        field_of(s)
    assert e.value.code == -5
    with pytest.raises(PsimError) as e:   # and a sharded build itself refuses that mode
        sharded_build(sims[:1] + sims[1:], 0, np.float32(0), np.float32(0), LoopbackComm(), torch)
    assert e.value.code == -2
    # a plain build on the same context makes everything available again
    s._call("psim_build", 0, np.float32(0), np.float32(0))
    s._call("psim_get_keys", keys.ctypes.data)
    assert np.all(keys[:-1] <= keys[1:])
    for x in sims:
        x.close()

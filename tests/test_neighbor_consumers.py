"""update_surrounded_flags / metal_neighbor_count and enforce_metal_z_boundaries.
CPU: the oracle restatement on hand-derived cases.  GPU: the device kernels against the oracle, bit-exact
(flags, state, z, vz), on the clustered LithiumMetal set."""
import numpy as np
import pytest

from helpers import clustered, oracle_for


def hex_patch():
    """one Li+ at the origin ringed by `k` LithiumMetal bodies at distance 3 A, plus a far bystander"""
    def make(k):
        ang = np.arange(k) * (2 * np.pi / max(k, 1))
        pos = np.concatenate([[[0.0, 0.0]], 3.0 * np.stack([np.cos(ang), np.sin(ang)], 1), [[40.0, 40.0]]]).astype(np.float32)
        n = len(pos)
        species = np.array([0] + [1] * k + [0], np.uint8)
        return dict(pos=pos, vel=np.zeros((n, 2), np.float32), mass=np.ones(n, np.float32),
                    radius=np.where(species == 1, 1.52, 0.76).astype(np.float32), charge=np.zeros(n, np.float32),
                    species=species, hw=45.0, hh=45.0, ebody=np.zeros(0, np.uint32), erel=np.zeros((0, 2), np.float32))
    return make


def test_oracle_surrounded_threshold_and_schedule():
    """radius 0.76 * 4 = 3.04 A reaches the ring at 3 A: 8 metals -> surrounded, 7 -> not (config.rs:182-184);
    an unmoved body is only re-checked every 10 frames (config.rs:186-188)"""
    for k, want in ((8, 1), (7, 0)):
        o = oracle_for(hex_patch()(k))
        o.update_surrounded_flags(45.0, 45.0, frame=10)
        flags, pos, frame = o.surrounded()
        assert flags[0] == want and flags[-1] == 0
        assert np.all(frame == 10)
    # below the cell-list density threshold the reference walks the tree instead (and the build permutes the
    # bodies): same counts
    o = oracle_for(hex_patch()(8))
    o.update_surrounded_flags(60.0, 60.0, frame=10)
    ids = o.get_bodies()["id"]
    assert o.surrounded()[0][int(np.argmax(ids == 0))] == 1 and o.surrounded()[0].sum() == 1
    o = oracle_for(hex_patch()(8))
    o.update_surrounded_flags(45.0, 45.0, frame=3)     # frame_diff 3 < 10 and nobody moved: nothing happens
    assert o.surrounded()[0].sum() == 0 and np.all(o.surrounded()[2] == 0)
    o.update_surrounded_flags(45.0, 45.0, frame=10)
    assert o.surrounded()[0][0] == 1
    o.update_surrounded_flags(45.0, 45.0, frame=12, neighbor_threshold=9)   # not due: flag kept
    assert o.surrounded()[0][0] == 1 and o.surrounded()[2][0] == 10


def test_oracle_metal_z_boundaries():
    """a body 3 A from a metal (reach 0.76 + 1.52 + 1.52 = 3.8 A) is confined to |z| <= 1.52 + 0.01; the far
    bystander is only clamped to max_z"""
    b = hex_patch()(3)
    o = oracle_for(b)
    z = np.array([5.0, 0, 0, 0, -9.0], np.float32)
    vz = np.array([1.0, 0, 0, 0, -2.0], np.float32)
    o.set_bodies(b["pos"], z=z, vz=vz, mass=b["mass"], radius=b["radius"], charge=b["charge"], species=b["species"])
    o.enforce_metal_z_boundaries(6.0, 45.0, 45.0)
    ob = o.get_bodies()   # 5 bodies in 90 x 90 A: tree branch, the build has permuted them
    zz, vv = np.zeros(5, np.float32), np.zeros(5, np.float32)
    zz[ob["id"]], vv[ob["id"]] = ob["z"], ob["vz"]
    assert zz[0] == np.float32(np.float32(0.0) + np.float32(1.52) + np.float32(0.01)) and vv[0] == 0.0
    assert zz[-1] == -6.0 and vv[-1] == 0.0
    assert np.all(zz[1:4] == 0.0)


def _device(bodies, z=None, vz=None):
    from particlesim_b200 import Bodies, Simulation
    b = Bodies(bodies["pos"], vel=bodies.get("vel"), mass=bodies.get("mass"), radius=bodies.get("radius"),
               charge=bodies.get("charge"), species=bodies.get("species"), z=z, vz=vz)
    return Simulation(b, bodies["hw"], bodies["hh"])


@pytest.mark.gpu
def test_surrounded_flags_match_the_oracle(cuda_device):
    bodies = clustered(40_000)
    hw, hh = bodies["hw"], bodies["hh"]
    sim, o = _device(bodies), oracle_for(bodies)
    rng = np.random.default_rng(3)
    for frame in (1, 4, 10, 13, 25):
        sim.frame = frame
        flags = sim.update_surrounded_flags()
        o.update_surrounded_flags(hw, hh, frame)
        of, op, ofr = o.surrounded()
        df, dp, dfr = sim.surrounded()
        assert np.array_equal(flags, of) and np.array_equal(df, of)
        assert np.array_equal(dp, op) and np.array_equal(dfr, ofr)
        # move a third of the bodies by up to one radius so that the "moved" rule fires for some of them
        step = (rng.uniform(-1, 1, bodies["pos"].shape) * bodies["radius"][:, None] * (rng.random(len(flags)) < 0.33)[:, None]).astype(np.float32)
        bodies["pos"] = np.clip(bodies["pos"] + step, -hw, hw).astype(np.float32)
        sim._call("psim_update_positions", len(flags), bodies["pos"].ctypes.data)
        o.set_positions(bodies["pos"])
    assert 0 < int(of.sum()) < len(of)
    sim.close()


@pytest.mark.gpu
def test_metal_z_boundaries_match_the_oracle(cuda_device):
    bodies = clustered(40_000)
    n = len(bodies["pos"])
    rng = np.random.default_rng(4)
    z = rng.uniform(-3, 3, n).astype(np.float32)
    vz = rng.uniform(-1, 1, n).astype(np.float32)
    z[bodies["species"] == 1] = 0.0
    sim = _device(bodies, z=z, vz=vz)
    o = oracle_for(bodies)
    o.set_bodies(bodies["pos"], z=z, vz=vz, vel=bodies.get("vel"), mass=bodies["mass"], radius=bodies["radius"],
                 charge=bodies["charge"], species=bodies["species"])
    sim.enforce_metal_z_boundaries(2.5)
    o.enforce_metal_z_boundaries(2.5, bodies["hw"], bodies["hh"])
    ob = o.get_bodies()
    assert np.array_equal(sim.bodies.z, ob["z"]) and np.array_equal(sim.bodies.vz, ob["vz"])
    assert np.any(ob["z"] != np.clip(z, -2.5, 2.5))   # some body was constrained by a metal, not just by max_z
    sim.close()


def test_oracle_matches_the_committed_fixture():
    """tests/golden/make_golden.py 'consumers': catches silent drift of the oracle's restatement"""
    import json
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    meta = json.load(open(os.path.join(gold, "golden.json")))["consumers"]
    z = np.load(os.path.join(gold, "golden.npz"))
    g = lambda k: z[f"consumers/{k}"]
    b = dict(pos=g("pos"), radius=g("radius"), species=g("species"), mass=g("mass"), charge=g("charge"),
             hw=meta["hw"], hh=meta["hh"])
    o = oracle_for(b)
    o.update_surrounded_flags(meta["hw"], meta["hh"], frame=meta["frames"][0])
    o.set_positions(g("nudged"))
    o.update_surrounded_flags(meta["hw"], meta["hh"], frame=meta["frames"][1])
    flags, last_pos, last_frame = o.surrounded()
    assert np.array_equal(flags, g("flags")) and np.array_equal(last_pos, g("last_pos"))
    assert np.array_equal(last_frame, g("last_frame"))
    assert 0 < flags.sum() < len(flags) and set(np.unique(last_frame)) == {10, 13}
    o = oracle_for(b)
    o.set_bodies(b["pos"], z=g("z0"), vz=g("vz0"), mass=b["mass"], radius=b["radius"], charge=b["charge"], species=b["species"])
    o.enforce_metal_z_boundaries(meta["max_z"], meta["hw"], meta["hh"])
    ob = o.get_bodies()
    assert np.array_equal(ob["z"], g("z")) and np.array_equal(ob["vz"], g("vz"))

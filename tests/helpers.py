"""Shared test helpers: synthetic body sets, canonical tree forms, error norms.

Test infrastructure only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.pyoracle import CANON_DTYPE, NODE_DTYPE, OracleSim  # noqa: E402

KE = np.float32(0.138935)  # f32 of units.rs:32-34
SEED = 0xC0FFEE
RHO = 0.0625  # bodies per square angstrom


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.sqrt((b ** 2).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / den) if den > 0 else float(np.sqrt(((a - b) ** 2).sum()))


def canonical_from_nodes(nodes: np.ndarray) -> np.ndarray:
    """DFS pre-order canonical listing of a reference-shaped node array (children/next layout)."""
    out = []
    if len(nodes) == 0:
        return np.zeros(0, dtype=CANON_DTYPE)
    stack = [(0, 0, 0, 0)]
    children = nodes["children"]
    while stack:
        idx, depth, hi, lo = stack.pop()
        nd = nodes[idx]
        out.append((hi, lo, depth, int(children[idx] == 0), int(nd["bodies_start"]), int(nd["bodies_end"]),
                    tuple(nd["pos"]), float(nd["mass"]), float(nd["charge"]), tuple(nd["quad_center"]),
                    float(nd["quad_size"]), 0))
        c = int(children[idx])
        if c != 0:
            level = depth + 1
            for q in (3, 2, 1, 0):
                h2, l2 = hi, lo
                if level <= 32:
                    l2 |= q << (64 - 2 * level)
                elif level <= 64:
                    h2 |= q << (64 - 2 * (level - 32))
                stack.append((c + q, level, h2, l2))
    return np.array(out, dtype=CANON_DTYPE)


def check_next_pointers(nodes: np.ndarray):
    """`next` must drive the reference's stackless walk through exactly the DFS pre-order."""
    if len(nodes) == 0:
        return
    order_walk = []
    node = 0
    while True:
        order_walk.append(node)
        if nodes["children"][node] != 0:
            node = int(nodes["children"][node])
        else:
            # climb by next pointers like acc_pos does when every node is opened
            nxt = int(nodes["next"][node])
            if nxt == 0:
                break
            node = nxt
    canon = []
    stack = [0]
    while stack:
        i = stack.pop()
        canon.append(i)
        c = int(nodes["children"][i])
        if c:
            stack.extend([c + 3, c + 2, c + 1, c])
    assert order_walk == canon


TOPO_FIELDS = ["path_hi", "path_lo", "depth", "is_leaf", "start", "end"]


def assert_same_topology(a: np.ndarray, b: np.ndarray):
    assert len(a) == len(b), f"node count {len(a)} vs {len(b)}"
    for f in TOPO_FIELDS:
        if not np.array_equal(a[f], b[f]):
            bad = np.nonzero(a[f] != b[f])[0][:5]
            raise AssertionError(f"topology field {f} differs at canonical rows {bad}: {a[f][bad]} vs {b[f][bad]}")
    for f in ["quad_center", "quad_size"]:
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f"{f} not bit-identical"


# ------------------------------------------------------------------------------------------------
# synthetic sets: xoshiro is not needed for parity — numpy's PCG64 seeded with the
# reference's seed constant is documented here and committed with the golden fixtures.
def uniform_pm1(n, seed=SEED, rho=RHO):
    """config 2: uniform, q = ±1 alternating, radius 0.76 / 2.0 by sign."""
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(n / rho))
    pos = rng.uniform(-L / 2, L / 2, (n, 2)).astype(np.float32)
    pos = _dedupe(pos, rng, L)
    q = np.where(np.arange(n) % 2 == 0, 1.0, -1.0).astype(np.float32)
    radius = np.where(q > 0, 0.76, 2.0).astype(np.float32)
    species = np.where(q > 0, 0, 3).astype(np.uint8)
    mass = np.where(q > 0, 6.94, 145.0).astype(np.float32)
    return dict(pos=pos, charge=q, radius=radius, species=species, mass=mass, hw=L / 2, hh=L / 2)


def fractional(n, seed=SEED, rho=RHO):
    """uniform_pm1 with non-integer charges (some zero): node charges then need the bottom-up sweep in the reference's
    child order; integer charges take the exact prefix-difference path of the emit kernel"""
    b = uniform_pm1(n, seed=seed, rho=rho)
    rng = np.random.default_rng(seed + 1)
    f = rng.uniform(0.05, 3.0, n).astype(np.float32)
    f[rng.random(n) < 0.3] = 0.0
    b["charge"] = (b["charge"] * f).astype(np.float32)
    return b


def big_integer(n, seed=SEED, rho=RHO):
    """uniform_pm1 with integer charges whose partial sums leave the exactly representable range (sum |q| >= 2^24):
    the integer-prefix path must stand down"""
    b = uniform_pm1(n, seed=seed, rho=rho)
    b["charge"] = (b["charge"] * 1000003.0).astype(np.float32)
    return b


def electrolyte(n, seed=SEED, rho=RHO):
    """config 4: Li+ / PF6- / EC / DMC at 342:342:2393:2394 (scenario.rs:180-200)."""
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(n / rho))
    pos = rng.uniform(-L / 2, L / 2, (n, 2)).astype(np.float32)
    pos = _dedupe(pos, rng, L)
    w = np.array([342, 342, 2393, 2394], np.float64)
    kind = rng.choice(4, size=n, p=w / w.sum())
    species = np.array([0, 3, 4, 5], np.uint8)[kind]
    charge = np.array([1.0, -1.0, 0.0, 0.0], np.float32)[kind]
    radius = np.array([0.76, 2.0, 2.5, 2.5], np.float32)[kind]
    mass = np.array([6.94, 145.0, 88.06, 90.08], np.float32)[kind]
    polar = np.array([0.0, 0.3, 0.85, 0.60], np.float32)[kind]
    has_e = kind != 0
    ebody = np.nonzero(has_e)[0].astype(np.uint32)
    ang = rng.uniform(0, 2 * np.pi, len(ebody))
    rr = np.sqrt(rng.uniform(0, 1, len(ebody))) * polar[ebody] * radius[ebody]
    erel = np.stack([rr * np.cos(ang), rr * np.sin(ang)], 1).astype(np.float32)
    vel = rng.normal(0, 0.01, (n, 2)).astype(np.float32)
    return dict(pos=pos, charge=charge, radius=radius, species=species, mass=mass, vel=vel,
                ebody=ebody, erel=erel, hw=L / 2, hh=L / 2)


def clustered(n, seed=SEED, rho=RHO):
    """config 3: LithiumMetal clusters on a jittered 3.04 A lattice + dendrite filaments + electrolyte."""
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(n / rho))
    n_cl = n // 2
    n_fil = n // 10
    n_bg = n - n_cl - n_fil
    ncent = max(1, n_cl // 1000)
    cent = rng.uniform(-L / 2 + 60, L / 2 - 60, (ncent, 2))
    which = rng.integers(0, ncent, n_cl)
    r = np.abs(rng.lognormal(np.log(20.0), 0.5, n_cl))
    ang = rng.uniform(0, 2 * np.pi, n_cl)
    p_cl = cent[which] + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    # one body per lattice site (the reference fills lattices site by site, app/spawn.rs:141-195);
    # bodies that land on an occupied site are moved to a free site of a growing ring around it
    site = np.round(p_cl / 3.04).astype(np.int64)
    site = _unique_sites(site, rng)
    p_cl = site * 3.04 + rng.normal(0, 0.05, (n_cl, 2))
    # filaments: random walks on the lattice
    p_f = []
    left = n_fil
    while left > 0:
        ln = int(min(left, rng.integers(200, 2000)))
        start = rng.uniform(-L / 2 + 100, L / 2 - 100, 2)
        steps = rng.integers(0, 4, ln)
        d = np.array([[1, 0], [-1, 0], [0, 1], [0, 1]])[steps]
        walk = np.round(start / 3.04).astype(np.int64) + np.cumsum(d, 0)
        walk = walk[np.sort(np.unique(walk, axis=0, return_index=True)[1])]  # self-avoiding
        p_f.append(walk * 3.04 + 1.52 + rng.normal(0, 0.05, walk.shape))
        ln = max(len(walk), 1)
        left -= ln
    p_f = np.concatenate(p_f) if p_f else np.zeros((0, 2))
    n_fil = len(p_f)
    n_bg = n - n_cl - n_fil
    p_bg = rng.uniform(-L / 2, L / 2, (n_bg, 2))
    pos = np.clip(np.concatenate([p_cl, p_f, p_bg]), -L / 2, L / 2).astype(np.float32)
    pos = _dedupe(pos, rng, L)
    kind_bg = rng.choice(4, size=n_bg, p=np.array([342, 342, 2393, 2394]) / 5471.0)
    species = np.concatenate([np.full(n_cl + n_fil, 1, np.uint8), np.array([0, 3, 4, 5], np.uint8)[kind_bg]])
    charge = np.concatenate([np.zeros(n_cl + n_fil, np.float32), np.array([1.0, -1.0, 0.0, 0.0], np.float32)[kind_bg]])
    radius = np.concatenate([np.full(n_cl + n_fil, 1.52, np.float32), np.array([0.76, 2.0, 2.5, 2.5], np.float32)[kind_bg]])
    mass = np.concatenate([np.full(n_cl + n_fil, 6.94, np.float32), np.array([6.94, 145.0, 88.06, 90.08], np.float32)[kind_bg]])
    sh = rng.permutation(n)
    return dict(pos=pos[sh], charge=charge[sh], radius=radius[sh], species=species[sh], mass=mass[sh],
                hw=L / 2, hh=L / 2)


def _unique_sites(site, rng):
    """make integer lattice sites unique: duplicates random-walk until they find a free site"""
    site = site.copy()
    for _ in range(200):
        key = site[:, 0] * 4_000_003 + site[:, 1]
        _, first = np.unique(key, return_index=True)
        if len(first) == len(site):
            break
        dup = np.ones(len(site), bool)
        dup[first] = False
        site[dup] += rng.integers(-2, 3, (int(dup.sum()), 2))
    return site


def _dedupe(pos, rng, L):
    """re-draw exact duplicate positions (a C=1 tree turns them into refused leaves)"""
    for _ in range(8):
        v = pos.view(np.uint64).ravel() if pos.dtype == np.float32 else None
        _, first = np.unique(v, return_index=True)
        if len(first) == len(pos):
            break
        dup = np.ones(len(pos), bool)
        dup[first] = False
        pos[dup] = rng.uniform(-L / 2, L / 2, (int(dup.sum()), 2)).astype(np.float32)
    return pos


def oracle_for(bodies, theta=1.0, epsilon=2.0, leaf=1, thread=1024, variant=""):
    s = OracleSim(theta, epsilon, leaf, thread, variant=variant)
    s.set_bodies(bodies["pos"], z=bodies.get("z"), vel=bodies.get("vel"), vz=bodies.get("vz"), mass=bodies.get("mass"),
                 radius=bodies.get("radius"), charge=bodies.get("charge"), species=bodies.get("species"))
    if "ebody" in bodies:
        s.set_electrons(bodies["ebody"], bodies["erel"])
    return s


# ------------------------------------------------------------------------------------------------
# host emulation of the device construction logic (tests/emu/emu.cpp)
class Emu:
    def __init__(self):
        here = os.path.join(ROOT, "tests", "emu")
        so = os.path.join(here, "_build", "libemu.so")
        src = os.path.join(here, "emu.cpp")
        deps = [src] + [os.path.join(ROOT, "particlesim_b200", "csrc", f) for f in ("psim_core.cuh", "tree_logic.cuh", "shard_logic.cuh", "strict_logic.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                            "-I/usr/local/cuda/include", "-o", so, src], check=True)
        self.lib = C.CDLL(so)
        L = self.lib
        L.emu_create.restype = C.c_void_p
        L.emu_destroy.argtypes = [C.c_void_p]
        pf = C.POINTER(C.c_float)
        L.emu_build.restype = C.c_uint32
        L.emu_build.argtypes = [C.c_void_p, C.c_uint32, pf, pf, pf, pf, C.c_int, C.c_float, C.c_float,
                                C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emu_reference_node_count.restype = C.c_uint64
        L.emu_reference_node_count.argtypes = [C.c_void_p]
        L.emu_meta.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_export_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.emu_set_params.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.emu_set_slab.argtypes = [C.c_void_p, C.c_uint32]
        L.emu_walk.argtypes = [C.c_void_p, C.c_uint32, pf, pf, pf, C.c_float, pf, C.c_void_p, C.c_void_p]
        L.emu_sorted_bodies.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_walk_signatures.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_group_walk.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_group_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.emu_strict_sum.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
        L.emu_shard_check.restype = C.c_uint32
        L.emu_shard_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        self.h = L.emu_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.emu_destroy(self.h)
            self.h = None

    def strict_sum(self, a, block=512, s0=0.0):
        """serial f32 sum of a from s0 three ways (plain loop, block functions, composed block functions)"""
        a = np.ascontiguousarray(a, np.float32)
        out, st = np.zeros(3, np.float32), np.zeros(2, np.uint64)
        self.lib.emu_strict_sum(a.ctypes.data, len(a), block, np.float32(s0), out.ctypes.data, st.ctypes.data)
        return out, int(st[0]), int(st[1])

    def walk_signatures(self, pts, radius=None, theta=1.0, epsilon=2.0):
        """per-target interaction signatures of the reference-order walk (see emu.cpp)"""
        pts = np.ascontiguousarray(pts, np.float32)
        m = len(pts)
        rad = None if radius is None else np.ascontiguousarray(radius, np.float32)
        self.lib.emu_set_params(self.h, np.float32(theta), np.float32(epsilon))
        sig = np.zeros((m, 3), np.uint64)
        self.lib.emu_walk_signatures(self.h, m, pts.ctypes.data, None if rad is None else rad.ctypes.data, sig.ctypes.data)
        return sig

    def group_walk(self, pts, q=None, radius=None, k_e=KE, theta=1.0, epsilon=2.0):
        """serial port of the device's group walk: (fields, signatures, nodes visited, lifo rounds)"""
        pts = np.ascontiguousarray(pts, np.float32)
        m = len(pts)
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        q, rad = f(q), f(radius)
        p = lambda a: None if a is None else a.ctypes.data
        self.lib.emu_set_params(self.h, np.float32(theta), np.float32(epsilon))
        out, sig, stats = np.zeros((m, 2), np.float32), np.zeros((m, 3), np.uint64), np.zeros(2, np.uint64)
        self.lib.emu_group_walk(self.h, m, pts.ctypes.data, p(q), p(rad), np.float32(k_e), np.float32(theta),
                                out.ctypes.data, sig.ctypes.data, stats.ctypes.data)
        assert stats[0] != np.uint64(0xFFFFFFFFFFFFFFFF), "ring buffer overflow"
        return out, sig, int(stats[0]), int(stats[1])

    def group_stats(self, pts, radius=None, theta=1.0, epsilon=2.0, nsub=1):
        """interaction-list statistics of the group walk over the charged nodes (emu_group_stats)"""
        pts = np.ascontiguousarray(pts, np.float32)
        rad = None if radius is None else np.ascontiguousarray(radius, np.float32)
        self.lib.emu_set_params(self.h, np.float32(theta), np.float32(epsilon))
        out = np.zeros(14, np.uint64)
        self.lib.emu_group_stats(self.h, len(pts), pts.ctypes.data, None if rad is None else rad.ctypes.data,
                                 np.float32(theta), nsub, out.ctypes.data)
        return out

    def shard_check(self, world, leaf=1, thread=1024):
        """replay a `world`-rank sharded build of the last build's bodies; returns (node total, mismatch
        counts against the single build: nodeA, nodeB, rec, ndepth, sentinel depth, multi-writer heap slots,
        travA, travB)"""
        out = np.zeros(8, np.uint64)
        total = self.lib.emu_shard_check(self.h, world, leaf, thread, out.ctypes.data)
        return int(total), out

    def build(self, bodies, mode=0, leaf=1, thread=1024, slab=0):
        """slab > 0 replays tree_emit_kernel's slabs: every `slab` bodies sum the cells inside their slab
        right after emitting them, the level sweeps only see the cells that straddle a slab boundary"""
        self.lib.emu_set_slab(self.h, slab)
        pos = np.ascontiguousarray(bodies["pos"], np.float32)
        n = len(pos)
        f = lambda k: None if bodies.get(k) is None else np.ascontiguousarray(bodies[k], np.float32)
        mass, radius, charge = f("mass"), f("radius"), f("charge")
        p = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
        self.perm = np.zeros(n, np.uint32)
        self.keys = np.zeros(n, np.uint64)
        self.n = n
        self.M = self.lib.emu_build(self.h, n, p(pos), p(mass), p(radius), p(charge), mode,
                                    bodies.get("hw", 0.0), bodies.get("hh", 0.0), leaf, thread,
                                    self.perm.ctypes.data, self.keys.ctypes.data)
        return self.M

    def meta(self):
        m = np.zeros(8, np.uint32)
        r = np.zeros(3, np.float32)
        self.lib.emu_meta(self.h, m.ctypes.data, r.ctypes.data)
        return dict(num_nodes=int(m[0]), num_internal=int(m[1]), max_depth=int(m[2]), dcap=int(m[3]),
                    zero_leaves=int(m[4]), cap_leaves=int(m[5]), err=int(m[6]), root=r)

    def nodes(self):
        cnt = int(self.lib.emu_reference_node_count(self.h))
        out = np.zeros(cnt, dtype=NODE_DTYPE)
        if cnt:
            self.lib.emu_export_nodes(self.h, out.ctypes.data, cnt)
        return out

    def sorted_bodies(self):
        out = np.zeros((self.n, 4), np.float32)
        self.lib.emu_sorted_bodies(self.h, out.ctypes.data)
        return out

    def walk(self, pts, q=None, radius=None, k_e=KE, theta=1.0, epsilon=2.0):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        m = len(pts)
        q = None if q is None else np.ascontiguousarray(q, np.float32)
        radius = None if radius is None else np.ascontiguousarray(radius, np.float32)
        p = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
        out = np.zeros((m, 2), np.float32)
        steps, pairs = C.c_uint64(), C.c_uint64()
        self.lib.emu_set_params(self.h, theta, epsilon)
        self.lib.emu_walk(self.h, m, p(pts), p(q), p(radius), k_e, p(out), C.byref(steps), C.byref(pairs))
        return out, steps.value, pairs.value


def slab(n, seed=SEED, rho=RHO):
    """config 5: mixed-species slab.  20 % solid-electrolyte scaffold (LLZO / LLZT / S40B, LJ-enabled, neutral) on a
    jittered 5 A lattice in a central band, 5 % LithiumMetal in two lattice slabs (3.04 A) at the +-x edges, 75 %
    electrolyte mix (Li+ / PF6- / EC / DMC 342:342:2393:2394) everywhere; z ~ U(-1, 1) for the 2.5-D path."""
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(n / rho))
    n_sc, n_li = n // 5, n // 20
    n_el = n - n_sc - n_li

    def lattice(count, a, x0, width):
        """`count` jittered sites of spacing a filling x in [x0, x0 + width), y over the whole height"""
        nx = max(1, int(width / a))
        ny = (count + nx - 1) // nx
        a_y = min(a, L / max(ny, 1))
        k = np.arange(count)
        p = np.stack([x0 + (k % nx + 0.5) * a, -L / 2 + (k // nx + 0.5) * a_y], 1)
        return p + rng.normal(0, 0.05, p.shape)

    w_sc = n_sc * 25.0 / L
    p_sc = lattice(n_sc, 5.0, -w_sc / 2, w_sc)
    w_li = (n_li // 2) * 3.04 ** 2 / L
    p_li = np.concatenate([lattice(n_li // 2, 3.04, -L / 2, w_li), lattice(n_li - n_li // 2, 3.04, L / 2 - w_li, w_li)])
    p_el = rng.uniform(-L / 2, L / 2, (n_el, 2))
    pos = np.clip(np.concatenate([p_sc, p_li, p_el]), -L / 2, np.nextafter(np.float32(L / 2), np.float32(0))).astype(np.float32)
    pos = _dedupe(pos, rng, L)
    kind = rng.choice(4, size=n_el, p=np.array([342, 342, 2393, 2394]) / 5471.0)
    sc_species = rng.choice(np.array([9, 10, 11], np.uint8), n_sc)
    species = np.concatenate([sc_species, np.full(n_li, 1, np.uint8), np.array([0, 3, 4, 5], np.uint8)[kind]])
    table = {0: (6.94, 0.76), 1: (6.94, 1.52), 3: (145.0, 2.0), 4: (88.06, 2.5), 5: (90.08, 2.5), 9: (840.0, 4.5),
             10: (865.0, 4.7), 11: (340.0, 4.2)}
    mass = np.zeros(n, np.float32)
    radius = np.zeros(n, np.float32)
    for s, (m, r) in table.items():
        sel = species == s
        mass[sel], radius[sel] = m, r
    charge = np.zeros(n, np.float32)
    charge[species == 0], charge[species == 3] = 1.0, -1.0
    polar = np.zeros(n, np.float32)
    polar[species == 3], polar[species == 4], polar[species == 5] = 0.3, 0.85, 0.60
    ebody = np.nonzero(polar > 0)[0].astype(np.uint32)
    ang = rng.uniform(0, 2 * np.pi, len(ebody))
    rr = np.sqrt(rng.uniform(0, 1, len(ebody))) * polar[ebody] * radius[ebody]
    erel = np.stack([rr * np.cos(ang), rr * np.sin(ang)], 1).astype(np.float32)
    sh = rng.permutation(n)
    inv = np.empty(n, np.int64)
    inv[sh] = np.arange(n)
    ebody = np.sort(inv[ebody]).astype(np.uint32)  # electrons keep their body; rel_pos is i.i.d., so re-pairing by rank is fine
    vel = rng.normal(0, 0.01, (n, 2)).astype(np.float32)
    z = rng.uniform(-1, 1, n).astype(np.float32)
    return dict(pos=pos[sh], charge=charge[sh], radius=radius[sh], species=species[sh], mass=mass[sh], vel=vel, z=z,
                ebody=ebody, erel=erel, hw=L / 2, hh=L / 2, hd=1.0, enable_out_of_plane=True)

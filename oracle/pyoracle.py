"""ctypes binding of the CPU oracle (oracle/oracle.h).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under particlesim_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


class OrcNode(C.Structure):
    _fields_ = [
        ("children", C.c_uint64), ("next", C.c_uint64), ("pos", C.c_float * 2), ("mass", C.c_float),
        ("quad_center", C.c_float * 2), ("quad_size", C.c_float), ("bodies_start", C.c_uint64),
        ("bodies_end", C.c_uint64), ("charge", C.c_float), ("_pad", C.c_uint32),
    ]


NODE_DTYPE = np.dtype([
    ("children", "<u8"), ("next", "<u8"), ("pos", "<f4", (2,)), ("mass", "<f4"),
    ("quad_center", "<f4", (2,)), ("quad_size", "<f4"), ("bodies_start", "<u8"),
    ("bodies_end", "<u8"), ("charge", "<f4"), ("_pad", "<u4"),
], align=True)
CANON_DTYPE = np.dtype([
    ("path_hi", "<u8"), ("path_lo", "<u8"), ("depth", "<u4"), ("is_leaf", "<u4"),
    ("start", "<u8"), ("end", "<u8"), ("pos", "<f4", (2,)), ("mass", "<f4"), ("charge", "<f4"),
    ("quad_center", "<f4", (2,)), ("quad_size", "<f4"), ("_pad", "<u4"),
], align=True)
SPECIES_DTYPE = np.dtype([
    ("mass", "<f4"), ("radius", "<f4"), ("damping", "<f4"), ("lj_epsilon", "<f4"),
    ("lj_sigma", "<f4"), ("lj_cutoff", "<f4"), ("polar_offset", "<f4"), ("polar_charge", "<f4"),
    ("repulsion_strength", "<f4"), ("repulsion_cutoff", "<f4"), ("lj_enabled", "<u4"),
    ("repulsion_enabled", "<u4"),
], align=True)
assert NODE_DTYPE.itemsize == 64 and CANON_DTYPE.itemsize == 72 and SPECIES_DTYPE.itemsize == 48


class Counters(C.Structure):
    _fields_ = [("visits", C.c_uint64), ("accepts", C.c_uint64), ("pairs", C.c_uint64)]


def build(native: bool = False) -> None:
    target = ["native"] if native else []
    subprocess.run(["make", "-C", _HERE, "-s"] + target, check=True, capture_output=not native)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


_LIBS: dict[str, C.CDLL] = {}


def load(variant: str = "") -> C.CDLL:
    """variant: "" (plain), "uvfma" (fused Vec2::mag_sq), "hp" (f64 node-centre sums; not the
    reference), "native" (-O3 -march=native, timing)."""
    if variant in _LIBS:
        return _LIBS[variant]
    name = "liboracle.so" if not variant else f"liboracle_{variant}.so"
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build(native=(variant == "native"))
    lib = C.CDLL(path)
    vp, f, u64, i, u32 = C.c_void_p, C.c_float, C.c_uint64, C.c_int, C.c_uint32
    pf, pd = C.POINTER(C.c_float), C.POINTER(C.c_double)
    pu64, pu32, pu8 = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    pc = C.POINTER(Counters)
    sig = {
        "orc_create": (vp, [f, f, u64, u64]),
        "orc_destroy": (None, [vp]),
        "orc_set_species_table": (None, [vp, vp, u32]),
        "orc_default_species_table": (None, [vp]),
        "orc_set_bodies": (None, [vp, u64, pf, pf, pf, pf, pf, pf, pf, pu8]),
        "orc_set_electrons": (None, [vp, u64, pu32, pf, pf]),
        "orc_set_positions": (None, [vp, pf]),
        "orc_num_bodies": (u64, [vp]),
        "orc_num_electrons": (u64, [vp]),
        "orc_get_bodies": (None, [vp, pu64, pf, pf, pf, pf, pf, pf, pf, pf, pf, pu8, pf]),
        "orc_get_electrons": (None, [vp, pu32, pf, pf]),
        "orc_build": (None, [vp, i, f, f, i]),
        "orc_num_nodes": (u64, [vp]),
        "orc_get_nodes": (None, [vp, vp]),
        "orc_canonical": (u64, [vp, vp, u64]),
        "orc_set_canonical_pos": (None, [vp, pf, u64]),
        "orc_max_depth": (u32, [vp]),
        "orc_flags": (u32, [vp]),
        "orc_field": (None, [vp, f, i, pc]),
        "orc_acc_points": (None, [vp, u64, pf, pf, pf, f, pf, i, pc]),
        "orc_tree_neighbors": (u64, [vp, u64, f, pu64, u64]),
        "orc_cell_set_domain": (None, [vp, f, f]),
        "orc_cell_rebuild": (None, [vp, f]),
        "orc_cell_dims": (None, [vp, pu64, pu64]),
        "orc_cell_contents": (u64, [vp, u64, pu64, u64]),
        "orc_cell_neighbors": (u64, [vp, u64, f, pu64, u64]),
        "orc_cell_metal_neighbor_count": (u64, [vp, u64, f]),
        "orc_use_cell_list": (i, [vp, f, f, f]),
        "orc_reset_acc": (None, [vp]),
        "orc_prepare_spatial_structures": (None, [vp, f, f, f, i]),
        "orc_attract": (None, [vp, f, f, f, i]),
        "orc_apply_polar_forces": (None, [vp, i, f, i]),
        "orc_apply_lj_forces": (None, [vp, i, f, u32]),
        "orc_apply_repulsive_forces": (None, [vp, i]),
        "orc_apply_stack_pressure": (None, [vp, i, f, f, f]),
        "orc_iterate": (None, [vp, f, f, f, f, f, i, i]),
        "orc_update_electrons": (None, [vp, f, f, f, f, i]),
        "orc_update_surrounded_flags": (None, [vp, f, f, f, u64, f, u64]),
        "orc_get_surrounded": (None, [vp, vp, vp, vp]),
        "orc_enforce_metal_z_boundaries": (None, [vp, f, f, f, f]),
        "orc_direct_f64": (None, [vp, u64, pf, pf, C.c_double, C.c_double, pd, i]),
        "orc_collide": (u64, [vp, f, u32, f, i, i]),
        "orc_hop_alignment": (None, [vp, u64, vp, vp, vp, f, f, f, f, vp, vp]),
        "orc_max_threads": (i, []),
        "orc_uv_fma": (i, []),
    }
    for name_, (res, args) in sig.items():
        fn = getattr(lib, name_)
        fn.restype = res
        fn.argtypes = args
    _LIBS[variant] = lib
    return lib


def default_species_table(variant: str = "") -> np.ndarray:
    t = np.zeros(21, dtype=SPECIES_DTYPE)
    load(variant).orc_default_species_table(t.ctypes.data)
    return t


class OracleSim:
    """The slice of `Simulation` the hot path touches: Vec<Body> + Quadtree + CellList."""

    def __init__(self, theta=1.0, epsilon=2.0, leaf_capacity=1, thread_capacity=1024, variant=""):
        self.lib = load(variant)
        self.h = self.lib.orc_create(theta, epsilon, leaf_capacity, thread_capacity)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    # ---- state
    def set_species_table(self, table: np.ndarray):
        t = np.ascontiguousarray(table, dtype=SPECIES_DTYPE)
        self.lib.orc_set_species_table(self.h, t.ctypes.data, len(t))

    def set_bodies(self, pos, z=None, vel=None, vz=None, mass=None, radius=None, charge=None,
                   species=None):
        pos = _f32(pos, (-1, 2))
        n = len(pos)
        z, vz, mass, radius, charge = (_f32(a) for a in (z, vz, mass, radius, charge))
        vel = _f32(vel, (-1, 2))
        sp = None if species is None else np.ascontiguousarray(species, dtype=np.uint8)
        self.lib.orc_set_bodies(self.h, n, _ptr(pos), _ptr(z), _ptr(vel), _ptr(vz), _ptr(mass),
                                _ptr(radius), _ptr(charge), _ptr(sp, C.c_uint8))

    def set_positions(self, pos):
        pos = _f32(pos, (-1, 2))
        assert len(pos) == self.n
        self.lib.orc_set_positions(self.h, _ptr(pos))

    def set_electrons(self, body, rel, vel=None):
        body = np.ascontiguousarray(body, dtype=np.uint32)
        rel = _f32(rel, (-1, 2))
        vel = _f32(vel, (-1, 2))
        self.lib.orc_set_electrons(self.h, len(body), _ptr(body, C.c_uint32), _ptr(rel), _ptr(vel))

    @property
    def n(self):
        return int(self.lib.orc_num_bodies(self.h))

    def get_bodies(self) -> dict:
        n = self.n
        out = {
            "id": np.zeros(n, np.uint64), "pos": np.zeros((n, 2), np.float32),
            "z": np.zeros(n, np.float32), "vel": np.zeros((n, 2), np.float32),
            "vz": np.zeros(n, np.float32), "acc": np.zeros((n, 2), np.float32),
            "az": np.zeros(n, np.float32), "mass": np.zeros(n, np.float32),
            "radius": np.zeros(n, np.float32), "charge": np.zeros(n, np.float32),
            "species": np.zeros(n, np.uint8), "e_field": np.zeros((n, 2), np.float32),
        }
        o = out
        self.lib.orc_get_bodies(self.h, _ptr(o["id"], C.c_uint64), _ptr(o["pos"]), _ptr(o["z"]),
                                _ptr(o["vel"]), _ptr(o["vz"]), _ptr(o["acc"]), _ptr(o["az"]),
                                _ptr(o["mass"]), _ptr(o["radius"]), _ptr(o["charge"]),
                                _ptr(o["species"], C.c_uint8), _ptr(o["e_field"]))
        return out

    def get_electrons(self):
        m = int(self.lib.orc_num_electrons(self.h))
        body = np.zeros(m, np.uint32)
        rel = np.zeros((m, 2), np.float32)
        vel = np.zeros((m, 2), np.float32)
        self.lib.orc_get_electrons(self.h, _ptr(body, C.c_uint32), _ptr(rel), _ptr(vel))
        return body, rel, vel

    # ---- quadtree
    def build(self, threads=1):
        self.lib.orc_build(self.h, 0, 0.0, 0.0, threads)

    def build_with_domain(self, hw, hh, threads=1):
        self.lib.orc_build(self.h, 1, hw, hh, threads)

    def permutation(self) -> np.ndarray:
        """perm[i] = upload-order index (id) of the body now at position i."""
        return self.get_bodies()["id"].astype(np.int64)

    def nodes(self) -> np.ndarray:
        m = int(self.lib.orc_num_nodes(self.h))
        out = np.zeros(m, dtype=NODE_DTYPE)
        if m:
            self.lib.orc_get_nodes(self.h, out.ctypes.data)
        return out

    def canonical(self) -> np.ndarray:
        m = int(self.lib.orc_canonical(self.h, None, 0))
        out = np.zeros(m, dtype=CANON_DTYPE)
        if m:
            self.lib.orc_canonical(self.h, out.ctypes.data, m)
        return out

    def set_canonical_pos(self, pos):
        pos = _f32(pos, (-1, 2))
        self.lib.orc_set_canonical_pos(self.h, _ptr(pos), len(pos))

    def max_depth(self):
        return int(self.lib.orc_max_depth(self.h))

    def flags(self):
        return int(self.lib.orc_flags(self.h))

    def field(self, k_e, threads=0):
        c = Counters()
        self.lib.orc_field(self.h, k_e, threads, C.byref(c))
        return self.get_bodies()["e_field"], (c.visits, c.accepts, c.pairs)

    def acc_points(self, pts, q=None, radius=None, k_e=0.138935, threads=0):
        pts = _f32(pts, (-1, 2))
        m = len(pts)
        q, radius = _f32(q), _f32(radius)
        out = np.zeros((m, 2), np.float32)
        c = Counters()
        self.lib.orc_acc_points(self.h, m, _ptr(pts), _ptr(q), _ptr(radius), k_e, _ptr(out),
                                threads, C.byref(c))
        return out, (c.visits, c.accepts, c.pairs)

    def direct_f64(self, pts, target_radius=None, k_e=0.138935, epsilon=2.0, threads=0):
        pts = _f32(pts, (-1, 2))
        tr = _f32(target_radius)
        out = np.zeros((len(pts), 2), np.float64)
        self.lib.orc_direct_f64(self.h, len(pts), _ptr(pts), _ptr(tr), k_e, epsilon,
                                _ptr(out, C.c_double), threads)
        return out

    def tree_neighbors(self, i, cutoff, cap=4096):
        out = np.zeros(cap, np.uint64)
        k = int(self.lib.orc_tree_neighbors(self.h, i, cutoff, _ptr(out, C.c_uint64), cap))
        return out[:min(k, cap)].astype(np.int64)

    # ---- cell list
    def cell_set_domain(self, hw, hh):
        self.lib.orc_cell_set_domain(self.h, hw, hh)

    def cell_rebuild(self, cell_size):
        self.lib.orc_cell_rebuild(self.h, cell_size)

    def cell_dims(self):
        gx, gy = C.c_uint64(), C.c_uint64()
        self.lib.orc_cell_dims(self.h, C.byref(gx), C.byref(gy))
        return gx.value, gy.value

    def cell_contents(self, cell, cap=4096):
        out = np.zeros(cap, np.uint64)
        k = int(self.lib.orc_cell_contents(self.h, cell, _ptr(out, C.c_uint64), cap))
        return out[:min(k, cap)].astype(np.int64)

    def cell_neighbors(self, i, cutoff, cap=4096):
        out = np.zeros(cap, np.uint64)
        k = int(self.lib.orc_cell_neighbors(self.h, i, cutoff, _ptr(out, C.c_uint64), cap))
        return out[:min(k, cap)].astype(np.int64)

    def cell_metal_neighbor_count(self, i, cutoff):
        return int(self.lib.orc_cell_metal_neighbor_count(self.h, i, cutoff))

    # ---- force phase + integrator
    def use_cell_list(self, hw, hh, threshold=0.001):
        return bool(self.lib.orc_use_cell_list(self.h, hw, hh, threshold))

    def reset_acc(self):
        self.lib.orc_reset_acc(self.h)

    def prepare_spatial_structures(self, hw, hh, threshold=0.001, threads=1):
        self.lib.orc_prepare_spatial_structures(self.h, hw, hh, threshold, threads)

    def attract(self, k_e, bg=(0.0, 0.0), threads=0):
        self.lib.orc_attract(self.h, k_e, bg[0], bg[1], threads)

    def apply_polar_forces(self, k_e, use_cell=True, dipole_model=1):
        self.lib.orc_apply_polar_forces(self.h, int(use_cell), k_e, dipole_model)

    def apply_lj_forces(self, use_cell=True, lj_force_max=200.0, collision_passes=7):
        self.lib.orc_apply_lj_forces(self.h, int(use_cell), lj_force_max, collision_passes)

    def apply_repulsive_forces(self, use_cell=True):
        self.lib.orc_apply_repulsive_forces(self.h, int(use_cell))

    def apply_stack_pressure(self, enabled, pressure, decay, hw):
        self.lib.orc_apply_stack_pressure(self.h, int(enabled), pressure, decay, hw)

    def iterate(self, dt, damping_base, hw, hh, hd=1.0, enable_out_of_plane=False, threads=0):
        self.lib.orc_iterate(self.h, dt, damping_base, hw, hh, hd, int(enable_out_of_plane), threads)

    def collide(self, domain_depth=1.0, num_passes=7, li_collision_softness=0.8, soft_collision_lithium_ion=True,
                soft_collision_anion=False):
        """one pass of collision::collide (pairs resolved in index order); returns the touching pairs"""
        return int(self.lib.orc_collide(self.h, domain_depth, num_passes, li_collision_softness,
                                        int(soft_collision_lithium_ion), int(soft_collision_anion)))

    def hop_alignment(self, src_idx, candidates, k_e, bg=(0.0, 0.0), alignment_bias=1.0):
        """simulation/electron_hopping.rs:283-329 per candidate: (local_field per donor, alignment per candidate)"""
        src = np.ascontiguousarray(src_idx, np.uint32)
        off = np.zeros(len(src) + 1, np.uint32)
        off[1:] = np.cumsum([len(c) for c in candidates])
        dst = (np.ascontiguousarray(np.concatenate([np.asarray(c, np.uint32) for c in candidates]), np.uint32)
               if len(src) and off[-1] else np.zeros(0, np.uint32))
        field = np.zeros((len(src), 2), np.float32)
        al = np.zeros(int(off[-1]), np.float32)
        self.lib.orc_hop_alignment(self.h, len(src), src.ctypes.data, off.ctypes.data, dst.ctypes.data, k_e, bg[0], bg[1],
                                   alignment_bias, field.ctypes.data, al.ctypes.data)
        return field, [al[off[i]:off[i + 1]] for i in range(len(src))]

    def update_electrons(self, bg, dt, k_e, threads=1):
        self.lib.orc_update_electrons(self.h, bg[0], bg[1], dt, k_e, threads)

    # ---- neighbour-count consumers
    def update_surrounded_flags(self, hw, hh, frame, radius_factor=4.0, neighbor_threshold=8, threshold=0.001):
        self.lib.orc_update_surrounded_flags(self.h, hw, hh, threshold, int(frame), radius_factor, int(neighbor_threshold))

    def surrounded(self):
        n = self.n
        flags, pos, frame = np.zeros(n, np.uint8), np.zeros((n, 2), np.float32), np.zeros(n, np.uint64)
        self.lib.orc_get_surrounded(self.h, flags.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p),
                                    frame.ctypes.data_as(C.c_void_p))
        return flags, pos, frame

    def enforce_metal_z_boundaries(self, max_z, hw, hh, threshold=0.001):
        self.lib.orc_enforce_metal_z_boundaries(self.h, max_z, hw, hh, threshold)

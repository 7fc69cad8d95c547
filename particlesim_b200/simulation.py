"""Host-side mirror of the reference's interface for the force hot path, over the C ABI.

Same names, argument meaning and side effects as the reference (paths relative to the reference):
  Quadtree        src/quadtree/quadtree.rs      new / build / build_with_domain / field / acc_pos /
                                                field_at_point / find_neighbors_within / nodes
  CellList        src/cell_list.rs              rebuild / find_neighbors_within / metal_neighbor_count
  forces.*        src/simulation/forces.rs      prepare_spatial_structures / attract / apply_lj_forces /
                                                apply_repulsive_forces / apply_stack_pressure
  Simulation      src/simulation/simulation.rs  iterate (:1437-1486), use_cell_list (:1798-1802), and
                                                the electron loop of step() (:1186-1196)

The reference is compiled Rust; with no Rust toolchain in this image the host side is Python over
ctypes (the Rust FFI crate a maintainer would use is under rust/).  Bodies live in a `Bodies` SoA on
the host, mirrored on the device by one context; like the reference's in-place partition,
`Quadtree.build` permutes `Bodies`.

All arithmetic happens in libpsim_b200.so on the GPU.  Nothing here computes forces on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as dc_field

import numpy as np

from . import _lib
from ._lib import PsimError

COULOMB_CONSTANT = np.float32(0.138935)  # f32 value of units.rs:32-34 (see species/units notes)


def coulomb_constant() -> np.float32:
    # units.rs:32-34 evaluated in f64 then cast
    k = 8.9875517923e9 * 1.602176634e-19 * 1.602176634e-19 * 1.0e-15 * 1.0e-15 / (
        1.66053906660e-27 * 1.0e-10 * 1.0e-10 * 1.0e-10)
    return np.float32(k)


def _p(a):
    return None if a is None else a.ctypes.data


@dataclass
class SimConfig:
    """The SimConfig fields the path reads (config.rs:337-339,409-417,467-469,512)."""
    coulomb_constant: float = float(coulomb_constant())
    damping_base: float = 1.0
    cell_list_density_threshold: float = 0.001
    stack_pressure_enabled: bool = False
    stack_pressure: float = 0.0
    stack_pressure_decay: float = 1.0
    enable_out_of_plane: bool = False


class Bodies:
    """Vec<Body> as a struct of arrays (body/types.rs:38-62, hot fields only) + flattened electrons."""

    FIELDS = ("pos", "z", "vel", "vz", "acc", "az", "mass", "radius", "charge", "species", "e_field", "id")

    def __init__(self, pos, z=None, vel=None, vz=None, mass=None, radius=None, charge=None, species=None,
                 ebody=None, erel=None, evel=None):
        self.pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2).copy()
        n = len(self.pos)
        f = lambda a, shape, fill=0.0: (np.full(shape, fill, np.float32) if a is None
                                        else np.ascontiguousarray(a, np.float32).reshape(shape).copy())
        self.z, self.vz = f(z, (n,)), f(vz, (n,))
        self.vel = f(vel, (n, 2))
        self.acc, self.az = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
        self.mass, self.radius, self.charge = f(mass, (n,), 1.0), f(radius, (n,)), f(charge, (n,))
        self.species = (np.zeros(n, np.uint8) if species is None
                        else np.ascontiguousarray(species, np.uint8).copy())
        self.e_field = np.zeros((n, 2), np.float32)
        self.id = np.arange(n, dtype=np.uint64)
        self.ebody = np.zeros(0, np.uint32) if ebody is None else np.ascontiguousarray(ebody, np.uint32).copy()
        self.erel = (np.zeros((0, 2), np.float32) if erel is None
                     else np.ascontiguousarray(erel, np.float32).reshape(-1, 2).copy())
        self.evel = (np.zeros((len(self.ebody), 2), np.float32) if evel is None
                     else np.ascontiguousarray(evel, np.float32).reshape(-1, 2).copy())

    def __len__(self):
        return len(self.pos)


class Quadtree:
    """quadtree.rs:11-34.  `nodes` is materialised on demand from the device tree."""
    ROOT = 0

    def __init__(self, theta, epsilon, leaf_capacity, thread_capacity):
        self.t_sq = np.float32(theta) * np.float32(theta)
        self.e_sq = np.float32(epsilon) * np.float32(epsilon)
        self.theta, self.epsilon = float(theta), float(epsilon)
        self.leaf_capacity, self.thread_capacity = int(leaf_capacity), int(thread_capacity)
        self._sim = None

    # the methods below are bound to a Simulation (which owns the device context)
    def build(self, bodies: "Bodies"):
        self._sim._build(_lib.BUILD_CONTAINING, 0.0, 0.0)

    def build_with_domain(self, bodies: "Bodies", domain_width, domain_height):
        self._sim._build(_lib.BUILD_DOMAIN, domain_width, domain_height)

    def field(self, bodies: "Bodies", k_e):
        """e_field[i] = acc_pos(pos_i, 1.0, radius_i)   (quadtree.rs:418-427)"""
        s = self._sim
        s._call("psim_field", np.float32(k_e), 0.0, 0.0, 0, _p(s.bodies.e_field), None)

    def acc_pos(self, pos, q, radius, bodies, k_e):
        """Batch form of quadtree.rs:350-407: pos (m,2); q, radius scalars or (m,) arrays."""
        s = self._sim
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        m = len(pos)
        qa = np.ascontiguousarray(np.broadcast_to(np.float32(q), (m,)), np.float32)
        ra = np.ascontiguousarray(np.broadcast_to(np.float32(radius), (m,)), np.float32)
        out = np.zeros((m, 2), np.float32)
        s._call("psim_acc_points", m, _p(pos), _p(qa), _p(ra), np.float32(k_e), _p(out))
        return out

    def field_at_point(self, bodies, pos, k_e):
        """quadtree.rs:504-507: acc_pos(pos, 1.0, 0.0)"""
        s = self._sim
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        out = np.zeros((len(pos), 2), np.float32)
        s._call("psim_acc_points", len(pos), _p(pos), None, None, np.float32(k_e), _p(out))
        return out

    def find_neighbors_within(self, bodies, i, cutoff):
        """quadtree.rs:430-501.  Same result set as the grid query; served from the cell grid."""
        return self._sim._neighbors([i], cutoff, False)[0]

    @property
    def nodes(self) -> np.ndarray:
        s = self._sim
        cnt = C.c_uint64(0)
        s._call("psim_download_nodes", None, 0, C.byref(cnt))
        out = np.zeros(cnt.value, dtype=_lib.NODE_DTYPE)
        if cnt.value:
            s._call("psim_download_nodes", out.ctypes.data, cnt.value, C.byref(cnt))
        return out

    def keys(self) -> np.ndarray:
        s = self._sim
        out = np.zeros(len(s.bodies), np.uint64)
        s._call("psim_get_keys", _p(out))
        return out


class CellList:
    """cell_list.rs:4-25"""

    def __init__(self, domain_width, domain_height, cell_size):
        self.domain_width, self.domain_height, self.cell_size = float(domain_width), float(domain_height), float(cell_size)
        self._sim = None

    def update_domain_size(self, domain_width, domain_height):
        self.domain_width, self.domain_height = float(domain_width), float(domain_height)

    def rebuild(self, bodies: "Bodies"):
        self._sim._call("psim_cell_build", self.domain_width, self.domain_height, self.cell_size)

    def find_neighbors_within(self, bodies, i, cutoff):
        return self._sim._neighbors([i], cutoff, False)[0]

    def metal_neighbor_count(self, bodies, i, cutoff):
        return len(self._sim._neighbors([i], cutoff, True)[0])

    def cells(self):
        """(grid_size_x, grid_size_y, offsets, indices): per-cell body lists as CSR"""
        s = self._sim
        gx, gy = C.c_uint64(), C.c_uint64()
        s._call("psim_cell_download", C.byref(gx), C.byref(gy), None, None)
        off = np.zeros(gx.value * gy.value + 1, np.uint32)
        idx = np.zeros(len(s.bodies), np.uint32)
        s._call("psim_cell_download", C.byref(gx), C.byref(gy), _p(off), _p(idx))
        return gx.value, gy.value, off, idx


class Simulation:
    """The slice of `Simulation` the hot path touches (simulation.rs:83-139 fields)."""

    def __init__(self, bodies: Bodies, domain_width, domain_height, domain_depth=1.0, dt=5.0,
                 theta=1.0, epsilon=2.0, leaf_capacity=1, thread_capacity=1024, config: SimConfig | None = None,
                 device=0, parity_mode=True, node_factor=4.0, species_table=None, max_bodies=None,
                 max_electrons=None, stream=0, strict_centres=True):
        self.lib = _lib.load()
        self.bodies = bodies
        self.domain_width, self.domain_height, self.domain_depth = float(domain_width), float(domain_height), float(domain_depth)
        self.dt = float(dt)
        self.config = config or SimConfig()
        self.background_e_field = (0.0, 0.0)
        self.quadtree = Quadtree(theta, epsilon, leaf_capacity, thread_capacity)
        self.cell_list = CellList(domain_width, domain_height, 1.0)
        self.quadtree._sim = self
        self.cell_list._sim = self
        cfg = _lib.default_config(theta=theta, epsilon=epsilon, leaf_capacity=leaf_capacity,
                                  thread_capacity=thread_capacity, parity_mode=int(parity_mode),
                                  node_factor=node_factor, strict_centres=int(strict_centres),
                                  stack_pressure_enabled=int(self.config.stack_pressure_enabled),
                                  stack_pressure=self.config.stack_pressure,
                                  stack_pressure_decay=self.config.stack_pressure_decay)
        self._cfg = cfg
        h = C.c_void_p()
        nmax = max(int(max_bodies or 0), len(bodies), 1)
        emax = max(int(max_electrons or 0), len(bodies.ebody), 1)
        rc = self.lib.psim_create(device, nmax, emax, C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PsimError(rc, "psim_create failed (no CUDA device or out of memory); there is no CPU fallback")
        self.h = h
        if stream:
            self._call("psim_set_stream", int(stream))
        self.species_table = _lib.default_species_table() if species_table is None else np.ascontiguousarray(species_table, _lib.SPECIES_DTYPE)
        self._call("psim_upload_species_table", self.species_table.ctypes.data, len(self.species_table))
        self.upload()

    def close(self):
        if getattr(self, "h", None):
            self.lib.psim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def _call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise PsimError(rc, self.lib.psim_last_error(self.h).decode())
        return rc

    def upload(self):
        b = self.bodies
        self._call("psim_upload_bodies", len(b), _p(b.pos), _p(b.z), _p(b.vel), _p(b.vz), _p(b.mass),
                   _p(b.radius), _p(b.charge), _p(b.species))
        if len(b.ebody):
            self._call("psim_upload_electrons", len(b.ebody), _p(b.ebody), _p(b.erel), _p(b.evel))
        self._orig = np.arange(len(b), dtype=np.uint32)

    def download(self, fields=("pos", "vel", "acc", "e_field")):
        b = self.bodies
        want = lambda k: _p(getattr(b, k)) if k in fields else None
        self._call("psim_download_bodies", want("pos"), want("z"), want("vel"), want("vz"), want("acc"),
                   want("az"), want("mass"), want("radius"), want("charge"), want("species"),
                   want("e_field"), None)

    def download_electrons(self):
        b = self.bodies
        if len(b.ebody):
            self._call("psim_download_electrons", _p(b.ebody), _p(b.erel), _p(b.evel))

    def _build(self, mode, hw, hh):
        self._call("psim_build", mode, np.float32(hw), np.float32(hh))
        self._apply_permutation()

    def _apply_permutation(self):
        """reorder the host Vec<Body> the way the reference's in-place partition would"""
        b = self.bodies
        n = len(b)
        if n == 0:
            return
        perm = np.zeros(n, np.uint32)
        self._call("psim_get_permutation", _p(perm))
        for k in Bodies.FIELDS:
            setattr(b, k, np.ascontiguousarray(getattr(b, k)[perm]))
        if len(b.ebody):
            self.download_electrons()
        self.last_permutation = perm

    def _neighbors(self, idx, cutoff, metals_only):
        q = np.ascontiguousarray(idx, np.uint32)
        m = len(q)
        off = np.zeros(m + 1, np.uint32)
        tot = C.c_uint64()
        self._call("psim_neighbors_within", m, _p(q), np.float32(cutoff), int(metals_only), _p(off), None, 0, C.byref(tot))
        ind = np.zeros(max(tot.value, 1), np.uint32)
        if tot.value:
            self._call("psim_neighbors_within", m, _p(q), np.float32(cutoff), int(metals_only), _p(off), _p(ind),
                       tot.value, C.byref(tot))
        return [ind[off[k]:off[k + 1]].astype(np.int64) for k in range(m)]

    def force_cell_size(self) -> float:
        """forces.rs:17-22: max(3 * max_lj_cutoff, max_repulsion_cutoff, max_lj_cutoff)"""
        t = self.species_table
        lj = np.float32(0.0)
        rep = np.float32(0.0)
        for r in t:
            if r["lj_enabled"]:
                lj = max(lj, np.float32(r["lj_cutoff"]) * np.float32(r["lj_sigma"]))
            if r["repulsion_enabled"]:
                rep = max(rep, np.float32(r["repulsion_cutoff"]))
        return float(max(np.float32(3.0) * lj, rep, lj))

    def step_cell_size(self, do_polar: bool = False) -> float:
        """the cell size psim_step bins at: the largest cutoff its short-range passes use (the pair sets of the polar /
        LJ / repulsion passes do not depend on the cell size; the polar pass reaches 3 x the radius of EC / DMC)"""
        t = self.species_table
        lj = np.float32(0.0)
        rep = np.float32(0.0)
        for r in t:
            if r["lj_enabled"]:
                lj = max(lj, np.float32(r["lj_cutoff"]) * np.float32(r["lj_sigma"]))
            if r["repulsion_enabled"]:
                rep = max(rep, np.float32(r["repulsion_cutoff"]))
        if not do_polar:
            return float(max(rep, lj))
        polar = max(np.float32(3.0) * np.float32(t[s]["radius"]) for s in (4, 5) if s < len(t))  # EC / DMC, forces.rs:74
        return float(max(rep, lj, polar))

    def stats(self) -> dict:
        st = _lib.Stats()
        self._call("psim_stats_get", C.byref(st))
        d = {k: getattr(st, k) for k, _ in _lib.Stats._fields_}
        d["root_center"] = tuple(st.root_center)
        return d

    def build_info(self) -> dict:
        """how the last build made the node charges (psim_build_info)"""
        out = (C.c_uint64 * 4)()
        self._call("psim_build_info", out)
        return {"integer_charges": bool(out[0]), "charged_bodies": int(out[1]), "abs_charge_sum": int(out[2]),
                "non_integer_charges": int(out[3])}

    def sync(self):
        self._call("psim_sync")

    # ---- Simulation methods on the path
    def use_cell_list(self) -> bool:
        """simulation.rs:1798-1802"""
        return bool(self.lib.psim_use_cell_list(self.h, self.domain_width, self.domain_height,
                                                self.config.cell_list_density_threshold))

    def reset_acc(self):
        """simulation.rs:1000-1003"""
        self._call("psim_reset_acc")
        self.bodies.acc[:] = 0
        self.bodies.az[:] = 0

    def iterate(self):
        """simulation.rs:1437-1486"""
        self._call("psim_iterate", self.dt, self.config.damping_base, self.domain_width, self.domain_height,
                   self.domain_depth, int(self.config.enable_out_of_plane))
        self.download(("pos", "vel", "z", "vz"))

    def update_surrounded_flags(self, radius_factor=4.0, neighbor_threshold=8):
        """simulation.rs:1893-1918 (+ Body::maybe_update_surrounded, CellList::metal_neighbor_count).  Uses
        self.frame like the reference; returns the flags in the current body order."""
        self._call("psim_update_surrounded_flags", self.domain_width, self.domain_height,
                   int(getattr(self, "frame", 0)), np.float32(radius_factor), int(neighbor_threshold))
        return self.surrounded()[0]

    def surrounded(self):
        """(surrounded_by_metal, last_surround_pos, last_surround_frame) per body, current order"""
        n = len(self.bodies)
        flags, pos, frame = np.zeros(n, np.uint8), np.zeros((n, 2), np.float32), np.zeros(n, np.uint64)
        self._call("psim_get_surrounded", _p(flags), _p(pos), _p(frame))
        return flags, pos, frame

    def enforce_metal_z_boundaries(self, max_z):
        """simulation/out_of_plane.rs:140-254"""
        self._call("psim_enforce_metal_z_boundaries", np.float32(max_z), self.domain_width, self.domain_height)
        self.download(("z", "vz"))

    def collide(self, passes=None, num_passes=7, li_collision_softness=0.8, soft_collision_lithium_ion=True,
                soft_collision_anion=False):
        """collision::collide, `passes` passes (default: num_passes of them, like Simulation::step,
        simulation.rs:1025-1028); returns the number of touching pairs the last pass found"""
        tp = C.c_uint64()
        self._call("psim_collide", self.domain_width, self.domain_height, self.domain_depth,
                   int(num_passes if passes is None else passes), int(num_passes), np.float32(li_collision_softness),
                   int(soft_collision_lithium_ion), int(soft_collision_anion), C.byref(tp))
        self.download(("pos", "vel", "z", "vz"))
        return int(tp.value)

    def hop_alignment(self, src_idx, candidates, alignment_bias=1.0):
        """simulation/electron_hopping.rs:283-329 for a batch: `candidates[i]` lists the acceptor indices of donor
        src_idx[i] (current body order).  Returns (local_field per donor, list of alignment arrays)."""
        src = np.ascontiguousarray(src_idx, np.uint32)
        off = np.zeros(len(src) + 1, np.uint32)
        off[1:] = np.cumsum([len(c) for c in candidates])
        dst = np.ascontiguousarray(np.concatenate([np.asarray(c, np.uint32) for c in candidates]) if len(src) and off[-1]
                                   else np.zeros(0, np.uint32), np.uint32)
        field = np.zeros((len(src), 2), np.float32)
        al = np.zeros(int(off[-1]), np.float32)
        bg = self.background_e_field
        self._call("psim_hop_alignment", len(src), _p(src), _p(off), _p(dst), self.config.coulomb_constant,
                   bg[0], bg[1], np.float32(alignment_bias), _p(field), _p(al))
        return field, [al[off[i]:off[i + 1]] for i in range(len(src))]

    def update_electrons(self):
        """the loop at simulation.rs:1186-1196 over Body::update_electrons (body/electron.rs:19-46)"""
        self._call("psim_update_electrons", self.background_e_field[0], self.background_e_field[1], self.dt,
                   self.config.coulomb_constant)
        self.download_electrons()

    def step_params(self, do_short_range=True, do_electrons=True, do_iterate=True, do_polar=True) -> "_lib.StepParams":
        p = _lib.StepParams()
        p.hw, p.hh, p.hd = self.domain_width, self.domain_height, self.domain_depth
        p.dt, p.damping_base = self.dt, self.config.damping_base
        p.k_e = self.config.coulomb_constant
        p.bg_x, p.bg_y = self.background_e_field
        p.density_threshold = self.config.cell_list_density_threshold
        p.enable_out_of_plane = int(self.config.enable_out_of_plane)
        p.do_short_range, p.do_electrons, p.do_iterate = int(do_short_range), int(do_electrons), int(do_iterate)
        p.do_polar = int(do_polar)
        return p

    def step_device(self, params=None):
        """The hot path of Simulation::step (simulation.rs:1000-1196) without host round trips."""
        p = params or self.step_params()
        self._call("psim_step", C.byref(p))

    def step_host(self, pos, vel=None, charge=None, params=None, out=None):
        """psim_step_host: state refresh from host arrays + the hot path + read-back, pipelined.
        Arrays are float32 numpy arrays in the order of the previous step_host outputs (first call: the
        device's body order); `out` may hold preallocated (ideally page-locked) 'pos', 'vel', 'e_field',
        'orig' arrays.  Returns the dict of outputs, rows in the order of the step's first build."""
        p = params or self.step_params()
        n = len(self.bodies)

        def ptr(a, dt, shape):
            if a is None:
                return None
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"] and a.shape == shape, "step_host: bad array"
            return a.ctypes.data_as(C.c_void_p)

        if out is None:
            out = dict(pos=np.empty((n, 2), np.float32), vel=np.empty((n, 2), np.float32),
                       e_field=np.empty((n, 2), np.float32), orig=np.empty(n, np.uint32))
        self._call("psim_step_host", C.byref(p), n, ptr(pos, np.float32, (n, 2)), ptr(vel, np.float32, (n, 2)),
                   ptr(charge, np.float32, (n,)), ptr(out.get("pos"), np.float32, (n, 2)),
                   ptr(out.get("vel"), np.float32, (n, 2)), ptr(out.get("e_field"), np.float32, (n, 2)),
                   ptr(out.get("orig"), np.uint32, (n,)))
        return out


class forces:
    """src/simulation/forces.rs — free functions taking the simulation, like the reference."""

    @staticmethod
    def prepare_spatial_structures(sim: Simulation):
        sim._call("psim_prepare_spatial_structures", sim.domain_width, sim.domain_height,
                  sim.config.cell_list_density_threshold)
        sim._apply_permutation()
        st = sim.stats()
        sim.cell_list.cell_size = 0.0 if st["grid_x"] == 0 else sim.cell_list.cell_size

    @staticmethod
    def attract(sim: Simulation):
        bg = sim.background_e_field
        sim._call("psim_field", sim.config.coulomb_constant, bg[0], bg[1], 1, _p(sim.bodies.e_field), _p(sim.bodies.acc))

    @staticmethod
    def apply_polar_forces(sim: Simulation, dipole_model=1):
        """forces.rs:52-175 (ConjugatePair by default, config.rs:278-281)"""
        sim._call("psim_apply_polar_forces", sim.config.coulomb_constant, int(dipole_model))
        sim.download(("acc",))

    @staticmethod
    def apply_lj_forces(sim: Simulation):
        sim._call("psim_short_range", _lib.SR_LJ)
        sim.download(("acc",))

    @staticmethod
    def apply_repulsive_forces(sim: Simulation):
        sim._call("psim_short_range", _lib.SR_REPULSION)
        sim.download(("acc",))

    @staticmethod
    def apply_stack_pressure(sim: Simulation):
        sim._call("psim_short_range", _lib.SR_STACK_PRESSURE)
        sim.download(("acc",))

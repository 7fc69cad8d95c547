"""Multi-GPU hot path: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Body state is replicated; the WORK is sharded two ways:
  targets  rank r owns a contiguous slice of the Morton order (a multiple of 64 bodies) and computes field /
           short-range forces / integrator only for those targets, plus an equal slice of the bound
           electrons; the owned slices are all-gathered in place on the library's device arrays after the
           integrator (velocities behind the next build's position-only phases) and after the electron update.
           No force reduction is needed: a target's field depends only on the read-only source tree.
  build    rank r owns a contiguous range of the KEY order (bins of 65 536 top-level cells) and builds the
           part of the tree that starts in it (csrc/shard.cuh, psim_shard_phase); `sharded_build` runs the
           phases and the exchanges between them: sorted-index segments and traversal-node segments
           (uneven all-gathers through one padded ncclAllGather each), two per-bin tables and the top heap
           (all-reduce).  The pieces concatenate into exactly the single-GPU arrays.
`local_build=False` keeps the first scheme (every rank builds the whole tree) for comparison.

The partition and exchange helpers are pure functions / small classes so that the plumbing is testable with
the gloo backend on the CPU (tests/test_parallel_cpu.py); LoopbackComm runs all ranks in one process for the
single-GPU tests of the sharded build (tests/test_gpu_shard.py).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .simulation import Simulation


def shard_width(n: int, world: int) -> int:
    """bodies per rank (the last ranks may own fewer); arrays are padded to world * width.  A multiple of
    64 (two groups), so that the 32-target groups of the traversal are the ones the single-GPU run forms and the sharded
    result stays bit-identical (the order of a target's additions depends on its group)"""
    return ((n + world - 1) // world + 63) // 64 * 64 if n else 0


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """(first, count) of the slice of [0, n) that `rank` owns"""
    w = shard_width(n, world)
    first = min(rank * w, n)
    return first, min(w, n - first)


def all_gather_slices(full, width: int, rank: int, world: int, dist, scratch=None):
    """In-place all-gather of equal slices of a (world * width, ...) tensor: rank r contributes rows
    [r * width, (r + 1) * width).  Works for CUDA tensors over NCCL and CPU tensors over gloo."""
    mine = full[rank * width:(rank + 1) * width]
    if scratch is None:
        scratch = mine.clone()
    else:
        scratch.copy_(mine)
    dist.all_gather_into_tensor(full[:world * width], scratch)
    return full


def all_gatherv(full, offsets, rank: int, dist):
    """In-place all-gather of UNEVEN row segments: rank r contributes rows [offsets[r], offsets[r+1]) of
    `full`.  One broadcast per non-empty segment (every rank knows all the sizes, so no size exchange)."""
    for r in range(len(offsets) - 1):
        a, b = int(offsets[r]), int(offsets[r + 1])
        if b > a:
            dist.broadcast(full[a:b], src=r)
    return full


def all_gatherv_padded_begin(full, offsets, rank: int, dist, stage_cache: dict, async_op: bool = False):
    """The same exchange as ONE all-gather: every rank contributes a block of max-segment rows starting at
    its segment (rows past its count are padding), the blocks land in a staging tensor and the valid rows
    are copied into place by all_gatherv_padded_end.  One ncclAllGather at full NVSwitch bandwidth instead
    of world broadcasts; with async_op the caller can queue independent work before the end call."""
    world = len(offsets) - 1
    counts = [int(offsets[r + 1]) - int(offsets[r]) for r in range(world)]
    seg = max(counts)
    if seg == 0:
        return None
    seg = (seg + 1023) // 1024 * 1024  # few distinct staging sizes across builds
    key = (full.data_ptr(), full.dtype, tuple(full.shape[1:]))
    need = world * seg
    stage = stage_cache.get(key)
    if stage is None or stage.shape[0] < need + seg:
        stage = full.new_empty((int((need + seg) * 1.25),) + tuple(full.shape[1:]))
        stage_cache[key] = stage
    a = int(offsets[rank])
    if a + seg <= full.shape[0]:
        mine = full[a:a + seg]
    else:  # the padded block would run past the end of the array: go through the spare block
        mine = stage[need:need + seg]
        mine[:counts[rank]].copy_(full[a:a + counts[rank]])
    work = dist.all_gather_into_tensor(stage[:need], mine, async_op=async_op)
    return full, offsets, rank, counts, seg, stage, work


def all_gatherv_padded_end(pending):
    if pending is None:
        return
    full, offsets, rank, counts, seg, stage, work = pending
    if work is not None:
        work.wait()
    for r in range(len(counts)):
        if r != rank and counts[r]:
            full[int(offsets[r]):int(offsets[r + 1])].copy_(stage[r * seg:r * seg + counts[r]])


def all_gatherv_padded(full, offsets, rank: int, dist, stage_cache: dict):
    all_gatherv_padded_end(all_gatherv_padded_begin(full, offsets, rank, dist, stage_cache))
    return full


class DistComm:
    """exchanges of the sharded build over torch.distributed (NCCL on GPUs, gloo in the CPU tests); the
    lists hold one tensor per rank that lives in this process, i.e. exactly one"""

    def __init__(self, dist, rank: int, padded: bool = True):
        self.dist, self.rank, self.padded = dist, rank, padded
        self._stage = {}

    def all_gatherv(self, tensors, offsets):
        self.all_gatherv_end(self.all_gatherv_begin(tensors, offsets, async_op=False))

    def all_gatherv_begin(self, tensors, offsets, async_op=True):
        """start the exchange; with the padded NCCL path it runs beside whatever is queued before _end"""
        if self.padded:
            return all_gatherv_padded_begin(tensors[0], offsets, self.rank, self.dist, self._stage, async_op)
        all_gatherv(tensors[0], offsets, self.rank, self.dist)
        return None

    def all_gatherv_end(self, pending):
        all_gatherv_padded_end(pending)

    def all_reduce(self, tensors):
        self.dist.all_reduce(tensors[0])


class LoopbackComm:
    """all ranks live in this process (one context each, e.g. on one GPU): the exchanges are copies.  Used
    by the single-GPU tests of the sharded build."""

    def all_gatherv(self, tensors, offsets):
        for r, src in enumerate(tensors):
            a, b = int(offsets[r]), int(offsets[r + 1])
            for q, dst in enumerate(tensors):
                if q != r and b > a:
                    dst[a:b].copy_(src[a:b])

    def all_gatherv_begin(self, tensors, offsets, async_op=True):
        self.all_gatherv(tensors, offsets)
        return None

    def all_gatherv_end(self, pending):
        pass

    def all_reduce(self, tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t
        for t in tensors:
            t.copy_(total)


def shard_views(sim, torch):
    """torch views of the exchange buffers of psim_shard_ptrs (bit containers: int32 / int64)"""
    p = np.zeros(8, np.uint64)
    sim._call("psim_shard_ptrs", p.ctypes.data)
    mk = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device="cuda")
    cap = int(p[7])
    return dict(order=mk(p[0], (int(sim._shard_nb),), "<i4"), xbuf=mk(p[1], (int(p[5]),), "<i8"),
                heap=mk(p[2], (int(p[6]),), "<i8"), travA=mk(p[3], (cap, 4), "<i4"), travB=mk(p[4], (cap, 4), "<i4"))


def sharded_build(sims, mode: int, hw: float, hh: float, comm, torch, mark=None, overlap=None, before_gather=None):
    """Quadtree::build / build_with_domain across ranks (include/psim_b200.h, psim_shard_phase): `sims`
    are the contexts of the ranks that live in this process (one under torchrun).  `mark(name)`, if given,
    is called at every phase / exchange boundary (tools/shard_phases.py records CUDA events there);
    `overlap()`, if given, queues work that does not need the tree (the cell-list rebuild) while the
    traversal segments travel; `before_gather()` runs before the phase that permutes the body arrays (the
    sharded step waits there for the velocities it all-gathers behind the first two phases)."""
    mark = mark or (lambda name: None)
    world = sims[0].world
    lo = [np.zeros(world + 1, np.uint32) for _ in sims]
    views = [s._shard_view_cache if getattr(s, "_shard_view_cache", None) else None for s in sims]
    for k, s in enumerate(sims):
        if views[k] is None:
            views[k] = s._shard_view_cache = shard_views(s, torch)

    def phase(k, out=None):
        for i, s in enumerate(sims):
            s._call("psim_shard_phase", k, mode, hw, hh, out[i].ctypes.data if out is not None else None)

    mark("start")
    phase(0, lo)
    mark("p0 keys+bins (replicated)")
    phase(1)
    mark("p1 select+sort")
    comm.all_gatherv([v["order"] for v in views], lo[0])
    mark("x1 order all-gather")
    if before_gather is not None:
        before_gather()
    phase(2)
    mark("p2 gather (replicated)+levels")
    comm.all_reduce([v["xbuf"] for v in views])
    mark("x2 table")
    phase(3)
    mark("p3 emit+sweeps")
    comm.all_reduce([v["heap"] for v in views])
    mark("x3 heap")
    phase(4)
    mark("p4 top+finalize+scan")
    comm.all_reduce([v["xbuf"] for v in views])
    mark("x4 table")
    tl = [np.zeros(world + 1, np.uint32) for _ in sims]
    phase(5, tl)
    mark("p5 compact")
    pa = comm.all_gatherv_begin([v["travA"] for v in views], tl[0])
    pb = comm.all_gatherv_begin([v["travB"] for v in views], tl[0])
    if overlap is not None:
        overlap()
    comm.all_gatherv_end(pa)
    comm.all_gatherv_end(pb)
    mark("x5 traversal all-gather")
    phase(6)
    mark("p6 children links")
    return lo[0], tl[0]


class _DevArray:
    """zero-copy view of library-owned device memory for torch (CUDA array interface v2)"""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class ShardedSimulation(Simulation):
    """One rank of the multi-GPU hot path.  orchestration = "library" (default): psim_comm_init + psim_step_sharded,
    the library runs the phases and the NCCL exchanges itself (the unique id travels over torch.distributed);
    "python": the same sequence driven from here through psim_shard_phase and torch.distributed collectives (kept as
    the readable reference of the protocol and for the comparison in tools/)."""

    def __init__(self, bodies, domain_width, domain_height, *, rank: int, world: int, local_build: bool = True,
                 orchestration: str = "library", **kw):
        n, m = len(bodies), len(bodies.ebody)
        self.rank, self.world = rank, world
        self.wb, self.we = shard_width(n, world), shard_width(m, world)
        kw["max_bodies"] = max(self.wb * world, 1)
        kw["max_electrons"] = max(self.we * world, 1)
        super().__init__(bodies, domain_width, domain_height, **kw)
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self._scratch_b = torch.empty((self.wb, 4), dtype=torch.float32, device="cuda")
        self._scratch_v = torch.empty((self.wb, 4), dtype=torch.float32, device="cuda")
        self._scratch_e = torch.empty((self.we, 2), dtype=torch.float32, device="cuda")
        self._shard_nb = kw["max_bodies"]
        self.local_build = local_build and world > 1
        self.library = orchestration == "library" and self.local_build
        if self.library:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = np.zeros(128, np.uint8)
                rc = self.lib.psim_comm_unique_id(buf.ctypes.data)
                if rc != 0:
                    raise _lib.PsimError(rc, "psim_comm_unique_id: NCCL not available")
                uid = torch.from_numpy(buf)
            uid = uid.cuda()
            dist.broadcast(uid, src=0)
            self._uid = uid.cpu().numpy().copy()
            self._call("psim_comm_init", self._uid.ctypes.data, rank, world)
        elif self.local_build:
            self._call("psim_shard_init", rank, world)
            self._comm = DistComm(dist, rank)
        f, c = shard_range(n, world, rank)
        self._call("psim_set_target_range", f, c)
        if m:
            f, c = shard_range(m, world, rank)
            self._call("psim_set_electron_range", f, c)

    def _views(self):
        p = np.zeros(8, np.uint64)
        self._call("psim_device_ptrs", p.ctypes.data)
        t = self.torch
        nb, ne = self.wb * self.world, self.we * self.world
        mk = lambda ptr, rows, cols: t.as_tensor(_DevArray(ptr, (rows, cols)), device="cuda")
        return dict(pqr=mk(p[0], nb, 4), velz=mk(p[1], nb, 4), erel=mk(p[4], ne, 2) if ne else None,
                    evel=mk(p[5], ne, 2) if ne else None)

    def _build(self, mode, hw, hh, overlap=None, before_gather=None):
        """each rank builds the part of the tree that starts in its key range (local_build), or every rank
        builds the whole tree (the v1 scheme, kept for comparison)"""
        if self.local_build:
            sharded_build([self], mode, hw, hh, self._comm, self.torch, overlap=overlap, before_gather=before_gather)
        else:
            if before_gather is not None:
                before_gather()
            self._call("psim_build_async", mode, hw, hh)
            if overlap is not None:
                overlap()

    def step_device(self, params=None, record=False):
        """Simulation::step's hot path (simulation.rs:1000-1196), sharded.  record=True keeps CUDA events
        at the phase boundaries (see phase_ms)."""
        p = params or self.step_params()
        C = self._call
        if self.library:
            import ctypes
            C("psim_step_sharded", ctypes.byref(p))
            self._events = None
            return
        ev = []

        def mark():
            if record:
                e = self.torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append(e)

        mark()
        pending_vel = None
        C("psim_reset_acc")
        cell = self.step_cell_size(bool(getattr(p, "do_polar", 0)))  # what psim_step uses: same addition order
        do_cells = bool(p.do_short_range and cell > 0.0)
        # the cell-list rebuild only needs the sorted bodies: it runs while the tree pieces travel
        self._build(_lib.BUILD_CONTAINING, 0.0, 0.0,
                    overlap=(lambda: C("psim_cell_build", p.hw, p.hh, cell)) if do_cells else None)
        mark()
        mark()
        C("psim_field", p.k_e, p.bg_x, p.bg_y, 1, None, None)
        if p.do_polar and do_cells:  # forces::apply_polar_forces (simulation.rs:1007), between attract and the LJ pass
            C("psim_apply_polar_forces", p.k_e, 1)
        mark()
        if p.do_short_range:
            C("psim_short_range", _lib.SR_LJ | _lib.SR_REPULSION | _lib.SR_STACK_PRESSURE)
        mark()
        if p.do_iterate:
            C("psim_iterate", p.dt, p.damping_base, p.hw, p.hh, p.hd, int(p.enable_out_of_plane))
            v = self._views()
            all_gather_slices(v["pqr"], self.wb, self.rank, self.world, self.dist, self._scratch_b)
            # the velocities are not needed before the next build permutes the bodies: their all-gather runs
            # behind that build's first two phases (bins, own keys + sort), which only read positions
            self._scratch_v.copy_(v["velz"][self.rank * self.wb:(self.rank + 1) * self.wb])
            pending_vel = self.dist.all_gather_into_tensor(v["velz"][:self.world * self.wb], self._scratch_v,
                                                           async_op=True)
            C("psim_mark_positions_changed")
        mark()
        wait_vel = (lambda: pending_vel.wait()) if pending_vel is not None else None
        if p.do_electrons:
            self._build(_lib.BUILD_DOMAIN, p.hw, p.hh, before_gather=wait_vel)
            mark()
            C("psim_update_electrons", p.bg_x, p.bg_y, p.dt, p.k_e)
            if self.we:
                v = self._views()
                all_gather_slices(v["erel"], self.we, self.rank, self.world, self.dist, self._scratch_e)
                all_gather_slices(v["evel"], self.we, self.rank, self.world, self.dist, self._scratch_e)
        else:
            if wait_vel is not None:
                wait_vel()
            mark()
        mark()
        self._events = ev

    def phase_ms(self):
        """device ms of the phases of the last recorded step (same order as psim_phase_times; the two
        exchanges are inside `iterate` and `electron_updates`)"""
        ev = self._events
        if ev is None:  # library orchestration: the context's own phase events
            ph = np.zeros(8, np.float32)
            self._call("psim_phase_times", ph.ctypes.data)
            return [float(v) for v in ph]
        self.torch.cuda.synchronize()
        out = [ev[k].elapsed_time(ev[k + 1]) for k in range(7)]
        out.append(ev[0].elapsed_time(ev[7]))
        return out

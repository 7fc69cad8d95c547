#!/usr/bin/env python
"""bench.py — the hot path of Simulation::step on synthetic charged-particle sets.

A "step" = one pass of the hot path over the whole body set, in Simulation::step's order
(reference simulation.rs:1000-1196): reset acc, quadtree build (tight AABB), cell list, Coulomb field +
attract, LJ / repulsion / stack pressure, integrator, domain-bounded quadtree build, electron field
sampling + drift.  Metric: Mparticles/s = bodies / step time (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--bodies BODIES] [--theta T]
  python bench.py --impl reference ...    # the C++ restatement of the reference's rayon path on the host cores

Workload at N=1: BASELINE.json configs[3], "N=16M uniform electrolyte with electron polarization
field sampling" (the configuration the 100x target is quoted on); theta is the reference default 1.0.
Timing: CUDA events on the stream the kernels are launched on, W warm-up steps, K timed steps between
barrier + synchronize, max over ranks.  The body set (16 M x ~140 B of device state plus ~28 M tree nodes)
is far larger than the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "particle-force evals/sec (Mparticles/s per step)"
UNIT = "Mparticles/s"
PHASES = ["quadtree_build", "cell_list_rebuild", "quadtree_field", "forces_lj", "iterate",
          "quadtree_build_domain", "electron_updates", "step"]


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed
    ncu --set full capture of the same workload (profiles/); null for other sizes"""
    try:
        import glob
        t = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]))  # latest round
        if int(t["n_bodies"]) == int(n):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(n, seed=0xC0FFEE):
    from helpers import electrolyte
    return electrolyte(n, seed=seed)


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(bodies, theta, threads, steps=1, warmup=0, variant="native", all_parallel=False):
    """One hot-path step with the oracle, structured like the reference: rayon-parallel field / iterate
    / tree build workers, SERIAL propagate, LJ, repulsion and electron loop (as in the reference).
    all_parallel also runs the electron loop on every thread (reported beside the reference-shaped
    number so that the ratio is not inflated by the reference's serial sections; propagate and the LJ loop
    stay serial).  Returns (seconds per step, threads)."""
    from helpers import KE
    from oracle import pyoracle
    try:
        pyoracle.load(variant)
    except Exception:
        variant = ""
    from helpers import oracle_for
    o = oracle_for(bodies, theta=theta, variant=variant)
    hw, hh = bodies["hw"], bodies["hh"]
    T = threads
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        o.reset_acc()
        o.prepare_spatial_structures(hw, hh, threads=T)
        o.attract(KE, threads=T)
        o.apply_lj_forces(True)
        o.apply_repulsive_forces(True)
        o.iterate(5.0, 1.0, hw, hh, 1.0, False, threads=T)
        o.build_with_domain(hw, hh, threads=T)
        o.update_electrons((0.0, 0.0), 5.0, KE, threads=T if all_parallel else 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times)), T


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be
    built in this image (no cargo, un-vendored quarkstrom), so this is the C++ restatement (oracle/),
    kind "port", with every host thread the reference's rayon pool would use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    try:
        pyoracle.build(native=True)
        variant = "native"
    except Exception:
        variant = ""
    threads = pyoracle.load(variant).orc_max_threads()
    # bounded sample: a smaller instance of the same generator, sized so K + W steps take ~2 minutes
    n_probe = min(args.n, 100_000)
    t_probe, _ = cpu_reference_step(make_workload(n_probe), args.theta, threads, 1, 0, variant)
    per_body = t_probe / n_probe
    budget = 120.0 / max(1, args.steps + args.warmup)
    n_s = int(min(args.n, max(100_000, min(2_000_000, budget / (per_body * 1.3)))))
    bodies = make_workload(n_s)
    t_step, _ = cpu_reference_step(bodies, args.theta, threads, args.steps, args.warmup, variant)
    value = n_s / t_step / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_s),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_s}-body instance of the same generator (full {args.n} would take minutes per step); "
                                   "C++ restatement of the reference's rayon path: parallel field/iterate/build workers, "
                                   "serial propagate, LJ and electron loop as in the reference"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n):
    return {"workload": "configs[3]: uniform electrolyte (Li+/PF6-/EC/DMC 342:342:2393:2394), electron polarization "
                        "field sampling, full hot-path step", "n_bodies": int(n), "theta": args.theta, "epsilon": 2.0,
            "leaf_capacity": 1, "density_per_A2": 0.0625, "seed": "0xC0FFEE", "parity_mode": int(args.ieee),
            "l2": "working set >> 126 MB L2, no flush needed", "parallelism": f"morton-sharded x{args.gpus}",
            "multi_gpu_build": "n/a" if args.gpus == 1 else ("replicated" if args.replicated_build else
                                                            "sharded by key range (65536 top-level cells), tree pieces all-gathered")}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from helpers import KE
    from particlesim_b200 import Bodies, Simulation, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.n
    bd = make_workload(n)
    hw, hh = bd["hw"], bd["hh"]
    stream = torch.cuda.current_stream().cuda_stream
    b = Bodies(bd["pos"], vel=bd["vel"], mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
               species=bd["species"], ebody=bd["ebody"], erel=bd["erel"])
    if world > 1:
        from particlesim_b200.parallel import ShardedSimulation
        sim = ShardedSimulation(b, hw, hh, theta=args.theta, parity_mode=int(args.ieee), device=local_rank,
                                stream=stream, rank=rank, world=world, local_build=not args.replicated_build)
    else:
        sim = Simulation(b, hw, hh, theta=args.theta, parity_mode=int(args.ieee), device=local_rank, stream=stream)
    sim.config.coulomb_constant = float(KE)
    params = sim.step_params()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sim.step_device(params)
    barrier()
    sim._call("psim_reset_counters")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    phase_ms = np.zeros((args.steps, 8), np.float64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        sim.step_device(params)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    launches = sim.stats()["kernel_launches"]

    # per-phase device times (separate pass so the event reads do not sit inside the timed region)
    ph = np.zeros(8, np.float32)
    acc = np.zeros(8, np.float64)
    reps = min(args.steps, 5)
    for _ in range(reps):
        if world > 1:
            sim.step_device(params, record=True)
            acc += np.array(sim.phase_ms())
        else:
            sim.step_device(params)
            sim._call("psim_phase_times", ph.ctypes.data)
            acc += ph
    phase = {nm: float(v / reps) for nm, v in zip(PHASES, acc)}

    line = None
    if rank == 0:
        value = n / (ms_step * 1e-3) / 1e6
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
                "gpu_launches": int(launches), "clocks": clocks, "phase_ms": phase}

    if world == 1:
        # ---- roofline of the dominant kernel (bh_group_bodies_kernel = phase quadtree_field) --------
        sim._call("psim_build", 0, 0.0, 0.0)
        cnt = np.zeros(4, np.uint64)
        sim._call("psim_field_counters", cnt.ctypes.data)
        opened, accepted, pairs, wsteps = (int(x) for x in cnt)
        visits = n + 4 * opened  # the reference evaluates the opening test on all 4 children of every opened node
        flops = 12.0 * visits + 14.0 * (accepted + pairs)
        tf = C.c_float()
        sms = C.c_int32()
        sim._call("psim_fp32_peak", C.byref(tf), C.byref(sms))
        t_field = phase["quadtree_field"] * 1e-3
        achieved = flops / t_field / 1e12
        st = sim.stats()
        line["roofline"] = {
            "kernel": "bh_group_bodies_kernel", "bound": "fp32",
            "achieved": achieved, "peak": float(tf.value), "unit": "TFLOP/s", "frac": achieved / float(tf.value),
            "traffic": ncu_traffic(n),
            "peak_source": "measured live: FP32 FMA microbenchmark on this GPU (MEASURED_PEAKS.json holds HBM and bf16 "
                           "tensor peaks only; the traversal is FP32-pipe bound and uses no tensor cores, SURVEY.md 8d)",
            "algorithmic": {"node_visits_per_body": visits / n, "monopoles_per_body": accepted / n,
                            "direct_terms_per_body": pairs / n, "flops_per_body": flops / n,
                            "rule": "12 flops per opening test + 14 per monopole or direct term (SURVEY.md 8d)",
                            "warp_steps_per_32_bodies": wsteps / ((n + 31) // 32)},
            "avg_launch_ms": phase["quadtree_field"],
            "algorithmic_bytes": int(st["compact_nodes"] * 0 + n * (16 + 8 + 16)),
        }
        # HBM roofline of the build pipeline (keys, sort, gather, nodes, aggregation), for the explanation
        peaks = measured_peaks()
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        M = st["compact_nodes"]
        # keys 16 r + 16 w; radix passes on the upper key word 4 + 4 x 16; sorted keys + run fix-up 28; body
        # gather 2 x 61; levels 2 + node scan 4 + emit 16 r; per node: records 32 + 4 + 4, sums 64, compaction 48
        build_bytes = n * (16 + 16 + 4 + 4 * 16 + 28 + 2 * 61 + 2 + 4 + 16) + M * (32 + 4 + 4 + 64 + 48)
        line["roofline_build"] = {"bound": "hbm", "achieved": build_bytes / (phase["quadtree_build"] * 1e-3) / 1e9,
                                  "peak": hbm, "unit": "GB/s",
                                  "frac": build_bytes / (phase["quadtree_build"] * 1e-3) / 1e9 / hbm,
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                                  "bytes_per_body": build_bytes / n}
        line["tree"] = {"compact_nodes": int(M), "reference_nodes": int(st["reference_nodes"]), "max_depth": int(st["max_depth"])}

        # ---- the same step with the other traversal arithmetic, for the explanation ------------------
        other = 0 if args.ieee else 1
        sim._cfg.parity_mode = other
        sim._call("psim_set_config", C.byref(sim._cfg))
        for _ in range(2):
            sim.step_device(params)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(min(args.steps, 5)):
            sim.step_device(params)
        e1.record()
        torch.cuda.synchronize()
        ms_other = e0.elapsed_time(e1) / min(args.steps, 5)
        line["other_arithmetic"] = {"parity_mode": other, "ms_per_step": ms_other, "value": n / ms_other / 1e3, "unit": UNIT,
                                    "note": "parity_mode 1 = IEEE div/sqrt, no FMA contraction (the reference's per-term "
                                            "arithmetic); parity_mode 0 = MUFU rsqrt/rcp + FMA on the same interaction sets"}
        sim._cfg.parity_mode = int(args.ieee)
        sim._call("psim_set_config", C.byref(sim._cfg))

        # ---- e2e: host buffers in, host buffers out, every step -------------------------------------
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_pos, h_vel, h_q = pin(bd["pos"]), pin(bd["vel"]), pin(bd["charge"])
        o_pos, o_vel = torch.empty_like(h_pos).pin_memory(), torch.empty_like(h_vel).pin_memory()
        o_ef = torch.empty_like(h_pos).pin_memory()
        o_orig = torch.empty(n, dtype=torch.int32).pin_memory()
        assert all(t.is_pinned() for t in (h_pos, h_vel, h_q, o_pos, o_vel, o_ef, o_orig))
        ksteps = max(3, min(args.steps, 5))
        # The host's copy of the state is rewritten between steps; whatever was written last still sits
        # dirty in the CPU's last-level cache (60 MB here) and DMA reads of it run at a quarter of the
        # PCIe rate.  The reference's host step touches far more memory than that between two calls, so
        # the cache is evicted with a scratch write before each timed call (both variants).
        scratch = torch.empty(256 << 20, dtype=torch.uint8)

        def run_e2e(pipelined):
            ts = []
            for k in range(ksteps + 1):
                scratch.add_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if pipelined:
                    sim._call("psim_step_host", C.byref(params), n, h_pos.data_ptr(), h_vel.data_ptr(), h_q.data_ptr(),
                              o_pos.data_ptr(), o_vel.data_ptr(), o_ef.data_ptr(), o_orig.data_ptr())
                else:
                    sim._call("psim_update_state", n, h_pos.data_ptr(), h_vel.data_ptr(), h_q.data_ptr())
                    sim.step_device(params)
                    sim._call("psim_download_bodies", o_pos.data_ptr(), None, o_vel.data_ptr(), None, None, None, None,
                              None, None, None, o_ef.data_ptr(), o_orig.data_ptr())
                torch.cuda.synchronize()
                if k > 0:
                    ts.append(time.perf_counter() - t0)
                # next step's input = this step's output (host owns the state)
                h_pos.copy_(o_pos)
                h_vel.copy_(o_vel)
                h_q.copy_(torch.from_numpy(bd["charge"])[o_orig.long()])
            return float(np.mean(ts))

        t_seq = run_e2e(False)
        t_e2e = run_e2e(True)
        line["e2e"] = {"value": n / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(n * 20),
                       "d2h_bytes_per_step": int(n * 28), "ms_per_step": t_e2e * 1e3,
                       "api": "psim_step_host: pos, vel, charge from pinned host -> hot-path step -> pos, vel, e_field, "
                              "orig_index to pinned host; copies pipelined against the device work on two copy streams",
                       "host_cache": "256 MB scratch write before each timed call (evicts the 60 MB CPU LLC)",
                       "unpipelined_ms_per_step": t_seq * 1e3,
                       "unpipelined_api": "psim_update_state + psim_step + psim_download_bodies"}
        sim.close()
        # ---- CPU baseline on a bounded sample --------------------------------------------------------
        if not args.no_cpu:
            from oracle import pyoracle
            try:
                pyoracle.build(native=True)
                variant = "native"
            except Exception:
                variant = ""
            threads = pyoracle.load(variant).orc_max_threads()
            n_s = min(n, args.cpu_n)
            w_s = make_workload(n_s)
            t_cpu, _ = cpu_reference_step(w_s, args.theta, threads, 1, 0, variant)
            t_par, _ = cpu_reference_step(w_s, args.theta, threads, 1, 0, variant, all_parallel=True)
            line["cpu_baseline"] = {"value": n_s / t_cpu / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                    "all_parallel_value": n_s / t_par / 1e6,
                                    "all_parallel_note": "same step with the electron loop on every thread too (the "
                                                         "reference runs it serially, simulation.rs:1186-1196)",
                                    "sample": f"one step of a {n_s}-body instance of the same generator "
                                              f"({t_cpu:.1f} s of CPU work); C++ restatement of the reference's rayon path "
                                              "(serial propagate / LJ / electron loop as in the reference)"}
    else:
        # ---- e2e at N GPUs: the host owns the state; every step each rank moves ITS slice over PCIe (pinned
        # host -> device: pos, vel, charge; device -> pinned host: pos, vel, e_field), the slices are
        # all-gathered over NVLink into the replicated device state, and the sharded step runs in between
        from particlesim_b200.parallel import all_gather_slices, shard_range
        f, c = shard_range(n, world, rank)
        wb = sim.wb
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        st0 = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
        sim._call("psim_download_bodies", st0[0].ctypes.data, None, st0[1].ctypes.data, None, None, None, None, None,
                  st0[2].ctypes.data, None, None, None)   # the device's current order
        h_pos, h_vel, h_q = pin(st0[0][f:f + c]), pin(st0[1][f:f + c]), pin(st0[2][f:f + c])
        o_pos, o_vel, o_ef = (torch.empty_like(h_pos).pin_memory() for _ in range(3))
        ksteps = max(3, min(args.steps, 5))
        ts = []
        for k in range(ksteps + 1):
            barrier()
            t0 = time.perf_counter()
            v = sim._views()
            v["pqr"][f:f + c, 0:2].copy_(h_pos, non_blocking=True)
            v["pqr"][f:f + c, 2].copy_(h_q, non_blocking=True)
            v["velz"][f:f + c, 0:2].copy_(h_vel, non_blocking=True)
            all_gather_slices(v["pqr"], wb, rank, world, dist, sim._scratch_b)
            all_gather_slices(v["velz"], wb, rank, world, dist, sim._scratch_b)
            sim._call("psim_mark_positions_changed")
            sim.step_device(params)
            v = sim._views()
            p8 = np.zeros(8, np.uint64)
            sim._call("psim_device_ptrs", p8.ctypes.data)
            from particlesim_b200.parallel import _DevArray
            ef = torch.as_tensor(_DevArray(p8[3], (world * wb, 2)), device="cuda")
            o_pos.copy_(v["pqr"][f:f + c, 0:2], non_blocking=True)
            o_vel.copy_(v["velz"][f:f + c, 0:2], non_blocking=True)
            o_ef.copy_(ef[f:f + c], non_blocking=True)
            barrier()
            if k > 0:
                ts.append(time.perf_counter() - t0)
            h_pos.copy_(o_pos)
            h_vel.copy_(o_vel)
        t = torch.tensor([float(np.mean(ts))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            t_e2e = float(t.item())
            line["e2e"] = {"value": n / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(n * 20),
                           "d2h_bytes_per_step": int(n * 24), "ms_per_step": t_e2e * 1e3,
                           "api": "per rank: its slice of pos / vel / charge from pinned host, all-gather over NVLink, "
                                  "ShardedSimulation.step_device, its slice of pos / vel / e_field back to pinned host; "
                                  "bytes are the sum over ranks"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", dest="n", type=int, default=16_000_000)
    ap.add_argument("--theta", type=float, default=1.0)
    ap.add_argument("--ieee", type=int, default=0,
                    help="1: traversal with IEEE div/sqrt and no FMA contraction (parity_mode 1) instead of the default "
                         "MUFU rsqrt/rcp arithmetic (parity_mode 0); both sum the reference's interaction sets and "
                         "both are tested against the oracle at the 1e-5 tolerance")
    ap.add_argument("--cpu-n", type=int, default=8_000_000,
                    help="bodies in the CPU-baseline sample (one step of 8 M bodies is 10-20 s of work on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--replicated-build", action="store_true",
                    help="multi-GPU: every rank builds the whole tree (the v1 scheme) instead of its key range")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

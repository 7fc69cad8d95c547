#!/usr/bin/env python
"""bench.py — the hot path of Simulation::step on synthetic charged-particle sets.

A "step" = one pass of the hot path over the whole body set, in Simulation::step's order
(reference simulation.rs:1000-1196): reset acc, quadtree build (tight AABB), cell list, Coulomb field +
attract, LJ / repulsion / stack pressure, integrator, domain-bounded quadtree build, electron field
sampling + drift.  Metric: Mparticles/s = bodies / step time (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config {2,3,4,5}] [--bodies BODIES] [--theta T]
  python bench.py --impl reference ...    # the C++ restatement of the reference's rayon path on the host cores

Default workload (--config 4): BASELINE.json configs[3], "N=16M uniform electrolyte with electron polarization
field sampling" (the configuration the 100x target is quoted on); theta is the reference default 1.0.
--config 2 / 3 / 5 run BASELINE.json configs[1] / [2] / [4] (1 M uniform +-1 charges at theta 0.5, Coulomb only;
4 M clustered with LJ; 64 M mixed-species slab with the out-of-plane integrator) through the same code.
Node centres are the reference's serial f32 sums (psim_config.strict_centres = 1), so the timed mode is the one
whose fields meet the 1e-5 bar against the strict oracle; the line carries that parity figure ("parity").
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


# BASELINE.json configs[1..4] (SURVEY.md 8d "Config 2..5"); configs[0] is the reference's CPU-only scenario
CONFIGS = {
    2: dict(gen="uniform_pm1", n=1_000_000, theta=0.5, short=False, electrons=False, iterate=False,
            name="configs[1]: uniform random +-1 charges, Coulomb-only Barnes-Hut step (build + field)"),
    3: dict(gen="clustered", n=4_000_000, theta=1.0, short=True, electrons=False, iterate=False,
            name="configs[2]: clustered / dendrite-like LithiumMetal + electrolyte (deep unbalanced tree): build + field "
                 "+ cell list + polar + LJ"),
    4: dict(gen="electrolyte", n=16_000_000, theta=1.0, short=True, electrons=True, iterate=True,
            name="configs[3]: uniform electrolyte (Li+/PF6-/EC/DMC 342:342:2393:2394), electron polarization "
                 "field sampling, full hot-path step"),
    5: dict(gen="slab", n=64_000_000, theta=1.0, short=True, electrons=True, iterate=True,
            name="configs[4]: mixed-species slab (20 % LLZO/LLZT/S40B scaffold, 5 % LithiumMetal, 75 % electrolyte), "
                 "2.5-D out-of-plane integrator, full hot-path step"),
}


def make_workload(n, seed=0xC0FFEE, gen="electrolyte"):
    import helpers
    return getattr(helpers, gen)(n, seed=seed)


def host_threads():
    """every core this process may run on (torchrun exports OMP_NUM_THREADS=1; the reference's rayon pool would
    still use all of them)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(bodies, theta, threads, steps=1, warmup=0, variant="native", all_parallel=False, cfg=None):
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
    cfg = cfg or CONFIGS[4]
    o = oracle_for(bodies, theta=theta, variant=variant)
    hw, hh = bodies["hw"], bodies["hh"]
    oop = bool(bodies.get("enable_out_of_plane", False))
    T = threads
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        o.reset_acc()
        if cfg["short"]:
            o.prepare_spatial_structures(hw, hh, threads=T)
        else:
            o.build(threads=T)
        o.attract(KE, threads=T)
        if cfg["short"]:
            o.apply_polar_forces(KE, True, 1)   # serial in the reference (forces.rs:52-175)
            o.apply_lj_forces(True)
            o.apply_repulsive_forces(True)
        if cfg["iterate"]:
            o.iterate(5.0, 1.0, hw, hh, float(bodies.get("hd", 1.0)), oop, threads=T)
        if cfg["electrons"]:
            o.build_with_domain(hw, hh, threads=T)
            o.update_electrons((0.0, 0.0), 5.0, KE, threads=T if all_parallel else 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times)), T


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be
    built in this image (no cargo, un-vendored quarkstrom), so this is the C++ restatement (oracle/),
    kind "port", with every host thread the reference's rayon pool would use, on the SAME body set as the
    GPU arm (same generator, same n, same seed).  A 16 M-body step is ~30 s of CPU work, so the number of
    timed steps is capped by a wall-clock budget (reported as "steps"); the first step is never a warm-up
    casualty: if the budget allows a warm-up step it is taken, otherwise the timed steps include first touch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    try:
        pyoracle.build(native=True)
        variant = "native"
    except Exception:
        variant = ""
    cfg = CONFIGS[args.config]
    threads = host_threads()
    n_probe = min(args.n, 200_000)
    t_probe, _ = cpu_reference_step(make_workload(n_probe, gen=cfg["gen"]), args.theta, threads, 1, 0, variant, cfg=cfg)
    est = t_probe / n_probe * args.n * 1.25  # N log N growth + cache effects
    budget = float(args.ref_budget_s)
    steps = int(max(1, min(args.steps, budget // max(est, 1e-3))))
    warmup = int(min(args.warmup, 1)) if est * (steps + 1) <= budget * 1.25 else 0
    bodies = make_workload(args.n, gen=cfg["gen"])
    t_step, _ = cpu_reference_step(bodies, args.theta, threads, steps, warmup, variant, cfg=cfg)
    value = args.n / t_step / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "requested": {"steps": args.steps, "warmup": args.warmup},
        "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"the full {args.n}-body set of the GPU arm (same generator and seed), {steps} timed "
                                   f"step(s) of ~{t_step:.0f} s inside a {budget:.0f} s budget; C++ restatement of the "
                                   "reference's rayon path: parallel field / iterate / build workers on all host cores, "
                                   "serial propagate, polar, LJ and electron loop as in the reference"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n):
    cfg = CONFIGS[args.config]
    return {"workload": cfg["name"], "n_bodies": int(n), "theta": args.theta, "epsilon": 2.0,
            "leaf_capacity": 1, "density_per_A2": 0.0625, "seed": "0xC0FFEE", "parity_mode": int(args.ieee),
            "strict_centres": int(args.strict), "do_polar": int(cfg["short"]),
            "l2": "device working set (bodies 2 x 61 B + 4 n node slots x ~185 B) >> 126 MB L2, no flush needed",
            "parallelism": f"morton-sharded x{args.gpus}",
            "multi_gpu_build": "n/a" if args.gpus == 1 else ("replicated" if args.replicated_build else
                                                            "psim_step_sharded: builds sharded by key range (65536 top-level "
                                                            "cells), locally-essential-tree exchange of the traversal records, "
                                                            "body state replicated")}


# ------------------------------------------------------------------------------------------------
def parity_block(sim, bd, cfg, args, threads):
    """Field of the timed configuration against the oracle on the SAME bodies (north star: rel-L2 <= 1e-5 at the
    same theta, bit-exact permutation): the strict oracle (the reference restatement with its serial f32 node-centre
    sums), the f64-centre variant, and an FP64 direct sum on a sample of targets."""
    from helpers import KE, oracle_for, rel_l2
    n = len(bd["pos"])
    sim._call("psim_build", 0, 0.0, 0.0)
    sim._call("psim_field", float(KE), 0.0, 0.0, 0, None, None)
    e_dev = np.zeros((n, 2), np.float32)
    orig = np.zeros(n, np.uint32)
    pos = np.zeros((n, 2), np.float32)
    radius = np.zeros(n, np.float32)
    sim._call("psim_download_bodies", pos.ctypes.data, None, None, None, None, None, None, radius.ctypes.data, None,
              None, e_dev.ctypes.data, orig.ctypes.data)
    out = {"mode": f"strict_centres={int(args.strict)}, parity_mode={int(args.ieee)}", "n": int(n), "theta": args.theta,
           "quantity": "e_field of Quadtree::field after Quadtree::build, all bodies"}
    for key, variant in (("rel_l2_vs_strict", ""), ("rel_l2_vs_f64centre", "hp")):
        t0 = time.perf_counter()
        o = oracle_for(bd, theta=args.theta, variant=variant)
        o.build(threads=threads)
        e_ref, _ = o.field(KE, threads=threads)
        same = bool(np.array_equal(o.permutation(), orig.astype(np.int64)))
        out[key] = rel_l2(e_dev, e_ref) if same else None
        if variant == "":
            out["permutation_equal"] = same
            rng = np.random.default_rng(7)
            pick = rng.choice(n, min(n, 256), replace=False)
            direct = o.direct_f64(pos[pick], target_radius=radius[pick], k_e=float(KE), epsilon=2.0, threads=threads)
            out["rel_l2_vs_direct_f64"] = rel_l2(e_dev[pick], direct)
            out["oracle_rel_l2_vs_direct_f64"] = rel_l2(e_ref[pick], direct)
            out["direct_sample"] = int(len(pick))
        out.setdefault("oracle_seconds", {})[variant or "strict"] = round(time.perf_counter() - t0, 1)
        del o
    out["note"] = ("rel_l2_vs_strict is the bar (<= 1e-5); rel_l2_vs_direct_f64 is the Barnes-Hut truncation error at this "
                   "theta and must equal the oracle's")
    return out


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
    cfg = CONFIGS[args.config]
    n = args.n
    bd = make_workload(n, gen=cfg["gen"])
    hw, hh = bd["hw"], bd["hh"]
    stream = torch.cuda.current_stream().cuda_stream
    b = Bodies(bd["pos"], z=bd.get("z"), vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
               species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
    if world > 1:
        from particlesim_b200.parallel import ShardedSimulation
        sim = ShardedSimulation(b, hw, hh, domain_depth=float(bd.get("hd", 1.0)), theta=args.theta,
                                parity_mode=int(args.ieee), device=local_rank, stream=stream, rank=rank, world=world,
                                local_build=not args.replicated_build, strict_centres=bool(args.strict))
    else:
        sim = Simulation(b, hw, hh, domain_depth=float(bd.get("hd", 1.0)), theta=args.theta, parity_mode=int(args.ieee),
                         device=local_rank, stream=stream, strict_centres=bool(args.strict))
    sim.config.coulomb_constant = float(KE)
    sim.config.enable_out_of_plane = bool(bd.get("enable_out_of_plane", False))
    params = sim.step_params(do_short_range=cfg["short"], do_electrons=cfg["electrons"], do_iterate=cfg["iterate"],
                             do_polar=cfg["short"])

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
        try:  # which way the build made the node charges (psim_build_info): exact integer prefix or level sweeps
            line["build_path"] = sim.build_info()
        except Exception as ex:  # diagnostics only
            line["build_path"] = {"error": str(ex)}
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
        peaks = measured_peaks()
        sm_max = float((clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        line["roofline"] = {
            "kernel": "bh_group_bodies_kernel", "bound": "fp32",
            "achieved": achieved, "peak": float(tf.value), "unit": "TFLOP/s", "frac": achieved / float(tf.value),
            "traffic": ncu_traffic(n),
            "peak_source": "measured live: FP32 FMA microbenchmark on this GPU (MEASURED_PEAKS.json holds HBM and bf16 "
                           "tensor peaks only; the traversal is FP32-pipe bound and uses no tensor cores, SURVEY.md 8d)",
            "peak_formula": {"value": 2 * 128 * int(sms.value) * sm_max * 1e6 / 1e12,
                             "rule": f"2 flops x 128 FP32 lanes x {int(sms.value)} SMs x {sm_max:.0f} MHz"},
            "algorithmic": {"node_visits_per_body": visits / n, "monopoles_per_body": accepted / n,
                            "direct_terms_per_body": pairs / n, "flops_per_body": flops / n,
                            "rule": "12 flops per opening test + 14 per monopole or direct term (SURVEY.md 8d)",
                            "warp_steps_per_32_bodies": wsteps / ((n + 31) // 32)},
            "avg_launch_ms": phase["quadtree_field"],
            "algorithmic_bytes": int(n * (16 + 8 + 16)),
        }
        # HBM roofline of the build pipeline (keys, sort, gather, nodes, aggregation)
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        M = st["compact_nodes"]
        d_sig = int(st["max_depth"])
        passes = -(-2 * d_sig // 8)
        # keygen 16, sort 8 + passes x 2 x 12 with passes = ceil(2 D_sig / 8), gather 32, nodes M x (32 + 16)
        survey_bytes = n * (16 + 8 + 24 * passes + 32) + M * 48
        # what this pipeline moves: keys 16 r + 16 w; radix passes on the upper key word 4 + 4 x 16; sorted keys + run
        # fix-up 28; body gather 2 x 61; levels 2 + node scan 4 + emit 16 r; per node: records 32 + 4 + 4, sums 64, compaction 48
        model_bytes = n * (16 + 16 + 4 + 4 * 16 + 28 + 2 * 61 + 2 + 4 + 16) + M * (32 + 4 + 4 + 64 + 48)
        t_build = phase["quadtree_build"] * 1e-3
        line["roofline_build"] = {"bound": "hbm", "achieved": survey_bytes / t_build / 1e9, "peak": hbm, "unit": "GB/s",
                                  "frac": survey_bytes / t_build / 1e9 / hbm,
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                                  "bytes_per_body": survey_bytes / n,
                                  "rule": f"SURVEY.md 8d algorithmic bytes with the measured M = {M} nodes and D_sig = {d_sig} "
                                          f"({passes} 8-bit passes over 12-byte pairs)",
                                  "own_traffic_model": {"bytes_per_body": model_bytes / n,
                                                        "frac": model_bytes / t_build / 1e9 / hbm}}
        if cfg["short"] and phase["forces_lj"] > 0.02:
            # short_range_kernel (+ the polar pass): SURVEY.md 8d puts cell build + LJ at ~350 B per body
            t_sr = (phase["forces_lj"] + phase["cell_list_rebuild"]) * 1e-3
            line["roofline_short_range"] = {"bound": "hbm", "kernels": "cell_id / onesweep / cell_ranges / polar / short_range",
                                            "achieved": 350.0 * n / t_sr / 1e9, "peak": hbm, "unit": "GB/s",
                                            "frac": 350.0 * n / t_sr / 1e9 / hbm, "bytes_per_body": 350.0,
                                            "ms": t_sr * 1e3}
        line["tree"] = {"compact_nodes": int(M), "reference_nodes": int(st["reference_nodes"]), "max_depth": int(st["max_depth"])}

        # ---- the same step in the other modes, for the explanation -----------------------------------
        def timed_steps(k=5):
            for _ in range(2):
                sim.step_device(params)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(k):
                sim.step_device(params)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k

        other = 0 if args.ieee else 1
        sim._cfg.parity_mode = other
        sim._call("psim_set_config", C.byref(sim._cfg))
        ms_other = timed_steps(min(args.steps, 5))
        line["other_arithmetic"] = {"parity_mode": other, "ms_per_step": ms_other, "value": n / ms_other / 1e3, "unit": UNIT,
                                    "note": "parity_mode 1 = IEEE div/sqrt, no FMA contraction (the reference's per-term "
                                            "arithmetic); parity_mode 0 = MUFU rsqrt/rcp + FMA on the same interaction sets"}
        sim._cfg.parity_mode = int(args.ieee)
        sim._cfg.strict_centres = 0 if args.strict else 1
        sim._call("psim_set_config", C.byref(sim._cfg))
        ms_centres = timed_steps(min(args.steps, 5))
        line["other_centres"] = {"strict_centres": int(sim._cfg.strict_centres), "ms_per_step": ms_centres,
                                 "value": n / ms_centres / 1e3, "unit": UNIT,
                                 "note": "strict_centres 1 = node centres by the reference's serial f32 sums (bit-identical "
                                         "nodes, fields within 1e-5 of the reference); 0 = f64 sums carried up the tree "
                                         "(faster, ~1e-4 from the reference: its own summation noise)"}
        sim._cfg.strict_centres = int(args.strict)
        sim._call("psim_set_config", C.byref(sim._cfg))

        # ---- e2e: host buffers in, host buffers out, every step -------------------------------------
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_pos, h_vel, h_q = pin(bd["pos"]), pin(b.vel), pin(bd["charge"])
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
        # ---- parity of the timed mode against the oracle, same bodies ---------------------------------
        if not args.no_parity:
            sim.upload()  # back to the generator's state (the e2e steps moved the bodies)
            line["parity"] = parity_block(sim, bd, cfg, args, host_threads())
        sim.close()
        # ---- CPU baseline on a bounded sample --------------------------------------------------------
        if not args.no_cpu:
            from oracle import pyoracle
            try:
                pyoracle.build(native=True)
                variant = "native"
            except Exception:
                variant = ""
            threads = host_threads()
            n_s = min(n, args.cpu_n)
            w_s = bd if n_s == n else make_workload(n_s, gen=cfg["gen"])
            t_cpu, _ = cpu_reference_step(w_s, args.theta, threads, 1, 0, variant, cfg=cfg)
            t_par, _ = cpu_reference_step(w_s, args.theta, threads, 1, 0, variant, all_parallel=True, cfg=cfg)
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
        cs = np.zeros(4, np.uint64)
        if getattr(sim, "library", False):
            sim._call("psim_comm_stats", cs.ctypes.data)
        if rank == 0:
            line["let"] = {"enabled": bool(cs[2]), "records_sent_by_rank0": int(cs[0]), "full_allgather_would_send": int(cs[1]),
                           "fraction": float(cs[0]) / max(float(cs[1]), 1.0)}
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
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS),
                    help="SURVEY.md 8d config number: 2 = BASELINE configs[1] ... 5 = configs[4]; default 4 (16 M electrolyte)")
    ap.add_argument("--bodies", dest="n", type=int, default=None)
    ap.add_argument("--theta", type=float, default=None)
    ap.add_argument("--strict", type=int, default=1,
                    help="1 (default): node centres by the reference's serial f32 sums (psim_config.strict_centres); "
                         "0: f64 sums carried up the tree")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed mode")
    ap.add_argument("--ref-budget-s", type=float, default=240.0,
                    help="--impl reference: wall-clock budget for the timed CPU steps on the full body set")
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
    if args.n is None:
        args.n = CONFIGS[args.config]["n"]
    if args.theta is None:
        args.theta = CONFIGS[args.config]["theta"]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

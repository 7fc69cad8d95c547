"""torchrun --nproc-per-node G tools/check_sharded.py [n] [library|python|repl] [electrolyte|clustered|uniform_pm1] [theta]: three sharded hot-path steps (both
builds, field, polar, LJ, integrator, electrons) over real NCCL must leave exactly the single-GPU state (default
configuration: the reference's serial-sum node centres on both sides).  Exits non-zero on a mismatch; tests/test_gpu_multi.py runs it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from helpers import KE  # noqa: E402
from particlesim_b200 import Bodies, Simulation  # noqa: E402
from particlesim_b200.parallel import ShardedSimulation  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_001
gen = sys.argv[3] if len(sys.argv) > 3 else "electrolyte"
theta = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
bd = getattr(helpers, gen)(n)
if os.environ.get("NO_LJ") != "1" and gen == "electrolyte":
    bd["species"][: n // 10] = 1  # some LJ bodies


def mk(cls, **kw):
    b = Bodies(bd["pos"], vel=bd.get("vel"), mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
               species=bd["species"], ebody=bd.get("ebody"), erel=bd.get("erel"))
    s = cls(b, bd["hw"], bd["hh"], device=lr, stream=torch.cuda.current_stream().cuda_stream, theta=theta, **kw)
    s.config.coulomb_constant = float(KE)
    return s


def state(sim):
    nb, m = len(sim.bodies), len(sim.bodies.ebody)
    pos, vel = np.zeros((nb, 2), np.float32), np.zeros((nb, 2), np.float32)
    orig = np.zeros(nb, np.uint32)
    sim._call("psim_download_bodies", pos.ctypes.data, None, vel.ctypes.data, None, None, None, None, None, None,
              None, None, orig.ctypes.data)
    eb, er, ev = np.zeros(m, np.uint32), np.zeros((m, 2), np.float32), np.zeros((m, 2), np.float32)
    sim._call("psim_download_electrons", eb.ctypes.data, er.ctypes.data, ev.ctypes.data)
    return pos, vel, orig, eb, er, ev


how = sys.argv[2] if len(sys.argv) > 2 else "library"
sh = mk(ShardedSimulation, rank=rank, world=world, local_build=how != "repl",
        orchestration="python" if how == "python" else "library")
for _ in range(3):
    sh.step_device()
torch.cuda.synchronize()
a = state(sh)
if how == "library":
    cs = np.zeros(4, np.uint64)
    sh._call("psim_comm_stats", cs.ctypes.data)
    print(f"rank {rank}: LET exchange sent {int(cs[0])} of {int(cs[1])} records ({100.0 * cs[0] / max(int(cs[1]), 1):.1f} %), "
          f"enabled {int(cs[2])}, tree is LET {int(cs[3])}", flush=True)
ok = True
if rank == 0:
    one = mk(Simulation)
    for _ in range(3):
        one.step_device()
    b = state(one)
    names = ["pos", "vel", "orig", "ebody", "erel", "evel"]
    ok = True
    for nm, x, y in zip(names, a, b):
        same = np.array_equal(x, y)
        ok &= same
        extra = ""
        if not same:
            d = np.abs(x.astype(np.float64) - y.astype(np.float64))
            rows = np.unique(np.argwhere(d > 0)[:, 0])
            extra = f"  differing rows {len(rows)} of {len(x)} (first {rows[:4].tolist()}, last {rows[-1]}), max |diff| {d.max():.3e}"
        print(f"{nm}: identical={same}{extra}")
    print(f"SHARDED ({how}, {world} ranks, n = {n}, {gen}, theta {theta}) == SINGLE:", ok, flush=True)
flag = torch.tensor([1 if (rank != 0 or ok) else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)

"""torchrun --nproc-per-node G tools/shard_phases.py [n]: device time of every phase and exchange of the
sharded build (max over ranks), 16 M-body electrolyte by default."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import KE, electrolyte  # noqa: E402
from particlesim_b200 import Bodies, _lib  # noqa: E402
from particlesim_b200.parallel import ShardedSimulation, sharded_build  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
bd = electrolyte(n)
b = Bodies(bd["pos"], vel=bd["vel"], mass=bd["mass"], radius=bd["radius"], charge=bd["charge"],
           species=bd["species"], ebody=bd["ebody"], erel=bd["erel"])
sim = ShardedSimulation(b, bd["hw"], bd["hh"], device=lr, stream=torch.cuda.current_stream().cuda_stream,
                        rank=rank, world=world, parity_mode=0, orchestration="python")
sim.config.coulomb_constant = float(KE)
for _ in range(3):
    sim.step_device()
names, acc = [], None
reps = 5
for _ in range(reps):
    ev = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append((name, e))

    sharded_build([sim], _lib.BUILD_CONTAINING, 0.0, 0.0, sim._comm, torch, mark)
    torch.cuda.synchronize()
    ms = np.array([ev[k][1].elapsed_time(ev[k + 1][1]) for k in range(len(ev) - 1)])
    names = [ev[k + 1][0] for k in range(len(ev) - 1)]
    acc = ms if acc is None else acc + ms
t = torch.tensor(acc / reps, device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    for nm, v in zip(names, t.tolist()):
        print(f"{nm:36s} {v:8.3f} ms")
    print(f"{'total':36s} {sum(t.tolist()):8.3f} ms   (n = {n}, world = {world})")
dist.barrier()
dist.destroy_process_group()

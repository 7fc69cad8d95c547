"""CPU estimate of the locally essential tree (LET) a rank would need under distributed body ownership
(DESIGN.md 8, item 1): for `world` ranks owning equal contiguous slices of the Morton order, the number of
CHARGED tree nodes that some target of the rank can visit, found with the same conservative box test the
device walk uses (a node may be opened by a target in the box iff size >= theta * (dmin - rmax)), with the
rank's region described by the bounding boxes of its runs of `cells`-th of the slice.  Uses the CPU
emulation's tree (tests/emu), no GPU.   python tools/let_estimate.py [n] [world] [theta]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Emu, electrolyte  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
sub = 64  # region = union of the boxes of 64 consecutive sub-slices (a Morton range is not a box)

bodies = electrolyte(n)
emu = Emu()
emu.build(bodies, 0)
nodes = emu.nodes()               # reference-shaped: 4 contiguous children, 0 = leaf
sb = emu.sorted_bodies()          # x, y, q, r in tree order
children = nodes["children"].astype(np.int64)
pos, size, charge = nodes["pos"].astype(np.float64), nodes["quad_size"].astype(np.float64), nodes["charge"]
b0, b1 = nodes["bodies_start"].astype(np.int64), nodes["bodies_end"].astype(np.int64)
absq = np.concatenate([[0.0], np.cumsum(np.abs(sb[:, 2]))])
charged = (absq[b1] - absq[b0]) > 0          # some charged body below the node
total_charged = int(charged.sum())
print(f"n = {n}, world = {world}, theta = {theta}: {len(nodes)} reference nodes, {total_charged} charged (traversal) nodes")

per = (n + world - 1) // world
for r in range(world):
    lo, hi = r * per, min(n, (r + 1) * per)
    edges = np.linspace(lo, hi, sub + 1).astype(np.int64)
    boxes = []
    for a, b in zip(edges[:-1], edges[1:]):
        if b > a:
            p = sb[a:b]
            boxes.append((p[:, 0].min(), p[:, 0].max(), p[:, 1].min(), p[:, 1].max(), p[:, 3].max()))
    boxes = np.array(boxes, np.float64)
    need = np.zeros(len(nodes), bool)
    frontier = np.array([0], np.int64)
    while len(frontier):
        need[frontier] = True
        c = pos[frontier]
        dx = np.maximum(np.maximum(boxes[None, :, 0] - c[:, None, 0], c[:, None, 0] - boxes[None, :, 1]), 0.0)
        dy = np.maximum(np.maximum(boxes[None, :, 2] - c[:, None, 1], c[:, None, 1] - boxes[None, :, 3]), 0.0)
        dmin = np.sqrt(dx * dx + dy * dy) - boxes[None, :, 4]          # distance after the target radius
        opened = (size[frontier][:, None] >= theta * np.maximum(dmin, 0.0) * 0.9999).any(axis=1)
        inner = frontier[opened & (children[frontier] != 0) & charged[frontier]]
        kids = (children[inner][:, None] + np.arange(4)[None, :]).ravel()
        frontier = kids[charged[kids]]
    let_nodes = int((need & charged).sum())
    own = int(((b0 >= lo) & (b1 <= hi) & charged).sum())
    print(f"rank {r}: LET {let_nodes} nodes = {100 * let_nodes / total_charged:5.1f} % of the traversal tree "
          f"(its own subtrees: {100 * own / total_charged:5.1f} %, imported: {100 * (let_nodes - own) / total_charged:5.1f} %)")

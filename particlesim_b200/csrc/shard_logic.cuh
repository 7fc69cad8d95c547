// shard_logic.cuh — the host+device part of the sharded build (shard.cuh): constants, the top-heap record
// and its sweep, the subtree-end rule of a rank's piece.  tests/emu compiles it as plain C++ and replays a
// G-rank build serially (tests/test_emulation.py) before any GPU time is spent.
#pragma once
#include <stdint.h>

#include "tree_logic.cuh"

namespace psim {

constexpr int kShardDepth = 8;
constexpr uint32_t kBins = 1u << (2 * kShardDepth);
constexpr uint32_t kTopSlots = ((1u << (2 * (kShardDepth + 1))) - 1u) / 3u;  // cells of depth 0..8
constexpr int kMaxRanks = 64;
constexpr uint32_t kHaloMax = 2050;  // >= effective leaf capacity + 1

PSIM_HD uint32_t top_base(int d) { return ((1u << (2 * d)) - 1u) / 3u; }
PSIM_HD uint32_t top_slot(int d, uint64_t key) { return top_base(d) + (d ? (uint32_t)(key >> (64 - 2 * d)) : 0u); }

struct TopRec {  // 40 bytes = 5 u64 words: exactly one rank writes a slot, the others leave zeros
  NodeRec r;
  uint32_t node;   // global pre-order index
  uint32_t state;  // 0 absent, 1 complete, 2 internal above the bins, 3 internal bin (complete after the local sweep)
};
constexpr uint32_t kTopAbsent = 0, kTopComplete = 1, kTopInternal = 2, kTopBin = 3, kTopComputed = 4;

struct ShardPlan {       // device + host copy
  uint32_t bin_lo[kMaxRanks + 1];
  uint32_t body_lo[kMaxRanks + 1];
};

struct ShardMeta {  // device side, one per context
  uint32_t rank, world;
  uint32_t n_local, hl, L;  // local bodies, left halo length, hl + n_local + hr
  uint32_t body_base;       // global body index of local array slot 0
  uint32_t node_off, M_local, M_total;
  uint32_t trav_off, T_local, T_total;
  uint32_t node_lo[kMaxRanks + 1], trav_lo[kMaxRanks + 1];
};

// one cell (d, p) of the top heap: children in quadrant order, the reference's ((c0 + c1) + c2) + c3
// (quadtree.rs:142-149); the skip pointer is the last present child's
PSIM_HD void heap_sweep_cell(TopRec* heap, int d, uint32_t p) {
  TopRec* me = &heap[top_base(d) + p];
  if (me->state != kTopInternal) return;
  double aq = 0.0, aqx = 0.0, aqy = 0.0;
  float charge = 0.0f;
  uint32_t next = 0;
  for (uint32_t q = 0; q < 4; ++q) {
    const TopRec* c = &heap[top_base(d + 1) + 4 * p + q];
    if (c->state == kTopAbsent) continue;
    charge = f_add(charge, c->r.charge);
    aq += c->r.aq, aqx += c->r.aqx, aqy += c->r.aqy;
    next = c->r.next & kNextMask;
  }
  bool last = true;  // no later sibling under my parent
  if (d > 0)
    for (uint32_t q = (p & 3u) + 1; q < 4; ++q)
      if (heap[top_base(d) + (p & ~3u) + q].state != kTopAbsent) last = false;
  me->r.aq = aq, me->r.aqx = aqx, me->r.aqy = aqy, me->r.charge = charge;
  me->r.next = next | (last ? kLastSibling : 0u);
  me->state = kTopComputed;
}

struct SubtreeEndShard {  // the subtree ends where a remote node starts: at a bin boundary
  uint32_t node_lo, node_hi, M, n_bodies, body_base;
  const uint4* nodeB;  // pre-offset
  const uint64_t* lkeys;
  const uint32_t* binprefix;
  PSIM_HD uint32_t operator()(uint32_t c, const uint4& nb) const {
    if (c >= M) return n_bodies;
    if (c < node_hi) return nodeB[c].y;
    const int d = (int)(nb.w & kNodeDepthMask), de = d < kShardDepth ? d : kShardDepth;
    const uint64_t key = lkeys[nb.y - body_base];
    const uint32_t prefix = de ? (uint32_t)(key >> (64 - 2 * de)) : 0u;
    return binprefix[(prefix + 1u) << (2 * (kShardDepth - de))];
  }
};

}  // namespace psim

// let.cuh — locally essential traversal tree: which of a rank's traversal records another rank can ever touch.
//
// SURVEY.md 8e ("LET exchange"): after the key-range sharded build (shard.cuh) every rank holds its own piece of the
// traversal tree; instead of all-gathering all pieces, rank S sends rank R only the records R's walks can reach.
// A walk reaches node n only by OPENING n's parent, and a target t opens a cell of size s only if
// s^2 >= dist_adj(t, centre)^2 * theta^2 (quadtree.rs:361-371), with dist_adj >= dist(t, cell) - radius_t.  So
//     needed(n, R)  <=  size(parent(n)) >= theta * (dist(cell(parent(n)), region(R)) - margin)
// is a superset test that only needs geometry: the parent's cell from the key prefix of the node's first body, and
// region(R) = the bins (depth-8 cells) that hold R's targets, a contiguous Morton range and hence <= ~48 aligned
// squares.  `margin` covers the targets' radii, the electrons' offsets from their bodies and the rounding of the fp32
// cell recurrence.  The test is monotone up the tree (a parent's cell contains the child's), so every needed node's
// ancestors, siblings and skip-pointer target are needed too: the records keep their GLOBAL traversal indices and land
// at the same positions of travA / travB on the receiver; positions nobody sends hold stale records no walk can reach.
// Bodies stay replicated (leaf terms read them by global index).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "shard.cuh"

namespace psim {

constexpr int kLetSquares = 64;

struct LetSquare {
  float x0, y0, size;  // normalised to the root square [0, 1)^2
};
struct LetRegions {
  uint32_t count[kMaxRanks];
  LetSquare sq[kMaxRanks][kLetSquares];
};
struct LetRec {  // one traversal record on the wire
  float4 a;
  uint4 b;
  uint32_t g;  // global traversal index
  uint32_t pad[3];
};

// x (even) / y (odd) bits of a `levels`-digit quadrant prefix (digit = qy << 1 | qx, first level most significant)
__device__ __forceinline__ void let_deinterleave(uint64_t prefix, int levels, uint32_t& ix, uint32_t& iy) {
  ix = iy = 0;
  for (int l = 0; l < levels; ++l) {
    const uint32_t d = (uint32_t)(prefix >> (2 * (levels - 1 - l))) & 3u;
    ix = (ix << 1) | (d & 1u), iy = (iy << 1) | (d >> 1);
  }
}

// One thread per rank: the bins that hold the rank's body targets (its slice of the sorted order) and the bodies of
// its electron targets, as aligned squares.
__global__ void let_regions_kernel(const uint32_t* __restrict__ binprefix, uint32_t n, uint32_t m, uint32_t wb,
                                   uint32_t we, const uint32_t* __restrict__ ebody, uint32_t world,
                                   LetRegions* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= world) return;
  uint32_t b0 = kBins, b1 = 0;
  auto cover = [&](uint32_t first, uint32_t last) {  // inclusive body range
    const uint32_t lo = bin_of_body(binprefix, first), hi = bin_of_body(binprefix, last);
    b0 = lo < b0 ? lo : b0, b1 = hi > b1 ? hi : b1;
  };
  {
    const uint32_t f = min(r * wb, n), c = min(wb, n - f);
    if (c) cover(f, f + c - 1);
  }
  if (m && ebody) {
    const uint32_t f = min(r * we, m), c = min(we, m - f);
    if (c) cover(min(ebody[f], n - 1), min(ebody[f + c - 1], n - 1));
  }
  uint32_t cnt = 0;
  if (b0 <= b1) {
    uint32_t lo = b0;
    while (lo <= b1 && cnt < kLetSquares) {
      int lvl = 0;  // block of 4^lvl bins
      while (lvl < kShardDepth) {
        const uint32_t big = 1u << (2 * (lvl + 1));
        if ((lo & (big - 1u)) != 0 || (uint64_t)lo + big - 1u > b1) break;
        ++lvl;
      }
      const int depth = kShardDepth - lvl;
      uint32_t ix, iy;
      let_deinterleave((uint64_t)(lo >> (2 * lvl)), depth, ix, iy);
      const float s = ldexpf(1.0f, -depth);
      out->sq[r][cnt++] = LetSquare{ix * s, iy * s, s};
      lo += 1u << (2 * lvl);
    }
    if (lo <= b1) {  // more squares than slots (cannot happen for a contiguous range): cover everything
      cnt = 1;
      out->sq[r][0] = LetSquare{0.0f, 0.0f, 1.0f};
    }
  }
  out->count[r] = cnt;
}

// Every record of this rank's traversal segment goes to the send area of each rank that may reach it.
__global__ void __launch_bounds__(256)
    let_select_kernel(const ShardMeta* __restrict__ sm, const float4* __restrict__ travA, const uint4* __restrict__ travB,
                      const uint64_t* __restrict__ lkeys, const LetRegions* __restrict__ regions,
                      const TreeMeta* __restrict__ meta, float theta, float margin_abs, uint32_t cap_per_rank,
                      LetRec* __restrict__ send, uint32_t* __restrict__ cnt) {
  // targets' radii + electron offsets + the fp32 cell recurrence's rounding, in units of the root square
  const float margin = margin_abs / meta->root.size + 1e-4f;
  __shared__ LetSquare s_sq[kLetSquares];
  __shared__ uint32_t s_n;
  const uint32_t T = sm->T_local, toff = sm->trav_off, world = sm->world, me = sm->rank, body_base = sm->body_base;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t dst = 0; dst < world; ++dst) {
    if (dst == me) continue;
    __syncthreads();
    if (threadIdx.x == 0) s_n = regions->count[dst];
    if (threadIdx.x < kLetSquares) s_sq[threadIdx.x] = regions->sq[dst][threadIdx.x];
    __syncthreads();
    const uint32_t nsq = s_n;
    for (uint32_t base = blockIdx.x * blockDim.x; base < T; base += gridDim.x * blockDim.x) {
      const uint32_t k = base + threadIdx.x;
      bool need = false;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      uint4 b = make_uint4(0, 0, 0, 0);
      if (k < T && nsq) {
        a = travA[toff + k], b = travB[toff + k];
        const int d = (int)(b.w & kNodeDepthMask);
        if (d <= 1) {
          need = true;  // the root and its children: every walk starts there
        } else {
          const uint64_t key = lkeys[b.y - body_base];
          uint32_t ix, iy;
          let_deinterleave(key >> (64 - 2 * (d - 1)), d - 1, ix, iy);
          const float s = ldexpf(1.0f, -(d - 1));
          const float x0 = ix * s, y0 = iy * s, x1 = x0 + s, y1 = y0 + s;
          float best = 3.0f;
          for (uint32_t q = 0; q < nsq; ++q) {
            const LetSquare g = s_sq[q];
            const float dx = fmaxf(fmaxf(g.x0 - x1, x0 - (g.x0 + g.size)), 0.0f);
            const float dy = fmaxf(fmaxf(g.y0 - y1, y0 - (g.y0 + g.size)), 0.0f);
            best = fminf(best, dx * dx + dy * dy);
          }
          const float dist = fmaxf(sqrtf(best) - margin, 0.0f);
          need = s * 1.0001f >= theta * dist;
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, need);
      uint32_t slot0 = 0;
      if (lane == 0 && bal) slot0 = atomicAdd(&cnt[dst], (uint32_t)__popc(bal));
      slot0 = __shfl_sync(0xffffffffu, slot0, 0);
      if (need) {
        const uint32_t slot = slot0 + __popc(bal & ((1u << lane) - 1u));
        if (slot < cap_per_rank) {
          LetRec r;
          r.a = a, r.b = b, r.g = toff + k, r.pad[0] = r.pad[1] = r.pad[2] = 0;
          send[(size_t)dst * cap_per_rank + slot] = r;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
    let_scatter_kernel(const LetRec* __restrict__ recv, uint32_t count, uint32_t node_cap, float4* __restrict__ travA,
                       uint4* __restrict__ travB) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const LetRec r = recv[k];
    if (r.g < node_cap) travA[r.g] = r.a, travB[r.g] = r.b;
  }
}

// test aid (PSIM_LET_POISON=1): records nobody sent must be unreachable, so they may hold anything
__global__ void __launch_bounds__(256)
    let_poison_kernel(const ShardMeta* __restrict__ sm, float4* __restrict__ travA, uint4* __restrict__ travB) {
  const uint32_t T = sm->T_total, lo = sm->trav_off, hi = sm->trav_off + sm->T_local;
  const uint32_t stride = gridDim.x * blockDim.x;
  const float nanv = __int_as_float(0x7fc00000);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < T; k += stride)
    if (k < lo || k >= hi) travA[k] = make_float4(nanv, nanv, nanv, nanv), travB[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffu);
}

}  // namespace psim

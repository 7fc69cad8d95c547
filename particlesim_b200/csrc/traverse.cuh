// traverse.cuh — Barnes-Hut Coulomb field / force traversal.
//
// Replaces Quadtree::acc_pos (src/quadtree/quadtree.rs:350-407) driven by Quadtree::field
// (:418-427), Quadtree::field_at_point (:504-507), the `attract` epilogue
// (src/simulation/forces.rs:37-43) and the per-electron drift of Body::update_electrons
// (src/body/electron.rs:19-46).
//
// One warp walks the compact pre-order tree for 32 spatially adjacent targets (targets are in
// Morton order).  The walk visits the union of the 32 reference traversals in DFS order; every
// lane applies the reference's opening test with ITS OWN position and radius, so each target sums
// exactly the reference's interaction set, in the reference's order:
//   - a lane that accepts a node (or direct-sums a leaf) records the node's skip pointer; it ignores
//     every node below that index, i.e. the subtree it has already accounted for;
//   - the warp descends (n+1) if any still-active lane rejects an internal node, else follows the
//     skip pointer.
// Node records are two 16-byte loads at a warp-uniform address (one L1 broadcast each); descending
// is a sequential walk through memory because of the pre-order layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tree.cuh"

namespace psim {

struct FieldParams {
  float t_sq, e_sq, k_e;
  float bg_x, bg_y;
  float inv_theta;  // 1 / theta, only used by conservative pre-tests (never decides a borderline case)
  int one_mufu;     // fast mode only: theta <= 1, epsilon and root size in the range where 1 / ((d^2 + e^2) d) is taken
                    // as ONE rsqrt of (d^2 + e^2)^2 d^2 (see bh_group_walk)
};

// PARITY = true: IEEE sqrt/div, no FMA contraction, the reference's operation order.
// PARITY = false: rsqrt / fast divide / FMA (reported separately; still checked against the oracle).
template <bool PARITY>
struct Arith {
  static __device__ __forceinline__ float add(float a, float b) { return PARITY ? __fadd_rn(a, b) : a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return PARITY ? __fsub_rn(a, b) : a - b; }
  static __device__ __forceinline__ float mul(float a, float b) { return PARITY ? __fmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ float div(float a, float b) { return PARITY ? __fdiv_rn(a, b) : __fdividef(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return PARITY ? __fsqrt_rn(a) : __fsqrt_rn(a); }
};

// MUFU approximations without the denormal pre-scaling of rsqrtf() / __fdividef() (inputs here are
// squared distances and softened cubes of them, never denormal unless exactly zero)
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct TraverseCounters {
  unsigned long long steps;  // warp steps (nodes visited by a warp)
};

// Shared walk.  (px, py, q, radius) are per-lane; `live` is false for tail lanes.
// COUNT adds per-lane interaction counters (cnt[0] = internal nodes this lane opened, cnt[1] =
// monopoles accepted, cnt[2] = direct body terms): the reference visits 1 + 4 * opened nodes.
template <bool PARITY, bool COUNT = false>
__device__ __forceinline__ float2 bh_walk(const float4* __restrict__ nodeA,
                                          const uint4* __restrict__ nodeB,
                                          const float4* __restrict__ pqr, uint32_t M, float px,
                                          float py, float q, float radius, bool live,
                                          const FieldParams P, uint32_t& steps_out,
                                          uint32_t* cnt = nullptr) {
  using A = Arith<PARITY>;
  float ax = 0.0f, ay = 0.0f;
  uint32_t skip = live ? 0u : 0xffffffffu;
  const float kq = A::mul(P.k_e, q);
  uint32_t n = 0, steps = 0;
  while (n < M) {
    const float4 na = __ldg(&nodeA[n]);
    const uint4 nb = __ldg(&nodeB[n]);
    ++steps;
    const bool active = n >= skip;
    const float dx = A::sub(px, na.x), dy = A::sub(py, na.y);
    const float d_sq = A::add(A::mul(dx, dx), A::mul(dy, dy));
    const float dist = A::sqrt(d_sq);
    const float dist_adj = fmaxf(A::sub(dist, radius), 0.0f);
    const bool accept = A::mul(na.w, na.w) < A::mul(A::mul(dist_adj, dist_adj), P.t_sq);
    const bool leaf = (nb.w & kNodeLeaf) != 0;
    if (active) {
      if (accept) {
        if (na.z != 0.0f) {  // a zero-charge monopole adds exactly +-0
          const float r_eff = fmaxf(dist, A::add(radius, A::mul(na.w, 0.5f)));
          const float denom = A::mul(A::add(A::mul(r_eff, r_eff), P.e_sq), r_eff);
          const float s = A::div(A::mul(kq, na.z), denom);
          ax = A::add(ax, A::mul(dx, s));
          ay = A::add(ay, A::mul(dy, s));
        }
        skip = nb.x;
        if (COUNT) cnt[1]++;
      } else if (leaf) {
        for (uint32_t b = nb.y; b < nb.y + nb.z; ++b) {
          const float4 s4 = __ldg(&pqr[b]);
          if (!COUNT && s4.z == 0.0f) continue;  // a zero-charge body adds exactly +-0
          const float ex = A::sub(s4.x, px), ey = A::sub(s4.y, py);
          if (A::add(A::mul(ex, ex), A::mul(ey, ey)) < 1e-6f) continue;  // positional self skip
          const float bx = A::sub(px, s4.x), by = A::sub(py, s4.y);
          const float bd = A::sqrt(A::add(A::mul(bx, bx), A::mul(by, by)));
          const float r_eff = fmaxf(bd, A::add(radius, s4.w));
          const float denom = A::mul(A::add(A::mul(r_eff, r_eff), P.e_sq), r_eff);
          const float s = fminf(A::div(A::mul(kq, s4.z), denom), 3.402823466e+38f);
          ax = A::add(ax, A::mul(bx, s));
          ay = A::add(ay, A::mul(by, s));
          if (COUNT) cnt[2]++;
        }
        skip = nb.x;
      } else if (COUNT) {
        cnt[0]++;
      }
    }
    const bool descend = __any_sync(0xffffffffu, active && !accept && !leaf);
    n = descend ? n + 1 : nb.x;
  }
  steps_out = steps;
  return make_float2(ax, ay);
}

// Quadtree::field + attract: targets are the bodies themselves (q_test = 1, radius_i).
template <bool PARITY>
__global__ void __launch_bounds__(128)
    bh_field_bodies_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                           const uint4* __restrict__ nodeB, const float4* __restrict__ pqr,
                           const float4* __restrict__ acc_mass /* {ax, ay, az, mass} */, uint32_t n,
                           FieldParams P, float2* __restrict__ e_field,
                           float4* __restrict__ acc_mass_out, int write_acc,
                           unsigned long long* __restrict__ step_counter) {
  const uint32_t M = meta->err ? 0u : meta->num_nodes;  // arena overflow: walk nothing
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_block;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t groups = (n + 31) / 32;
  unsigned long long my_steps = 0;
  for (uint32_t g = warp_global; g < groups; g += total_warps) {
    const uint32_t i = g * 32 + lane;
    const bool live = i < n;
    float4 me = make_float4(0, 0, 0, 0);
    if (live) me = pqr[i];
    uint32_t steps;
    float2 e = bh_walk<PARITY>(nodeA, nodeB, pqr, M, me.x, me.y, 1.0f, me.w, live, P, steps);
    my_steps += steps;
    if (live) {
      // forces.rs:37-43: e_field += background; acc = (charge * e_field) / mass
      e.x = __fadd_rn(e.x, P.bg_x);
      e.y = __fadd_rn(e.y, P.bg_y);
      e_field[i] = e;
      if (write_acc) {
        float4 am = acc_mass[i];
        am.x = __fdiv_rn(__fmul_rn(me.z, e.x), am.w);
        am.y = __fdiv_rn(__fmul_rn(me.z, e.y), am.w);
        acc_mass_out[i] = am;
      }
    }
  }
  if (step_counter && lane == 0 && my_steps) atomicAdd(step_counter, my_steps);
}

// diagnostic: the interaction counters of Quadtree::field for the roofline's algorithmic flops
__global__ void __launch_bounds__(128)
    bh_count_bodies_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                           const uint4* __restrict__ nodeB, const float4* __restrict__ pqr, uint32_t n,
                           FieldParams P, unsigned long long* __restrict__ out4) {
  const uint32_t M = meta->err ? 0u : meta->num_nodes;  // arena overflow: walk nothing
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_block;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t groups = (n + 31) / 32;
  unsigned long long opened = 0, acc = 0, pairs = 0, wsteps = 0;
  for (uint32_t g = warp_global; g < groups; g += total_warps) {
    const uint32_t i = g * 32 + lane;
    const bool live = i < n;
    float4 me = make_float4(0, 0, 0, 0);
    if (live) me = pqr[i];
    uint32_t steps, cnt[3] = {0, 0, 0};
    bh_walk<true, true>(nodeA, nodeB, pqr, M, me.x, me.y, 1.0f, me.w, live, P, steps, cnt);
    opened += cnt[0], acc += cnt[1], pairs += cnt[2];
    if (lane == 0) wsteps += steps;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    opened += __shfl_xor_sync(0xffffffffu, opened, off);
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    pairs += __shfl_xor_sync(0xffffffffu, pairs, off);
  }
  if (lane == 0) {
    atomicAdd(&out4[0], opened);
    atomicAdd(&out4[1], acc);
    atomicAdd(&out4[2], pairs);
    atomicAdd(&out4[3], wsteps);
  }
}

// FP32 FMA pipe peak: 8 independent FMA chains per thread (no memory traffic)
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b), x1 = fmaf(x1, a, b), x2 = fmaf(x2, a, b), x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b), x5 = fmaf(x5, a, b), x6 = fmaf(x6, a, b), x7 = fmaf(x7, a, b);
    }
  }
  const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 123.456f) out[0] = r;
}

// acc_pos at arbitrary points.  q / radius may be null (1 and 0: field_at_point).
template <bool PARITY>
__global__ void __launch_bounds__(128)
    bh_field_points_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                           const uint4* __restrict__ nodeB, const float4* __restrict__ pqr,
                           const float2* __restrict__ pts, const float* __restrict__ q,
                           const float* __restrict__ radius, uint32_t m, FieldParams P,
                           float2* __restrict__ out, unsigned long long* __restrict__ step_counter) {
  const uint32_t M = meta->err ? 0u : meta->num_nodes;  // arena overflow: walk nothing
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_block;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t groups = (m + 31) / 32;
  unsigned long long my_steps = 0;
  for (uint32_t g = warp_global; g < groups; g += total_warps) {
    const uint32_t i = g * 32 + lane;
    const bool live = i < m;
    float2 p = make_float2(0, 0);
    float qq = 1.0f, rr = 0.0f;
    if (live) {
      p = pts[i];
      if (q) qq = q[i];
      if (radius) rr = radius[i];
    }
    uint32_t steps;
    const float2 e = bh_walk<PARITY>(nodeA, nodeB, pqr, M, p.x, p.y, qq, rr, live, P, steps);
    my_steps += steps;
    if (live) out[i] = e;
  }
  if (step_counter && lane == 0 && my_steps) atomicAdd(step_counter, my_steps);
}


// ------------------------------------------------------------------------------------------------
// Group walk: the warp's 32 targets share one walk of the tree.  Lanes classify 32 tree nodes per
// round (one node per lane) against the group's bounding box; a node is
//   - accepted by every target   -> monopole term for all (shared interaction entry),
//   - rejected by every target   -> opened (children pushed) or, for a leaf, direct-summed by all,
//   - otherwise                  -> the node is broadcast and every lane evaluates the reference's opening
//                                   test EXACTLY for its own target; accepting targets take the monopole,
//                                   the rest carry on below the node (per-target mask).
// The box tests are conservative (2e-5 margin, far above the worst-case rounding of the exact test)
// and never decide a borderline case, so every target sums exactly the reference's interaction set
// with the reference's per-term arithmetic; only the order of the additions differs from acc_pos.
// Pending nodes live in a per-warp ring buffer in shared memory: entry = (node, mask of targets that
// reach the node).  A lane that opens a node finds all its children at once (child k+1 is the skip
// pointer of child k, until it equals the parent's), so the frontier widens by the branching factor
// every round and a walk takes about as many rounds as the tree is deep.
constexpr int kStackCap = 512;               // a power of two: ring indices wrap with a mask
constexpr int kLifoAbove = kStackCap - 192;  // see the capacity argument in bh_group_walk
struct WarpShared {
  uint2 st[kStackCap];  // {node, mask of targets that reach it}
  // the round's nodes that some target may accept, staged for broadcast reads
  float4 l_node[32];  // {centre.x, centre.y, charge, size}
  uint32_t l_mask[32]; // mask of targets that reach the node
  uint32_t l_acc[32];  // result: mask of targets that accepted
  // the round's charged nodes that every target reaching them accepts: no test, no result
  float4 s_node[32];
  uint32_t s_mask[32];
};

// ONE (fast mode, theta <= 1): a target that accepts a node has dist - radius > size (quadtree.rs:361-371 with
// theta <= 1), so r_eff = max(dist, radius + size / 2) IS dist and the monopole weight q / ((d^2 + e^2) d) is one
// MUFU rsqrt of (d^2 + e^2)^2 d^2; k_e q_target is factored out of the whole sum.  The host only selects it where that
// product can neither overflow nor underflow (FieldParams::one_mufu).
template <bool PARITY, bool ONE = false>
__device__ __forceinline__ float2 bh_group_walk(const float4* __restrict__ nodeA,
                                                const uint4* __restrict__ nodeB,
                                                const float4* __restrict__ pqr, uint32_t M, float px,
                                                float py, float q, float radius, bool live,
                                                const FieldParams P, WarpShared& ws,
                                                uint32_t& nodes_out, float root_size) {
  using A = Arith<PARITY>;
  const uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  float ax = 0.0f, ay = 0.0f;
  nodes_out = 0;
  const float INF = __int_as_float(0x7f800000);
  // A target at a NaN / infinite position never passes an opening test, visits every body of the tree and comes
  // back with NaN (quadtree.rs:361-394: d, d_sq and every term are NaN).  That result is returned at once; the
  // lane sits out the shared walk instead of dragging its 31 neighbours through all N leaves.
  const bool lost = live && !(fabsf(px) < INF && fabsf(py) < INF) && M != 0;
  if (lost) live = false;
  const float lost_value = __int_as_float(0x7fc00000);
  const uint32_t live_mask = __ballot_sync(FULL, live);
  if (live_mask == 0 || M == 0) return lost ? make_float2(lost_value, lost_value) : make_float2(0.f, 0.f);
  // bounding box and radius range of the live targets
  float bx0 = live ? px : INF, bx1 = live ? px : -INF, by0 = live ? py : INF, by1 = live ? py : -INF;
  float rmin = live ? radius : INF, rmax = live ? radius : -INF;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    bx0 = fminf(bx0, __shfl_xor_sync(FULL, bx0, off));
    bx1 = fmaxf(bx1, __shfl_xor_sync(FULL, bx1, off));
    by0 = fminf(by0, __shfl_xor_sync(FULL, by0, off));
    by1 = fmaxf(by1, __shfl_xor_sync(FULL, by1, off));
    rmin = fminf(rmin, __shfl_xor_sync(FULL, rmin, off));
    rmax = fmaxf(rmax, __shfl_xor_sync(FULL, rmax, off));
  }
  // a target with a NaN / infinite coordinate or radius makes every box test meaningless: such a
  // group takes the exact test everywhere
  const bool finite_me = !live || (fabsf(px) < INF && fabsf(py) < INF && fabsf(radius) < INF);
  const bool box_ok = __all_sync(FULL, finite_me);
  const float kq = ONE ? 1.0f : A::mul(P.k_e, q);
  const float my_lim_r = radius;
  const uint32_t lanebit = 1u << lane;
  if (lane == 0) ws.st[0] = make_uint2(0u, live_mask);
  __syncwarp();
  // Ring buffer served first-in-first-out (wide rounds).  Capacity: a FIFO round pops k <= 32 entries
  // and pushes <= 4k, so starting at size <= kLifoAbove it ends at <= kLifoAbove + 96; above that the
  // walk goes last-in-first-out one node per round, where at most 3 siblings per level stay pending
  // (<= 96 more for 32 levels): kLifoAbove + 96 + 96 <= kStackCap (512).
  int head = 0, size = 1;
  uint32_t visited = 0;
  while (size > 0) {
    const bool lifo = size > kLifoAbove;
    const int k = lifo ? 1 : (size < 32 ? size : 32);
    const bool has = lane < k;
    uint32_t node = 0, mask = 0;
    if (has) {
      const uint2 ent = ws.st[(lifo ? head + size - 1 : head + lane) & (kStackCap - 1)];
      node = ent.x, mask = ent.y;
    }
    __syncwarp();
    if (!lifo) head = (head + k) & (kStackCap - 1);
    size -= k;
    visited += k;
    float4 na = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 nb = make_uint4(0, 0, 0, 0);
    bool leaf = false;
    uint32_t c3 = 0;
    int cls = 2;  // 0 ambiguous, 1 every target accepts, 2 every target rejects
    if (has) {
      na = __ldg(&nodeA[node]);
      nb = __ldg(&nodeB[node]);
      leaf = (nb.w & kNodeLeaf) != 0;
      // traversal records of internal nodes carry their children (link_children_kernel): B.y, B.z and the
      // bits of A.w are children 1..3; the cell size is the root's halved `depth` times, exactly
      c3 = __float_as_uint(na.w);
      {
        // root_size * 2^-depth: an exponent subtraction while the result stays a normal number (always, for any
        // root a simulation can have: depth <= 32), ldexpf otherwise
        const uint32_t rb = __float_as_uint(root_size), dep = nb.w & kNodeDepthMask;
        const uint32_t ex = (rb >> 23) & 0xffu;
        na.w = (ex > dep && ex != 0xffu) ? __uint_as_float(rb - (dep << 23)) : ldexpf(root_size, -(int)dep);
      }
      cls = 0;
      if (box_ok) {
        const float s_t = na.w * P.inv_theta;
        const float ddx = fmaxf(fmaxf(bx0 - na.x, na.x - bx1), 0.0f);
        const float ddy = fmaxf(fmaxf(by0 - na.y, na.y - by1), 0.0f);
        const float fx = fmaxf(na.x - bx0, bx1 - na.x), fy = fmaxf(na.y - by0, by1 - na.y);
        const float dmin2 = ddx * ddx + ddy * ddy, dmax2 = fx * fx + fy * fy;
        const float la = s_t + rmax, lr = s_t + rmin;
        if (dmin2 > la * la * 1.00002f) cls = 1;
        else if (dmax2 < lr * lr * 0.99998f) cls = 2;
      }
    }
    // Nodes that every reaching target accepts (and that carry charge) go to the "sure" list: monopole
    // only.  Undecided nodes go to the second list: each lane applies the reference's opening test
    // with its own position and radius (quadtree.rs:361-371) to every staged node (broadcast reads)
    // and, if it accepts, adds the monopole (:372-375); the ballot is the node's accept mask.
    uint32_t acc_mask = (has && cls == 1) ? mask : 0u;
    const bool in_sure = has && cls == 1 && na.z != 0.0f;  // charge 0 adds exactly +-0
    const bool in_list = has && cls == 0;
    const uint32_t sm = __ballot_sync(FULL, in_sure);
    const uint32_t lm = __ballot_sync(FULL, in_list);
    const int scnt = __popc(sm), cnt = __popc(lm);
    const int my_slot = __popc(lm & lt);
    if (in_sure) {
      const int sl = __popc(sm & lt);
      if (ONE) {
        ws.s_node[sl] = make_float4(na.x, na.y, na.z, __uint_as_float(mask));  // the size is not needed: r_eff = dist
      } else {
        ws.s_node[sl] = na;
        ws.s_mask[sl] = mask;
      }
    }
    if (in_list) {
      ws.l_node[my_slot] = na;
      ws.l_mask[my_slot] = mask;
    }
    __syncwarp();
    auto monopole = [&](const float4& nd, float dx, float dy, float d_sq, float dist, bool have_dist) {
      if (PARITY) {
        if (!have_dist) dist = __fsqrt_rn(d_sq);
        const float r_eff = fmaxf(dist, __fadd_rn(radius, __fmul_rn(nd.w, 0.5f)));
        const float denom = __fmul_rn(__fadd_rn(__fmul_rn(r_eff, r_eff), P.e_sq), r_eff);
        const float sc = __fdiv_rn(__fmul_rn(kq, nd.z), denom);
        ax = __fadd_rn(ax, __fmul_rn(dx, sc));
        ay = __fadd_rn(ay, __fmul_rn(dy, sc));
      } else {
        // fast arithmetic on the same interaction set: rsqrt / rcp approximations (2 ulp each), FMA
        const float r_eff = fmaxf(d_sq * rsqrt_ftz(d_sq), fmaf(nd.w, 0.5f, radius));  // fmaxf drops the NaN of d_sq = 0
        const float denom = fmaf(r_eff, r_eff, P.e_sq) * r_eff;
        const float sc = (kq * nd.z) * rcp_ftz(denom);
        ax = fmaf(dx, sc, ax);
        ay = fmaf(dy, sc, ay);
      }
    };
    if (ONE) {
#pragma unroll 2
      for (int it = 0; it < scnt; ++it) {
        const float4 nd = ws.s_node[it];
        const float dx = px - nd.x, dy = py - nd.y;
        const float d_sq = fmaf(dx, dx, dy * dy);
        const float t = d_sq + P.e_sq;
        const float w = rsqrt_ftz((t * t) * d_sq);
        // a lane the node does not reach weighs its term with zero (whatever came out of the rsqrt)
        const float sc = (__float_as_uint(nd.w) & lanebit) ? nd.z * w : 0.0f;
        ax = fmaf(dx, sc, ax);
        ay = fmaf(dy, sc, ay);
      }
    }
    for (int it = 0; !ONE && it < scnt; ++it) {
      const float4 nd = ws.s_node[it];
      const uint32_t m = ws.s_mask[it];
      if (PARITY) {
        if ((m >> lane) & 1u) {
          const float dx = A::sub(px, nd.x), dy = A::sub(py, nd.y);
          const float d_sq = A::add(A::mul(dx, dx), A::mul(dy, dy));
          monopole(nd, dx, dy, d_sq, 0.0f, false);
        }
      } else {
        // fast mode: no branch per entry - a lane the node does not reach weighs its term with zero (a sure node is
        // accepted by every target that reaches it, so d_sq > 0 there; elsewhere the select drops whatever came out)
        const float dx = px - nd.x, dy = py - nd.y;
        const float d_sq = fmaf(dx, dx, dy * dy);
        const float r_eff = fmaxf(d_sq * rsqrt_ftz(d_sq), fmaf(nd.w, 0.5f, radius));
        const float denom = fmaf(r_eff, r_eff, P.e_sq) * r_eff;
        const float sc = ((m >> lane) & 1u) ? (kq * nd.z) * rcp_ftz(denom) : 0.0f;
        ax = fmaf(dx, sc, ax);
        ay = fmaf(dy, sc, ay);
      }
    }
    if (ONE) {
#pragma unroll 2
      for (int it = 0; it < cnt; ++it) {
        const float4 nd = ws.l_node[it];
        const uint32_t m = ws.l_mask[it];
        const float dx = px - nd.x, dy = py - nd.y;
        const float d_sq = fmaf(dx, dx, dy * dy);
        const bool reach = (m & lanebit) != 0;
        // conservative pre-test with this target's own radius (its 2e-5 margins cover the contraction of d_sq); only
        // the band in between takes the reference's test on the reference's d_sq, which alone decides a borderline case
        const float lim = fmaf(nd.w, P.inv_theta, my_lim_r);
        const float lim2 = lim * lim;
        bool acc = reach && d_sq > lim2 * 1.00002f;
        if (reach && !acc && !(d_sq < lim2 * 0.99998f)) {
          const float ex = __fsub_rn(px, nd.x), ey = __fsub_rn(py, nd.y);
          const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
          const float dist_adj = fmaxf(__fsub_rn(dist, radius), 0.0f);
          acc = __fmul_rn(nd.w, nd.w) < __fmul_rn(__fmul_rn(dist_adj, dist_adj), P.t_sq);
        }
        const uint32_t am = __ballot_sync(FULL, acc);
        if (lane == 0) ws.l_acc[it] = am;
        const float t = d_sq + P.e_sq;
        const float w = rsqrt_ftz((t * t) * d_sq);
        const float sc = acc ? nd.z * w : 0.0f;
        ax = fmaf(dx, sc, ax);
        ay = fmaf(dy, sc, ay);
      }
    }
    for (int it = 0; !ONE && it < cnt; ++it) {
      const float4 nd = ws.l_node[it];
      const uint32_t m = ws.l_mask[it];
      // the distance feeds the opening decision: reference arithmetic in every mode
      const float dx = __fsub_rn(px, nd.x), dy = __fsub_rn(py, nd.y);
      const float d_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      const bool reach = ((m >> lane) & 1u) != 0;
      // conservative pre-test with this target's own radius; only the band in between takes the
      // reference's test, which alone decides a borderline case
      const float lim = fmaf(nd.w, P.inv_theta, my_lim_r);
      const float lim2 = lim * lim;
      bool acc = reach && d_sq > lim2 * 1.00002f;
      float dist = 0.0f;
      bool have_dist = false;
      if (reach && !acc && !(d_sq < lim2 * 0.99998f)) {
        dist = __fsqrt_rn(d_sq);
        have_dist = true;
        const float dist_adj = fmaxf(__fsub_rn(dist, radius), 0.0f);
        acc = __fmul_rn(nd.w, nd.w) < __fmul_rn(__fmul_rn(dist_adj, dist_adj), P.t_sq);
      }
      const uint32_t am = __ballot_sync(FULL, acc);
      if (lane == 0) ws.l_acc[it] = am;
      if (PARITY) {
        if (acc && nd.z != 0.0f) monopole(nd, dx, dy, d_sq, dist, have_dist);
      } else {
        // fast mode: branch-free like the sure list (a chargeless node weighs zero by itself)
        const float r_eff = fmaxf(d_sq * rsqrt_ftz(d_sq), fmaf(nd.w, 0.5f, radius));
        const float denom = fmaf(r_eff, r_eff, P.e_sq) * r_eff;
        const float sc = acc ? (kq * nd.z) * rcp_ftz(denom) : 0.0f;
        ax = fmaf(dx, sc, ax);
        ay = fmaf(dy, sc, ay);
      }
    }
    __syncwarp();
    if (in_list) acc_mask = ws.l_acc[my_slot];
    const uint32_t rem = mask & ~acc_mask;
    // a node some target still has to look below: all its children go onto the buffer
    uint32_t kid[4] = {0u, 0u, 0u, 0u};
    int nk = 0;
    if (has && rem != 0 && !leaf) {
      // a missing child is marked by the parent's own skip pointer
      kid[0] = node + 1, nk = 1;
      if (nb.y != nb.x) {
        kid[1] = nb.y, nk = 2;
        if (nb.z != nb.x) {
          kid[2] = nb.z, nk = 3;
          if (c3 != nb.x) kid[3] = c3, nk = 4;
        }
      }
    }
    if (__any_sync(FULL, nk != 0)) {  // rounds near the leaves often push nothing
      int incl = nk;  // inclusive prefix of the child counts over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      const int off = incl - nk;
      const int total = __shfl_sync(FULL, incl, 31);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nk) ws.st[(head + size + off + j) & (kStackCap - 1)] = make_uint2(kid[j], rem);
      size += total;
      __syncwarp();
    }
    // direct terms of this round (quadtree.rs:381-395)
    uint32_t nm = __ballot_sync(FULL, has && rem != 0 && leaf);
    while (nm) {
      const int src = __ffs(nm) - 1;
      nm &= nm - 1;
      const uint32_t b0 = __shfl_sync(FULL, nb.y, src), bn = __shfl_sync(FULL, nb.z, src);
      const uint32_t m = __shfl_sync(FULL, rem, src);
      if ((m >> lane) & 1u) {
        for (uint32_t b = b0; b < b0 + bn; ++b) {
          const float4 s4 = __ldg(&pqr[b]);
          if (s4.z == 0.0f) continue;  // adds exactly +-0
          const float ex = A::sub(s4.x, px), ey = A::sub(s4.y, py);
          if (A::add(A::mul(ex, ex), A::mul(ey, ey)) < 1e-6f) continue;  // positional self skip
          const float bx = A::sub(px, s4.x), by = A::sub(py, s4.y);
          const float bd = A::sqrt(A::add(A::mul(bx, bx), A::mul(by, by)));
          const float r_eff = fmaxf(bd, A::add(radius, s4.w));
          const float denom = A::mul(A::add(A::mul(r_eff, r_eff), P.e_sq), r_eff);
          const float s = fminf(A::div(A::mul(kq, s4.z), denom), 3.402823466e+38f);
          ax = A::add(ax, A::mul(bx, s));
          ay = A::add(ay, A::mul(by, s));
        }
      }
    }
  }
  nodes_out = visited;
  if (lost) return make_float2(lost_value, lost_value);
  if (ONE) {
    const float kq1 = P.k_e * q;
    ax *= kq1, ay *= kq1;
  }
  return make_float2(ax, ay);
}

// Warp-uniform: may this group take the one-rsqrt monopole?  (t^2 d^2 with t = d^2 + e^2 must stay a normal number for
// every distance a walk of this tree can meet: d <= ~1.3e6 from the bounds below, d > the smallest cell otherwise.)
__device__ __forceinline__ bool one_mufu_ok(const FieldParams& P, const RootQuad& root, float px, float py, bool live) {
  const bool tree_ok = P.one_mufu && root.size >= 1e-3f && root.size <= 262144.0f && fabsf(root.cx) <= 262144.0f &&
                       fabsf(root.cy) <= 262144.0f;
  const bool me_ok = !live || (fabsf(px) <= 524288.0f && fabsf(py) <= 524288.0f);
  return __all_sync(0xffffffffu, tree_ok && me_ok);
}

template <bool PARITY>
__global__ void __launch_bounds__(128)
    bh_group_bodies_kernel(const float4* __restrict__ nodeA, const uint4* __restrict__ nodeB,
                           const uint32_t* __restrict__ num_nodes, const float4* __restrict__ pqr,
                           const float4* __restrict__ acc_mass, uint32_t first, uint32_t n,
                           FieldParams P, float2* __restrict__ e_field,
                           float4* __restrict__ acc_mass_out, int write_acc,
                           unsigned long long* __restrict__ step_counter, const TreeMeta* __restrict__ meta) {
  __shared__ WarpShared ws[4];
  const uint32_t M = *num_nodes;
  const float root_size = meta->root.size;
  const uint32_t g = blockIdx.x * 4 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t i = first + g * 32 + lane;
  const bool live = g * 32 + lane < n;
  float4 me = make_float4(0, 0, 0, 0);
  if (live) me = pqr[i];
  uint32_t visited;
  float2 e;
  if (!PARITY && one_mufu_ok(P, meta->root, me.x, me.y, live))
    e = bh_group_walk<PARITY, !PARITY>(nodeA, nodeB, pqr, M, me.x, me.y, 1.0f, me.w, live, P, ws[threadIdx.x >> 5], visited,
                                       root_size);
  else
    e = bh_group_walk<PARITY>(nodeA, nodeB, pqr, M, me.x, me.y, 1.0f, me.w, live, P, ws[threadIdx.x >> 5], visited,
                              root_size);
  if (live) {
    e.x = __fadd_rn(e.x, P.bg_x);
    e.y = __fadd_rn(e.y, P.bg_y);
    e_field[i] = e;
    if (write_acc) {
      float4 am = acc_mass[i];
      am.x = __fdiv_rn(__fmul_rn(me.z, e.x), am.w);
      am.y = __fdiv_rn(__fmul_rn(me.z, e.y), am.w);
      acc_mass_out[i] = am;
    }
  }
  if (step_counter && lane == 0 && visited) atomicAdd(step_counter, (unsigned long long)visited);
}

template <bool PARITY>
__global__ void __launch_bounds__(128)
    bh_group_points_kernel(const float4* __restrict__ nodeA, const uint4* __restrict__ nodeB,
                           const uint32_t* __restrict__ num_nodes, const float4* __restrict__ pqr,
                           const float2* __restrict__ pts, const float* __restrict__ q,
                           const float* __restrict__ radius, uint32_t first, uint32_t m, FieldParams P,
                           float2* __restrict__ out, unsigned long long* __restrict__ step_counter,
                           const TreeMeta* __restrict__ meta) {
  __shared__ WarpShared ws[4];
  const uint32_t M = *num_nodes;
  const float root_size = meta->root.size;
  const uint32_t g = blockIdx.x * 4 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t i = first + g * 32 + lane;
  const bool live = g * 32 + lane < m;
  float2 p = make_float2(0, 0);
  float qq = 1.0f, rr = 0.0f;
  if (live) {
    p = pts[i];
    if (q) qq = q[i];
    if (radius) rr = radius[i];
  }
  uint32_t visited;
  float2 e;
  if (!PARITY && one_mufu_ok(P, meta->root, p.x, p.y, live))
    e = bh_group_walk<PARITY, !PARITY>(nodeA, nodeB, pqr, M, p.x, p.y, qq, rr, live, P, ws[threadIdx.x >> 5], visited,
                                       root_size);
  else
    e = bh_group_walk<PARITY>(nodeA, nodeB, pqr, M, p.x, p.y, qq, rr, live, P, ws[threadIdx.x >> 5], visited, root_size);
  if (live) out[i] = e;
  if (step_counter && lane == 0 && visited) atomicAdd(step_counter, (unsigned long long)visited);
}

// ------------------------------------------------------------------------------------------------
// electrons (body/electron.rs:19-46).  Electrons are stored grouped by body, in body order, so the
// sample points of consecutive electrons are spatially adjacent.
__global__ void __launch_bounds__(256)
    electron_points_kernel(const float4* __restrict__ pqr, const uint32_t* __restrict__ ebody,
                           const float2* __restrict__ erel, uint32_t m, float2* __restrict__ pts) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const float4 b = pqr[ebody[k]];
    const float2 r = erel[k];
    pts[k] = make_float2(__fadd_rn(b.x, r.x), __fadd_rn(b.y, r.y));
  }
}

struct SpeciesRow {  // mirrors psim_species in include/psim_b200.h
  float mass, radius, damping;
  float lj_epsilon, lj_sigma, lj_cutoff;
  float polar_offset, polar_charge;
  float repulsion_strength, repulsion_cutoff;
  uint32_t lj_enabled, repulsion_enabled;
};
constexpr int kMaxSpecies = 32;

__global__ void __launch_bounds__(256)
    electron_drift_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                          const SpeciesRow* __restrict__ table, const uint32_t* __restrict__ ebody,
                          float2* __restrict__ erel, float2* __restrict__ evel,
                          const float2* __restrict__ field, uint32_t m, float bg_x, float bg_y,
                          float dt, float spring_k, float max_speed_factor) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const uint32_t b = ebody[k];
    const float radius = pqr[b].w;
    const float2 f = field[k];
    const float lx = __fadd_rn(f.x, bg_x), ly = __fadd_rn(f.y, bg_y);
    const float accx = __fmul_rn(-lx, spring_k), accy = __fmul_rn(-ly, spring_k);
    float2 v = evel[k];
    v.x = __fadd_rn(v.x, __fmul_rn(accx, dt));
    v.y = __fadd_rn(v.y, __fmul_rn(accy, dt));
    const float speed = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
    const float max_speed = __fdiv_rn(__fmul_rn(max_speed_factor, radius), dt);
    if (speed > max_speed) {
      v.x = __fmul_rn(__fdiv_rn(v.x, speed), max_speed);
      v.y = __fmul_rn(__fdiv_rn(v.y, speed), max_speed);
    }
    float2 r = erel[k];
    r.x = __fadd_rn(r.x, __fmul_rn(v.x, dt));
    r.y = __fadd_rn(r.y, __fmul_rn(v.y, dt));
    const float max_dist = __fmul_rn(table[species[b] < kMaxSpecies ? species[b] : 0].polar_offset, radius);
    const float rm = __fsqrt_rn(__fadd_rn(__fmul_rn(r.x, r.x), __fmul_rn(r.y, r.y)));
    if (rm > max_dist) {
      const float inv = __fdiv_rn(1.0f, rm);  // Vec2::normalized: multiply by 1/mag
      r.x = __fmul_rn(__fmul_rn(r.x, inv), max_dist);
      r.y = __fmul_rn(__fmul_rn(r.y, inv), max_dist);
    }
    evel[k] = v;
    erel[k] = r;
  }
}

}  // namespace psim

// tree.cuh — quadtree construction on sorted quadrant keys, bottom-up aggregation, export.
//
// Replaces src/quadtree/quadtree.rs:153-348 (build / build_with_domain / build_internal),
// :40-101 (subdivide) and :103-151 (propagate) of the reference.
//
// Device tree ("compact pre-order tree"): only non-empty cells are stored, in DFS pre-order, so
//   first child of an internal node n  = n + 1
//   skip pointer (reference `next`)     = NodeB.x  (index of the first node after n's subtree;
//                                                   num_nodes means "end", the reference's 0)
// The reference's empty children carry charge 0 and contribute exactly nothing to a traversal;
// they are materialised only by the export kernel, which emits reference-shaped nodes (4
// contiguous children, `next`, Quad) for psim_download_nodes.
//
// Construction (Karras-style: every node is derived independently from the sorted keys):
//   λ_i  = quadrant levels body i shares with body i-1           (tree_count_kernel)
//   ℓ_i  = depth of the leaf cell holding body i                  (psim_core.cuh: leaf_depth)
//   body i starts a leaf iff λ_i < ℓ_i; it then also starts the cells at depths λ_i+1 .. ℓ_i-1,
//   which are exactly the internal nodes whose first body is i.  An exclusive scan of
//   (ℓ_i - λ_i) over leaf heads gives every node its pre-order index; each node finds the end of
//   its body range by a galloping search on the keys, and its skip pointer is the pre-order base
//   of the body that follows the range.
// The per-body / per-node logic lives in tree_logic.cuh; the kernels here are the parallel drivers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "psim_core.cuh"
#include "sort.cuh"
#include "tree_logic.cuh"

namespace psim {

// ------------------------------------------------------------------------------------------------
// root quad: Quad::new_containing (quad.rs:11-35) as a two-stage min/max reduction
__global__ void __launch_bounds__(256)
    bounds_partial_kernel(const float4* __restrict__ pqr, uint32_t n, float4* __restrict__ partial) {
  float mnx = 3.402823466e+38f, mny = 3.402823466e+38f, mxx = -3.402823466e+38f, mxy = -3.402823466e+38f;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 p = pqr[i];
    mnx = fminf(mnx, p.x);  // fminf/fmaxf ignore a NaN operand, like f32::min/max
    mny = fminf(mny, p.y);
    mxx = fmaxf(mxx, p.x);
    mxy = fmaxf(mxy, p.y);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, off));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, off));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, off));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, off));
  }
  __shared__ float4 s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = make_float4(mnx, mny, mxx, mxy);
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 r = s[0];
    for (int w = 1; w < 8; ++w) {
      r.x = fminf(r.x, s[w].x);
      r.y = fminf(r.y, s[w].y);
      r.z = fmaxf(r.z, s[w].z);
      r.w = fmaxf(r.w, s[w].w);
    }
    partial[blockIdx.x] = r;
  }
}

// mode 0: finish the reduction; mode 1: Quad::new_for_domain (quad.rs:38-43).  Resets the meta.
__global__ void root_quad_kernel(const float4* __restrict__ partial, int nblocks, int mode, float hw,
                                 float hh, uint32_t n, TreeMeta* __restrict__ meta) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  RootQuad r;
  if (mode == 0) {
    float4 b = partial[0];
    for (int i = 1; i < nblocks; ++i) {
      b.x = fminf(b.x, partial[i].x);
      b.y = fminf(b.y, partial[i].y);
      b.z = fmaxf(b.z, partial[i].z);
      b.w = fmaxf(b.w, partial[i].w);
    }
    r = root_from_bounds(b.x, b.y, b.z, b.w);
  } else {
    r = root_for_domain(hw, hh);
  }
  meta_reset(meta, r, n);
}

__global__ void __launch_bounds__(256)
    keygen_kernel(const float4* __restrict__ pqr, uint32_t n, const TreeMeta* __restrict__ meta,
                  uint64_t* __restrict__ keys, uint32_t* __restrict__ khi, uint32_t* __restrict__ idx) {
  const RootQuad r = meta->root;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 p = pqr[i];
    const uint64_t k = morton_key(p.x, p.y, r);
    keys[i] = k;
    khi[i] = (uint32_t)(k >> 32);  // what the radix passes sort (sort.cuh, two-tier key sort)
    idx[i] = i;
  }
}

// ------------------------------------------------------------------------------------------------
// per body: λ_i and ℓ_i (packed), the maximum depth, and per-depth counts of the internal cells the level
// sweeps will visit: those that straddle a boundary of the emit kernel's slabs of `per_block` bodies (a
// cell that lies inside one slab is summed by that slab's CTA, see tree_emit_kernel)
__global__ void __launch_bounds__(256)
    tree_count_kernel(const uint64_t* __restrict__ keys0, const uint64_t* __restrict__ keys1,
                      const SortPlan* __restrict__ plan, int npass, const float4* __restrict__ pqr,
                      uint32_t n, uint32_t c_eff, uint32_t per_block, TreeMeta* __restrict__ meta,
                      uint16_t* __restrict__ le, uint32_t leaf_capacity, uint32_t thread_capacity) {
  const uint64_t* __restrict__ keys = plan->src[npass] ? keys1 : keys0;
  __shared__ uint32_t s_cnt[kLevels];
  __shared__ uint32_t s_maxd, s_internal, s_zero_agg;
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_maxd = 0, s_internal = 0, s_zero_agg = 0;
  __syncthreads();
  const int dcap = (int)meta->dcap;
  // a leaf of at least this many bodies is not aggregated (leaf_is_aggregated)
  const uint64_t min_zero = (uint64_t)leaf_capacity + 1 < (uint64_t)thread_capacity ? (uint64_t)leaf_capacity + 1
                                                                                    : (uint64_t)thread_capacity;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint16_t lev = body_levels(keys, pqr, n, i, c_eff, dcap);
    le[i] = lev;
    const int lam = le_lambda(lev), ell = le_ell(lev);
    if (lam < ell) {
      if (ell - lam > 1) {
        atomicAdd(&s_internal, (uint32_t)(ell - lam - 1));
        const uint64_t slab_end = ((uint64_t)(i / per_block) + 1) * per_block;
        const int straddle = slab_end < n ? lcp_levels(keys[i], keys[slab_end]) : -1;
        const int top = (ell - 1 < straddle) ? ell - 1 : straddle;
        for (int d = lam + 1; d <= top; ++d) atomicAdd(&s_cnt[d], 1u);
      }
      atomicMax(&s_maxd, (uint32_t)ell);
      if (min_zero <= 1 || ((uint64_t)i + min_zero - 1 < n && lcp_levels(keys[i], keys[i + (uint32_t)min_zero - 1]) >= ell))
        s_zero_agg = 1u;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_zero_agg) meta->zero_agg_hint = 1u;
  if (threadIdx.x < kLevels && s_cnt[threadIdx.x]) atomicAdd(&meta->level_count[threadIdx.x], s_cnt[threadIdx.x]);
  if (threadIdx.x == 0 && s_maxd) atomicMax(&meta->max_depth, s_maxd);
  if (threadIdx.x == 0 && s_internal) atomicAdd(&meta->internal_total, s_internal);
}

struct LeCountFn {
  const uint16_t* le;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return le_nodes(le[i]); }
};

__global__ void level_scan_kernel(TreeMeta* __restrict__ meta, uint32_t node_cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  level_scan(meta, node_cap);
  meta->num_internal = meta->internal_total;  // the buckets only hold the cells the level sweeps visit
}

// level-bucket slots.  A CTA first counts, from the packed (λ, ℓ) bytes alone, how many internal nodes
// of each depth its bodies start, reserves one contiguous range per depth with a single global atomic,
// and then hands slots out of shared-memory counters while it emits.
struct StrictEmit {        // psim_config.strict_centres (strict.cuh); direct == 0: off
  uint32_t direct;         // nodes of at most this many bodies are summed by finalize_node
  const uint32_t* cidx;    // see StrictDirect
  const float4* cw;
  uint4* cand;             // chains with larger nodes: {head body, shallowest node, big nodes - 1, bodies of the largest}
  uint32_t* cand_count;
  uint32_t cand_cap;
  // Integer charges (what Body::update_charge_from_electrons produces): while every partial sum of the reference's
  // nested f32 additions ((c0 + c1) + c2) + c3 (quadtree.rs:142-149) is an integer below 2^24 it is exact, so a node's
  // charge is the difference of an integer prefix over its charged bodies, whatever the order - no bottom-up sweep.
  const unsigned long long* qstat;  // strict_addends_kernel; null: never
  const uint32_t* qc;               // qc[k] = sum of the charges of the first k charged bodies (mod 2^32)
};
__device__ __forceinline__ bool integer_charges(const unsigned long long* qstat, const TreeMeta* meta) {
  return qstat && qstat[1] == 0ull && qstat[0] < (1ull << 24) && meta->zero_agg_hint == 0u;
}

struct DeviceSink {
  static constexpr bool kTop = false;
  static constexpr bool kStrict = true;
  StrictEmit strict;
  __device__ __forceinline__ uint32_t strict_direct() const { return strict.direct ? strict.direct : 0xffffffffu; }
  __device__ __forceinline__ void strict_chain(uint32_t i, uint32_t base, uint32_t big, uint32_t cnt_top) {
    const uint32_t slot = atomicAdd(strict.cand_count, 1u);
    if (slot < strict.cand_cap) strict.cand[slot] = make_uint4(i, base, big - 1u, cnt_top);
  }
  __device__ __forceinline__ void top_leaf(int, uint64_t, uint32_t, const NodeRec&) {}
  __device__ __forceinline__ void top_internal(int, uint64_t, uint32_t) {}
  TreeMeta* meta;
  uint32_t* s_cursor;  // [kLevels] next free slot per depth (absolute index into level_nodes)
  uint32_t* s_local;   // [kLevels] next free slot per depth of the CTA's own list
  uint32_t* local_nodes;
  __device__ __forceinline__ uint32_t level_slot(int d) { return atomicAdd(&s_cursor[d], 1u); }
  __device__ __forceinline__ void local_node(int d, uint32_t node) { local_nodes[atomicAdd(&s_local[d], 1u)] = node; }
  __device__ __forceinline__ void zero_leaf() { atomicAdd(&meta->num_zero_leaves, 1u); }
  __device__ __forceinline__ void cap_leaf() { atomicAdd(&meta->num_cap_leaves, 1u); }
};

// Emit + in-CTA sums.  Each CTA owns a contiguous slab of bodies and emits the nodes that start in it.
// An internal cell that lies inside the slab (it does not contain the first body of the next slab) has
// all its descendants among the CTA's own nodes, so the CTA also runs the bottom-up sweep for those cells
// itself, level by level, right after writing the leaves - the records are still in L1 / L2 - and finishes
// their centres (finalize_node).  Only the cells that straddle a slab boundary (about depth x 2 per CTA)
// are left to the level sweeps.
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
    tree_emit_kernel(const uint64_t* __restrict__ keys0, const uint64_t* __restrict__ keys1,
                     const SortPlan* __restrict__ plan, int npass, uint32_t n, uint32_t per_block,
                     const uint16_t* __restrict__ le, const uint32_t* __restrict__ nodebase,
                     const float4* __restrict__ pqr, const float4* __restrict__ accm,
                     uint32_t leaf_capacity, uint32_t thread_capacity, TreeMeta* __restrict__ meta,
                     TreeArrays t, StrictEmit strict) {
  const uint64_t* __restrict__ keys = plan->src[npass] ? keys1 : keys0;
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;  // arena overflow, flagged by level_scan_kernel
  __shared__ uint32_t s_cnt[kLevels], s_cursor[kLevels];
  __shared__ uint32_t s_in[kLevels], s_in_off[kLevels + 1], s_local[kLevels];
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0, s_in[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * per_block;
  if (lo >= n) return;
  const uint32_t hi = (lo + per_block < n) ? lo + per_block : n;
  const bool has_next = hi < n;
  const uint64_t key_hi = has_next ? keys[hi] : 0ull;
  // integer charges: every internal cell is finished where it is emitted (no lists, no sweep, no NodeRec)
  const bool fastq = strict.direct != 0 && integer_charges(strict.qstat, meta);
  for (uint32_t i = lo + threadIdx.x; !fastq && i < hi; i += blockDim.x) {
    const uint16_t lev = le[i];
    const int lam = le_lambda(lev), ell = le_ell(lev);
    if (ell - lam > 1) {
      const int straddle = has_next ? lcp_levels(keys[i], key_hi) : -1;
      for (int d = lam + 1; d < ell; ++d) atomicAdd(d <= straddle ? &s_cnt[d] : &s_in[d], 1u);
    }
  }
  __syncthreads();
  const uint32_t local_base = nodebase[lo];  // the CTA's list lives where its own nodes' indices start
  if (threadIdx.x < kLevels) {
    const uint32_t c = s_cnt[threadIdx.x];
    s_cursor[threadIdx.x] =
        meta->level_start[threadIdx.x] + (c ? atomicAdd(&meta->level_cursor[threadIdx.x], c) : 0u);
  }
  if (threadIdx.x == 0) {
    uint32_t run = local_base;
    for (int l = 0; l < kLevels; ++l) s_in_off[l] = run, s_local[l] = run, run += s_in[l];
    s_in_off[kLevels] = run;
  }
  __syncthreads();
  const float root_size = meta->root.size;
  const int dcap = (int)meta->dcap;
  DeviceSink sink{strict, meta, s_cursor, s_local, t.local_nodes};
  // Leaves: one thread per body.  Internal cells: the chains of a warp's 32 bodies (0 .. 31 cells each, ~0.7 on
  // average) are dealt out to its lanes one cell at a time, so that a body at a large cell boundary does not keep 31
  // lanes waiting; every cell finds its own range end by a galloping search that starts at its leaf's end.
  __shared__ uint32_t s_big[4][32], s_top[4][32];
  const uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const uint32_t strict_limit = sink.strict_direct();
  for (uint32_t i0 = lo; i0 < hi; i0 += blockDim.x) {  // trip count uniform over the CTA
    const uint32_t i = i0 + threadIdx.x;
    const bool valid = i < hi;
    const uint16_t lev = valid ? le[i] : (uint16_t)0;
    const int lam = le_lambda(lev), ell = le_ell(lev);
    const bool head = valid && lam < ell;
    const int straddle = (head && has_next) ? lcp_levels(keys[i], key_hi) : -1;
    uint32_t base = 0, jleaf = 0;
    if (head) {
      base = nodebase[i];
      jleaf = emit_leaf_for_body(keys, n, i, lam, ell, base, nodebase, M, pqr, leaf_capacity, thread_capacity, root_size,
                                 dcap, t, sink, !fastq);
    }
    const int len = head ? ell - lam - 1 : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int off = incl - len, total = __shfl_sync(FULL, incl, 31);
    if (total == 0) continue;  // warp-uniform
    s_big[wrp][lane] = 0;
    __syncwarp();
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int tt = t0 + lane;
      // owner = the last lane whose exclusive offset is <= tt (lanes without cells share their successor's offset)
      int o = 0;
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        const int cand = o + sft;
        const int v = __shfl_sync(FULL, off, cand & 31);
        if (cand < 32 && v <= tt) o = cand;
      }
      const int o_off = __shfl_sync(FULL, off, o);
      const uint32_t o_lev = __shfl_sync(FULL, (uint32_t)lev, o);
      const uint32_t o_base = __shfl_sync(FULL, base, o);
      const uint32_t o_jleaf = __shfl_sync(FULL, jleaf, o);
      const int o_straddle = __shfl_sync(FULL, straddle, o);
      if (tt < total) {
        const uint32_t oi = i - (uint32_t)lane + (uint32_t)o;
        const int o_lam = le_lambda((uint16_t)o_lev), o_ell = le_ell((uint16_t)o_lev);
        const int d = o_ell - 1 - (tt - o_off);
        const uint32_t node = o_base + (uint32_t)(d - o_lam - 1);
        const uint32_t jd = run_end(keys, n, oi, o_jleaf, d);
        const uint32_t nx = (jd < n) ? nodebase[jd] : M;
        const uint32_t cnt = jd - oi;
        if (fastq) {
          const uint32_t c0 = strict.cidx[oi], c1 = strict.cidx[jd];
          const uint32_t chg = c1 > c0 ? 1u : 0u;
          const float q = (float)(int32_t)(strict.qc[c1] - strict.qc[c0]);
          float px = 0.0f, py = 0.0f;
          if (cnt <= strict_limit && chg) {  // the reference's serial sums over the cell's charged bodies (finalize_node)
            float total_abs = 0.0f, wx = 0.0f, wy = 0.0f;
            for (uint32_t k = c0; k < c1; ++k) {
              const float4 w = strict.cw[k];
              total_abs = f_add(total_abs, w.x), wx = f_add(wx, w.y), wy = f_add(wy, w.z);
            }
            px = f_div(wx, total_abs), py = f_div(wy, total_abs);  // total_abs >= 1: integer charges
          }
          t.nodeA[node] = make_float4(px, py, q, ldexpf(root_size, -d));
          t.nodeB[node] = make_uint4(nx, oi, cnt, (uint32_t)d | (chg ? kNodeCharged : 0u));
          t.ndepth[node] = (uint8_t)((uint32_t)d | (chg ? kDepthCharged : 0u));
        } else {
          t.nodeB[node] = make_uint4(nx, oi, cnt, (uint32_t)d);
          t.ndepth[node] = (uint8_t)d;
          if (d <= o_straddle) t.level_nodes[sink.level_slot(d)] = node;
          else sink.local_node(d, node);
        }
        if (cnt > strict_limit) atomicAdd(&s_big[wrp][o], 1u);
        if (d == o_lam + 1) s_top[wrp][o] = cnt;
      }
    }
    __syncwarp();
    if (head && len > 0) {
      const uint32_t big = s_big[wrp][lane];
      if (big) sink.strict_chain(i, base, big, s_top[wrp][lane]);
    }
    __syncwarp();
  }
  // the slab's own cells, deepest level first
  for (int level = kLevels - 1; !fastq && level >= 0; --level) {
    const uint32_t begin = s_in_off[level], end = s_in_off[level + 1];
    if (begin == end) continue;  // uniform over the CTA
    __threadfence_block();
    __syncthreads();
    for (uint32_t k = begin + threadIdx.x; k < end; k += blockDim.x) {
      const uint32_t node = t.local_nodes[k];
      aggregate_node_ranged(node, level, t);
      finalize_node(node, root_size, pqr, accm, t, SubtreeEndCount{}, StrictDirect{strict.direct, strict.cidx, strict.cw});
    }
  }
}

// one level of the build's bottom-up sweep (deepest level first) over the cells that straddle slab
// boundaries; their centres are finished in the same visit
__global__ void __launch_bounds__(128)
    aggregate_level_kernel(int level, const TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                           const float4* __restrict__ accm, TreeArrays t, StrictDirect strict_direct) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const uint32_t begin = meta->level_start[level], end = meta->level_start[level + 1];
  const float root_size = meta->root.size;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += stride) {
    const uint32_t node = t.level_nodes[k];
    aggregate_node_ranged(node, level, t);
    finalize_node(node, root_size, pqr, accm, t, SubtreeEndCount{}, strict_direct);
  }
}

// The same sweep as ONE launch of one CTA per SM (all resident), with a grid-wide barrier between the levels: the
// straddling cells are ~2 x depth per slab, far too few per level to amortise 32 launch latencies each build.
// `barrier` is a zeroed counter; empty levels are skipped by every CTA alike.
__global__ void __launch_bounds__(128)
    aggregate_levels_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                            const float4* __restrict__ accm, TreeArrays t, StrictDirect strict_direct,
                            unsigned int* __restrict__ barrier, const unsigned long long* __restrict__ qstat) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  if (strict_direct.limit && integer_charges(qstat, meta)) return;  // the emit kernel has finished every cell
  const float root_size = meta->root.size;
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned int target = 0;
  for (int level = kMaxLevels - 1; level >= 0; --level) {
    const uint32_t begin = meta->level_start[level], end = meta->level_start[level + 1];
    if (begin == end) continue;
    for (uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += stride) {
      const uint32_t node = t.level_nodes[k];
      aggregate_node_ranged(node, level, t);
      finalize_node(node, root_size, pqr, accm, t, SubtreeEndCount{}, strict_direct);
    }
    // grid barrier: the next (shallower) level reads this level's records
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      __threadfence();
      atomicAdd(barrier, 1u);
      while (*reinterpret_cast<volatile unsigned int*>(barrier) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  }
}

// psim_download_nodes only: the export sweep visits EVERY internal node level by level, so the buckets are
// rebuilt to hold them all (the build only buckets the cells its level sweeps visit)
__global__ void __launch_bounds__(256)
    export_count_levels_kernel(TreeMeta* __restrict__ meta, const uint4* __restrict__ nodeB) {
  const uint32_t M = meta->num_nodes;
  __shared__ uint32_t s_cnt[kLevels];
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < M; node += stride) {
    const uint32_t w = nodeB[node].w;
    if (!(w & kNodeLeaf)) atomicAdd(&s_cnt[w & kNodeDepthMask], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kLevels && s_cnt[threadIdx.x]) atomicAdd(&meta->level_count[threadIdx.x], s_cnt[threadIdx.x]);
}
__global__ void export_reset_levels_kernel(TreeMeta* __restrict__ meta, int stage, uint32_t node_cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (stage == 0) {
    for (int l = 0; l < kLevels; ++l) meta->level_count[l] = 0, meta->level_cursor[l] = 0;
  } else {
    const uint32_t keep = meta->num_internal;
    level_scan(meta, node_cap);
    meta->num_internal = keep;
  }
}
__global__ void __launch_bounds__(256)
    export_fill_levels_kernel(TreeMeta* __restrict__ meta, TreeArrays t) {
  const uint32_t M = meta->num_nodes;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < M; node += stride) {
    const uint32_t w = t.nodeB[node].w;
    if (w & kNodeLeaf) continue;
    const uint32_t d = w & kNodeDepthMask;
    t.level_nodes[meta->level_start[d] + atomicAdd(&meta->level_cursor[d], 1u)] = node;
  }
}

// one level of the export sweep (psim_download_nodes only)
__global__ void __launch_bounds__(128)
    export_level_kernel(int level, const TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                        const float4* __restrict__ accm, TreeArrays t, bool write_chargeless_centres) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const uint32_t begin = meta->level_start[level], end = meta->level_start[level + 1];
  const float root_size = meta->root.size;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += stride)
    aggregate_node(t.level_nodes[k], root_size, pqr, accm, t, write_chargeless_centres);
}

// leaf masses for a tree that is a single leaf (no internal node ever visits it)
__global__ void export_root_leaf_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ accm,
                                        TreeArrays t) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (meta->num_nodes == 0 || meta->num_nodes > t.node_cap) return;
  const uint4 nb = t.nodeB[0];
  if (!(nb.w & kNodeLeaf)) return;
  float lm = 0.0f;
  if (!(nb.w & kNodeZeroAgg))
    for (uint32_t b = nb.y; b < nb.y + nb.z; ++b) lm = f_add(lm, accm[b].w);
  t.node_mass[0] = lm;
  t.parent[0] = 0xffffffffu;
}

// Traversal arrays: the nodes that can contribute to a field sum, i.e. those with a charged body
// below them, compacted in pre-order (skip pointers remapped).  A node without charge adds exactly
// +-0 to every acc_pos sum whatever the opening test says, so the reference's result is unchanged.
struct ChargedFlagFn {
  const uint8_t* ndepth;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return ndepth[i] >> 7; }
};

__global__ void __launch_bounds__(256)
    compact_traversal_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                             const uint4* __restrict__ nodeB, const uint32_t* __restrict__ rank,
                             uint32_t* __restrict__ total, uint32_t node_cap,
                             float4* __restrict__ travA, uint4* __restrict__ travB) {
  const uint32_t M = meta->num_nodes;
  if (M > node_cap) {
    // arena overflow (flagged in meta->err): leave an empty traversal tree so that a chained field
    // pass walks nothing instead of walking stale records
    if (blockIdx.x == 0 && threadIdx.x == 0) *total = 0;
    return;
  }
  const uint32_t T = *total;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += stride) {
    uint4 nb = nodeB[n];
    if (!(nb.w & kNodeCharged)) continue;
    const uint32_t r = rank[n];
    nb.x = nb.x < M ? rank[nb.x] : T;
    travA[r] = nodeA[n];
    travB[r] = nb;
  }
}

// Children of every internal traversal node, so that the walk finds them without chasing skip pointers
// through dependent loads: child 0 is r + 1, children 1..3 go into B.y, B.z and the bits of A.w (body
// range and cell size of an internal node are not read by the walk; the size is recomputed from the depth).
// A missing child is marked by the node's own skip pointer.  Only .y / .z / A.w are written, .x is only read.
__global__ void __launch_bounds__(256)
    link_children_kernel(const uint32_t* __restrict__ total, float4* __restrict__ travA, uint4* __restrict__ travB) {
  const uint32_t T = *total;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < T; r += stride) {
    const uint32_t w = travB[r].w;
    if (w & kNodeLeaf) continue;
    const uint32_t skip = travB[r].x;
    uint32_t c1 = skip, c2 = skip, c3 = skip;
    if (r + 1 < T) c1 = travB[r + 1].x;
    if (c1 != skip && c1 < T) c2 = travB[c1].x;
    if (c2 != skip && c2 < T) c3 = travB[c2].x;
    travB[r].y = c1;
    travB[r].z = c2;
    travA[r].w = __uint_as_float(c3);
  }
}

struct InternalFlagFn {
  const uint4* nodeB;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
    return (nodeB[i].w & kNodeLeaf) ? 0u : 1u;
  }
};

__global__ void __launch_bounds__(128)
    export_nodes_kernel(const uint64_t* __restrict__ keys0, const uint64_t* __restrict__ keys1,
                        const SortPlan* __restrict__ plan, int npass,
                        const TreeMeta* __restrict__ meta, TreeArrays t,
                        const uint32_t* __restrict__ irank, PsimNodeOut* __restrict__ out,
                        uint64_t out_cap) {
  const uint64_t* __restrict__ keys = plan->src[npass] ? keys1 : keys0;
  const uint32_t M = meta->num_nodes;
  const RootQuad root = meta->root;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < M; node += stride)
    export_node(node, keys, root, t, irank, out, out_cap);
}

}  // namespace psim

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
// per body: λ_i and ℓ_i (packed), plus per-depth internal-node counts and the maximum depth
__global__ void __launch_bounds__(256)
    tree_count_kernel(const uint64_t* __restrict__ keys0, const uint64_t* __restrict__ keys1,
                      const SortPlan* __restrict__ plan, int npass, const float4* __restrict__ pqr,
                      uint32_t n, uint32_t c_eff, TreeMeta* __restrict__ meta,
                      uint16_t* __restrict__ le) {
  const uint64_t* __restrict__ keys = plan->src[npass] ? keys1 : keys0;
  __shared__ uint32_t s_cnt[kLevels];
  __shared__ uint32_t s_maxd;
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_maxd = 0;
  __syncthreads();
  const int dcap = (int)meta->dcap;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint16_t lev = body_levels(keys, pqr, n, i, c_eff, dcap);
    le[i] = lev;
    const int lam = le_lambda(lev), ell = le_ell(lev);
    if (lam < ell) {
      for (int d = lam + 1; d < ell; ++d) atomicAdd(&s_cnt[d], 1u);
      atomicMax(&s_maxd, (uint32_t)ell);
    }
  }
  __syncthreads();
  if (threadIdx.x < kLevels && s_cnt[threadIdx.x]) atomicAdd(&meta->level_count[threadIdx.x], s_cnt[threadIdx.x]);
  if (threadIdx.x == 0 && s_maxd) atomicMax(&meta->max_depth, s_maxd);
}

struct LeCountFn {
  const uint16_t* le;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return le_nodes(le[i]); }
};

__global__ void level_scan_kernel(TreeMeta* __restrict__ meta, uint32_t node_cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  level_scan(meta, node_cap);
}

// level-bucket slots.  A CTA first counts, from the packed (λ, ℓ) bytes alone, how many internal nodes
// of each depth its bodies start, reserves one contiguous range per depth with a single global atomic,
// and then hands slots out of shared-memory counters while it emits.
struct DeviceSink {
  static constexpr bool kTop = false;
  __device__ __forceinline__ void top_leaf(int, uint64_t, uint32_t, const NodeRec&) {}
  __device__ __forceinline__ void top_internal(int, uint64_t, uint32_t) {}
  TreeMeta* meta;
  uint32_t* s_cursor;  // [kLevels] next free slot per depth (absolute index into level_nodes)
  __device__ __forceinline__ uint32_t level_slot(int d) { return atomicAdd(&s_cursor[d], 1u); }
  __device__ __forceinline__ void zero_leaf() { atomicAdd(&meta->num_zero_leaves, 1u); }
  __device__ __forceinline__ void cap_leaf() { atomicAdd(&meta->num_cap_leaves, 1u); }
};

__global__ void __launch_bounds__(128)
    tree_emit_kernel(const uint64_t* __restrict__ keys0, const uint64_t* __restrict__ keys1,
                     const SortPlan* __restrict__ plan, int npass, uint32_t n,
                     const uint16_t* __restrict__ le, const uint32_t* __restrict__ nodebase,
                     const float4* __restrict__ pqr, const float4* __restrict__ accm,
                     uint32_t leaf_capacity, uint32_t thread_capacity, TreeMeta* __restrict__ meta,
                     TreeArrays t) {
  const uint64_t* __restrict__ keys = plan->src[npass] ? keys1 : keys0;
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;  // arena overflow, flagged by level_scan_kernel
  __shared__ uint32_t s_cnt[kLevels];
  __shared__ uint32_t s_cursor[kLevels];
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  // each CTA owns a contiguous slab of bodies (keeps its nodes, and the level buckets, local)
  const uint32_t per_block = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = blockIdx.x * per_block;
  const uint32_t hi = (lo + per_block < n) ? lo + per_block : n;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint16_t lev = le[i];
    const int lam = le_lambda(lev), ell = le_ell(lev);
    for (int d = lam + 1; d < ell; ++d) atomicAdd(&s_cnt[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kLevels) {
    const uint32_t c = s_cnt[threadIdx.x];
    s_cursor[threadIdx.x] =
        meta->level_start[threadIdx.x] + (c ? atomicAdd(&meta->level_cursor[threadIdx.x], c) : 0u);
  }
  __syncthreads();
  const float root_size = meta->root.size;
  const int dcap = (int)meta->dcap;
  DeviceSink sink{meta, s_cursor};
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
    emit_nodes_for_body(keys, n, i, le[i], nodebase, M, pqr, accm, leaf_capacity, thread_capacity,
                        root_size, dcap, t, sink);
}

// one level of the build's bottom-up sweep (deepest level first)
__global__ void __launch_bounds__(128)
    aggregate_level_kernel(int level, const TreeMeta* __restrict__ meta, TreeArrays t) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const uint32_t begin = meta->level_start[level], end = meta->level_start[level + 1];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += stride)
    aggregate_node_lean(t.level_nodes[k], level, M, t);
}

__global__ void __launch_bounds__(256)
    finalize_nodes_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                          const float4* __restrict__ accm, TreeArrays t) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const float root_size = meta->root.size;
  const uint32_t n_bodies = meta->n;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < M; node += stride)
    finalize_node(node, root_size, pqr, accm, t, SubtreeEndLocal{M, n_bodies, t.nodeB});
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

// psim_config.strict_centres: the reference's own arithmetic for internal centres - three serial f32
// running sums over the node's body range in sorted order (quadtree.rs:114-139).  One thread per node.
__global__ void __launch_bounds__(128)
    strict_centres_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                          const float4* __restrict__ accm, TreeArrays t) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const uint32_t count = meta->num_internal, n_bodies = meta->n;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint32_t node = t.level_nodes[k];
    const uint4 nb = t.nodeB[node];
    const uint32_t b0 = nb.y, b1 = nb.x < M ? t.nodeB[nb.x].y : n_bodies;
    float total_mass = 0.0f, total_abs = 0.0f;
    for (uint32_t b = b0; b < b1; ++b) total_mass = f_add(total_mass, accm[b].w);
    for (uint32_t b = b0; b < b1; ++b) total_abs = f_add(total_abs, fabsf(pqr[b].z));
    float wx = 0.0f, wy = 0.0f;
    if (total_abs > 1e-6f) {
      for (uint32_t b = b0; b < b1; ++b) {
        const float4 p = pqr[b];
        wx = f_add(wx, f_mul(p.x, fabsf(p.z))), wy = f_add(wy, f_mul(p.y, fabsf(p.z)));
      }
      wx = f_div(wx, total_abs), wy = f_div(wy, total_abs);
    } else if (total_mass > 1e-6f) {
      for (uint32_t b = b0; b < b1; ++b) {
        const float4 p = pqr[b];
        const float m = accm[b].w;
        wx = f_add(wx, f_mul(p.x, m)), wy = f_add(wy, f_mul(p.y, m));
      }
      wx = f_div(wx, total_mass), wy = f_div(wy, total_mass);
    } else if (b1 > b0) {
      for (uint32_t b = b0; b < b1; ++b) wx = f_add(wx, pqr[b].x), wy = f_add(wy, pqr[b].y);
      wx = f_div(wx, (float)(b1 - b0)), wy = f_div(wy, (float)(b1 - b0));
    }
    float4 a = t.nodeA[node];
    a.x = wx, a.y = wy;
    t.nodeA[node] = a;
  }
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

// shard.cuh — Morton-range sharded quadtree build (multi-GPU).
//
// Replaces, across G ranks, the single work queue of Quadtree::build_internal
// (src/quadtree/quadtree.rs:197-345): every rank owns a CONTIGUOUS RANGE OF THE KEY ORDER and builds
// the part of the tree whose cells start in it; the pieces concatenate into exactly the tree the
// single-GPU build makes (same nodes, same pre-order indices, same aggregates), so every downstream
// result is bit-identical.
//
// Shard bins.  The cells at depth kShardDepth = 8 ("bins", 65 536 of them: the top 16 key bits) are the
// unit of ownership: rank r owns the bins [bin_lo[r], bin_lo[r+1]) chosen from the bin histogram so that
// body counts balance.  Every rank computes the histogram of ALL bodies itself (positions are
// replicated), so the splitters need no communication.  Consequences:
//   * a cell deeper than a bin lies inside one bin, hence inside one rank: its subtree is local;
//   * a cell at depth <= 8 ends at a bin boundary, so anything a local node needs to know about a
//     remote neighbour is a per-bin quantity: the key prefix of the bodies beyond the range (virtual
//     halo keys carrying the exact top 16 bits), the pre-order index of the first node of a bin, the
//     traversal rank of that node - two 65 536-entry tables exchanged by all-reduce;
//   * the aggregates of the at most 21 845 cells above the bins are finished by every rank from the
//     all-reduced records of the complete cells below them (top heap: slot (4^d - 1) / 3 + prefix),
//     children in quadrant order, i.e. the reference's ((c0 + c1) + c2) + c3 (quadtree.rs:142-149).
// Exchanges per build (host side: particlesim_b200/parallel.py): sorted index segments (all-gather),
// table 1, top heap, table 2 (all-reduce, <= 3.5 MB), traversal node segments (all-gather).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "shard_logic.cuh"
#include "sort.cuh"
#include "strict.cuh"
#include "tree.cuh"

namespace psim {

// ---- replicated: shard bin (first 8 key levels) of all bodies + bin histogram ----------------------
__global__ void __launch_bounds__(256)
    bins_kernel(const float4* __restrict__ pqr, uint32_t n, const TreeMeta* __restrict__ meta,
                uint16_t* __restrict__ bins, uint32_t* __restrict__ binhist) {
  const RootQuad r = meta->root;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
    const uint32_t i = base + threadIdx.x;
    const bool live = i < n;
    uint32_t bin = 0xffffffffu;
    if (live) {
      const float4 p = pqr[i];
      bin = morton_prefix(p.x, p.y, r, kShardDepth);
      bins[i] = (uint16_t)bin;
    }
    // bodies arrive nearly sorted: one atomic per distinct bin in the warp
    const uint32_t m = __match_any_sync(0xffffffffu, bin);
    if (live && lane == (uint32_t)(__ffs(m) - 1)) atomicAdd(&binhist[bin], (uint32_t)__popc(m));
  }
}

// one CTA, 1024 threads: exclusive prefix of the histogram and the balanced bin splitters
__global__ void __launch_bounds__(1024)
    bin_split_kernel(const uint32_t* __restrict__ binhist, uint32_t n, uint32_t world,
                     uint32_t* __restrict__ binprefix /*[kBins + 1]*/, ShardPlan* __restrict__ plan) {
  __shared__ uint32_t s_part[1024];
  const int t = threadIdx.x;
  constexpr int per = kBins / 1024;
  uint32_t local[per];
  uint32_t sum = 0;
  for (int j = 0; j < per; ++j) local[j] = binhist[t * per + j], sum += local[j];
  s_part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t add = t >= off ? s_part[t - off] : 0u;
    __syncthreads();
    s_part[t] += add;
    __syncthreads();
  }
  uint32_t run = s_part[t] - sum;
  for (int j = 0; j < per; ++j) binprefix[t * per + j] = run, run += local[j];
  if (t == 1023) binprefix[kBins] = run;
  __syncthreads();
  __threadfence();
  if (t <= (int)world) {
    // bin_lo[r] = first bin at or after which at least r * n / world bodies precede
    const uint64_t target = (uint64_t)n * (uint32_t)t / world;
    uint32_t lo = 0, hi = kBins;  // first b with binprefix[b] >= target
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (binprefix[mid] >= target) hi = mid; else lo = mid + 1;
    }
    if (t == 0) lo = 0;
    if (t == (int)world) lo = kBins;
    plan->bin_lo[t] = lo;
    plan->body_lo[t] = binprefix[lo];
  }
}

struct InRangeFn {  // 1 for a body whose bin this rank owns
  const uint16_t* bins;
  uint32_t bin_lo, bin_hi;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
    const uint32_t b = bins[i];
    return (b >= bin_lo && b < bin_hi) ? 1u : 0u;
  }
};

// ordered compaction of the owned bodies with their full 32-level keys: upper key word + compact slot as
// the input of the radix passes, 64-bit key and global body index by slot
__global__ void __launch_bounds__(256)
    select_owned_kernel(const float4* __restrict__ pqr, const uint16_t* __restrict__ bins,
                        const uint32_t* __restrict__ off, uint32_t n, uint32_t bin_lo, uint32_t bin_hi,
                        const TreeMeta* __restrict__ meta, uint64_t* __restrict__ keys,
                        uint32_t* __restrict__ khi, uint32_t* __restrict__ slot, uint32_t* __restrict__ body) {
  const RootQuad r = meta->root;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t b = bins[i];
    if (b >= bin_lo && b < bin_hi) {
      const float4 p = pqr[i];
      const uint64_t k = morton_key(p.x, p.y, r);
      const uint32_t o = off[i];
      keys[o] = k;
      khi[o] = (uint32_t)(k >> 32);
      slot[o] = o;
      body[o] = i;
    }
  }
}

__global__ void __launch_bounds__(256)
    copy_sorted_idx_kernel(const uint32_t* __restrict__ idx0, const uint32_t* __restrict__ idx1,
                           const SortPlan* __restrict__ plan, int npass, uint32_t n,
                           const uint32_t* __restrict__ body, uint32_t* __restrict__ out) {
  const uint32_t* __restrict__ idx = plan->src[npass] ? idx1 : idx0;  // sorted compact slots
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = body[idx[i]];
}

// bin that holds the body at global sorted position g: binprefix[b] <= g < binprefix[b + 1]
__device__ __forceinline__ uint32_t bin_of_body(const uint32_t* __restrict__ binprefix, uint32_t g) {
  uint32_t lo = 0, hi = kBins;  // first b with binprefix[b] > g, minus one
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (binprefix[mid] > g) hi = mid; else lo = mid + 1;
  }
  return lo - 1;
}

// virtual halo keys (exact top 16 bits of the bodies beyond the range) and "starts no node" levels
__global__ void __launch_bounds__(256)
    halo_kernel(const uint32_t* __restrict__ binprefix, const ShardMeta* __restrict__ sm,
                uint64_t* __restrict__ lkeys, uint16_t* __restrict__ le) {
  const uint32_t hl = sm->hl, nl = sm->n_local, L = sm->L, base = sm->body_base;
  const uint32_t halo = L - nl;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < halo; t += gridDim.x * blockDim.x) {
    const uint32_t a = t < hl ? t : nl + t;  // slot in the local array
    lkeys[a] = (uint64_t)bin_of_body(binprefix, base + a) << 48;
    le[a] = 1;  // lambda 0, ell 0: no nodes
  }
}

// ---- per-rank tree pieces ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    tree_count_range_kernel(const uint64_t* __restrict__ keys, const float4* __restrict__ pqr, uint32_t L,
                            uint32_t first, uint32_t count, uint32_t c_eff, int min_depth,
                            TreeMeta* __restrict__ meta, uint16_t* __restrict__ le) {
  __shared__ uint32_t s_cnt[kLevels];
  __shared__ uint32_t s_maxd;
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_maxd = 0;
  __syncthreads();
  const int dcap = (int)meta->dcap;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint32_t i = first + k;
    const uint16_t lev = body_levels(keys, pqr, L, i, c_eff, dcap);
    le[i] = lev;
    const int lam = le_lambda(lev), ell = le_ell(lev);
    if (lam < ell) {
      for (int d = (lam + 1 > min_depth ? lam + 1 : min_depth); d < ell; ++d) atomicAdd(&s_cnt[d], 1u);
      atomicMax(&s_maxd, (uint32_t)ell);
    }
  }
  __syncthreads();
  if (threadIdx.x < kLevels && s_cnt[threadIdx.x]) atomicAdd(&meta->level_count[threadIdx.x], s_cnt[threadIdx.x]);
  if (threadIdx.x == 0 && s_maxd) atomicMax(&meta->max_depth, s_maxd);
}

// table 1: (local pre-order index of the first node of each owned, non-empty bin) + 1, and M_local
__global__ void __launch_bounds__(256)
    table_nodes_kernel(const uint32_t* __restrict__ binhist, const uint32_t* __restrict__ binprefix,
                       const ShardPlan* __restrict__ plan, ShardMeta* __restrict__ sm,
                       const uint32_t* __restrict__ nodebase, const uint32_t* __restrict__ m_local,
                       unsigned long long* __restrict__ xbuf) {
  const uint32_t r = sm->rank;
  const uint32_t b0 = plan->bin_lo[r], b1 = plan->bin_lo[r + 1];
  for (uint32_t b = b0 + blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += gridDim.x * blockDim.x)
    if (binhist[b]) xbuf[b] = (unsigned long long)nodebase[binprefix[b] - sm->body_base] + 1ull;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sm->M_local = *m_local;
    xbuf[kBins + r] = *m_local;
  }
}

// after the all-reduce: rank offsets, and per bin the GLOBAL index of the first node at or after it
// (tab[kBins] = total).  Empty bins take the value of the next non-empty one, so tab is monotone.
// which = 0: node table (also makes the local nodebase global), 1: traversal table.
__global__ void __launch_bounds__(256)
    resolve_table_kernel(const uint32_t* __restrict__ binhist, const uint32_t* __restrict__ binprefix,
                         const ShardPlan* __restrict__ plan, ShardMeta* __restrict__ sm,
                         const unsigned long long* __restrict__ xbuf, uint32_t n_total, int which,
                         uint32_t* __restrict__ tab) {
  __shared__ uint32_t s_lo[kMaxRanks + 1];
  const uint32_t world = sm->world;
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t r = 0; r < world; ++r) s_lo[r] = run, run += (uint32_t)xbuf[kBins + r];
    s_lo[world] = run;
  }
  __syncthreads();
  const uint32_t total = s_lo[world];
  if (blockIdx.x == 0 && threadIdx.x <= world) {
    if (which == 0) sm->node_lo[threadIdx.x] = s_lo[threadIdx.x];
    else sm->trav_lo[threadIdx.x] = s_lo[threadIdx.x];
    if (threadIdx.x == 0) {
      if (which == 0) sm->node_off = s_lo[sm->rank], sm->M_total = total;
      else sm->trav_off = s_lo[sm->rank], sm->T_total = total;
    }
  }
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b <= kBins; b += gridDim.x * blockDim.x) {
    uint32_t v = total;
    if (b < kBins) {
      const uint32_t g = binprefix[b];  // first body at or after bin b
      if (g < n_total) {
        const uint32_t bb = binhist[b] ? b : bin_of_body(binprefix, g);
        uint32_t owner = 0;
        while (owner + 1 < world && plan->bin_lo[owner + 1] <= bb) ++owner;
        v = s_lo[owner] + (uint32_t)(xbuf[bb] - 1ull);
      }
    }
    tab[b] = v;
  }
}

// nodebase of the local slice becomes global; the right halo's entries are the first nodes of their bins
__global__ void __launch_bounds__(256)
    globalize_nodebase_kernel(const ShardMeta* __restrict__ sm, const uint64_t* __restrict__ lkeys,
                              const uint32_t* __restrict__ nb_bin, uint32_t* __restrict__ nodebase,
                              uint8_t* __restrict__ ndepth_local) {
  const uint32_t hl = sm->hl, nl = sm->n_local, L = sm->L, off = sm->node_off;
  for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < L; a += gridDim.x * blockDim.x) {
    if (a < hl) continue;
    if (a < hl + nl) nodebase[a] += off;
    else nodebase[a] = nb_bin[(uint32_t)(lkeys[a] >> 48)];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && nl > 0 && hl + nl < L) {
    // depth of the first node after the local piece: it starts the cells below the level the first remote
    // body shares with the last local one
    ndepth_local[sm->M_local] = (uint8_t)(lcp_levels(lkeys[hl + nl - 1], lkeys[hl + nl]) + 1);
  }
}

struct ShardSink {
  static constexpr bool kTop = true;
  static constexpr bool kStrict = false;
  __device__ __forceinline__ uint32_t strict_direct() const { return 0xffffffffu; }
  __device__ __forceinline__ void strict_chain(uint32_t, uint32_t, uint32_t, uint32_t) {}
  TreeMeta* meta;
  uint32_t* s_cursor;
  TopRec* heap;
  __device__ __forceinline__ uint32_t level_slot(int d) { return atomicAdd(&s_cursor[d], 1u); }
  __device__ __forceinline__ void local_node(int, uint32_t) {}
  __device__ __forceinline__ void zero_leaf() { atomicAdd(&meta->num_zero_leaves, 1u); }
  __device__ __forceinline__ void cap_leaf() { atomicAdd(&meta->num_cap_leaves, 1u); }
  __device__ __forceinline__ void top_leaf(int d, uint64_t key, uint32_t node, const NodeRec& r) {
    TopRec t;
    t.r = r, t.node = node, t.state = kTopComplete;
    heap[top_slot(d, key)] = t;
  }
  __device__ __forceinline__ void top_internal(int d, uint64_t key, uint32_t node) {
    TopRec* t = &heap[top_slot(d, key)];
    t->node = node;
    t->state = d == kShardDepth ? kTopBin : kTopInternal;
  }
};

// t's node arrays are pre-offset by -node_off (global node indices address the local arrays)
__global__ void __launch_bounds__(128)
    tree_emit_range_kernel(const uint64_t* __restrict__ keys, const ShardMeta* __restrict__ sm,
                           const uint16_t* __restrict__ le, const uint32_t* __restrict__ nodebase,
                           const float4* __restrict__ pqr, uint32_t leaf_capacity, uint32_t thread_capacity,
                           TreeMeta* __restrict__ meta, TreeArrays t, TopRec* __restrict__ heap) {
  const uint32_t M = sm->M_total;
  if (sm->M_local > t.node_cap) return;
  const uint32_t first = sm->hl, count = sm->n_local, L = sm->L, body_base = sm->body_base;
  const uint32_t noff = sm->node_off;
  t.nodeA -= noff, t.nodeB -= noff, t.rec -= noff, t.ndepth -= noff;
  __shared__ uint32_t s_cnt[kLevels];
  __shared__ uint32_t s_cursor[kLevels];
  if (threadIdx.x < kLevels) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t per_block = (count + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = first + blockIdx.x * per_block;
  const uint32_t hi = (lo + per_block < first + count) ? lo + per_block : first + count;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint16_t lev = le[i];
    const int lam = le_lambda(lev), ell = le_ell(lev);
    for (int d = (lam + 1 > kShardDepth ? lam + 1 : kShardDepth); d < ell; ++d) atomicAdd(&s_cnt[d], 1u);
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
  ShardSink sink{meta, s_cursor, heap};
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
    emit_nodes_for_body(keys, L, i, le[i], nodebase, M, pqr, (const float4*)nullptr, leaf_capacity,
                        thread_capacity, root_size, dcap, t, sink, body_base, kShardDepth);
}

__global__ void level_scan_shard_kernel(TreeMeta* __restrict__ meta, const ShardMeta* __restrict__ sm,
                                        uint32_t node_cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  meta->num_nodes = sm->M_local;
  level_scan(meta, node_cap);
}

// one level of the local bottom-up sweep (levels >= kShardDepth: whole subtrees are local)
__global__ void __launch_bounds__(128)
    aggregate_level_shard_kernel(int level, const TreeMeta* __restrict__ meta, const ShardMeta* __restrict__ sm,
                                 TreeArrays t) {
  if (sm->M_local > t.node_cap) return;
  const uint32_t noff = sm->node_off, M = sm->M_total;
  t.nodeA -= noff, t.nodeB -= noff, t.rec -= noff, t.ndepth -= noff;
  const uint32_t begin = meta->level_start[level], end = meta->level_start[level + 1];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += stride)
    aggregate_node_lean(t.level_nodes[k], level, M, t);
}

// the records of the owned internal bins go into the top heap once their subtrees are summed
__global__ void __launch_bounds__(256)
    heap_bins_kernel(const ShardPlan* __restrict__ plan, const ShardMeta* __restrict__ sm, TreeArrays t,
                     TopRec* __restrict__ heap) {
  if (sm->M_local > t.node_cap) return;
  const uint32_t r = sm->rank, noff = sm->node_off;
  const uint32_t b0 = plan->bin_lo[r], b1 = plan->bin_lo[r + 1];
  for (uint32_t b = b0 + blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += gridDim.x * blockDim.x) {
    TopRec* h = &heap[top_base(kShardDepth) + b];
    if (h->state == kTopBin) {
      h->r = t.rec[h->node - noff];
      h->state = kTopComplete;
    }
  }
}

// after the all-reduce: the cells above the bins, one launch per level, deepest level first
__global__ void __launch_bounds__(256) heap_sweep_kernel(TopRec* __restrict__ heap, int d) {
  const uint32_t cells = 1u << (2 * d);
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < cells; p += gridDim.x * blockDim.x)
    heap_sweep_cell(heap, d, p);
}

__global__ void __launch_bounds__(256)
    heap_writeback_kernel(const ShardMeta* __restrict__ sm, const TopRec* __restrict__ heap, TreeArrays t) {
  if (sm->M_local > t.node_cap) return;
  const uint32_t noff = sm->node_off, ml = sm->M_local;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < top_base(kShardDepth); s += gridDim.x * blockDim.x) {
    const TopRec h = heap[s];
    if (h.state != kTopComputed || h.node < noff || h.node - noff >= ml) continue;
    t.rec[h.node - noff] = h.r;
    if (h.r.aq > 0.0) t.ndepth[h.node - noff] |= (uint8_t)kDepthCharged;
  }
}

__global__ void __launch_bounds__(256)
    finalize_nodes_shard_kernel(const TreeMeta* __restrict__ meta, const ShardMeta* __restrict__ sm,
                                const float4* __restrict__ pqr, const float4* __restrict__ accm,
                                const uint64_t* __restrict__ lkeys, const uint32_t* __restrict__ binprefix,
                                TreeArrays t, StrictDirect strict_direct) {
  const uint32_t ml = sm->M_local, noff = sm->node_off;
  if (ml > t.node_cap) return;
  t.nodeA -= noff, t.nodeB -= noff, t.rec -= noff, t.ndepth -= noff;
  const float root_size = meta->root.size;
  const SubtreeEndShard end{noff, noff + ml, sm->M_total, meta->n, sm->body_base, t.nodeB, lkeys, binprefix};
  const uint32_t stride = gridDim.x * blockDim.x;
  // the subtree end of a node is read from its successor's record: a node's own B.z may be written concurrently,
  // B.y (what SubtreeEndShard reads) never is
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < ml; k += stride)
    finalize_node(noff + k, root_size, pqr, accm, t, end, strict_direct, strict_direct.limit != 0);
}

// psim_config.strict_centres in the sharded build: the chains (nodes that start at the same body) of this rank's
// piece that own a node of more than `direct` bodies, in the format the single-GPU emit kernel reports them
// (tree.cuh StrictEmit; node and body indices are global)
__global__ void __launch_bounds__(256)
    strict_candidates_shard_kernel(const ShardMeta* __restrict__ sm, const uint16_t* __restrict__ le,
                                   const uint32_t* __restrict__ nodebase, TreeArrays t, uint32_t direct,
                                   uint4* __restrict__ cand, uint32_t* __restrict__ cand_count, uint32_t cand_cap) {
  const uint32_t first = sm->hl, count = sm->n_local, noff = sm->node_off, body_base = sm->body_base;
  if (sm->M_local > t.node_cap) return;
  t.nodeB -= noff;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint32_t a = first + k;
    const uint16_t lev = le[a];
    const int K = le_ell(lev) - le_lambda(lev) - 1;  // internal nodes base .. base + K - 1, shallowest first
    if (K < 1) continue;
    const uint32_t base = nodebase[a];  // global (globalize_nodebase_kernel)
    uint32_t big = 0;
    while ((int)big < K && t.nodeB[base + big].z > direct) ++big;  // counts shrink with depth
    if (big == 0) continue;
    const uint32_t slot = atomicAdd(cand_count, 1u);
    // the strict kernels index the rank's own node arrays: local node index
    if (slot < cand_cap) cand[slot] = make_uint4(a + body_base, base - noff, big - 1u, t.nodeB[base].z);
  }
}

// table 2: (local traversal rank of the first node of each owned, non-empty bin) + 1, and T_local
__global__ void __launch_bounds__(256)
    table_trav_kernel(const uint32_t* __restrict__ binhist, const ShardPlan* __restrict__ plan,
                      ShardMeta* __restrict__ sm, const uint32_t* __restrict__ nb_bin,
                      const uint32_t* __restrict__ trav_rank, const uint32_t* __restrict__ t_local,
                      unsigned long long* __restrict__ xbuf) {
  const uint32_t r = sm->rank, noff = sm->node_off;
  const uint32_t b0 = plan->bin_lo[r], b1 = plan->bin_lo[r + 1];
  for (uint32_t b = b0 + blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += gridDim.x * blockDim.x)
    if (binhist[b]) {
      // a bin whose bodies all sit in a leaf that started earlier has no node of its own: its "first
      // node" is the one after the local piece
      const uint32_t k = nb_bin[b] - noff;
      xbuf[b] = (unsigned long long)(k < sm->M_local ? trav_rank[k] : *t_local) + 1ull;
    }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sm->T_local = *t_local;
    xbuf[kBins + r] = *t_local;
  }
}

// the rank's charged nodes into its segment of the global traversal arrays, skip pointers remapped
__global__ void __launch_bounds__(256)
    compact_traversal_shard_kernel(const ShardMeta* __restrict__ sm, const float4* __restrict__ nodeA,
                                   const uint4* __restrict__ nodeB, const uint32_t* __restrict__ rank,
                                   const uint32_t* __restrict__ nb_bin, const uint32_t* __restrict__ trav_bin,
                                   uint32_t node_cap, float4* __restrict__ travA, uint4* __restrict__ travB,
                                   uint32_t* __restrict__ trav_count) {
  const uint32_t ml = sm->M_local, noff = sm->node_off, M = sm->M_total, T = sm->T_total, toff = sm->trav_off;
  if (blockIdx.x == 0 && threadIdx.x == 0) *trav_count = (ml > node_cap || T > node_cap) ? 0u : T;
  if (ml > node_cap || T > node_cap) return;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < ml; k += stride) {
    uint4 nb = nodeB[k];
    if (!(nb.w & kNodeCharged)) continue;
    const uint32_t x = nb.x;
    if (x >= M) {
      nb.x = T;
    } else if (x - noff < ml) {
      nb.x = toff + rank[x - noff];
    } else {
      uint32_t lo = 0, hi = kBins;  // first bin whose first node is >= x (x is the first node of a bin)
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (nb_bin[mid] >= x) hi = mid; else lo = mid + 1;
      }
      nb.x = trav_bin[lo];
    }
    const uint32_t r = toff + rank[k];
    travA[r] = nodeA[k];
    travB[r] = nb;
  }
}

}  // namespace psim

// strict.cuh — internal node centres by the reference's own arithmetic, at build speed.
//
// Replaces the centre part of Quadtree::propagate (src/quadtree/quadtree.rs:114-139): for every internal
// node, three serial f32 running sums over the node's body range, sum |q|, fold(acc + pos * |q|), then
// pos / sum|q|.  The result is bit-identical to that loop (leaf capacity 1: the sorted order IS the
// reference's order); the work is organised so that no thread walks more than ~8k addends:
//
//   * a body with q == 0 adds exactly +-0 to all three sums, so only CHARGED bodies are visited
//     (cidx = exclusive scan of the charged flags, cw = their {|q|, x|q|, y|q|} in sorted order);
//   * the nested nodes that START at the same body share one running sum ("chain": the node at depth d
//     is a prefix of the node at depth d - 1), so a chain costs the length of its largest node;
//   * chains are ordered by length class (counting sort, longest first) and walked one per thread
//     (strict_chain_kernel), so the lanes of a warp do similar amounts of work;
//   * a chain longer than kStrictT1 hands its accumulators over at a block boundary; the rest is cut
//     into blocks of kStrictBlock addends whose effect on the accumulator is computed independently
//     (strict_blockfn_kernel: strict_logic.cuh, exact integer mantissa offsets under a speculated
//     binade) and chained by one warp per sum (strict_compose_kernel: warp scan over 32 blocks at a
//     time, serial f32 additions for the few blocks where the accumulator changes binade).
//
// Nodes whose sum |q| is <= 1e-6 although they hold charge take the reference's mass / centroid
// fall-backs in strict_slow_kernel (one thread per such node; none exist for charges of order 1).
// Nodes without any charged body never enter a field sum; their centres are only needed by the export
// (psim_download_nodes), which runs strict_chargeless_kernel first.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sort.cuh"
#include "strict_logic.cuh"
#include "tree_logic.cuh"

namespace psim {

constexpr uint32_t kStrictT1 = 8192;      // addends a single thread may walk
constexpr int kStrictClasses = 15;        // length classes 1..14 (class c: 2^(c-1) <= len < 2^c, capped)
constexpr uint32_t kSlowSentinel = 0x7fc0deadu;

struct StrictLong {  // a chain handed over to the block machinery
  uint32_t body;     // chain head (sorted body index)
  uint32_t base;     // pre-order index of the chain's shallowest node
  int32_t k_next;    // nodes base + k_next .. base + 0 are still open (k_next: the smallest of them)
  uint32_t blk0;     // first block (global index into cw / kStrictBlock) not yet summed
  uint32_t cend;     // end of the chain in cw
  float s[3];        // accumulators at blk0 * kStrictBlock
};

struct StrictArrays {
  uint32_t* cidx;        // n + 1
  float4* cw;            // charged bodies: {|q|, x|q|, y|q|, 0}
  uint32_t* chains;      // chain heads by length class, longest first
  uint32_t* hist;        // [0, 32): class counts, [32]: chains, [64, 96): scatter cursors
  StrictLong* longs;
  uint32_t* counters;    // [0] long chains, [1] items, [2] slow nodes, [3] error bits
  uint32_t* item_first;  // long_cap + 1
  double* pblk;          // 3 arrays of (blocks + 1): exclusive f64 prefix of the block sums
  BlockFn* fns;          // 3 per item
  uint32_t long_cap, item_cap, blk_cap;
};

struct ChargedBodyFn {
  const float4* pqr;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return pqr[i].z != 0.0f ? 1u : 0u; }
};

__global__ void __launch_bounds__(256)
    strict_addends_kernel(const float4* __restrict__ pqr, uint32_t n, const uint32_t* __restrict__ cidx,
                          float4* __restrict__ cw, uint32_t* __restrict__ hist, uint32_t* __restrict__ counters) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < 96) hist[tid] = 0;
  if (tid < 4) counters[tid] = 0;
  for (uint32_t i = tid; i < n; i += stride) {
    const float4 p = pqr[i];
    if (p.z != 0.0f) {
      const float a = fabsf(p.z);
      cw[cidx[i]] = make_float4(a, f_mul(p.x, a), f_mul(p.y, a), 0.0f);
    }
  }
}

// length (in charged bodies) of the chain that starts at body i, 0 if none
__device__ __forceinline__ uint32_t chain_length(uint32_t i, const uint16_t* __restrict__ le,
                                                 const uint32_t* __restrict__ nodebase,
                                                 const uint4* __restrict__ nodeB,
                                                 const uint32_t* __restrict__ cidx) {
  const uint16_t lev = le[i];
  if (le_ell(lev) - le_lambda(lev) < 2) return 0;  // the body starts no internal node
  const uint32_t cnt = nodeB[nodebase[i]].z;       // the shallowest node of the chain is the largest
  return cidx[i + cnt] - cidx[i];
}
__device__ __forceinline__ int chain_class(uint32_t len) {
  const int c = 32 - __clz(len);
  return c < kStrictClasses - 1 ? c : kStrictClasses - 1;
}

__global__ void __launch_bounds__(256)
    strict_chain_count_kernel(uint32_t n, const uint16_t* __restrict__ le, const uint32_t* __restrict__ nodebase,
                              const TreeMeta* __restrict__ meta, TreeArrays t, const uint32_t* __restrict__ cidx,
                              uint32_t* __restrict__ hist) {
  if (meta->num_nodes > t.node_cap) return;
  __shared__ uint32_t s_cnt[kStrictClasses];
  if (threadIdx.x < kStrictClasses) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t len = chain_length(i, le, nodebase, t.nodeB, cidx);
    if (len) atomicAdd(&s_cnt[chain_class(len)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kStrictClasses && s_cnt[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_cnt[threadIdx.x]);
}

// counting sort by class, longest class first.  Every CTA owns a contiguous slab of bodies, reserves one
// range per class with a single global atomic and scatters into it.
__global__ void __launch_bounds__(256)
    strict_chain_scatter_kernel(uint32_t n, uint32_t per_block, const uint16_t* __restrict__ le,
                                const uint32_t* __restrict__ nodebase, const TreeMeta* __restrict__ meta,
                                TreeArrays t, const uint32_t* __restrict__ cidx, uint32_t* __restrict__ hist,
                                uint32_t* __restrict__ chains) {
  if (meta->num_nodes > t.node_cap) return;
  __shared__ uint32_t s_cnt[kStrictClasses], s_base[kStrictClasses];
  if (threadIdx.x < kStrictClasses) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * per_block;
  const uint32_t hi = lo + per_block < n ? lo + per_block : n;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint32_t len = chain_length(i, le, nodebase, t.nodeB, cidx);
    if (len) atomicAdd(&s_cnt[chain_class(len)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kStrictClasses) {
    uint32_t start = 0;
    for (int c = kStrictClasses - 1; c > (int)threadIdx.x; --c) start += hist[c];
    const uint32_t mine = s_cnt[threadIdx.x];
    s_base[threadIdx.x] = start + (mine ? atomicAdd(&hist[64 + threadIdx.x], mine) : 0u);
    if (blockIdx.x == 0 && threadIdx.x == 0) hist[32] = start + hist[0];  // class 0 is empty: total chains
  }
  __syncthreads();
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint32_t len = chain_length(i, le, nodebase, t.nodeB, cidx);
    if (len) chains[atomicAdd(&s_base[chain_class(len)], 1u)] = i;
  }
}

__device__ __forceinline__ void strict_write_centre(float sa, float sx, float sy, uint32_t node, const TreeArrays& t,
                                                    uint32_t* counters) {
  float2 c;
  if (sa > 1e-6f) {
    c = make_float2(f_div(sx, sa), f_div(sy, sa));
  } else {  // mass-weighted / centroid fall-back: strict_slow_kernel
    c = make_float2(__uint_as_float(kSlowSentinel), __uint_as_float(kSlowSentinel));
    atomicAdd(&counters[2], 1u);
  }
  *reinterpret_cast<float2*>(&t.nodeA[node]) = c;
}

__device__ __forceinline__ void strict_acc(const float4 v, float& sa, float& sx, float& sy) {
  sa = f_add(sa, v.x), sx = f_add(sx, v.y), sy = f_add(sy, v.z);
}

// one chain per thread
__global__ void __launch_bounds__(128)
    strict_chain_kernel(const uint16_t* __restrict__ le, const uint32_t* __restrict__ nodebase,
                        const TreeMeta* __restrict__ meta, TreeArrays t, StrictArrays S) {
  if (meta->num_nodes > t.node_cap) return;
  const uint32_t nchains = S.hist[32];
  const uint32_t stride = gridDim.x * blockDim.x;
  const float4* __restrict__ cw = S.cw;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < nchains; g += stride) {
    const uint32_t i = S.chains[g];
    const uint16_t lev = le[i];
    const int K = le_ell(lev) - le_lambda(lev) - 1;  // internal nodes base .. base + K - 1, shallowest first
    const uint32_t base = nodebase[i];
    const uint32_t c0 = S.cidx[i];
    const uint32_t limit = ((c0 + kStrictT1) / kStrictBlock + 1) * kStrictBlock;
    uint32_t pos = c0;
    float sa = 0.0f, sx = 0.0f, sy = 0.0f;
    for (int k = K - 1; k >= 0; --k) {
      const uint32_t node = base + (uint32_t)k;
      const uint32_t cend = S.cidx[i + t.nodeB[node].z];
      const uint32_t stop = cend < limit ? cend : limit;
      while (pos < stop && (pos & 7u)) strict_acc(cw[pos++], sa, sx, sy);
      while (pos + 8 <= stop) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = cw[pos + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) strict_acc(v[j], sa, sx, sy);
        pos += 8;
      }
      while (pos < stop) strict_acc(cw[pos++], sa, sx, sy);
      if (stop < cend) {  // hand the rest over
        const uint32_t slot = atomicAdd(&S.counters[0], 1u);
        if (slot < S.long_cap) {
          StrictLong L;
          L.body = i, L.base = base, L.k_next = k, L.blk0 = limit / kStrictBlock;
          L.cend = S.cidx[i + t.nodeB[base].z];
          L.s[0] = sa, L.s[1] = sx, L.s[2] = sy;
          S.longs[slot] = L;
        } else {
          atomicOr(&S.counters[3], 1u);
        }
        break;
      }
      if (cend > c0) strict_write_centre(sa, sx, sy, node, t, S.counters);
    }
  }
}

// f64 sums of every block of kStrictBlock addends (speculation only), one warp per block
__global__ void __launch_bounds__(256)
    strict_blocksum_kernel(const float4* __restrict__ cw, const uint32_t* __restrict__ n_charged, uint32_t blk_cap,
                           double* __restrict__ pblk) {
  const uint32_t nc = *n_charged;
  const uint32_t nblk = (nc + kStrictBlock - 1) / kStrictBlock;
  const int lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nblk && b < blk_cap; b += warps) {
    double a = 0.0, x = 0.0, y = 0.0;
    const uint32_t end = (b + 1) * kStrictBlock < nc ? (b + 1) * kStrictBlock : nc;
    for (uint32_t i = b * kStrictBlock + lane; i < end; i += 32) {
      const float4 v = cw[i];
      a += (double)v.x, x += (double)v.y, y += (double)v.z;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      x += __shfl_xor_sync(0xffffffffu, x, off);
      y += __shfl_xor_sync(0xffffffffu, y, off);
    }
    if (lane == 0) {
      pblk[b + 1] = a;
      pblk[(size_t)(blk_cap + 1) + b + 1] = x;
      pblk[2 * (size_t)(blk_cap + 1) + b + 1] = y;
    }
  }
}

// single CTA: in-place inclusive prefix of the three block-sum arrays (entry 0 = 0), and the item offsets
// of the long chains
__global__ void __launch_bounds__(1024)
    strict_long_setup_kernel(const uint32_t* __restrict__ n_charged, StrictArrays S) {
  __shared__ double s_part[1024];
  __shared__ uint32_t s_u[1024];
  const uint32_t nc = *n_charged;
  uint32_t nblk = (nc + kStrictBlock - 1) / kStrictBlock;
  if (nblk > S.blk_cap) nblk = S.blk_cap;
  const uint32_t per = (nblk + 1023) / 1024;
  for (int k = 0; k < 3; ++k) {
    double* p = S.pblk + (size_t)k * (S.blk_cap + 1);
    const uint32_t lo = threadIdx.x * per, hi = (lo + per < nblk) ? lo + per : nblk;
    double sum = 0.0;
    for (uint32_t b = lo; b < hi; ++b) sum += p[b + 1];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double run = 0.0;
      for (int i = 0; i < 1024; ++i) {
        const double v = s_part[i];
        s_part[i] = run;
        run += v;
      }
      p[0] = 0.0;
    }
    __syncthreads();
    double run = s_part[threadIdx.x];
    for (uint32_t b = lo; b < hi; ++b) {
      run += p[b + 1];
      p[b + 1] = run;
    }
    __syncthreads();
  }
  // item offsets
  uint32_t nlong = S.counters[0];
  if (nlong > S.long_cap) nlong = S.long_cap;
  const uint32_t lper = (nlong + 1023) / 1024;
  const uint32_t llo = threadIdx.x * lper, lhi = (llo + lper < nlong) ? llo + lper : nlong;
  uint32_t cnt = 0;
  for (uint32_t c = llo; c < lhi; ++c) {
    const StrictLong L = S.longs[c];
    cnt += (L.cend - L.blk0 * kStrictBlock + kStrictBlock - 1) / kStrictBlock;
  }
  s_u[threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 1024; ++i) {
      const uint32_t v = s_u[i];
      s_u[i] = run;
      run += v;
    }
    S.counters[1] = run;
    S.item_first[nlong] = run;
    if (run > S.item_cap) atomicOr(&S.counters[3], 2u);
  }
  __syncthreads();
  uint32_t run = s_u[threadIdx.x];
  for (uint32_t c = llo; c < lhi; ++c) {
    const StrictLong L = S.longs[c];
    S.item_first[c] = run;
    run += (L.cend - L.blk0 * kStrictBlock + kStrictBlock - 1) / kStrictBlock;
  }
}

// one (chain, block) item per thread: the block's effect on each of the three accumulators
__global__ void __launch_bounds__(128) strict_blockfn_kernel(StrictArrays S) {
  if (S.counters[3]) return;
  const uint32_t nitems = S.counters[1];
  uint32_t nlong = S.counters[0];
  const uint32_t stride = gridDim.x * blockDim.x;
  const size_t pstride = (size_t)S.blk_cap + 1;
  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += stride) {
    uint32_t a = 0, b = nlong;  // last chain whose first item is <= it
    while (b - a > 1) {
      const uint32_t mid = a + ((b - a) >> 1);
      if (S.item_first[mid] <= it) a = mid; else b = mid;
    }
    const StrictLong L = S.longs[a];
    const uint32_t gb = L.blk0 + (it - S.item_first[a]);
    const uint32_t first = gb * kStrictBlock;
    const uint32_t last = first + kStrictBlock < L.cend ? first + kStrictBlock : L.cend;
    BlockFn f[3];
    float iu[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* p = S.pblk + k * pstride;
      const double s_in = (double)L.s[k] + (p[gb] - p[L.blk0]);
      const int e = spec_exponent(s_in + 0.5 * (p[gb + 1] - p[gb]));
      blockfn_init(f[k], e);
      iu[k] = e == kBadExp ? 0.0f : inv_ulp(e);
    }
    for (uint32_t i = first; i < last; ++i) {
      const float4 v = S.cw[i];
      if (f[0].e != kBadExp) blockfn_step(f[0], v.x, iu[0]);
      if (f[1].e != kBadExp) blockfn_step(f[1], v.y, iu[1]);
      if (f[2].e != kBadExp) blockfn_step(f[2], v.z, iu[2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) S.fns[3 * (size_t)it + k] = f[k];
  }
}

// s + cw[first .. last).comp in order, by one warp (every lane returns the result)
__device__ __forceinline__ float strict_serial_run(float s, const float4* __restrict__ cw, int comp, uint32_t first,
                                                   uint32_t last, int lane) {
  const float* __restrict__ w = reinterpret_cast<const float*>(cw) + comp;
  for (uint32_t base = first; base < last; base += 32) {
    const uint32_t idx = base + lane;
    const float v = idx < last ? w[4 * (size_t)idx] : 0.0f;
    const int cnt = last - base < 32u ? (int)(last - base) : 32;
    for (int j = 0; j < cnt; ++j) s = f_add(s, __shfl_sync(0xffffffffu, v, j));
  }
  return s;
}

// one CTA of three warps per long chain; warp w owns accumulator w
__global__ void __launch_bounds__(96) strict_compose_kernel(const TreeMeta* __restrict__ meta, TreeArrays t, StrictArrays S) {
  if (meta->num_nodes > t.node_cap || S.counters[3]) return;
  __shared__ uint32_t s_end[kMaxLevels + 1];
  __shared__ float s_res[kMaxLevels + 1][3];
  uint32_t nlong = S.counters[0];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t c = blockIdx.x; c < nlong; c += gridDim.x) {
    const StrictLong L = S.longs[c];
    __syncthreads();
    if (threadIdx.x <= (uint32_t)L.k_next) s_end[threadIdx.x] = S.cidx[L.body + t.nodeB[L.base + threadIdx.x].z];
    __syncthreads();
    const uint32_t item0 = S.item_first[c];
    const uint32_t nblk = S.item_first[c + 1] - item0;
    uint32_t blk = 0;
    int kk = L.k_next;
    float s = L.s[w];
    while (blk < nblk) {
      BlockFn f;
      const bool have = blk + lane < nblk;
      if (have) f = S.fns[3 * (size_t)(item0 + blk + lane) + w];
      else blockfn_init(f, kBadExp);
      int e = 0;
      int32_t M = 0;
      const bool ok = f32_split(s, e, M);
      int32_t i0 = f.o[0], i1 = f.o[1];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int32_t p0 = __shfl_up_sync(0xffffffffu, i0, off), p1 = __shfl_up_sync(0xffffffffu, i1, off);
        if (lane >= off) {
          int32_t n0, n1;
          blockfn_compose(n0, n1, p0, p1, i0, i1);
          i0 = n0, i1 = n1;
        }
      }
      int32_t x0 = __shfl_up_sync(0xffffffffu, i0, 1), x1 = __shfl_up_sync(0xffffffffu, i1, 1);
      if (lane == 0) x0 = 0, x1 = 0;
      const int32_t Mj = M + ((M & 1) ? x1 : x0);  // mantissa entering block blk + lane (if all before are valid)
      const bool valid = ok && have && blockfn_valid(f, e, Mj);
      const uint32_t bal = __ballot_sync(0xffffffffu, valid);
      const int jstar = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;
      const uint32_t left = nblk - blk;
      uint32_t avail = (uint32_t)jstar + 1 < left ? (uint32_t)jstar + 1 : left;
      if (avail > 32u) avail = 32u;
      // nodes that end inside a block whose entering accumulator is known
      while (kk >= 0) {
        const uint32_t bend = (s_end[kk] - 1) / kStrictBlock - L.blk0;
        if (bend >= blk + avail) break;
        const int j = (int)(bend - blk);
        const int32_t Mb = __shfl_sync(0xffffffffu, Mj, j);
        const float s_in = j == 0 ? s : f32_join(e, Mb);
        const float v = strict_serial_run(s_in, S.cw, w, (L.blk0 + bend) * kStrictBlock, s_end[kk], lane);
        if (lane == 0) s_res[kk][w] = v;
        --kk;
      }
      if (jstar > 0) {
        const int32_t tot = __shfl_sync(0xffffffffu, (M & 1) ? i1 : i0, 31);
        const int32_t Mn = jstar < 32 ? __shfl_sync(0xffffffffu, Mj, jstar & 31) : M + tot;
        s = f32_join(e, Mn);
        blk += (uint32_t)jstar;
      }
      if (jstar < 32 && blk < nblk) {
        const uint32_t first = (L.blk0 + blk) * kStrictBlock;
        const uint32_t last = first + kStrictBlock < L.cend ? first + kStrictBlock : L.cend;
        s = strict_serial_run(s, S.cw, w, first, last, lane);
        ++blk;
      }
    }
    __syncthreads();
    if (threadIdx.x <= (uint32_t)L.k_next)
      strict_write_centre(s_res[threadIdx.x][0], s_res[threadIdx.x][1], s_res[threadIdx.x][2], L.base + threadIdx.x, t,
                          S.counters);
  }
}

// the reference's loop for one node (quadtree.rs:114-139), all three cases
__device__ __forceinline__ float2 strict_node_centre(uint32_t b0, uint32_t b1, const float4* __restrict__ pqr,
                                                     const float4* __restrict__ accm) {
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
  return make_float2(wx, wy);
}

// charged nodes whose sum |q| did not pass 1e-6 (marked by strict_write_centre); exits at once if none
__global__ void __launch_bounds__(128)
    strict_slow_kernel(TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                       const float4* __restrict__ accm, TreeArrays t, const uint32_t* __restrict__ counters) {
  // scratch overflow of the long-chain machinery (never seen; the arenas are sized for 32 levels)
  if (blockIdx.x == 0 && threadIdx.x == 0 && counters[3]) atomicOr(&meta->err, 2u);
  if (counters[2] == 0) return;
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < M; node += stride) {
    if (__float_as_uint(t.nodeA[node].x) != kSlowSentinel) continue;
    const uint4 nb = t.nodeB[node];
    if (nb.w & kNodeLeaf) continue;
    *reinterpret_cast<float2*>(&t.nodeA[node]) = strict_node_centre(nb.y, nb.y + nb.z, pqr, accm);
  }
}

// export only: centres of the internal nodes without any charged body
__global__ void __launch_bounds__(128)
    strict_chargeless_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ pqr,
                             const float4* __restrict__ accm, TreeArrays t) {
  const uint32_t M = meta->num_nodes;
  if (M > t.node_cap) return;
  const uint32_t n_bodies = meta->n;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < M; node += stride) {
    const uint4 nb = t.nodeB[node];
    if ((nb.w & kNodeLeaf) || (nb.w & kNodeCharged)) continue;
    const uint32_t b1 = nb.x < M ? t.nodeB[nb.x].y : n_bodies;
    *reinterpret_cast<float2*>(&t.nodeA[node]) = strict_node_centre(nb.y, b1, pqr, accm);
  }
}

}  // namespace psim

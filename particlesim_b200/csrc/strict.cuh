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
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sort.cuh"
#include "strict_logic.cuh"
#include "tree_logic.cuh"

namespace psim {

constexpr uint32_t kStrictDirect = 64;    // nodes up to this many bodies are summed in the emit kernel (tree_logic.cuh)
constexpr uint32_t kStrictT1 = 1024;      // addends a single thread may walk
constexpr int kStrictClasses = 12;        // length classes 1..11 (class c: 2^(c-1) <= len < 2^c, capped)
constexpr uint32_t kSlowSentinel = 0x7fc0deadu;
constexpr uint32_t kChainLenCap = (1u << 27) - 1;  // chain lengths saturate here in the packed record

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
  float4* cw;            // charged bodies: {|q|, x|q|, y|q|, q}
  unsigned long long* qstat;  // 3 words, see strict_addends_kernel
  uint32_t* qc;          // charged bodies + 1: integer prefix of the charges
  uint4* chains;         // per chain {head body, shallowest node, first addend, length | (nodes - 1) << 27},
                         // by length class, longest first
  uint32_t* hist;        // [0, 32): class counts, [32]: chains, [64, 96): scatter cursors
  StrictLong* longs;
  uint32_t* counters;    // [0] long chains, [1] items, [2] slow nodes, [3] error bits, [4] candidates
  uint4* cand;           // chains reported by the emit kernel
  uint32_t cand_cap;
  uint32_t* item_first;  // long_cap + 1
  double* pblk;          // 3 arrays of (blocks + 1): exclusive f64 prefix of the block sums
  BlockFn* fns;          // 3 per item
  uint32_t long_cap, item_cap, blk_cap;
};

struct ChargedBodyFn {
  const float4* pqr;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return pqr[i].z != 0.0f ? 1u : 0u; }
};

// Besides the addends: cw.w carries the signed charge, and qstat collects what the emit kernel needs to know whether
// node charges can be taken as exact integer prefix differences (IntegerCharges below): [0] = sum of |q| over the
// bodies whose charge is an integer below 2^20, [1] = number of charged bodies whose charge is not, and the word at
// qstat + 2 = number of charged bodies + 1 (length of the prefix array).
__global__ void __launch_bounds__(256)
    strict_addends_kernel(const float4* __restrict__ pqr, uint32_t n, const uint32_t* __restrict__ cidx,
                          float4* __restrict__ cw, uint32_t* __restrict__ hist, uint32_t* __restrict__ counters,
                          unsigned long long* __restrict__ qstat) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < 96) hist[tid] = 0;
  if (tid < 4) counters[tid] = 0;
  if (tid == 0) *reinterpret_cast<uint32_t*>(qstat + 2) = cidx[n] + 1u;
  unsigned long long abs_sum = 0;
  uint32_t bad = 0;
  for (uint32_t i = tid; i < n; i += stride) {
    const float4 p = pqr[i];
    if (p.z != 0.0f) {
      const float a = fabsf(p.z);
      cw[cidx[i]] = make_float4(a, f_mul(p.x, a), f_mul(p.y, a), p.z);
      if (a < 1048576.0f && p.z == rintf(p.z)) abs_sum += (unsigned long long)a;
      else ++bad;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    abs_sum += __shfl_xor_sync(0xffffffffu, abs_sum, off);
    bad += __shfl_xor_sync(0xffffffffu, bad, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (abs_sum) atomicAdd(&qstat[0], abs_sum);
    if (bad) atomicAdd(&qstat[1], (unsigned long long)bad);
  }
}

// integer charge of the k-th charged body, for the prefix array qc (0 beyond the last one)
struct ChargeIntFn {
  const float4* cw;
  const uint32_t* ncharged;
  __device__ __forceinline__ uint32_t operator()(uint32_t k) const {
    return k < *ncharged ? (uint32_t)__float2int_rn(cw[k].w) : 0u;
  }
};

__device__ __forceinline__ int chain_class(uint32_t len) {
  const int c = 32 - __clz(len);
  return c < kStrictClasses - 1 ? c : kStrictClasses - 1;
}

// The emit kernel reports every chain that has a node of more than kStrictDirect bodies (tree.cuh
// StrictEmit); chains without a charged body drop out here.
__global__ void __launch_bounds__(256)
    strict_chain_count_kernel(const uint4* __restrict__ cand, const uint32_t* __restrict__ cand_count, uint32_t cand_cap,
                              const uint32_t* __restrict__ cidx, uint32_t* __restrict__ hist,
                              uint32_t* __restrict__ counters) {
  __shared__ uint32_t s_cnt[kStrictClasses];
  if (threadIdx.x < kStrictClasses) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t nc = *cand_count;
  if (nc > cand_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&counters[3], 4u);
    nc = cand_cap;
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < nc; g += stride) {
    const uint4 c = cand[g];
    const uint32_t len = cidx[c.x + c.w] - cidx[c.x];
    if (len) atomicAdd(&s_cnt[chain_class(len)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kStrictClasses && s_cnt[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_cnt[threadIdx.x]);
}

// counting sort by class, longest class first.  Every CTA owns a contiguous slab of candidates, reserves one
// range per class with a single global atomic and scatters into it.
__global__ void __launch_bounds__(256)
    strict_chain_scatter_kernel(const uint4* __restrict__ cand, const uint32_t* __restrict__ cand_count,
                                uint32_t cand_cap, uint32_t per_block, const uint32_t* __restrict__ cidx,
                                uint32_t* __restrict__ hist, uint4* __restrict__ chains) {
  __shared__ uint32_t s_cnt[kStrictClasses], s_base[kStrictClasses];
  if (threadIdx.x < kStrictClasses) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t nc = *cand_count;
  if (nc > cand_cap) nc = cand_cap;
  for (uint32_t slab = blockIdx.x; (uint64_t)slab * per_block < nc; slab += gridDim.x) {
    const uint32_t lo = slab * per_block;
    const uint32_t hi = lo + per_block < nc ? lo + per_block : nc;
    for (uint32_t g = lo + threadIdx.x; g < hi; g += blockDim.x) {
      const uint4 c = cand[g];
      const uint32_t len = cidx[c.x + c.w] - cidx[c.x];
      if (len) atomicAdd(&s_cnt[chain_class(len)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kStrictClasses) {
      uint32_t start = 0;
      for (int k = kStrictClasses - 1; k > (int)threadIdx.x; --k) start += hist[k];
      const uint32_t mine = s_cnt[threadIdx.x];
      s_base[threadIdx.x] = start + (mine ? atomicAdd(&hist[64 + threadIdx.x], mine) : 0u);
      if (slab == 0 && threadIdx.x == 0) hist[32] = start + hist[0];  // class 0 is empty: total chains
    }
    __syncthreads();
    for (uint32_t g = lo + threadIdx.x; g < hi; g += blockDim.x) {
      const uint4 c = cand[g];
      const uint32_t c0 = cidx[c.x], len = cidx[c.x + c.w] - c0;
      if (len)
        chains[atomicAdd(&s_base[chain_class(len)], 1u)] =
            make_uint4(c.x, c.y, c0, (len < kChainLenCap ? len : kChainLenCap) | (c.z << 27));
    }
    __syncthreads();
    if (threadIdx.x < kStrictClasses) s_cnt[threadIdx.x] = 0;
    __syncthreads();
  }
  if (nc == 0 && blockIdx.x == 0 && threadIdx.x == 0) hist[32] = 0;
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

// A thread's private sequential read of cw[first, last) through shared memory: cp.async copies of kStreamChunk
// addends each, kStreamDepth chunks in the ring, so that kStreamDepth - 1 chunks are in flight while one is
// consumed (a thread that walks a long run of addends one after the other would otherwise wait a full
// memory latency every few additions).  Element j of ring slot b lives at ring[(b * chunk + j) * blockDim.x].
constexpr int kStreamChunk = 8, kStreamDepth = 4;
constexpr int kStreamThreads = 128;
constexpr size_t kStreamSmem = (size_t)kStreamThreads * kStreamChunk * kStreamDepth * sizeof(float4);

struct AddendStream {
  float4* ring;  // this thread's column of the ring
  const float4* src;
  uint32_t next, last;  // next addend to request, end of the run
  int issue_slot, read_slot;
  __device__ __forceinline__ void issue() {
    float4* dst = ring + (size_t)issue_slot * kStreamChunk * kStreamThreads;
#pragma unroll
    for (int j = 0; j < kStreamChunk; ++j)
      if (next + j < last) __pipeline_memcpy_async(dst + j * kStreamThreads, src + next + j, sizeof(float4));
    __pipeline_commit();
    next = next + kStreamChunk < last ? next + kStreamChunk : last;
    issue_slot = issue_slot + 1 == kStreamDepth ? 0 : issue_slot + 1;
  }
  __device__ __forceinline__ void start(float4* smem, const float4* cw, uint32_t first, uint32_t end) {
    ring = smem + threadIdx.x, src = cw, next = first, last = end, issue_slot = 0, read_slot = 0;
#pragma unroll
    for (int k = 0; k < kStreamDepth - 1; ++k) issue();
  }
  // wait for the oldest chunk; at(j) reads its addends; release() recycles its slot
  __device__ __forceinline__ void acquire() { __pipeline_wait_prior(kStreamDepth - 2); }
  __device__ __forceinline__ float4 at(int j) const {
    return ring[((size_t)read_slot * kStreamChunk + j) * kStreamThreads];
  }
  __device__ __forceinline__ void release() {
    read_slot = read_slot + 1 == kStreamDepth ? 0 : read_slot + 1;
    issue();
  }
};

// one chain per thread.  The walk is ONE flat loop over chunks of the chain's addends (node boundaries are
// handled inside a chunk), so that the lanes of a warp (chains of the same length class) stay converged
// although their nodes end at different places.
__global__ void __launch_bounds__(kStreamThreads)
    strict_chain_kernel(const TreeMeta* __restrict__ meta, TreeArrays t, StrictArrays S) {
  extern __shared__ float4 s_ring[];
  if (meta->num_nodes > t.node_cap) return;
  const uint32_t nchains = S.hist[32];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < nchains; g += stride) {
    const uint4 rec = S.chains[g];
    const uint32_t i = rec.x, base = rec.y, c0 = rec.z;
    int k = (int)(rec.w >> 27);  // nodes base .. base + k are open, shallowest first
    const uint32_t len = rec.w & kChainLenCap;
    const uint32_t limit = ((c0 + kStrictT1) / kStrictBlock + 1) * kStrictBlock;
    const uint32_t all_end = (len < kChainLenCap && c0 + len < limit) ? c0 + len : limit;
    AddendStream st;
    st.start(s_ring, S.cw, c0, all_end);
    uint32_t pos = c0;
    // a single-node chain ends where the record says; otherwise the deepest open node's end is looked up
    uint32_t cend = (k == 0 && len < kChainLenCap) ? c0 + len : S.cidx[i + t.nodeB[base + (uint32_t)k].z];
    uint32_t stop = cend < limit ? cend : limit;
    float sa = 0.0f, sx = 0.0f, sy = 0.0f;
    bool open = true;
    while (open) {
      st.acquire();
      int j = 0;
      while (true) {
        if (pos == stop) {
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
            open = false;
            break;
          }
          if (cend > c0) strict_write_centre(sa, sx, sy, base + (uint32_t)k, t, S.counters);
          if (--k < 0) {
            open = false;
            break;
          }
          cend = S.cidx[i + t.nodeB[base + (uint32_t)k].z];
          stop = cend < limit ? cend : limit;
          continue;
        }
        if (j == kStreamChunk) break;
        int run = (int)(stop - pos < (uint32_t)(kStreamChunk - j) ? stop - pos : (uint32_t)(kStreamChunk - j));
        pos += (uint32_t)run;
        for (; run > 0; --run, ++j) strict_acc(st.at(j), sa, sx, sy);
      }
      st.release();
    }
    __pipeline_wait_prior(0);  // nothing of this chain may land in the ring after the next one has started
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
__global__ void __launch_bounds__(kStreamThreads) strict_blockfn_kernel(StrictArrays S) {
  extern __shared__ float4 s_ring[];
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
    AddendStream st;
    st.start(s_ring, S.cw, first, last);
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
    for (uint32_t pos = first; pos < last; pos += kStreamChunk) {
      st.acquire();
      const int cnt = last - pos < (uint32_t)kStreamChunk ? (int)(last - pos) : kStreamChunk;
      for (int j = 0; j < cnt; ++j) {
        const float4 v = st.at(j);
        if (f[0].e != kBadExp) blockfn_step(f[0], v.x, iu[0]);
        if (f[1].e != kBadExp) blockfn_step(f[1], v.y, iu[1]);
        if (f[2].e != kBadExp) blockfn_step(f[2], v.z, iu[2]);
      }
      st.release();
    }
    __pipeline_wait_prior(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) S.fns[3 * (size_t)it + k] = f[k];
  }
}

// s + cw[first .. last).comp in order, by one warp (every lane returns the result).  The addends of up to
// 512 positions are fetched first (independent loads), then the additions run as one shuffle + add chain.
__device__ __forceinline__ float strict_serial_run(float s, const float4* __restrict__ cw, int comp, uint32_t first,
                                                   uint32_t last, int lane) {
  const float* __restrict__ w = reinterpret_cast<const float*>(cw) + comp;
  for (uint32_t base = first; base < last; base += 512) {
    float v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const uint32_t idx = base + 32u * r + lane;
      v[r] = idx < last ? w[4 * (size_t)idx] : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const uint32_t b0 = base + 32u * r;
      if (b0 < last) {
        const int cnt = last - b0 < 32u ? (int)(last - b0) : 32;
        if (cnt == 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) s = f_add(s, __shfl_sync(0xffffffffu, v[r], j));
        } else {
          for (int j = 0; j < cnt; ++j) s = f_add(s, __shfl_sync(0xffffffffu, v[r], j));
        }
      }
    }
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
    BlockFn pre;  // the window that follows if all 32 blocks of the current one apply (the common case)
    uint32_t pre_blk = 0xffffffffu;
    while (blk < nblk) {
      BlockFn f;
      const bool have = blk + lane < nblk;
      if (blk == pre_blk) f = pre;
      else if (have) f = S.fns[3 * (size_t)(item0 + blk + lane) + w];
      else blockfn_init(f, kBadExp);
      pre_blk = blk + 32;
      if (pre_blk + lane < nblk) pre = S.fns[3 * (size_t)(item0 + pre_blk + lane) + w];
      else blockfn_init(pre, kBadExp);
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

// sort.cuh — hand-written onesweep LSD radix sort (8-bit digits) for (key, u32 payload) pairs,
// plus the exclusive-scan primitive the tree and cell-list builders use.
//
// Replaces: the recursive in-place 4-way partition of src/quadtree/quadtree.rs:56-63 /
// src/partition.rs:11-38 (a stable sort by quadrant key gives the same leaf membership and, for
// leaf_capacity 1 on duplicate-free input, the same body order) and the per-cell index pushes of
// src/cell_list.rs:33-38.
//
// Structure per sort: one histogram kernel over all digit positions, one tiny scan kernel, then one
// kernel per 8-bit digit: each CTA takes a tile ticket, ranks its keys with warp match_any /
// popc into per-warp shared-memory digit histograms, chains its per-digit totals to the preceding
// tiles by decoupled look-back, and writes the tile out through shared memory so global stores
// are contiguous per digit run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psim {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kRadix = 256;

constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;

// Plan written on the device by sort_scan_kernel: for each pass whether it can be skipped (one
// digit value holds every key) and which of the two ping-pong buffers it reads.
struct SortPlan {
  uint32_t skip[8];
  uint32_t src[9];  // src[p] = buffer index (0/1) pass p reads; src[npass] = where the result is
};

template <typename K>
__global__ void __launch_bounds__(256) sort_hist_kernel(const K* __restrict__ keys, uint32_t n,
                                                       int npass, int first_bit,
                                                       uint32_t* __restrict__ hist /*[npass][256]*/) {
  __shared__ uint32_t sh[8 * kRadix];
  for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const K k = keys[i];
    for (int p = 0; p < npass; ++p) {
      const uint32_t d = (uint32_t)(k >> (first_bit + 8 * p)) & 0xffu;
      atomicAdd(&sh[p * kRadix + d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// one CTA of 256 threads: exclusive scan of each pass's histogram in place, build the plan
__global__ void __launch_bounds__(256) sort_scan_kernel(uint32_t* __restrict__ hist, uint32_t n,
                                                        int npass, SortPlan* __restrict__ plan) {
  __shared__ uint32_t sh[kRadix];
  __shared__ uint32_t s_skip[8];
  const int t = threadIdx.x;
  for (int p = 0; p < npass; ++p) {
    const uint32_t v = hist[p * kRadix + t];
    if (t == 0) s_skip[p] = 0;
    __syncthreads();
    if (v == n && n > 0) s_skip[p] = 1;
    sh[t] = v;
    __syncthreads();
    // Hillis-Steele inclusive scan over 256 entries
    for (int off = 1; off < kRadix; off <<= 1) {
      uint32_t add = (t >= off) ? sh[t - off] : 0;
      __syncthreads();
      sh[t] += add;
      __syncthreads();
    }
    hist[p * kRadix + t] = sh[t] - v;
    __syncthreads();
  }
  if (t == 0) {
    uint32_t cur = 0;
    for (int p = 0; p < npass; ++p) {
      plan->skip[p] = s_skip[p];
      plan->src[p] = cur;
      if (!s_skip[p]) cur ^= 1u;
    }
    plan->src[npass] = cur;
    for (int p = npass; p < 8; ++p) plan->skip[p] = 1;
  }
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One digit pass.  keys/vals are the two ping-pong buffers; the plan says which one is the source.
// status: [num_tiles][256] words, zeroed before the sort; ticket: one counter per pass, zeroed.
template <typename K>
__global__ void __launch_bounds__(kSortThreads, 4)
    onesweep_pass_kernel(K* __restrict__ keys0, K* __restrict__ keys1, uint32_t* __restrict__ vals0,
                         uint32_t* __restrict__ vals1, uint32_t n, int pass, int shift,
                         const uint32_t* __restrict__ digit_base /*[256] exclusive, this pass*/,
                         uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                         const SortPlan* __restrict__ plan) {
  if (plan->skip[pass]) return;
  const K* __restrict__ kin = plan->src[pass] ? keys1 : keys0;
  K* __restrict__ kout = plan->src[pass] ? keys0 : keys1;
  const uint32_t* __restrict__ vin = plan->src[pass] ? vals1 : vals0;
  uint32_t* __restrict__ vout = plan->src[pass] ? vals0 : vals1;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* s_keys = reinterpret_cast<K*>(smem_raw);                                   // [kSortTile]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(K) * kSortTile);  // [kSortTile]
  __shared__ uint32_t s_whist[kSortWarps][kRadix];
  __shared__ uint32_t s_digit_start[kRadix];
  __shared__ uint32_t s_out_base[kRadix];
  __shared__ uint32_t s_warp_tot[kSortWarps];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&s_whist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t tile_base = (uint64_t)tile * kSortTile;
  const uint32_t count = (uint32_t)((n - tile_base) < (uint64_t)kSortTile ? (n - tile_base) : kSortTile);

  // ---- load (warp-striped: warp w owns a contiguous run, item j of lane l is element j*32+l of it)
  K key[kSortItems];
  uint32_t val[kSortItems];  // payloads travel with the keys: one round of global loads per tile
  uint32_t rank[kSortItems];
  const uint32_t warp_off = warp * (32 * kSortItems);
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint32_t loc = warp_off + j * 32 + lane;
    key[j] = (loc < count) ? kin[tile_base + loc] : (K)0;
    val[j] = (loc < count) ? vin[tile_base + loc] : 0u;
  }

  // ---- rank within the warp, in element order
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint32_t loc = warp_off + j * 32 + lane;
    const bool valid = loc < count;
    const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
    const uint32_t m = __match_any_sync(0xffffffffu, valid ? d : 0x100u);
    const int leader = __ffs(m) - 1;
    uint32_t prev = 0;
    if (lane == leader && valid) {
      prev = s_whist[warp][d];
      s_whist[warp][d] = prev + __popc(m);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[j] = prev + __popc(m & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit (thread d): warp-exclusive offsets, tile total, look-back, output base
  {
    const int d = tid;
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = s_whist[w][d];
      s_whist[w][d] = sum;
      sum += t;
    }
    uint32_t* my_status = status + (uint64_t)tile * kRadix + d;
    uint32_t excl = 0;
    if (tile == 0) {
      st_volatile_u32(my_status, kFlagIncl | sum);
    } else {
      st_volatile_u32(my_status, kFlagAgg | sum);
      int64_t t = (int64_t)tile - 1;
      while (true) {
        const uint32_t s = ld_volatile_u32(status + (uint64_t)t * kRadix + d);
        const uint32_t f = s & kFlagMask;
        if (f == 0) continue;  // predecessor has not published yet
        excl += s & kValMask;
        if (f == kFlagIncl) break;
        --t;
      }
      st_volatile_u32(my_status, kFlagIncl | (excl + sum));
    }
    // block-exclusive scan of `sum` over the 256 digits
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w)
      if (w < warp) wbase += s_warp_tot[w];
    const uint32_t dstart = wbase + incl - sum;
    s_digit_start[d] = dstart;
    s_out_base[d] = digit_base[d] + excl - dstart;  // global index = s_out_base[d] + tile position
  }
  __syncthreads();

  // ---- tile-local positions; stage keys and payloads in shared memory in sorted order
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const uint32_t loc = warp_off + j * 32 + lane;
    if (loc < count) {
      const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
      const uint32_t p = s_digit_start[d] + s_whist[warp][d] + rank[j];
      rank[j] = p;
      s_keys[p] = key[j];
      s_vals[p] = val[j];
    }
  }
  __syncthreads();
#pragma unroll 4
  for (uint32_t i = tid; i < count; i += kSortThreads) {
    const K k = s_keys[i];
    const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
    const uint32_t g = s_out_base[d] + i;
    kout[g] = k;
    vout[g] = s_vals[i];
  }
}

template <typename K>
constexpr size_t onesweep_smem_bytes() {
  return (sizeof(K) + sizeof(uint32_t)) * (size_t)kSortTile;
}

// Scratch the caller provides (device memory):
//   hist   : 8*256 u32
//   status : npass * num_tiles * 256 u32
//   ticket : 8 u32
//   plan   : SortPlan
struct SortScratch {
  uint32_t* hist;
  uint32_t* status;
  uint32_t* ticket;
  SortPlan* plan;
  size_t status_words;  // capacity
};

inline uint32_t sort_num_tiles(uint32_t n) { return (n + kSortTile - 1) / kSortTile; }
inline size_t sort_status_words(uint32_t n, int npass) {
  return (size_t)npass * sort_num_tiles(n) * kRadix;
}

// Sorts n (key, payload) pairs by bits [first_bit, first_bit + 8*npass) of the key, stable.
// Buffers 0 hold the input; the result lives in buffer scratch.plan->src[npass] (device side), so
// consumers read the plan on the device — no host synchronisation here.
template <typename K>
cudaError_t onesweep_sort(K* keys0, K* keys1, uint32_t* vals0, uint32_t* vals1, uint32_t n,
                          int first_bit, int npass, const SortScratch& sc, int sm_count,
                          cudaStream_t stream) {
  if (npass < 1 || npass > 8) return cudaErrorInvalidValue;
  const uint32_t tiles = sort_num_tiles(n);
  const size_t need = sort_status_words(n, npass);
  if (need > sc.status_words) return cudaErrorInvalidValue;
  cudaError_t e;
  if ((e = cudaMemsetAsync(sc.hist, 0, 8 * kRadix * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(sc.ticket, 0, 8 * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  if (need)
    if ((e = cudaMemsetAsync(sc.status, 0, need * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  int hist_blocks = sm_count * 4;
  if ((uint32_t)hist_blocks > (n + 255) / 256) hist_blocks = (int)((n + 255) / 256);
  if (hist_blocks < 1) hist_blocks = 1;
  sort_hist_kernel<K><<<hist_blocks, 256, 0, stream>>>(keys0, n, npass, first_bit, sc.hist);
  sort_scan_kernel<<<1, 256, 0, stream>>>(sc.hist, n, npass, sc.plan);
  const size_t smem = onesweep_smem_bytes<K>();
  e = cudaFuncSetAttribute(onesweep_pass_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
  if (e != cudaSuccess) return e;
  if (tiles) {
    for (int p = 0; p < npass; ++p) {
      onesweep_pass_kernel<K><<<tiles, kSortThreads, smem, stream>>>(
          keys0, keys1, vals0, vals1, n, p, first_bit + 8 * p, sc.hist + p * kRadix,
          sc.status + (size_t)p * tiles * kRadix, sc.ticket + p, sc.plan);
    }
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Two-tier key sort.  The quadrant keys are 64 bits (32 levels), but bodies are almost always told
// apart by the upper 32 (16 levels: cells of root / 65536), so the radix passes run on the upper word
// only (4 passes over 8-byte pairs instead of 8 over 12-byte pairs) and the few runs of bodies that
// share it are put in order of the full key afterwards:
//   gather_keys_kernel   sorted 64-bit keys through the sorted payload
//   fix_runs_kernel      one thread per run head; runs of <= kFixInsertion keys by insertion sort,
//                        longer ones are queued
//   sort_long_runs_kernel  one CTA per queued run: nothing to do if it is already in order (identical
//                        bodies), else a stable CTA-wide LSD radix sort of the lower word
// Both steps are stable, so the result equals a stable sort by the full key.
constexpr int kFixInsertion = 64;

__global__ void __launch_bounds__(256)
    gather_keys_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ idx0,
                       const uint32_t* __restrict__ idx1, const SortPlan* __restrict__ plan, int npass,
                       uint32_t n, uint64_t* __restrict__ out, uint32_t* __restrict__ long_count) {
  const uint32_t* __restrict__ idx = plan->src[npass] ? idx1 : idx0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *long_count = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = keys_in[idx[i]];
}

__global__ void __launch_bounds__(256)
    fix_runs_kernel(uint64_t* __restrict__ skeys, uint32_t* __restrict__ idx0, uint32_t* __restrict__ idx1,
                    const SortPlan* __restrict__ plan, int npass, uint32_t n,
                    uint32_t* __restrict__ long_count, uint32_t* __restrict__ long_start) {
  uint32_t* __restrict__ idx = plan->src[npass] ? idx1 : idx0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // a run's upper words never change while it is being reordered, so neighbours can be read freely
    const uint32_t hi = (uint32_t)(skeys[i] >> 32);
    if (i + 1 >= n || (uint32_t)(skeys[i + 1] >> 32) != hi) continue;
    if (i > 0 && (uint32_t)(skeys[i - 1] >> 32) == hi) continue;  // not the head of its run
    uint32_t e = i + 2;
    while (e < n && e - i <= (uint32_t)kFixInsertion && (uint32_t)(skeys[e] >> 32) == hi) ++e;
    if (e - i > (uint32_t)kFixInsertion) {
      long_start[atomicAdd(long_count, 1u)] = i;  // at most n / kFixInsertion runs exist
      continue;
    }
    for (uint32_t a = i + 1; a < e; ++a) {
      const uint64_t k = skeys[a];
      const uint32_t v = idx[a];
      uint32_t b = a;
      while (b > i && skeys[b - 1] > k) {
        skeys[b] = skeys[b - 1];
        idx[b] = idx[b - 1];
        --b;
      }
      skeys[b] = k;
      idx[b] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
    sort_long_runs_kernel(uint64_t* __restrict__ skeys, uint32_t* __restrict__ idx0, uint32_t* __restrict__ idx1,
                          const SortPlan* __restrict__ plan, int npass, uint32_t n,
                          const uint32_t* __restrict__ long_count, const uint32_t* __restrict__ long_start,
                          uint64_t* __restrict__ kscratch) {
  uint32_t* __restrict__ idx = plan->src[npass] ? idx1 : idx0;
  uint32_t* __restrict__ vscratch = plan->src[npass] ? idx0 : idx1;
  __shared__ uint32_t s_hist[kRadix], s_base[kRadix], s_wcnt[8][kRadix];
  __shared__ uint32_t s_len, s_flag;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t runs = *long_count;
  for (uint32_t r = blockIdx.x; r < runs; r += gridDim.x) {
    const uint32_t start = long_start[r];
    __syncthreads();
    if (t == 0) {
      // end of the run of equal upper words: galloping + binary search on the sorted upper words
      const uint32_t hi = (uint32_t)(skeys[start] >> 32);
      uint64_t in = start, step = 1;
      while (in + step < n && (uint32_t)(skeys[in + step] >> 32) == hi) in += step, step <<= 1;
      uint64_t a = in, b = in + step < n ? in + step : n;
      while (b - a > 1) {
        const uint64_t mid = a + ((b - a) >> 1);
        if ((uint32_t)(skeys[mid] >> 32) == hi) a = mid; else b = mid;
      }
      s_len = (uint32_t)(b - start);
      s_flag = 0;
    }
    __syncthreads();
    const uint32_t len = s_len;
    for (uint32_t j = t; j + 1 < len; j += 256)
      if (skeys[start + j] > skeys[start + j + 1]) s_flag = 1;
    __syncthreads();
    if (!s_flag) continue;  // already in order (the common case: identical bodies)
    uint64_t* ksrc = skeys + start;
    uint64_t* kdst = kscratch + start;
    uint32_t* vsrc = idx + start;
    uint32_t* vdst = vscratch + start;
    bool swapped = false;
    for (int shift = 0; shift < 32; shift += 8) {
      s_hist[t] = 0;
      __syncthreads();
      for (uint32_t j = t; j < len; j += 256) atomicAdd(&s_hist[(uint32_t)(ksrc[j] >> shift) & 0xffu], 1u);
      __syncthreads();
      if (t == 0) {
        uint32_t run = 0, uniform = 0;
        for (int d = 0; d < kRadix; ++d) {
          if (s_hist[d] == len) uniform = 1;
          s_base[d] = run;
          run += s_hist[d];
        }
        s_flag = uniform;
      }
      __syncthreads();
      if (s_flag) continue;  // every key has the same digit: nothing moves
      for (uint32_t c = 0; c < len; c += 256) {
#pragma unroll
        for (int w = 0; w < 8; ++w) s_wcnt[w][t] = 0;
        __syncthreads();
        const uint32_t j = c + t;
        const bool valid = j < len;
        const uint64_t k = valid ? ksrc[j] : 0ull;
        const uint32_t v = valid ? vsrc[j] : 0u;
        const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
        const uint32_t m = __match_any_sync(0xffffffffu, valid ? d : 0x100u);
        if (valid && lane == __ffs(m) - 1) s_wcnt[warp][d] = __popc(m);
        __syncthreads();
        if (valid) {
          uint32_t off = 0;
          for (int w = 0; w < warp; ++w) off += s_wcnt[w][d];
          const uint32_t pos = s_base[d] + off + __popc(m & lt);
          kdst[pos] = k;
          vdst[pos] = v;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_wcnt[w][t];
        s_base[t] += tot;
        __syncthreads();
      }
      __threadfence_block();
      uint64_t* kt = ksrc; ksrc = kdst; kdst = kt;
      uint32_t* vt = vsrc; vsrc = vdst; vdst = vt;
      swapped = !swapped;
    }
    __syncthreads();
    if (swapped) {
      for (uint32_t j = t; j < len; j += 256) kdst[j] = ksrc[j], vdst[j] = vsrc[j];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan of u32 values produced by a functor f(i), i in [0, n): reduce / scan-partials /
// downsweep.  total (if not null) receives the grand sum on the device.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename F>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(F f, uint32_t n,
                                                                  uint32_t* __restrict__ partials) {
  __shared__ uint32_t s_w[kScanThreads / 32];
  const uint32_t base = blockIdx.x * kScanTile;
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint32_t i = base + j * kScanThreads + threadIdx.x;
    if (i < n) sum += f(i);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_w[w];
    partials[blockIdx.x] = t;
  }
}

// single CTA, 1024 threads: exclusive scan of up to any number of partials (looped), writes total
__global__ void __launch_bounds__(1024) scan_partials_kernel(uint32_t* __restrict__ partials,
                                                            uint32_t count,
                                                            uint32_t* __restrict__ total) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < count; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = (i < count) ? partials[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_w[lane];
      uint32_t wi = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, off);
        if (lane >= off) wi += t;
      }
      s_w[lane] = wi - w;  // exclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t excl = carry + s_w[warp] + incl - v;
    if (i < count) partials[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = s_carry;
}

template <typename F>
__global__ void __launch_bounds__(kScanThreads)
    scan_downsweep_kernel(F f, uint32_t n, const uint32_t* __restrict__ partials,
                          uint32_t* __restrict__ out) {
  __shared__ uint32_t s_w[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;  // blocked arrangement
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint32_t i = base + j;
    v[j] = (i < n) ? f(i) : 0;
    sum += v[j];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w)
    if (w < warp) wbase += s_w[w];
  uint32_t run = partials[blockIdx.x] + wbase + incl - sum;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint32_t i = base + j;
    if (i < n) out[i] = run;
    run += v[j];
  }
}

inline uint32_t scan_num_tiles(uint32_t n) { return (n + kScanTile - 1) / kScanTile; }

// out[i] = sum_{j<i} f(j); partials must hold scan_num_tiles(n) words.
template <typename F>
cudaError_t exclusive_scan(F f, uint32_t n, uint32_t* out, uint32_t* partials, uint32_t* total,
                           cudaStream_t stream) {
  const uint32_t tiles = scan_num_tiles(n);
  if (tiles == 0) {
    if (total) return cudaMemsetAsync(total, 0, sizeof(uint32_t), stream);
    return cudaSuccess;
  }
  scan_reduce_kernel<F><<<tiles, kScanThreads, 0, stream>>>(f, n, partials);
  scan_partials_kernel<<<1, 1024, 0, stream>>>(partials, tiles, total);
  scan_downsweep_kernel<F><<<tiles, kScanThreads, 0, stream>>>(f, n, partials, out);
  return cudaGetLastError();
}

// Same, for a length that only exists on the device (*n_dev <= n_cap): the launches are sized for
// n_cap and every kernel clips to *n_dev.
template <typename F>
struct ClippedFn {
  F f;
  const uint32_t* n_dev;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return i < *n_dev ? f(i) : 0u; }
};
template <typename F>
__global__ void __launch_bounds__(kScanThreads)
    scan_downsweep_dyn_kernel(F f, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ partials,
                              uint32_t* __restrict__ out) {
  __shared__ uint32_t s_w[kScanThreads / 32];
  const uint32_t n = *n_dev;
  if (blockIdx.x * kScanTile >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint32_t i = base + j;
    v[j] = (i < n) ? f(i) : 0;
    sum += v[j];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w)
    if (w < warp) wbase += s_w[w];
  uint32_t run = partials[blockIdx.x] + wbase + incl - sum;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint32_t i = base + j;
    if (i < n) out[i] = run;
    run += v[j];
  }
}

template <typename F>
cudaError_t exclusive_scan_dyn(F f, const uint32_t* n_dev, uint32_t n_cap, uint32_t* out,
                               uint32_t* partials, uint32_t* total, cudaStream_t stream) {
  const uint32_t tiles = scan_num_tiles(n_cap);
  if (tiles == 0) return total ? cudaMemsetAsync(total, 0, sizeof(uint32_t), stream) : cudaSuccess;
  ClippedFn<F> cf{f, n_dev};
  scan_reduce_kernel<ClippedFn<F>><<<tiles, kScanThreads, 0, stream>>>(cf, n_cap, partials);
  scan_partials_kernel<<<1, 1024, 0, stream>>>(partials, tiles, total);
  scan_downsweep_dyn_kernel<F><<<tiles, kScanThreads, 0, stream>>>(f, n_dev, partials, out);
  return cudaGetLastError();
}

}  // namespace psim

// psim_core.cuh — per-element logic shared by the sm_100a kernels (and, compiled as plain C++,
// by the test-only host emulation in tests/emu/ that checks the construction algorithm against
// the oracle without a GPU).  No product code path runs these functions on the CPU.
//
// Reference semantics restated here (paths relative to /root/reference):
//   quadrant choice      src/quadtree/quadtree.rs:56-63  (y < cy, then x < cx, children 0..3)
//   centre recurrence    src/quadtree/quad.rs:45-50      (size *= 0.5; c += (±0.5) * size)
//   refusal rules        src/quadtree/quadtree.rs:44-54  (coincident bodies, size < 1e-6, len <= 1)
//   leaf / thread rules  src/quadtree/quadtree.rs:249-257,281
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PSIM_HD __host__ __device__ __forceinline__
#else
#define PSIM_HD inline
#endif

namespace psim {

constexpr int kMaxLevels = 32;  // 2 bits per level in a 64-bit key

// Node flags kept in the low byte group of NodeB.w
constexpr uint32_t kNodeLeaf = 1u << 8;      // no children
constexpr uint32_t kNodeZeroAgg = 1u << 9;   // refused / thread-capacity leaf: mass = charge = pos = 0
constexpr uint32_t kNodeDepthMask = 0xffu;

struct RootQuad {
  float cx, cy, size;
};

PSIM_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
PSIM_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}

// 32-level quadrant key of (x, y): replays the reference's comparisons against the fp32 centre
// recurrence.  Level 1 sits in bits 63:62; in each digit the y bit is above the x bit.  NaN
// coordinates compare false and fall into quadrant 3 at every level.
PSIM_HD uint64_t morton_key(float x, float y, RootQuad r) {
  float cx = r.cx, cy = r.cy, size = r.size;
  uint64_t key = 0;
#pragma unroll 4
  for (int l = 0; l < kMaxLevels; ++l) {
    const unsigned qx = (x < cx) ? 0u : 1u;
    const unsigned qy = (y < cy) ? 0u : 1u;
    key = (key << 2) | (uint64_t)((qy << 1) | qx);
    size = f_mul(size, 0.5f);
    const float h = f_mul(0.5f, size);  // (q - 0.5) * size, exact
    cx = f_add(cx, qx ? h : -h);
    cy = f_add(cy, qy ? h : -h);
  }
  return key;
}

// centre/size of the cell reached from the root by the first `depth` digits of `key`
PSIM_HD RootQuad quad_at(RootQuad r, uint64_t key, int depth) {
  for (int l = 0; l < depth; ++l) {
    const unsigned q = (unsigned)(key >> (62 - 2 * l)) & 3u;
    r.size = f_mul(r.size, 0.5f);
    const float h = f_mul(0.5f, r.size);
    r.cx = f_add(r.cx, (q & 1u) ? h : -h);
    r.cy = f_add(r.cy, (q >> 1) ? h : -h);
  }
  return r;
}
// Quad::into_quadrant for one more level
PSIM_HD RootQuad quad_child(RootQuad r, unsigned q) {
  r.size = f_mul(r.size, 0.5f);
  const float h = f_mul(0.5f, r.size);
  r.cx = f_add(r.cx, (q & 1u) ? h : -h);
  r.cy = f_add(r.cy, (q >> 1) ? h : -h);
  return r;
}

PSIM_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}

// number of leading 2-bit digits two keys share: 0..32
PSIM_HD int lcp_levels(uint64_t a, uint64_t b) { return clz64(a ^ b) >> 1; }

// quadrant digit of `key` that selects the child at depth `d` (1-based)
PSIM_HD unsigned digit_at(uint64_t key, int d) { return (unsigned)(key >> (64 - 2 * d)) & 3u; }

// Smallest depth whose cell is smaller than the reference's 1e-6 refusal threshold, capped at the
// key length.  A cell at this depth never splits.
PSIM_HD int depth_cap(float root_size) {
  float s = root_size;
  int d = 0;
  // `size < 1e-6` is false for NaN, so a NaN root size never stops early (like the reference)
  while (d < kMaxLevels && !(s < 1e-6f)) {
    s = f_mul(s, 0.5f);
    ++d;
  }
  return d;
}

// A node with `len` bodies tries to split iff len >= thread_capacity or len > leaf_capacity, and
// the attempt is refused when len <= 1.  So cells stop splitting at len <= c_eff:
PSIM_HD uint32_t effective_capacity(uint32_t leaf_capacity, uint32_t thread_capacity) {
  uint32_t c = leaf_capacity;
  if (thread_capacity >= 1 && thread_capacity - 1 < c) c = thread_capacity - 1;
  return c < 1 ? 1 : c;
}
// leaf aggregates are computed only on the depth-first path (len < thread_capacity) for len <= leaf_capacity
PSIM_HD bool leaf_is_aggregated(uint32_t len, uint32_t leaf_capacity, uint32_t thread_capacity) {
  return len <= leaf_capacity && len < thread_capacity;
}

// λ_i: levels shared with the previous body in sorted order (-1 for i == 0)
PSIM_HD int lambda_at(const uint64_t* keys, uint32_t i) {
  return i == 0 ? -1 : lcp_levels(keys[i - 1], keys[i]);
}

// Largest j in [lo, n] such that every body in [i, j) shares at least `d` levels with body i.
// Requires lo > i and that bodies (i, lo) already do.  Galloping + binary search on the sorted keys.
PSIM_HD uint32_t run_end(const uint64_t* keys, uint32_t n, uint32_t i, uint32_t lo, int d) {
  if (d <= 0) return n;
  const uint64_t ki = keys[i];
  if (lo >= n || lcp_levels(ki, keys[lo]) < d) return lo;
  // invariant: body `lo` is inside; find first outside
  uint32_t step = 1, in = lo;
  while (true) {
    const uint64_t probe = (uint64_t)in + step;
    if (probe >= n) break;
    if (lcp_levels(ki, keys[probe]) < d) break;
    in = (uint32_t)probe;
    step <<= 1;
  }
  uint64_t out = (uint64_t)in + step;  // first index known (or assumed) outside
  if (out > n) out = n;
  // binary search in (in, out): in is inside, out is outside (or n)
  uint32_t a = in, b = (uint32_t)out;
  while (b - a > 1) {
    const uint32_t mid = a + ((b - a) >> 1);
    if (lcp_levels(ki, keys[mid]) >= d)
      a = mid;
    else
      b = mid;
  }
  return b;
}

// Depth of the leaf cell that contains body i, and whether body i is the first body of that leaf.
//   lam   : λ_i
//   c_eff : effective_capacity()
//   dcap  : depth_cap(root size)
// A body inside a run of equal keys that is not the run's first body never starts a leaf.
// Device rule for coincident bodies: a run of equal 32-level keys that is larger than
// c_eff becomes one refused leaf at the depth where the run is alone in its cell.
PSIM_HD int leaf_depth(const uint64_t* keys, uint32_t n, uint32_t i, int lam, uint32_t c_eff, int dcap) {
  const uint64_t ki = keys[i];
  // the c_eff-th largest value among L_j = lcp(k_{i-j}, k_i) and R_j = lcp(k_i, k_{i+j})
  uint32_t l = 1, r = 1;
  int lv = lam;  // L_1
  int rv = (i + 1 < n) ? lcp_levels(ki, keys[i + 1]) : -1;
  int v = -1;
  for (uint32_t pick = 0; pick < c_eff; ++pick) {
    if (lv >= rv) {
      v = lv;
      if (lv < 0) break;
      ++l;
      lv = (l <= i) ? lcp_levels(keys[i - l], ki) : -1;
    } else {
      v = rv;
      ++r;
      rv = ((uint64_t)i + r < n) ? lcp_levels(ki, keys[i + r]) : -1;
    }
  }
  int d = v + 1;
  if (d > kMaxLevels) {
    // more than c_eff bodies share all 32 levels with body i: refused leaf where the run is alone
    const uint32_t rb = run_end(keys, n, i, i + 1, kMaxLevels);
    const int lam_rb = (rb < n) ? lcp_levels(ki, keys[rb]) : -1;
    d = (lam > lam_rb ? lam : lam_rb) + 1;
  }
  return d < dcap ? d : dcap;
}

}  // namespace psim

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
constexpr uint32_t kNodeCharged = 1u << 10;  // some body below the node has a non-zero charge
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
// first `levels` digits of the key, right-aligned (levels = 8: the shard bin of shard.cuh)
PSIM_HD uint32_t morton_prefix(float x, float y, RootQuad r, int levels) {
  float cx = r.cx, cy = r.cy, size = r.size;
  uint32_t key = 0;
  for (int l = 0; l < levels; ++l) {
    const unsigned qx = (x < cx) ? 0u : 1u;
    const unsigned qy = (y < cy) ? 0u : 1u;
    key = (key << 2) | ((qy << 1) | qx);
    size = f_mul(size, 0.5f);
    const float h = f_mul(0.5f, size);
    cx = f_add(cx, qx ? h : -h);
    cy = f_add(cy, qy ? h : -h);
  }
  return key;
}

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

// Smallest j in [0, hi] such that every body in [j, i] shares at least `d` levels with body i.
// Requires hi <= i and that bodies [hi, i] already do.
PSIM_HD uint32_t run_begin(const uint64_t* keys, uint32_t i, uint32_t hi, int d) {
  if (d <= 0) return 0;
  const uint64_t ki = keys[i];
  if (hi == 0 || lcp_levels(ki, keys[hi - 1]) < d) return hi;
  uint32_t step = 1, in = hi - 1;  // `in` is inside
  while (true) {
    if (step > in) break;
    if (lcp_levels(ki, keys[in - step]) < d) break;
    in -= step;
    step <<= 1;
  }
  // first index known outside is in - step (or "-1"); binary search in (out, in)
  int64_t a = (int64_t)in - (int64_t)step, b = in;  // a outside (or < 0), b inside
  if (a < -1) a = -1;
  while (b - a > 1) {
    const int64_t mid = a + ((b - a) >> 1);
    if (lcp_levels(ki, keys[mid]) >= d)
      b = mid;
    else
      a = mid;
  }
  return (uint32_t)b;
}

// the reference's coincidence test on one consecutive pair: (w[0].pos - w[1].pos).mag_sq() < 1e-12
// (quadtree.rs:49-51); pos4 = {x, y, ., .}
template <class P4>
PSIM_HD bool close_pair(const P4* pos4, uint32_t a, uint32_t b) {
  const float dx = f_add(pos4[a].x, -pos4[b].x), dy = f_add(pos4[a].y, -pos4[b].y);
  return f_add(f_mul(dx, dx), f_mul(dy, dy)) < 1e-12f;
}

constexpr uint32_t kCloseScanLimit = 1024;

// Depth of the leaf cell that contains body i.
//   lam   : λ_i
//   c_eff : effective_capacity()
//   dcap  : depth_cap(root size)
// Body i starts that leaf iff λ_i < ℓ_i.  Stop rules, in the reference's terms (quadtree.rs:44-54,
// 249-257, 281): the cell holds <= c_eff bodies; or every consecutive pair of its bodies is within
// 1e-6 A (refused: zero-aggregate leaf); or its size is below 1e-6 / the key is used up.
// The coincidence chain is evaluated in sorted order (the reference uses the partition's current
// order; the two agree unless a cell holds >= 3 bodies strung out at sub-1e-6 spacing), and runs
// longer than kCloseScanLimit are continued by key equality (exact for identical positions).
template <class P4>
PSIM_HD int leaf_depth(const uint64_t* keys, const P4* pos4, uint32_t n, uint32_t i, int lam,
                       uint32_t c_eff, int dcap) {
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
  const bool cl = i > 0 && close_pair(pos4, i - 1, i);
  const bool cr = i + 1 < n && close_pair(pos4, i, i + 1);
  if (cl || cr || d > kMaxLevels) {
    // run of consecutively coincident bodies around i: [ra, rb)
    uint32_t ra = i, rb = i + 1, steps = 0;
    while (ra > 0 && steps < kCloseScanLimit && close_pair(pos4, ra - 1, ra)) --ra, ++steps;
    if (steps == kCloseScanLimit) ra = run_begin(keys, i, ra, kMaxLevels);
    steps = 0;
    while (rb < n && steps < kCloseScanLimit && close_pair(pos4, rb - 1, rb)) ++rb, ++steps;
    if (steps == kCloseScanLimit) rb = run_end(keys, n, i, rb, kMaxLevels);
    if (d > kMaxLevels) {
      // bodies that share all 32 levels but are not coincident (sub-resolution separation, NaN):
      // the key cannot split them, so the whole equal-key run stops where it is alone in its cell
      const uint32_t ka = run_begin(keys, i, i, kMaxLevels), kb = run_end(keys, n, i, i + 1, kMaxLevels);
      if (ka < ra) ra = ka;
      if (kb > rb) rb = kb;
    }
    const int la = (ra > 0) ? lcp_levels(keys[ra - 1], ki) : -1;
    const int lb = (rb < n) ? lcp_levels(ki, keys[rb]) : -1;
    const int d_same = (la > lb ? la : lb) + 1;
    if (d_same < d) d = d_same;
  }
  return d < dcap ? d : dcap;
}

}  // namespace psim

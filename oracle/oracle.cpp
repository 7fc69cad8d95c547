/*
 * oracle.cpp — CPU restatement of ParticleSim's force hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may load it and for the parity-pinning statement).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Arithmetic is IEEE f32 with no implicit FMA contraction (build with
 * -ffp-contract=off; Rust never contracts).  Third-party arithmetic on the path is
 * `ultraviolet` 0.9.2 Vec2 (Cargo.lock:2267-2268, source not vendored): its two-component
 * formulas are restated in struct V2 below; whether mag_sq/dot fuse is the compile-time
 * switch ORC_UV_FMA (default: plain x*x + y*y).
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- ultraviolet::Vec2 (f32)
struct V2 {
  float x, y;
};
static inline V2 v2(float x, float y) { return V2{x, y}; }
static inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 operator-(V2 a) { return v2(-a.x, -a.y); }
static inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
static inline V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
static inline V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
static inline bool is_zero(V2 a) { return a.x == 0.0f && a.y == 0.0f; }
static inline float mag_sq(V2 a) {
#ifdef ORC_UV_FMA
  return fmaf(a.x, a.x, a.y * a.y);
#else
  return (a.x * a.x) + (a.y * a.y);
#endif
}
static inline float mag(V2 a) { return sqrtf(mag_sq(a)); }
static inline V2 normalized(V2 a) {
  float r_mag = 1.0f / mag(a);
  return v2(a.x * r_mag, a.y * r_mag);
}
// Rust f32::max / f32::min: a NaN operand is ignored
static inline float rmax(float a, float b) { return fmaxf(a, b); }
static inline float rmin(float a, float b) { return fminf(a, b); }
// Rust `f as isize` / `f as usize`: saturating, NaN -> 0
static inline int64_t as_isize(float f) {
  if (f != f) return 0;
  if (f >= 9.2233720368547758e18f) return INT64_MAX;
  if (f <= -9.2233720368547758e18f) return INT64_MIN;
  return (int64_t)f;
}
static inline uint64_t as_usize(float f) {
  if (f != f || f <= 0.0f) return 0;
  if (f >= 1.8446744073709552e19f) return UINT64_MAX;
  return (uint64_t)f;
}

// ---------------------------------------------------------------- Body (src/body/types.rs:38-62)
// Same field set as the reference so that the in-place partition moves a record of comparable
// size (~140 B); electrons are the SmallVec<[Electron;2]> inline part only.
struct Electron {
  V2 rel_pos, vel;
};
struct Body {
  V2 pos;
  float z;
  V2 vel;
  float vz;
  V2 acc;
  float az;
  float mass, radius, charge;
  uint64_t id;
  uint8_t species;
  uint8_t n_electrons;  // inline electrons in use (0..2); more live in OrcSim::elec
  Electron electrons[2];
  V2 e_field;
  bool surrounded_by_metal;
  V2 last_surround_pos;
  uint64_t last_surround_frame;
  float lithium_content;
  float species_lock_until;
};

// ---------------------------------------------------------------- Quad (src/quadtree/quad.rs)
struct Quad {
  V2 center;
  float size;
};

// quad.rs:11-35
static Quad quad_new_containing(const std::vector<Body> &bodies) {
  if (bodies.empty()) return Quad{v2(0, 0), 1.0f};
  float min_x = FLT_MAX, min_y = FLT_MAX, max_x = -FLT_MAX, max_y = -FLT_MAX;
  for (const Body &b : bodies) {
    min_x = rmin(min_x, b.pos.x);
    min_y = rmin(min_y, b.pos.y);
    max_x = rmax(max_x, b.pos.x);
    max_y = rmax(max_y, b.pos.y);
  }
  V2 center = v2(min_x + max_x, min_y + max_y) * 0.5f;
  float size = rmax(max_x - min_x, max_y - min_y);
  return Quad{center, size};
}
// quad.rs:38-43
static Quad quad_new_for_domain(float domain_width, float domain_height) {
  return Quad{v2(0, 0), rmax(2.0f * domain_width, 2.0f * domain_height)};
}
// quad.rs:45-50
static Quad quad_into_quadrant(Quad q, unsigned quadrant) {
  q.size *= 0.5f;
  q.center.x += ((float)(quadrant & 1) - 0.5f) * q.size;
  q.center.y += ((float)(quadrant >> 1) - 0.5f) * q.size;
  return q;
}

// ---------------------------------------------------------------- Node (src/quadtree/node.rs)
struct Node {
  size_t children, next;
  V2 pos;
  float mass;
  Quad quad;
  size_t b_start, b_end;
  float charge;
};
static const Node NODE_ZEROED = {0, 0, {0, 0}, 0.0f, {{0, 0}, 0.0f}, 0, 0, 0.0f};
static inline Node node_new(size_t next, Quad quad, size_t s, size_t e) {
  return Node{0, next, v2(0, 0), 0.0f, quad, s, e, 0.0f};
}
// Range<usize>: len() saturates at 0, is_empty() is start >= end
static inline size_t range_len(size_t s, size_t e) { return e > s ? e - s : 0; }

// ---------------------------------------------------------------- Partition (src/partition.rs:11-38)
template <class Pred>
static size_t partition_in_place(Body *a, size_t len, Pred pred) {
  if (len == 0) return 0;
  size_t l = 0, r = len - 1;
  for (;;) {
    while (l <= r && pred(a[l])) l += 1;
    while (l < r && !pred(a[r])) r -= 1;
    if (l >= r) return l;
    std::swap(a[l], a[r]);
    l += 1;
    r -= 1;  // r >= 1 here because l < r held before the swap
  }
}

// ---------------------------------------------------------------- Quadtree (src/quadtree/quadtree.rs)
struct Quadtree {
  float t_sq, e_sq;
  size_t leaf_capacity, thread_capacity;
  std::atomic<size_t> atomic_len{0};
  std::vector<Node> nodes;
  std::vector<size_t> parents;
  // diagnostics (not in the reference)
  std::atomic<uint32_t> flags{0};
  std::mutex grow_mutex;
  bool allow_grow = true;

  // quadtree.rs:40-101
  size_t subdivide(size_t node, std::vector<Body> &bodies, size_t r_start, size_t r_end) {
    V2 center = nodes[node].quad.center;
    if (r_start >= bodies.size() || r_end > bodies.size() || r_start >= r_end) return node;

    bool all_same_pos = true, all_identical = true;
    for (size_t i = r_start; i + 1 < r_end; ++i) {
      if (!(mag_sq(bodies[i].pos - bodies[i + 1].pos) < 1e-12f)) {
        all_same_pos = false;
        break;
      }
    }
    if (all_same_pos && r_end - r_start > 1) {
      for (size_t i = r_start; i + 1 < r_end; ++i)
        if (memcmp(&bodies[i].pos, &bodies[i + 1].pos, sizeof(V2)) != 0) all_identical = false;
      if (!all_identical) flags |= 4u;
    }
    if (all_same_pos || nodes[node].quad.size < 1e-6f || (r_end - r_start) <= 1) {
      if (r_end - r_start > 1) flags |= 2u;
      return node;
    }

    size_t split[5] = {r_start, 0, 0, 0, r_end};
    const float cy = center.y, cx = center.x;
    split[2] = split[0] + partition_in_place(&bodies[split[0]], split[4] - split[0],
                                             [cy](const Body &b) { return b.pos.y < cy; });
    split[1] = split[0] + partition_in_place(&bodies[split[0]], split[2] - split[0],
                                             [cx](const Body &b) { return b.pos.x < cx; });
    split[3] = split[2] + partition_in_place(&bodies[split[2]], split[4] - split[2],
                                             [cx](const Body &b) { return b.pos.x < cx; });

    size_t prev_len = atomic_len.fetch_add(1, std::memory_order_relaxed);
    size_t children = prev_len * 4 + 1;

    if (parents.size() <= prev_len || nodes.size() <= children + 3) {
      if (!allow_grow) {
        fprintf(stderr, "oracle: node arena exhausted in a multi-worker build\n");
        abort();
      }
      while (parents.size() <= prev_len) parents.resize((prev_len + 1) * 2, 0);
      while (nodes.size() <= children + 3) nodes.resize((children + 4) * 2, NODE_ZEROED);
    }
    parents[prev_len] = node;
    nodes[node].children = children;

    size_t nexts[4] = {children + 1, children + 2, children + 3, nodes[node].next};
    Quad pq = nodes[node].quad;
    for (unsigned i = 0; i < 4; ++i) {
      Quad q = quad_into_quadrant(pq, i);
      if (split[i] <= split[i + 1] && split[i + 1] <= bodies.size())
        nodes[children + i] = node_new(nexts[i], q, split[i], split[i + 1]);
      else
        nodes[children + i] = node_new(nexts[i], q, split[i], split[i]);
    }
    return children;
  }

  // quadtree.rs:103-151 (serial, reverse allocation order, re-reads the node's whole body range)
  void propagate(const std::vector<Body> &bodies) {
    size_t len = atomic_len.load();
    for (size_t k = len; k-- > 0;) {
      size_t node = parents[k];
      size_t i = nodes[node].children;
      size_t s = nodes[node].b_start, e = nodes[node].b_end;
      if (s >= bodies.size() || e > bodies.size() || s > e) continue;
#ifdef ORC_HP_AGG
      // variant (not the reference): the same three fallbacks with f64 running sums, used to
      // separate "summation-order noise in the node centres" from every other source of difference
      {
        double tm = 0, ta = 0, ax = 0, ay = 0, mx = 0, my = 0, cx = 0, cy = 0;
        for (size_t b = s; b < e; ++b) {
          const double q = fabs((double)bodies[b].charge), m = bodies[b].mass;
          const double x = bodies[b].pos.x, y = bodies[b].pos.y;
          tm += m, ta += q, ax += q * x, ay += q * y, mx += m * x, my += m * y, cx += x, cy += y;
        }
        V2 wp;
        if (ta > (double)1e-6f) wp = v2((float)(ax / ta), (float)(ay / ta));
        else if (tm > (double)1e-6f) wp = v2((float)(mx / tm), (float)(my / tm));
        else if (e - s > 0) wp = v2((float)(cx / (double)(e - s)), (float)(cy / (double)(e - s)));
        else wp = v2(0, 0);
        nodes[node].pos = wp;
        nodes[node].mass = nodes[i].mass + nodes[i + 1].mass + nodes[i + 2].mass + nodes[i + 3].mass;
        nodes[node].charge =
            nodes[i].charge + nodes[i + 1].charge + nodes[i + 2].charge + nodes[i + 3].charge;
        continue;
      }
#endif
      float total_mass = 0.0f;
      for (size_t b = s; b < e; ++b) total_mass += bodies[b].mass;
      float total_abs_charge = 0.0f;
      for (size_t b = s; b < e; ++b) total_abs_charge += fabsf(bodies[b].charge);
      V2 weighted_pos;
      if (total_abs_charge > 1e-6f) {
        V2 acc = v2(0, 0);
        for (size_t b = s; b < e; ++b) acc = acc + bodies[b].pos * fabsf(bodies[b].charge);
        weighted_pos = acc / total_abs_charge;
      } else if (total_mass > 1e-6f) {
        V2 acc = v2(0, 0);
        for (size_t b = s; b < e; ++b) acc = acc + bodies[b].pos * bodies[b].mass;
        weighted_pos = acc / total_mass;
      } else if (e - s > 0) {
        V2 acc = v2(0, 0);
        for (size_t b = s; b < e; ++b) acc = acc + bodies[b].pos;
        weighted_pos = acc / (float)(e - s);
      } else {
        weighted_pos = v2(0, 0);
      }
      nodes[node].pos = weighted_pos;
      nodes[node].mass = nodes[i].mass + nodes[i + 1].mass + nodes[i + 2].mass + nodes[i + 3].mass;
      nodes[node].charge =
          nodes[i].charge + nodes[i + 1].charge + nodes[i + 2].charge + nodes[i + 3].charge;
    }
  }

  // leaf aggregation, quadtree.rs:281-306
  void finish_leaf(size_t node, const std::vector<Body> &bodies) {
    size_t s = nodes[node].b_start, e = nodes[node].b_end;
    float total_mass = 0.0f, total_charge = 0.0f;
    V2 weighted_pos = v2(0, 0);
    if (s < bodies.size() && e <= bodies.size() && s <= e) {
      for (size_t b = s; b < e; ++b) {
        total_mass += bodies[b].mass;
        weighted_pos = weighted_pos + bodies[b].pos * bodies[b].charge;
        total_charge += bodies[b].charge;
      }
    }
    nodes[node].mass = total_mass;
    nodes[node].pos = fabsf(total_charge) > 1e-6f ? weighted_pos / total_charge : weighted_pos;
    nodes[node].charge = total_charge;
  }

  // One worker of build_internal (quadtree.rs:216-345).  `queue` stands for the crossbeam
  // channel; with a single worker this is a deterministic replay.
  struct Shared {
    std::deque<size_t> queue;
    std::mutex m;
    std::atomic<size_t> counter{0};
    std::vector<std::atomic<uint8_t>> claims;
    explicit Shared(size_t n) : claims(n) {
      for (auto &c : claims) c.store(0, std::memory_order_relaxed);
    }
    bool try_recv(size_t &out) {
      std::lock_guard<std::mutex> g(m);
      if (queue.empty()) return false;
      out = queue.front();
      queue.pop_front();
      return true;
    }
    void send(size_t v) {
      std::lock_guard<std::mutex> g(m);
      queue.push_back(v);
    }
    std::vector<uint8_t> extra_claims;  // nodes past the reference's fixed claim array (it would
                                        // panic there: more than 4N+1024 nodes); single worker only
    bool claim(size_t node) {
      if (node >= claims.size()) {
        const size_t k = node - claims.size();
        if (k >= extra_claims.size()) extra_claims.resize((k + 1) * 2, 0);
        if (extra_claims[k]) return false;
        extra_claims[k] = 1;
        return true;
      }
      uint8_t z = 0;
      return claims[node].compare_exchange_strong(z, 1);
    }
  };

  void worker(Shared &sh, std::vector<Body> &bodies) {
    std::vector<size_t> stack;
    size_t idle_iterations = 0;
    const size_t MAX_IDLE_ITERATIONS = 1000;
    for (;;) {
      if (sh.counter.load(std::memory_order_relaxed) >= bodies.size() ||
          idle_iterations > MAX_IDLE_ITERATIONS)
        break;
      bool work_done = false;
      size_t node;
      while (sh.try_recv(node)) {
        work_done = true;
        idle_iterations = 0;
        size_t s = nodes[node].b_start, e = nodes[node].b_end;
        size_t len = range_len(s, e);
        if (len >= thread_capacity) {  // quadtree.rs:249-274
          if (sh.claim(node)) {
            size_t children = subdivide(node, bodies, s, e);
            if (children != node) {
              for (unsigned i = 0; i < 4; ++i)
                if (range_len(nodes[children + i].b_start, nodes[children + i].b_end) != 0)
                  sh.send(children + i);
            } else {
              sh.counter.fetch_add(len, std::memory_order_relaxed);
            }
          }
          continue;
        }
        sh.counter.fetch_add(len, std::memory_order_relaxed);
        stack.push_back(node);
        while (!stack.empty()) {  // quadtree.rs:279-334
          size_t nd = stack.back();
          stack.pop_back();
          size_t ns = nodes[nd].b_start, ne = nodes[nd].b_end;
          if (range_len(ns, ne) <= leaf_capacity) {
            finish_leaf(nd, bodies);
            continue;
          }
          if (sh.claim(nd)) {
            // NB the reference does not test `children != node` here: a refused subdivision
            // makes it look at nodes nd..nd+3.  Replayed literally; it only re-visits nodes.
            size_t children = subdivide(nd, bodies, ns, ne);
            for (unsigned i = 0; i < 4; ++i)
              if (children + i < nodes.size() &&
                  range_len(nodes[children + i].b_start, nodes[children + i].b_end) != 0)
                stack.push_back(children + i);
          } else {
            size_t children = nodes[nd].children;
            if (children != 0)
              for (unsigned i = 0; i < 4; ++i)
                if (range_len(nodes[children + i].b_start, nodes[children + i].b_end) != 0)
                  stack.push_back(children + i);
          }
        }
      }
      if (!work_done) {
        idle_iterations += 1;
        std::this_thread::yield();
      }
    }
  }

  // quadtree.rs:197-348
  void build_internal(std::vector<Body> &bodies, size_t new_len, int threads) {
    Shared sh(new_len);
    sh.queue.push_back(0);
    if (threads <= 1) {
      allow_grow = true;
      worker(sh, bodies);
    } else {
      allow_grow = false;
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; ++t) pool.emplace_back([&] { worker(sh, bodies); });
      for (auto &t : pool) t.join();
      allow_grow = true;
    }
    propagate(bodies);
  }

  // quadtree.rs:153-195
  void build(std::vector<Body> &bodies, int mode, float hw, float hh, int threads) {
    flags = 0;
    atomic_len.store(0);
    if (bodies.empty()) return;
    size_t new_len = 4 * bodies.size() + 1024;
    // The reference keeps stale nodes from earlier builds (resize only grows, clear() only
    // resets atomic_len); a fresh oracle tree starts from ZEROED, which is what the first
    // build of a reference Quadtree sees.
    nodes.assign(new_len, NODE_ZEROED);
    parents.assign(new_len / 4, 0);
    Quad quad = mode == 0 ? quad_new_containing(bodies) : quad_new_for_domain(hw, hh);
    nodes[0] = node_new(0, quad, 0, bodies.size());
    build_internal(bodies, new_len, threads);
  }

  // quadtree.rs:350-407
  V2 acc_pos(V2 pos, float q, float radius, const std::vector<Body> &bodies, float k_e,
             OrcCounters *ctr) const {
    V2 acc = v2(0, 0);
    size_t node = 0;
    uint64_t V = 0, A = 0, P = 0;
    for (;;) {
      if (node >= nodes.size()) break;
      const Node n = nodes[node];
      V++;
      V2 d = pos - n.pos;
      float d_sq = mag_sq(d);
      float dist = sqrtf(d_sq);
      float node_radius = n.quad.size * 0.5f;
      float dist_adj = rmax(dist - radius, 0.0f);
      if (n.quad.size * n.quad.size < (dist_adj * dist_adj) * t_sq) {
        A++;
        float min_sep = radius + node_radius;
        float r_eff = rmax(dist, min_sep);
        float denom = (r_eff * r_eff + e_sq) * r_eff;
        acc = acc + d * (k_e * q * n.charge / denom);
        if (n.next == 0) break;
        node = n.next;
      } else if (n.children == 0) {
        for (size_t i = n.b_start; i < n.b_end; ++i) {
          const Body &body = bodies[i];
          if (mag_sq(body.pos - pos) < 1e-6f) continue;
          P++;
          V2 dd = pos - body.pos;
          float dist2 = mag(dd);
          float min_sep = radius + body.radius;
          float r_eff = rmax(dist2, min_sep);
          float denom = (r_eff * r_eff + e_sq) * r_eff;
          acc = acc + dd * rmin(k_e * q * body.charge / denom, FLT_MAX);
        }
        if (n.next == 0) break;
        node = n.next;
      } else {
        node = n.children;
      }
    }
    if (ctr) {
      ctr->visits += V;
      ctr->accepts += A;
      ctr->pairs += P;
    }
    return acc;
  }

  // quadtree.rs:430-501
  void find_neighbors_within(const std::vector<Body> &bodies, size_t i, float cutoff,
                             std::vector<size_t> &neighbors) const {
    neighbors.clear();
    if (nodes.empty()) return;
    if (i >= bodies.size() || !std::isfinite(cutoff) || cutoff <= 0.0f) return;
    V2 pos = bodies[i].pos;
    if (!std::isfinite(pos.x) || !std::isfinite(pos.y)) return;
    float cutoff_sq = cutoff * cutoff;
    std::vector<size_t> stack{0};
    while (!stack.empty()) {
      size_t node_idx = stack.back();
      stack.pop_back();
      if (node_idx >= nodes.size()) continue;
      const Node &node = nodes[node_idx];
      float half = node.quad.size * 0.5f;
      V2 mn = node.quad.center - v2(1, 1) * half;
      V2 mx = node.quad.center + v2(1, 1) * half;
      float d2 = 0.0f;
      for (int k = 0; k < 2; ++k) {
        float p = k == 0 ? pos.x : pos.y;
        float lo = k == 0 ? mn.x : mn.y;
        float hi = k == 0 ? mx.x : mx.y;
        if (p < lo) {
          float t = lo - p;
          d2 += t * t;
        } else if (p > hi) {
          float t = p - hi;
          d2 += t * t;
        }
      }
      if (d2 > cutoff_sq) continue;
      if (node.children == 0) {
        for (size_t idx = node.b_start; idx < node.b_end; ++idx) {
          if (idx < bodies.size() && idx != i) {
            V2 op = bodies[idx].pos;
            if (std::isfinite(op.x) && std::isfinite(op.y))
              if (mag_sq(op - pos) < cutoff_sq) neighbors.push_back(idx);
          }
        }
      } else {
        for (size_t c = 0; c < 4; ++c)
          if (node.children + c < nodes.size()) stack.push_back(node.children + c);
      }
    }
  }
};

// ---------------------------------------------------------------- CellList (src/cell_list.rs)
struct CellList {
  float domain_width = 0, domain_height = 0, cell_size = 1;
  size_t grid_size_x = 0, grid_size_y = 0;
  std::vector<std::vector<size_t>> cells;

  // cell_list.rs:27-39
  void rebuild(const std::vector<Body> &bodies) {
    grid_size_x = as_usize(ceilf((2.0f * domain_width) / cell_size)) + 1;
    grid_size_y = as_usize(ceilf((2.0f * domain_height) / cell_size)) + 1;
    cells.clear();
    cells.resize(grid_size_x * grid_size_y);
    for (size_t i = 0; i < bodies.size(); ++i) {
      size_t cx, cy;
      coord(bodies[i].pos, cx, cy);
      if (cx < grid_size_x && cy < grid_size_y) cells[cx + cy * grid_size_x].push_back(i);
    }
  }
  // cell_list.rs:47-55
  void coord(V2 pos, size_t &ox, size_t &oy) const {
    float min_x = -domain_width, min_y = -domain_height;
    int64_t x = as_isize(floorf((pos.x - min_x) / cell_size));
    int64_t y = as_isize(floorf((pos.y - min_y) / cell_size));
    x = std::min(std::max(x, (int64_t)0), (int64_t)grid_size_x - 1);
    y = std::min(std::max(y, (int64_t)0), (int64_t)grid_size_y - 1);
    ox = (size_t)x;
    oy = (size_t)y;
  }
  // cell_list.rs:57-85 (metals_only: cell_list.rs:92-127, returns matches in the same order)
  void find_neighbors_within(const std::vector<Body> &bodies, size_t i, float cutoff,
                             std::vector<size_t> &neighbors, bool metals_only = false) const {
    neighbors.clear();
    size_t cx, cy;
    coord(bodies[i].pos, cx, cy);
    int64_t range = as_isize(ceilf(cutoff / cell_size));
    float cutoff_sq = cutoff * cutoff;
    for (int64_t dy = -range; dy <= range; ++dy) {
      for (int64_t dx = -range; dx <= range; ++dx) {
        int64_t x = (int64_t)cx + dx, y = (int64_t)cy + dy;
        if (x < 0 || y < 0 || x >= (int64_t)grid_size_x || y >= (int64_t)grid_size_y) continue;
        size_t cell_idx = (size_t)x + (size_t)y * grid_size_x;
        for (size_t idx : cells[cell_idx]) {
          if (idx != i) {
            float r2 = mag_sq(bodies[idx].pos - bodies[i].pos);
            if (r2 < cutoff_sq) {
              if (!metals_only || bodies[idx].species == 1 || bodies[idx].species == 2)
                neighbors.push_back(idx);
            }
          }
        }
      }
    }
  }
};

}  // namespace

// ---------------------------------------------------------------- the simulation slice
struct OrcSim {
  std::vector<Body> bodies;
  // electrons beyond the inline pair are not modelled: every synthetic set carries <= 2
  Quadtree qt;
  CellList cl;
  OrcSpecies species[32];
  uint32_t n_species = 21;
};

namespace {

// src/species.rs:26-408 with constants from src/config.rs:40-48,106-126 and units.rs
void fill_default_species(OrcSpecies *t) {
  const double EV_TO_SIM = 1.602176634e-19 / (1.66053906660e-27 * 1.0e-10 * 1.0e-10 / (1.0e-15 * 1.0e-15));
  const float LJ_EPS = (float)((double)0.0103f * EV_TO_SIM);  // config.rs:113
  const float SIG = 1.80f, CUT = 2.2f;
  struct Row {
    float mass, radius, damping;
    int lj;
    float eps, polar_offset, polar_charge, rep_k, rep_cut;
  };
  const Row rows[21] = {
      /* LithiumIon       */ {6.94f, 0.76f, 1.0f, 0, 0.0f, 0.0f, 1.0f, 5.0f, 2.0f},
      /* LithiumMetal     */ {6.94f, 1.52f, 0.01f, 1, 0.1f, 1.0f, 1.0f, 5.0f, 2.0f},
      /* FoilMetal        */ {1.0e6f, 1.52f, 0.1f, 1, 10.0f, 1.0f, 1.0f, 5.0f, 2.0f},
      /* ElectrolyteAnion */ {145.0f, 2.0f, 1.0f, 0, 0.0f, 0.3f, 1.0f, 5.0f, 2.0f},
      /* EC               */ {88.06f, 2.5f, 1.0f, 0, 0.0f, 0.85f, 0.80f, 5.0f, 5.0f},
      /* DMC              */ {90.08f, 2.5f, 1.0f, 0, 0.0f, 0.60f, 0.20f, 5.0f, 5.0f},
      /* VC               */ {86.0f, 2.4f, 1.0f, 0, 0.0f, 0.85f, 0.80f, 5.0f, 5.0f},
      /* FEC              */ {107.0f, 2.5f, 0.8f, 0, 0.0f, 0.85f, 0.80f, 6.0f, 5.0f},
      /* EMC              */ {104.0f, 2.6f, 1.0f, 0, 0.0f, 0.60f, 0.20f, 4.5f, 5.5f},
      /* LLZO             */ {840.0f, 4.5f, 0.2f, 1, LJ_EPS, 0.20f, 0.05f, 5.0f, 2.0f},
      /* LLZT             */ {865.0f, 4.7f, 0.2f, 1, LJ_EPS, 0.20f, 0.06f, 5.0f, 2.0f},
      /* S40B             */ {340.0f, 4.2f, 0.25f, 1, LJ_EPS, 0.22f, 0.04f, 5.0f, 2.0f},
      /* SEI              */ {100.0f, 2.0f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* Graphite         */ {72.0f, 1.7f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* HardCarbon       */ {72.0f, 1.8f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* SiliconOxide     */ {60.0f, 2.0f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* LTO              */ {460.0f, 2.5f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* LFP              */ {158.0f, 2.2f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* LMFP             */ {158.0f, 2.2f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* NMC              */ {97.0f, 2.0f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
      /* NCA              */ {97.0f, 2.0f, 0.01f, 1, LJ_EPS, 0.0f, 0.0f, 5.0f, 2.0f},
  };
  for (int i = 0; i < 21; ++i) {
    t[i].mass = rows[i].mass;
    t[i].radius = rows[i].radius;
    t[i].damping = rows[i].damping;
    t[i].lj_enabled = rows[i].lj;
    t[i].lj_epsilon = rows[i].eps;
    t[i].lj_sigma = SIG;
    t[i].lj_cutoff = CUT;
    t[i].polar_offset = rows[i].polar_offset;
    t[i].polar_charge = rows[i].polar_charge;
    t[i].repulsion_enabled = 0;
    t[i].repulsion_strength = rows[i].rep_k;
    t[i].repulsion_cutoff = rows[i].rep_cut;
  }
}

inline const OrcSpecies &sp(const OrcSim *s, uint8_t k) { return s->species[k < 32 ? k : 0]; }

// species.rs:412-445
float max_lj_cutoff(const OrcSim *s) {
  float m = 0.0f;
  for (uint32_t i = 0; i < s->n_species; ++i)
    if (s->species[i].lj_enabled) m = rmax(m, s->species[i].lj_cutoff * s->species[i].lj_sigma);
  return m;
}
// species.rs:447-479
float max_repulsion_cutoff(const OrcSim *s) {
  float m = 0.0f;
  for (uint32_t i = 0; i < s->n_species; ++i)
    if (s->species[i].repulsion_enabled) m = rmax(m, s->species[i].repulsion_cutoff);
  return m;
}

void neighbors_of(const OrcSim *s, int use_cell, size_t i, float cutoff, std::vector<size_t> &out) {
  if (use_cell)
    s->cl.find_neighbors_within(s->bodies, i, cutoff, out);
  else
    s->qt.find_neighbors_within(s->bodies, i, cutoff, out);
}

int clamp_threads(int threads) {
#ifdef _OPENMP
  if (threads <= 0) return omp_get_max_threads();
  return threads;
#else
  (void)threads;
  return 1;
#endif
}

}  // namespace

extern "C" {

OrcSim *orc_create(float theta, float epsilon, uint64_t leaf_capacity, uint64_t thread_capacity) {
  OrcSim *s = new OrcSim();
  // Quadtree::new, quadtree.rs:24-34
  s->qt.t_sq = theta * theta;
  s->qt.e_sq = epsilon * epsilon;
  s->qt.leaf_capacity = leaf_capacity;
  s->qt.thread_capacity = thread_capacity;
  fill_default_species(s->species);
  s->n_species = 21;
  return s;
}
void orc_destroy(OrcSim *s) { delete s; }

void orc_set_species_table(OrcSim *s, const OrcSpecies *rows, uint32_t nrows) {
  if (nrows > 32) nrows = 32;
  memcpy(s->species, rows, nrows * sizeof(OrcSpecies));
  s->n_species = nrows;
}
void orc_default_species_table(OrcSpecies *rows21) { fill_default_species(rows21); }

void orc_set_bodies(OrcSim *s, uint64_t n, const float *pos_xy, const float *z, const float *vel_xy,
                    const float *vz, const float *mass, const float *radius, const float *charge,
                    const uint8_t *species) {
  s->bodies.assign(n, Body{});
  for (uint64_t i = 0; i < n; ++i) {
    Body &b = s->bodies[i];
    memset(&b, 0, sizeof(Body));
    b.pos = v2(pos_xy[2 * i], pos_xy[2 * i + 1]);
    b.z = z ? z[i] : 0.0f;
    b.vel = vel_xy ? v2(vel_xy[2 * i], vel_xy[2 * i + 1]) : v2(0, 0);
    b.vz = vz ? vz[i] : 0.0f;
    b.mass = mass ? mass[i] : 1.0f;
    b.radius = radius ? radius[i] : 0.0f;
    b.charge = charge ? charge[i] : 0.0f;
    b.species = species ? species[i] : 0;
    b.id = i;
    b.species_lock_until = -INFINITY;
    b.last_surround_pos = b.pos;  // Body::new, body/types.rs:111-113
    b.last_surround_frame = 0;
  }
}

// positions only (the rest of the state, e.g. the surround bookkeeping, is kept)
void orc_set_positions(OrcSim *s, const float *pos_xy) {
  for (size_t i = 0; i < s->bodies.size(); ++i) s->bodies[i].pos = v2(pos_xy[2 * i], pos_xy[2 * i + 1]);
}

void orc_set_electrons(OrcSim *s, uint64_t m, const uint32_t *body, const float *rel_xy,
                       const float *vel_xy) {
  for (Body &b : s->bodies) b.n_electrons = 0;
  for (uint64_t k = 0; k < m; ++k) {
    Body &b = s->bodies[body[k]];
    if (b.n_electrons >= 2) {
      fprintf(stderr, "oracle: more than 2 electrons on one body is not modelled\n");
      abort();
    }
    Electron &e = b.electrons[b.n_electrons++];
    e.rel_pos = v2(rel_xy[2 * k], rel_xy[2 * k + 1]);
    e.vel = vel_xy ? v2(vel_xy[2 * k], vel_xy[2 * k + 1]) : v2(0, 0);
  }
}

uint64_t orc_num_bodies(const OrcSim *s) { return s->bodies.size(); }
uint64_t orc_num_electrons(const OrcSim *s) {
  uint64_t m = 0;
  for (const Body &b : s->bodies) m += b.n_electrons;
  return m;
}

void orc_get_bodies(const OrcSim *s, uint64_t *id, float *pos_xy, float *z, float *vel_xy, float *vz,
                    float *acc_xy, float *az, float *mass, float *radius, float *charge,
                    uint8_t *species, float *e_field_xy) {
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    const Body &b = s->bodies[i];
    if (id) id[i] = b.id;
    if (pos_xy) pos_xy[2 * i] = b.pos.x, pos_xy[2 * i + 1] = b.pos.y;
    if (z) z[i] = b.z;
    if (vel_xy) vel_xy[2 * i] = b.vel.x, vel_xy[2 * i + 1] = b.vel.y;
    if (vz) vz[i] = b.vz;
    if (acc_xy) acc_xy[2 * i] = b.acc.x, acc_xy[2 * i + 1] = b.acc.y;
    if (az) az[i] = b.az;
    if (mass) mass[i] = b.mass;
    if (radius) radius[i] = b.radius;
    if (charge) charge[i] = b.charge;
    if (species) species[i] = b.species;
    if (e_field_xy) e_field_xy[2 * i] = b.e_field.x, e_field_xy[2 * i + 1] = b.e_field.y;
  }
}

void orc_get_electrons(const OrcSim *s, uint32_t *body, float *rel_xy, float *vel_xy) {
  uint64_t k = 0;
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    const Body &b = s->bodies[i];
    for (unsigned e = 0; e < b.n_electrons; ++e, ++k) {
      if (body) body[k] = (uint32_t)i;
      if (rel_xy) rel_xy[2 * k] = b.electrons[e].rel_pos.x, rel_xy[2 * k + 1] = b.electrons[e].rel_pos.y;
      if (vel_xy) vel_xy[2 * k] = b.electrons[e].vel.x, vel_xy[2 * k + 1] = b.electrons[e].vel.y;
    }
  }
}

void orc_build(OrcSim *s, int mode, float hw, float hh, int threads) {
  s->qt.build(s->bodies, mode, hw, hh, threads <= 1 ? 1 : threads);
}

uint64_t orc_num_nodes(const OrcSim *s) {
  if (s->bodies.empty()) return 0;
  return s->qt.atomic_len.load() * 4 + 1;
}

void orc_get_nodes(const OrcSim *s, OrcNode *out) {
  uint64_t m = orc_num_nodes(s);
  for (uint64_t i = 0; i < m; ++i) {
    const Node &n = s->qt.nodes[i];
    OrcNode &o = out[i];
    memset(&o, 0, sizeof(o));
    o.children = n.children;
    o.next = n.next;
    o.pos[0] = n.pos.x, o.pos[1] = n.pos.y;
    o.mass = n.mass;
    o.quad_center[0] = n.quad.center.x, o.quad_center[1] = n.quad.center.y;
    o.quad_size = n.quad.size;
    o.bodies_start = n.b_start, o.bodies_end = n.b_end;
    o.charge = n.charge;
  }
}

static void canon_walk(const Quadtree &qt, uint64_t &count, OrcCanon *out, uint64_t cap,
                       uint32_t &max_depth) {
  struct Item {
    size_t node;
    uint32_t depth;
    uint64_t hi, lo;
  };
  std::vector<Item> stack;
  stack.push_back({0, 0, 0, 0});
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    const Node &n = qt.nodes[it.node];
    if (it.depth > max_depth) max_depth = it.depth;
    if (out && count < cap) {
      OrcCanon &c = out[count];
      memset(&c, 0, sizeof(c));
      c.path_hi = it.hi, c.path_lo = it.lo;
      c.depth = it.depth;
      c.is_leaf = n.children == 0;
      c.start = n.b_start, c.end = n.b_end;
      c.pos[0] = n.pos.x, c.pos[1] = n.pos.y;
      c.mass = n.mass, c.charge = n.charge;
      c.quad_center[0] = n.quad.center.x, c.quad_center[1] = n.quad.center.y;
      c.quad_size = n.quad.size;
    }
    count++;
    if (n.children != 0) {
      for (int q = 3; q >= 0; --q) {
        Item ch{n.children + (size_t)q, it.depth + 1, it.hi, it.lo};
        uint32_t level = it.depth + 1;  // 1-based level of the child
        if (level <= 32)
          ch.lo |= (uint64_t)q << (64 - 2 * level);
        else if (level <= 64)
          ch.hi |= (uint64_t)q << (64 - 2 * (level - 32));
        stack.push_back(ch);
      }
    }
  }
}

uint64_t orc_canonical(const OrcSim *s, OrcCanon *out, uint64_t cap) {
  if (s->bodies.empty()) return 0;
  uint64_t count = 0;
  uint32_t md = 0;
  canon_walk(s->qt, count, out, cap, md);
  return count;
}

// Overwrite node centres, given in canonical DFS pre-order (test hook: lets a test run the
// reference traversal on a tree whose centres were computed elsewhere).
void orc_set_canonical_pos(OrcSim *s, const float *pos_xy, uint64_t count) {
  if (s->bodies.empty()) return;
  std::vector<size_t> stack{0};
  uint64_t k = 0;
  while (!stack.empty() && k < count) {
    size_t nd = stack.back();
    stack.pop_back();
    s->qt.nodes[nd].pos = v2(pos_xy[2 * k], pos_xy[2 * k + 1]);
    ++k;
    size_t c = s->qt.nodes[nd].children;
    if (c != 0)
      for (int q = 3; q >= 0; --q) stack.push_back(c + (size_t)q);
  }
}

uint32_t orc_max_depth(const OrcSim *s) {
  if (s->bodies.empty()) return 0;
  uint64_t count = 0;
  uint32_t md = 0;
  canon_walk(s->qt, count, nullptr, 0, md);
  return md;
}

uint32_t orc_flags(const OrcSim *s) {
  uint32_t f = s->qt.flags.load();
  if (!s->bodies.empty() && orc_max_depth(s) > 32) f |= 1u;
  if (!s->bodies.empty() && orc_num_nodes(s) > 4 * s->bodies.size() + 1024) f |= 8u;
  return f;
}

void orc_field(OrcSim *s, float k_e, int threads, OrcCounters *ctr) {
  const int nt = clamp_threads(threads);
  const int64_t n = (int64_t)s->bodies.size();
  uint64_t V = 0, A = 0, P = 0;
  (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt) reduction(+ : V, A, P)
  for (int64_t i = 0; i < n; ++i) {
    OrcCounters c{0, 0, 0};
    Body &b = s->bodies[i];
    b.e_field = s->qt.acc_pos(b.pos, 1.0f, b.radius, s->bodies, k_e, &c);  // quadtree.rs:418-427
    V += c.visits, A += c.accepts, P += c.pairs;
  }
  if (ctr) ctr->visits += V, ctr->accepts += A, ctr->pairs += P;
}

void orc_acc_points(const OrcSim *s, uint64_t m, const float *pts_xy, const float *q,
                    const float *radius, float k_e, float *out_xy, int threads, OrcCounters *ctr) {
  const int nt = clamp_threads(threads);
  uint64_t V = 0, A = 0, P = 0;
  (void)nt;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt) reduction(+ : V, A, P)
  for (int64_t i = 0; i < (int64_t)m; ++i) {
    OrcCounters c{0, 0, 0};
    V2 r = s->qt.acc_pos(v2(pts_xy[2 * i], pts_xy[2 * i + 1]), q ? q[i] : 1.0f,
                         radius ? radius[i] : 0.0f, s->bodies, k_e, &c);
    out_xy[2 * i] = r.x, out_xy[2 * i + 1] = r.y;
    V += c.visits, A += c.accepts, P += c.pairs;
  }
  if (ctr) ctr->visits += V, ctr->accepts += A, ctr->pairs += P;
}

uint64_t orc_tree_neighbors(const OrcSim *s, uint64_t i, float cutoff, uint64_t *out, uint64_t cap) {
  std::vector<size_t> nb;
  s->qt.find_neighbors_within(s->bodies, i, cutoff, nb);
  for (size_t k = 0; k < nb.size() && k < cap; ++k) out[k] = nb[k];
  return nb.size();
}

void orc_cell_set_domain(OrcSim *s, float hw, float hh) {
  s->cl.domain_width = hw;  // update_domain_size, cell_list.rs:41-45
  s->cl.domain_height = hh;
}
void orc_cell_rebuild(OrcSim *s, float cell_size) {
  s->cl.cell_size = cell_size;
  s->cl.rebuild(s->bodies);
}
void orc_cell_dims(const OrcSim *s, uint64_t *gx, uint64_t *gy) {
  *gx = s->cl.grid_size_x;
  *gy = s->cl.grid_size_y;
}
uint64_t orc_cell_contents(const OrcSim *s, uint64_t cell, uint64_t *out, uint64_t cap) {
  const auto &c = s->cl.cells[cell];
  for (size_t k = 0; k < c.size() && k < cap; ++k) out[k] = c[k];
  return c.size();
}
uint64_t orc_cell_neighbors(const OrcSim *s, uint64_t i, float cutoff, uint64_t *out, uint64_t cap) {
  std::vector<size_t> nb;
  s->cl.find_neighbors_within(s->bodies, i, cutoff, nb);
  for (size_t k = 0; k < nb.size() && k < cap; ++k) out[k] = nb[k];
  return nb.size();
}
uint64_t orc_cell_metal_neighbor_count(const OrcSim *s, uint64_t i, float cutoff) {
  std::vector<size_t> nb;
  s->cl.find_neighbors_within(s->bodies, i, cutoff, nb, true);
  return nb.size();
}

// simulation.rs:1798-1802
int orc_use_cell_list(const OrcSim *s, float hw, float hh, float density_threshold) {
  float area = (2.0f * hw) * (2.0f * hh);
  float density = (float)s->bodies.size() / area;
  return density > density_threshold;
}

// simulation.rs:1000-1003
void orc_reset_acc(OrcSim *s) {
  for (Body &b : s->bodies) {
    b.acc = v2(0, 0);
    b.az = 0.0f;
  }
}

// forces.rs:14-25
void orc_prepare_spatial_structures(OrcSim *s, float hw, float hh, float density_threshold,
                                    int threads) {
  s->qt.build(s->bodies, 0, 0, 0, threads <= 1 ? 1 : threads);
  if (orc_use_cell_list(s, hw, hh, density_threshold)) {
    float lj_cutoff = max_lj_cutoff(s);
    float repulsion_cutoff = max_repulsion_cutoff(s);
    float polar_cutoff = 3.0f * lj_cutoff;
    float max_cutoff = rmax(rmax(polar_cutoff, repulsion_cutoff), lj_cutoff);
    s->cl.domain_width = hw;
    s->cl.domain_height = hh;
    s->cl.cell_size = max_cutoff;
    s->cl.rebuild(s->bodies);
  }
}

// forces.rs:33-44
void orc_attract(OrcSim *s, float k_e, float bg_x, float bg_y, int threads) {
  orc_field(s, k_e, threads, nullptr);
  V2 bg = v2(bg_x, bg_y);
  for (Body &b : s->bodies) b.e_field = b.e_field + bg;
  for (Body &b : s->bodies) b.acc = (b.charge * b.e_field) / b.mass;
}

// forces.rs:52-175 (serial).  dipole_model 0 = SingleOffset, 1 = ConjugatePair (the default,
// config.rs:278-281).  epsilon is config::QUADTREE_EPSILON (config.rs:214), not the tree's.
void orc_apply_polar_forces(OrcSim *s, int use_cell, float k_e, int dipole_model) {
  std::vector<Body> &bodies = s->bodies;
  if (bodies.empty()) return;
  const float epsilon_sq = 2.0f * 2.0f;
  std::vector<size_t> neighbors;
  auto field_from_source = [&](V2 point, float point_radius, V2 src_pos, float src_radius, float src_charge) -> V2 {
    if (fabsf(src_charge) < FLT_EPSILON) return v2(0, 0);
    V2 d = point - src_pos;
    float dist = mag(d);
    float min_sep = point_radius + src_radius;
    float r_eff = rmax(dist, min_sep);
    // DOCUMENTED DIVERGENCE from forces.rs:91-96: two zero-radius sites at exactly the same place give
    // d * (k q / 0) = 0 * inf = NaN in the reference, after which its next propagate() makes every node centre NaN.
    // At 16 M bodies f32 positions sit on a ~1e-3 A grid and this happens a few times per step, so the bounded CPU
    // arm of the benchmark could never finish a step.  The term is taken as the zero vector d already is.
    if (r_eff == 0.0f) return v2(0, 0);
    float denom = (r_eff * r_eff + epsilon_sq) * r_eff;
    return d * (k_e * src_charge / denom);
  };
  for (size_t i = 0; i < bodies.size(); ++i) {
    if (!(bodies[i].species == 4 || bodies[i].species == 5)) continue;  // EC | DMC
    if (bodies[i].n_electrons == 0) continue;
    V2 e_pos = bodies[i].pos + bodies[i].electrons[0].rel_pos;
    float cutoff = 3.0f * bodies[i].radius;
    neighbors_of(s, use_cell, i, cutoff, neighbors);
    for (size_t j : neighbors) {
      V2 i_nuc_pos = bodies[i].pos;
      float i_nuc_rad = bodies[i].radius;
      V2 i_ele_pos = e_pos;
      V2 j_pos = bodies[j].pos;
      float j_rad = bodies[j].radius;
      float j_q = bodies[j].charge;
      bool j_has_dipole = (bodies[j].species == 4 || bodies[j].species == 5) && bodies[j].n_electrons != 0;
      V2 j_e_pos = j_has_dipole ? j_pos + bodies[j].electrons[0].rel_pos : j_pos;
      float j_q_eff = j_has_dipole ? sp(s, bodies[j].species).polar_charge : 0.0f;
      V2 fnuc = v2(0, 0), fele = v2(0, 0);
      fnuc = fnuc + field_from_source(i_nuc_pos, i_nuc_rad, j_pos, j_rad, j_q);
      fele = fele + field_from_source(i_ele_pos, 0.0f, j_pos, j_rad, j_q);
      if (dipole_model == 1 && j_has_dipole) {
        fnuc = fnuc + field_from_source(i_nuc_pos, i_nuc_rad, j_pos, j_rad, j_q_eff);
        fnuc = fnuc - field_from_source(i_nuc_pos, i_nuc_rad, j_e_pos, 0.0f, j_q_eff);
        fele = fele + field_from_source(i_ele_pos, 0.0f, j_pos, j_rad, j_q_eff);
        fele = fele - field_from_source(i_ele_pos, 0.0f, j_e_pos, 0.0f, j_q_eff);
      }
      if (is_zero(fnuc) && is_zero(fele)) continue;
      float q_eff_i = sp(s, bodies[i].species).polar_charge;
      V2 force = (fnuc - fele) * q_eff_i;
      Body &a = bodies[i];
      Body &b = bodies[j];
      a.acc = a.acc + force / a.mass;
      b.acc = b.acc - force / b.mass;
    }
  }
}

// forces.rs:182-231 (serial)
void orc_apply_lj_forces(OrcSim *s, int use_cell, float lj_force_max, uint32_t collision_passes) {
  float max_cutoff = max_lj_cutoff(s);
  std::vector<size_t> neighbors;
  std::vector<Body> &bodies = s->bodies;
  for (size_t i = 0; i < bodies.size(); ++i) {
    if (!sp(s, bodies[i].species).lj_enabled) continue;
    neighbors_of(s, use_cell, i, max_cutoff, neighbors);
    for (size_t j : neighbors) {
      if (j <= i) continue;
      if (!sp(s, bodies[i].species).lj_enabled || !sp(s, bodies[j].species).lj_enabled) continue;
      Body &a = bodies[i];
      Body &b = bodies[j];
      const OrcSpecies &sa = sp(s, a.species), &sb = sp(s, b.species);
      float sigma = (sa.lj_sigma + sb.lj_sigma) * 0.5f;
      float epsilon = sqrtf(sa.lj_epsilon * sb.lj_epsilon);
      float cutoff = 0.5f * (sa.lj_cutoff * sa.lj_sigma + sb.lj_cutoff * sb.lj_sigma);
      V2 r_vec = b.pos - a.pos;
      float r = mag(r_vec);
      if (r < cutoff && r > 1e-6f) {
        float x = sigma / r;
        float x2 = x * x, x4 = x2 * x2;
        float sr6 = x2 * x4;  // f32::powi(6): square-and-multiply, x^2 * x^4
        float max_lj_force = (float)collision_passes * lj_force_max;
        float unclamped = 24.0f * epsilon * (2.0f * sr6 * sr6 - sr6) / r;
        // f32::clamp(min, max)
        float force_mag = unclamped;
        if (force_mag < -max_lj_force) force_mag = -max_lj_force;
        if (force_mag > max_lj_force) force_mag = max_lj_force;
        V2 force = force_mag * normalized(r_vec);
        a.acc = a.acc - force / a.mass;
        b.acc = b.acc + force / b.mass;
      }
    }
  }
}

// forces.rs:234-247
static V2 compute_repulsive_force(const OrcSim *s, const Body &p1, const Body &p2, V2 r_vec, float r) {
  float r0 = 0.5f * (sp(s, p1.species).repulsion_cutoff + sp(s, p2.species).repulsion_cutoff);
  if (r >= r0 || r <= 0.0f) return v2(0, 0);
  float k = 0.5f * (sp(s, p1.species).repulsion_strength + sp(s, p2.species).repulsion_strength);
  float m = k * (1.0f - r / r0) / r;
  return r_vec * m;
}

// forces.rs:250-289 (serial)
void orc_apply_repulsive_forces(OrcSim *s, int use_cell) {
  float max_cutoff = max_repulsion_cutoff(s);
  if (max_cutoff <= 0.0f) return;
  std::vector<size_t> neighbors;
  std::vector<Body> &bodies = s->bodies;
  for (size_t i = 0; i < bodies.size(); ++i) {
    if (!sp(s, bodies[i].species).repulsion_enabled) continue;
    float cutoff = sp(s, bodies[i].species).repulsion_cutoff;
    neighbors_of(s, use_cell, i, cutoff, neighbors);
    for (size_t j : neighbors) {
      if (j <= i) continue;
      if (!sp(s, bodies[j].species).repulsion_enabled) continue;
      V2 r_vec = bodies[j].pos - bodies[i].pos;
      float r = mag(r_vec);
      V2 f = compute_repulsive_force(s, bodies[i], bodies[j], r_vec, r);
      if (!is_zero(f)) {
        Body &a = bodies[i];
        Body &b = bodies[j];
        a.acc = a.acc - f / a.mass;
        b.acc = b.acc + f / b.mass;
      }
    }
  }
}

// forces.rs:294-321
void orc_apply_stack_pressure(OrcSim *s, int enabled, float pressure, float decay, float hw) {
  if (!enabled || pressure <= 0.0f) return;
  float x_min = -hw, x_max = hw;
  for (Body &body : s->bodies) {
    float dist_left = body.pos.x - x_min;
    if (dist_left < decay && dist_left > 0.0f) {
      float force = pressure * (1.0f - dist_left / decay);
      body.acc.x += force / body.mass;
    }
    float dist_right = x_max - body.pos.x;
    if (dist_right < decay && dist_right > 0.0f) {
      float force = pressure * (1.0f - dist_right / decay);
      body.acc.x -= force / body.mass;
    }
  }
}

// simulation.rs:1437-1486
void orc_iterate(OrcSim *s, float dt, float damping_base, float hw, float hh, float hd,
                 int enable_out_of_plane, int threads) {
  const int nt = clamp_threads(threads);
  (void)nt;
  float base_damping = powf(damping_base, dt / 0.01f);
  const int64_t n = (int64_t)s->bodies.size();
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int64_t i = 0; i < n; ++i) {
    Body &body = s->bodies[i];
    body.vel = body.vel + body.acc * dt;
    float damping = base_damping * sp(s, body.species).damping;
    body.vel = body.vel * damping;
    body.pos = body.pos + body.vel * dt;
    if (enable_out_of_plane) {
      body.vz += body.az * dt;
      body.vz *= damping;
      body.z += body.vz * dt;
      if (body.z < -hd) {
        body.z = -hd;
        body.vz = -body.vz;
      } else if (body.z > hd) {
        body.z = hd;
        body.vz = -body.vz;
      }
    }
    if (body.pos.x < -hw) {
      body.pos.x = -hw;
      body.vel.x = -body.vel.x;
    } else if (body.pos.x > hw) {
      body.pos.x = hw;
      body.vel.x = -body.vel.x;
    }
    if (body.pos.y < -hh) {
      body.pos.y = -hh;
      body.vel.y = -body.vel.y;
    } else if (body.pos.y > hh) {
      body.pos.y = hh;
      body.vel.y = -body.vel.y;
    }
  }
}

// body/electron.rs:19-46 driven by simulation.rs:1186-1196 (serial in the reference)
void orc_update_electrons(OrcSim *s, float bg_x, float bg_y, float dt, float k_e, int threads) {
  const int nt = threads <= 1 ? 1 : clamp_threads(threads);
  (void)nt;
  const V2 background_field = v2(bg_x, bg_y);
  const float ELECTRON_SPRING_K = 5.0f;          // config.rs:6-9,27-35: every species maps to 5.0
  const float ELECTRON_MAX_SPEED_FACTOR = 10.2f;  // config.rs:49
  const int64_t n = (int64_t)s->bodies.size();
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
  for (int64_t i = 0; i < n; ++i) {
    Body &self = s->bodies[i];
    float k = ELECTRON_SPRING_K;
    for (unsigned ei = 0; ei < self.n_electrons; ++ei) {
      Electron &e = self.electrons[ei];
      V2 electron_pos = self.pos + e.rel_pos;
      V2 local_field =
          s->qt.acc_pos(electron_pos, 1.0f, 0.0f, s->bodies, k_e, nullptr) + background_field;
      V2 acc = (-local_field) * k;
      e.vel = e.vel + acc * dt;
      float speed = mag(e.vel);
      float max_speed = ELECTRON_MAX_SPEED_FACTOR * self.radius / dt;
      if (speed > max_speed) e.vel = e.vel / speed * max_speed;
      e.rel_pos = e.rel_pos + e.vel * dt;
      float max_dist = sp(s, self.species).polar_offset * self.radius;
      if (mag(e.rel_pos) > max_dist) e.rel_pos = normalized(e.rel_pos) * max_dist;
    }
  }
}

// simulation.rs:1893-1918 (update_surrounded_flags) with body/types.rs:243-286 (maybe_update_surrounded)
// and cell_list.rs:92-127 (metal_neighbor_count).  radius_factor / neighbor_threshold are the runtime
// values of renderer/state.rs:30-36 (defaults config.rs:182-184); the move threshold and the check
// interval are config.rs:186-188.
void orc_update_surrounded_flags(OrcSim *s, float hw, float hh, float density_threshold, uint64_t frame,
                                 float radius_factor, uint64_t neighbor_threshold) {
  const float SURROUND_MOVE_THRESHOLD = 0.5f;
  const uint64_t SURROUND_CHECK_INTERVAL = 10;
  if (s->bodies.empty()) return;
  const int use_cell = orc_use_cell_list(s, hw, hh, density_threshold);
  const float neighbor_radius = max_lj_cutoff(s);
  if (use_cell) {
    s->cl.domain_width = hw;
    s->cl.domain_height = hh;
    s->cl.cell_size = neighbor_radius;
    s->cl.rebuild(s->bodies);
  } else {
    s->qt.build(s->bodies, 1, hw, hh, 1);
  }
  std::vector<size_t> nb;
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    Body &self = s->bodies[i];
    const bool moved = mag(self.pos - self.last_surround_pos) > SURROUND_MOVE_THRESHOLD * self.radius;
    const uint64_t frame_diff =
        frame >= self.last_surround_frame ? frame - self.last_surround_frame : SURROUND_CHECK_INTERVAL;
    if (moved || frame_diff >= SURROUND_CHECK_INTERVAL) {
      const float radius = self.radius * radius_factor;
      uint64_t count = 0;
      if (use_cell) {
        s->cl.find_neighbors_within(s->bodies, i, radius, nb, true);
        count = nb.size();
      } else {
        s->qt.find_neighbors_within(s->bodies, i, radius, nb);
        for (size_t j : nb)
          if (s->bodies[j].species == 1 || s->bodies[j].species == 2) ++count;
      }
      self.surrounded_by_metal = count >= neighbor_threshold;
      self.last_surround_pos = self.pos;
      self.last_surround_frame = frame;
    }
  }
}
void orc_get_surrounded(const OrcSim *s, uint8_t *flags, float *last_pos_xy, uint64_t *last_frame) {
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    if (flags) flags[i] = s->bodies[i].surrounded_by_metal ? 1 : 0;
    if (last_pos_xy) last_pos_xy[2 * i] = s->bodies[i].last_surround_pos.x, last_pos_xy[2 * i + 1] = s->bodies[i].last_surround_pos.y;
    if (last_frame) last_frame[i] = s->bodies[i].last_surround_frame;
  }
}

// simulation/out_of_plane.rs:140-254 (enforce_metal_z_boundaries)
void orc_enforce_metal_z_boundaries(OrcSim *s, float max_z, float hw, float hh, float density_threshold) {
  if (!std::isfinite(max_z) || max_z <= 0.0f) return;
  bool any_metal = false;
  for (const Body &b : s->bodies)
    if (b.species == 1 || b.species == 2) any_metal = true;
  if (!any_metal) return;
  const int use_cell = orc_use_cell_list(s, hw, hh, density_threshold);
  const float metal_max_r = rmax(sp(s, 1).radius, sp(s, 2).radius);
  if (use_cell) {
    s->cl.domain_width = hw;
    s->cl.domain_height = hh;
    s->cl.cell_size = 4.0f * metal_max_r;
    s->cl.rebuild(s->bodies);
  } else {
    s->qt.build(s->bodies, 0, 0, 0, 1);
  }
  std::vector<size_t> neighbors;
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    if (s->bodies[i].species == 1 || s->bodies[i].species == 2) continue;
    const V2 body_pos = s->bodies[i].pos;
    const float body_radius = s->bodies[i].radius;
    const float cutoff = 3.0f * body_radius + metal_max_r;
    neighbors_of(s, use_cell, i, cutoff, neighbors);
    float min_z_constraint = -max_z, max_z_constraint = max_z;
    int constraints_applied = 0;
    for (size_t j : neighbors) {
      if (!(s->bodies[j].species == 1 || s->bodies[j].species == 2)) continue;
      if (j == i) continue;
      constraints_applied += 1;
      if (constraints_applied > 5) break;
      const V2 metal_pos = s->bodies[j].pos;
      const float metal_radius = s->bodies[j].radius;
      const float dx = body_pos.x - metal_pos.x, dy = body_pos.y - metal_pos.y;
      const float distance_sq = dx * dx + dy * dy;
      const float reach = body_radius + metal_radius + 2.0f * body_radius;
      const float thresh = reach * reach;  // powi(2)
      if (distance_sq > thresh) continue;
      const float distance_2d = std::sqrt(distance_sq);
      if (distance_2d < body_radius + metal_radius + 2.0f * body_radius) {
        const float metal_z = s->bodies[j].z;
        const float lower_bound = metal_z - metal_radius - 0.01f, upper_bound = metal_z + metal_radius + 0.01f;
        if (lower_bound < upper_bound) {
          min_z_constraint = rmax(min_z_constraint, lower_bound);
          max_z_constraint = rmin(max_z_constraint, upper_bound);
        }
      }
    }
    if (min_z_constraint > max_z_constraint) min_z_constraint = 0.0f - 0.1f, max_z_constraint = 0.0f + 0.1f;
    Body &body = s->bodies[i];
    if (body.z < min_z_constraint) {
      body.z = min_z_constraint;
      if (body.vz < 0.0f) body.vz = 0.0f;
    }
    if (body.z > max_z_constraint) {
      body.z = max_z_constraint;
      if (body.vz > 0.0f) body.vz = 0.0f;
    }
    if (body.z > max_z) body.z = max_z, body.vz = 0.0f;  // Body::clamp_z, body/types.rs:296-304
    else if (body.z < -max_z) body.z = -max_z, body.vz = 0.0f;
  }
}

// bin/physics_invariants.rs:2430-2450, generalised with a target radius
void orc_direct_f64(const OrcSim *s, uint64_t m, const float *pts_xy, const float *target_radius,
                    double k_e, double epsilon, double *out_xy, int threads) {
  const int nt = clamp_threads(threads);
  (void)nt;
  const double e_sq = epsilon * epsilon;
  const size_t n = s->bodies.size();
  std::vector<double> sx(n), sy(n), sq(n), sr(n);
  for (size_t j = 0; j < n; ++j) {
    sx[j] = s->bodies[j].pos.x, sy[j] = s->bodies[j].pos.y;
    sq[j] = s->bodies[j].charge, sr[j] = s->bodies[j].radius;
  }
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
  for (int64_t i = 0; i < (int64_t)m; ++i) {
    const float pxf = pts_xy[2 * i], pyf = pts_xy[2 * i + 1];
    const double tr = target_radius ? (double)target_radius[i] : 0.0;
    double fx = 0.0, fy = 0.0;
    for (size_t j = 0; j < n; ++j) {
      // the skip test is evaluated on the f32 difference, like the reference
      float dxf = pxf - (float)sx[j], dyf = pyf - (float)sy[j];
      double msq = (double)((dxf * dxf) + (dyf * dyf));
      if (msq < 1e-6) continue;
      double dx = (double)dxf, dy = (double)dyf;
      double dist = sqrt(msq);
      double r_eff = std::max(dist, tr + sr[j]);
      double denom = (r_eff * r_eff + e_sq) * r_eff;
      fx += k_e * sq[j] * dx / denom;
      fy += k_e * sq[j] * dy / denom;
    }
    out_xy[2 * i] = fx, out_xy[2 * i + 1] = fy;
  }
}

// ---- collision.rs ---------------------------------------------------------------------------------
namespace {
struct CollideCfg {
  float softness;
  bool soft_li, soft_an;
  float domain_depth;
  uint32_t num_passes;
};
inline bool finite_f(float v) { return std::isfinite(v); }
// apply_collision_modifiers, collision.rs:16-60
void collision_modifiers(const Body &bi, const Body &bj, float wi, float wj, const CollideCfg &c, float &mi, float &mj) {
  const bool i_metal = bi.species == 1 || bi.species == 2, j_metal = bj.species == 1 || bj.species == 2;
  const float stiffness = rmin(rmax(c.softness, 0.0f), 1.0f);
  if (i_metal && !j_metal) {
    mj = wj + (wi * stiffness);
    mi = wi * (1.0f - stiffness);
    return;
  }
  if (j_metal && !i_metal) {
    mi = wi + (wj * stiffness);
    mj = wj * (1.0f - stiffness);
    return;
  }
  const bool i_li = c.soft_li && bi.species == 0, j_li = c.soft_li && bj.species == 0;
  const bool i_an = c.soft_an && bi.species == 3, j_an = c.soft_an && bj.species == 3;
  if (i_li || j_li || i_an || j_an) {
    const float scale = 1.0f - stiffness;
    mi = wi * scale, mj = wj * scale;
    return;
  }
  mi = wi, mj = wj;
}
void sanitize_body(Body &b, bool with_az) {
  if (!finite_f(b.pos.x)) b.pos.x = 0.0f;
  if (!finite_f(b.pos.y)) b.pos.y = 0.0f;
  if (!finite_f(b.vel.x)) b.vel.x = 0.0f;
  if (!finite_f(b.vel.y)) b.vel.y = 0.0f;
  if (!finite_f(b.z)) b.z = 0.0f;
  if (!finite_f(b.vz)) b.vz = 0.0f;
  if (with_az && !finite_f(b.az)) b.az = 0.0f;
}
// resolve, collision.rs:158-372; returns whether the pair touched
bool collide_resolve(std::vector<Body> &bodies, size_t i, size_t j, const CollideCfg &c) {
  V2 p1 = bodies[i].pos, p2 = bodies[j].pos;
  float z1 = bodies[i].z, z2 = bodies[j].z;
  const float r1 = bodies[i].radius, r2 = bodies[j].radius;
  V2 d_xy = p2 - p1;
  float dz = z2 - z1;
  const float r = r1 + r2;
  float dist_sq = mag_sq(d_xy) + dz * dz;
  V2 v1 = bodies[i].vel, v2b = bodies[j].vel;
  float v1z = bodies[i].vz, v2z = bodies[j].vz;
  bool need = !(finite_f(p1.x) && finite_f(p1.y) && finite_f(z1) && finite_f(v1.x) && finite_f(v1.y) && finite_f(v1z)) ||
              !(finite_f(p2.x) && finite_f(p2.y) && finite_f(z2) && finite_f(v2b.x) && finite_f(v2b.y) && finite_f(v2z));
  if (need || !finite_f(dist_sq)) {
    sanitize_body(bodies[i], true), sanitize_body(bodies[j], true);
    p1 = bodies[i].pos, p2 = bodies[j].pos, z1 = bodies[i].z, z2 = bodies[j].z;
    d_xy = p2 - p1, dz = z2 - z1;
    dist_sq = mag_sq(d_xy) + dz * dz;
    v1 = bodies[i].vel, v2b = bodies[j].vel, v1z = bodies[i].vz, v2z = bodies[j].vz;
  }
  if (dist_sq > r * r) return false;
  const V2 v_xy = v2b - v1;
  const float vz = v2z - v1z;
  const float d_dot_v = (d_xy.x * v_xy.x + d_xy.y * v_xy.y) + dz * vz;
  const float m1 = bodies[i].mass, m2 = bodies[j].mass;
  const float weight1 = m2 / (m1 + m2), weight2 = m1 / (m1 + m2);
  if (d_dot_v >= 0.0f && dist_sq > 0.0f && finite_f(dist_sq)) {
    const float dist = sqrtf(dist_sq);
    const float corr = r / dist - 1.0f;
    const float sep_x = d_xy.x * corr, sep_y = d_xy.y * corr, sep_z = dz * corr;
    float mw1, mw2;
    collision_modifiers(bodies[i], bodies[j], weight1, weight2, c, mw1, mw2);
    bodies[i].pos.x -= mw1 * sep_x, bodies[i].pos.y -= mw1 * sep_y, bodies[i].z -= mw1 * sep_z;
    bodies[j].pos.x += mw2 * sep_x, bodies[j].pos.y += mw2 * sep_y, bodies[j].z += mw2 * sep_z;
    return true;
  }
  const float v_sq = mag_sq(v_xy) + vz * vz;
  const float d_sq = dist_sq;
  if (!finite_f(d_sq) || d_sq <= 1.0e-8f || !finite_f(v_sq)) {
    const uint64_t jj = (uint64_t)j;
    const float angle = (float)(((uint64_t)i) ^ ((jj << 13) | (jj >> 51))) * (6.28318530717958647692f / 1024.0f);
    const float s = sinf(angle), co = cosf(angle);
    const V2 dir = v2(co, s);
    const float sep = r * 1.001f;
    const V2 mid = (bodies[i].pos + bodies[j].pos) * 0.5f;
    bodies[i].pos = mid - dir * (sep * weight1);
    bodies[j].pos = mid + dir * (sep * weight2);
    const float depth = c.domain_depth;
    const float midz = rmin(rmax((bodies[i].z + bodies[j].z) * 0.5f, -depth), depth);
    bodies[i].z = midz, bodies[j].z = midz;
    for (size_t k : {i, j}) {
      Body &b = bodies[k];
      if (!finite_f(b.vel.x)) b.vel.x = 0.0f;
      if (!finite_f(b.vel.y)) b.vel.y = 0.0f;
      if (!finite_f(b.vz)) b.vz = 0.0f;
    }
    return true;
  }
  const float r_sq = r * r;
  const float correction_scale = 1.0f / (float)c.num_passes;
  const float disc_term = rmax(d_dot_v * d_dot_v - v_sq * (d_sq - r_sq), 0.0f);
  const float sqrt_disc = sqrtf(disc_term);
  const float numerator = d_dot_v + sqrt_disc;
  const float t = correction_scale * numerator / v_sq;
  if (!finite_f(t)) {
    const float dist = sqrtf(d_sq);
    if (finite_f(dist) && dist > 0.0f) {
      const float corr = r / dist - 1.0f;
      const float sep_x = d_xy.x * corr, sep_y = d_xy.y * corr, sep_z = dz * corr;
      bodies[i].pos.x -= weight1 * sep_x, bodies[i].pos.y -= weight1 * sep_y, bodies[i].z -= weight1 * sep_z;
      bodies[j].pos.x += weight2 * sep_x, bodies[j].pos.y += weight2 * sep_y, bodies[j].z += weight2 * sep_z;
    }
    return true;
  }
  bodies[i].pos = bodies[i].pos - v1 * t, bodies[i].z -= v1z * t;
  bodies[j].pos = bodies[j].pos - v2b * t, bodies[j].z -= v2z * t;
  const V2 q1 = bodies[i].pos, q2 = bodies[j].pos;
  const float zz1 = bodies[i].z, zz2 = bodies[j].z;
  const V2 d2 = q2 - q1;
  const float dz2 = zz2 - zz1;
  const float ddv = (d2.x * v_xy.x + d2.y * v_xy.y) + dz2 * vz;
  const float dsq2 = mag_sq(d2) + dz2 * dz2;
  float scale = (finite_f(dsq2) && dsq2 > 0.0f) ? 1.5f * ddv / dsq2 : 0.0f;
  if (!finite_f(scale)) scale = 0.0f;
  const float sep_x = d2.x * scale, sep_y = d2.y * scale, sep_z = dz2 * scale;
  float mw1, mw2;
  collision_modifiers(bodies[i], bodies[j], weight1, weight2, c, mw1, mw2);
  const float v1x = v1.x + sep_x * mw1, v1y = v1.y + sep_y * mw1, v1z_new = v1z + sep_z * mw1;
  const float v2x = v2b.x - sep_x * mw2, v2y = v2b.y - sep_y * mw2, v2z_new = v2z - sep_z * mw2;
  bodies[i].vel = v2(v1x, v1y), bodies[i].vz = v1z_new;
  bodies[j].vel = v2(v2x, v2y), bodies[j].vz = v2z_new;
  bodies[i].pos = bodies[i].pos + v2(v1x, v1y) * t, bodies[i].z += v1z_new * t;
  bodies[j].pos = bodies[j].pos + v2(v2x, v2y) * t, bodies[j].z += v2z_new * t;
  sanitize_body(bodies[i], false), sanitize_body(bodies[j], false);
  return true;
}
}  // namespace

// collide, collision.rs:62-156: the broad phase is broccoli's (un-vendored, Cargo.lock: broccoli 6.3) "all pairs of
// intersecting axis-aligned rectangles, each once"; here a uniform grid finds the same pairs and they are resolved in
// index order.
uint64_t orc_collide(OrcSim *s, float domain_depth, uint32_t num_passes, float li_collision_softness,
                     int soft_collision_lithium_ion, int soft_collision_anion) {
  auto &bodies = s->bodies;
  const size_t n = bodies.size();
  if (n == 0) return 0;
  CollideCfg c{li_collision_softness, soft_collision_lithium_ion != 0, soft_collision_anion != 0, domain_depth,
               num_passes ? num_passes : 1u};
  float rmaxv = 0.0f, mnx = 3.4e38f, mny = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f;
  for (const Body &b : bodies) {
    if (!(finite_f(b.pos.x) && finite_f(b.pos.y))) continue;
    rmaxv = rmax(rmaxv, b.radius);
    mnx = rmin(mnx, b.pos.x), mny = rmin(mny, b.pos.y), mxx = rmax(mxx, b.pos.x), mxy = rmax(mxy, b.pos.y);
  }
  if (!(rmaxv > 0.0f) || !(mxx >= mnx)) return 0;
  const double cell = 2.0 * (double)rmaxv;
  const int64_t gx = (int64_t)std::floor(((double)mxx - mnx) / cell) + 1, gy = (int64_t)std::floor(((double)mxy - mny) / cell) + 1;
  std::vector<std::vector<uint32_t>> cells((size_t)(gx * gy));
  auto cell_of = [&](const Body &b, int64_t &cx, int64_t &cy) {
    cx = (int64_t)std::floor(((double)b.pos.x - mnx) / cell), cy = (int64_t)std::floor(((double)b.pos.y - mny) / cell);
  };
  for (size_t i = 0; i < n; ++i) {
    if (!(finite_f(bodies[i].pos.x) && finite_f(bodies[i].pos.y))) continue;
    int64_t cx, cy;
    cell_of(bodies[i], cx, cy);
    cells[(size_t)(cy * gx + cx)].push_back((uint32_t)i);
  }
  // pairs from the positions at the start of the pass (the BVH is built once per pass, collision.rs:148)
  std::vector<std::pair<uint32_t, uint32_t>> pairs;
  for (size_t i = 0; i < n; ++i) {
    const Body &a = bodies[i];
    if (!(finite_f(a.pos.x) && finite_f(a.pos.y))) continue;
    int64_t cx, cy;
    cell_of(a, cx, cy);
    for (int64_t y = std::max<int64_t>(cy - 1, 0); y <= std::min<int64_t>(cy + 1, gy - 1); ++y)
      for (int64_t x = std::max<int64_t>(cx - 1, 0); x <= std::min<int64_t>(cx + 1, gx - 1); ++x)
        for (uint32_t j : cells[(size_t)(y * gx + x)]) {
          if (j <= i) continue;
          const Body &b = bodies[j];
          const float rr = a.radius + b.radius;
          if (fabsf(b.pos.x - a.pos.x) <= rr && fabsf(b.pos.y - a.pos.y) <= rr) pairs.emplace_back((uint32_t)i, j);
        }
  }
  std::sort(pairs.begin(), pairs.end());
  uint64_t touched = 0;
  for (auto &pr : pairs) touched += collide_resolve(bodies, pr.first, pr.second, c) ? 1 : 0;
  return touched;
}

// simulation/electron_hopping.rs:283-329: what the hopping loop computes per candidate before the rate tests
void orc_hop_alignment(const OrcSim *s, uint64_t m_src, const uint32_t *src_idx, const uint32_t *pair_offsets,
                       const uint32_t *dst_idx, float k_e, float bg_x, float bg_y, float alignment_bias,
                       float *local_field_xy, float *alignment) {
  auto electrode = [](uint8_t sp) { return sp >= 13 && sp <= 20; };             // :312-315
  auto metal_or_electrode = [&](uint8_t sp) { return sp == 1 || sp == 2 || electrode(sp); };  // :316-320
  for (uint64_t i = 0; i < m_src; ++i) {
    const auto &src = s->bodies[src_idx[i]];
    for (uint32_t k = pair_offsets[i]; k < pair_offsets[i + 1]; ++k) {
      const auto &dst = s->bodies[dst_idx[k]];
      const V2 hop_vec = dst.pos - src.pos;                                                    // :284
      const V2 hop_dir = mag(hop_vec) > 1e-6f ? normalized(hop_vec) : v2(0.0f, 0.0f);           // :285-289
      const V2 local_field = v2(bg_x, bg_y) + s->qt.acc_pos(src.pos, 1.0f, 0.0f, s->bodies, k_e, nullptr);  // :290-295
      const V2 field_dir = mag(local_field) > 1e-6f ? normalized(local_field) : v2(0.0f, 0.0f); // :296-300
      float a = rmax(-(hop_dir.x * field_dir.x + hop_dir.y * field_dir.y), 0.0f);              // :301
      if (field_dir.x == 0.0f && field_dir.y == 0.0f) a = 1.0f;                                 // :302-304
      a = a * rmax(alignment_bias, 0.0f);                                                       // :305-307
      const bool both = metal_or_electrode(src.species) && metal_or_electrode(dst.species);     // :323
      const bool involves = electrode(src.species) || electrode(dst.species);                   // :324
      if (both && involves) a = rmax(a, 0.5f);                                                  // :326-329
      alignment[k] = a;
      if (local_field_xy) local_field_xy[2 * i] = local_field.x, local_field_xy[2 * i + 1] = local_field.y;
    }
    if (local_field_xy && pair_offsets[i] == pair_offsets[i + 1]) {
      const V2 lf = v2(bg_x, bg_y) + s->qt.acc_pos(src.pos, 1.0f, 0.0f, s->bodies, k_e, nullptr);
      local_field_xy[2 * i] = lf.x, local_field_xy[2 * i + 1] = lf.y;
    }
  }
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_uv_fma(void) {
#ifdef ORC_UV_FMA
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"

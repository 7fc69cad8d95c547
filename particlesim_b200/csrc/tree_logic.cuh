// tree_logic.cuh — per-body / per-node bodies of the tree kernels, written once as host+device
// functions so that tests/emu can run the very same construction logic serially on the CPU and
// compare it with the oracle before any GPU time is spent.  The product only ever calls these from
// the __global__ wrappers in tree.cuh.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector_functions.h>
#include <vector_types.h>

#include "psim_core.cuh"

namespace psim {

constexpr int kLevels = kMaxLevels + 1;  // node depths 0..32

struct TreeMeta {
  RootQuad root;
  uint32_t n;
  uint32_t num_nodes;       // compact nodes
  uint32_t num_internal;    // internal nodes (== number of 4-child groups in the reference shape)
  uint32_t max_depth;
  uint32_t dcap;
  uint32_t err;             // bit0 node arena overflow
  uint32_t num_zero_leaves; // refused / thread-capacity leaves
  uint32_t num_cap_leaves;  // multi-body leaves stopped by the 32-level key or the 1e-6 size rule
  uint32_t level_count[kLevels];
  uint32_t level_start[kLevels + 1];
  uint32_t level_cursor[kLevels];
  uint32_t internal_total;  // all internal nodes (level_count only covers the ones the level sweeps visit)
  uint32_t zero_agg_hint;   // some leaf holds more bodies than it aggregates: set by tree_count_kernel
};

struct NodeSums {  // per internal node: running sums over its body range
  double aq, aqx, aqy;  // Σ|q|, Σ|q|x, Σ|q|y
  double m, mx, my;     // Σm, Σm x, Σm y
  double x, y;          // Σx, Σy
};

// What the bottom-up sweep of the build carries per node: one 32-byte sector.
struct NodeRec {
  double aq, aqx, aqy;  // Σ|q|, Σ|q|x, Σ|q|y over the node's bodies
  float charge;         // node charge in the reference's child order
  uint32_t next;        // skip pointer | kLastSibling if the node is its parent's last non-empty child
};
constexpr uint32_t kLastSibling = 1u << 31;
// ndepth[node] = depth | kDepthCharged if some body below the node carries charge (one byte per node:
// the traversal compaction scans these instead of the 32-byte records)
constexpr uint32_t kDepthCharged = 0x80u;
constexpr uint32_t kDepthMask8 = 0x7fu;
constexpr uint32_t kNextMask = ~kLastSibling;

struct TreeArrays {
  float4* nodeA;      // {pos.x, pos.y, charge, quad.size}
  uint4* nodeB;       // {next, body_start, body_count, depth | flags}; internal body_count is
                      // filled by the export sweep only (the traversal never reads it)
  NodeRec* rec;
  uint8_t* ndepth;
  float* node_mass;   // export sweep only
  uint32_t* parent;   // export sweep only: compact index of the parent (root: 0xffffffff)
  NodeSums* sums;     // export sweep only
  uint32_t* level_nodes;  // internal nodes bucketed by depth (those the level sweeps visit)
  uint32_t* local_nodes;  // internal nodes an emit CTA sums itself, at [first node of its slab ...)
  uint32_t node_cap;
};

// Node in the reference's field order (node.rs:6-14), 64 bytes; == psim_node of the C ABI
struct PsimNodeOut {
  uint64_t children, next;
  float pos[2];
  float mass;
  float quad_center[2];
  float quad_size;
  uint64_t bodies_start, bodies_end;
  float charge;
  uint32_t _pad;
};

PSIM_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
PSIM_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}

// root square: Quad::new_containing (quad.rs:31-34) from the reduced AABB, or new_for_domain (:38-43)
PSIM_HD RootQuad root_from_bounds(float mnx, float mny, float mxx, float mxy) {
  RootQuad r;
  r.cx = f_mul(f_add(mnx, mxx), 0.5f);
  r.cy = f_mul(f_add(mny, mxy), 0.5f);
  r.size = fmaxf(f_sub(mxx, mnx), f_sub(mxy, mny));
  return r;
}
PSIM_HD RootQuad root_for_domain(float hw, float hh) {
  RootQuad r;
  r.cx = 0.0f, r.cy = 0.0f;
  r.size = fmaxf(f_mul(2.0f, hw), f_mul(2.0f, hh));
  return r;
}

PSIM_HD void meta_reset(TreeMeta* meta, RootQuad r, uint32_t n) {
  meta->root = r;
  meta->n = n;
  meta->num_nodes = 0;
  meta->num_internal = 0;
  meta->max_depth = 0;
  meta->dcap = (uint32_t)depth_cap(r.size);
  meta->err = 0;
  meta->num_zero_leaves = 0;
  meta->num_cap_leaves = 0;
  meta->internal_total = 0;
  meta->zero_agg_hint = 0;
  for (int l = 0; l < kLevels; ++l) meta->level_count[l] = 0, meta->level_cursor[l] = 0;
}

// (λ_i + 1) | ℓ_i << 8
PSIM_HD uint16_t body_levels(const uint64_t* keys, const float4* pqr, uint32_t n, uint32_t i,
                             uint32_t c_eff, int dcap) {
  const int lam = lambda_at(keys, i);
  int ell = 0;
  if (lam < kMaxLevels) ell = leaf_depth(keys, pqr, n, i, lam, c_eff, dcap);
  return (uint16_t)((uint32_t)(lam + 1) | ((uint32_t)ell << 8));
}
PSIM_HD int le_lambda(uint16_t le) { return (int)(le & 0xff) - 1; }
PSIM_HD int le_ell(uint16_t le) { return (int)(le >> 8); }
PSIM_HD uint32_t le_nodes(uint16_t le) {
  const int lam = le_lambda(le), ell = le_ell(le);
  return lam < ell ? (uint32_t)(ell - lam) : 0u;
}

PSIM_HD void level_scan(TreeMeta* meta, uint32_t node_cap) {
  uint32_t run = 0;
  for (int l = 0; l < kLevels; ++l) {
    meta->level_start[l] = run;
    run += meta->level_count[l];
    meta->level_cursor[l] = 0;
  }
  meta->level_start[kLevels] = run;
  meta->num_internal = run;
  if (meta->num_nodes > node_cap) meta->err |= 1u;
}

// The leaf whose first body is i (the first half of emit_nodes_for_body, for the single-GPU emit kernel, which deals the
// internal cells of a warp's chains out to its lanes): returns the end of the leaf's body range.
template <class Sink>
PSIM_HD uint32_t emit_leaf_for_body(const uint64_t* keys, uint32_t n, uint32_t i, int lam, int ell, uint32_t base,
                                    const uint32_t* nodebase, uint32_t M, const float4* pqr, uint32_t leaf_capacity,
                                    uint32_t thread_capacity, float root_size, int dcap, const TreeArrays& t, Sink& sink,
                                    bool write_rec = true) {
  const int d = ell;
  const uint32_t j = run_end(keys, n, i, i + 1, d);
  const uint32_t node = base + (uint32_t)(d - lam - 1);
  const uint32_t next = (j < n) ? nodebase[j] : M;
  const uint32_t count = j - i;
  const float size = ldexpf(root_size, -d);  // size *= 0.5 per level, exact
  const bool agg = leaf_is_aggregated(count, leaf_capacity, thread_capacity);
  float tq = 0.0f, wx = 0.0f, wy = 0.0f;
  double aq = 0.0, aqx = 0.0, aqy = 0.0;
  for (uint32_t b = i; b < j; ++b) {
    const float4 p = pqr[b];
    const double a = fabs((double)p.z);
    aq += a, aqx += a * (double)p.x, aqy += a * (double)p.y;
    if (agg) {
      wx = f_add(wx, f_mul(p.x, p.z));
      wy = f_add(wy, f_mul(p.y, p.z));
      tq = f_add(tq, p.z);
    }
  }
  if (agg) {
    if (fabsf(tq) > 1e-6f) {
      wx = f_div(wx, tq);
      wy = f_div(wy, tq);
    }
  } else {
    sink.zero_leaf();
  }
  if (count > 1 && d == dcap) sink.cap_leaf();
  t.nodeA[node] = make_float4(wx, wy, tq, size);
  t.nodeB[node] = make_uint4(next, i, count, (uint32_t)d | kNodeLeaf | (agg ? 0u : kNodeZeroAgg) | (aq > 0.0 ? kNodeCharged : 0u));
  const bool last = (j >= n) || lcp_levels(keys[i], keys[j]) < d - 1;
  NodeRec r;
  r.aq = aq, r.aqx = aqx, r.aqy = aqy, r.charge = tq, r.next = next | (last ? kLastSibling : 0u);
  if (write_rec) t.rec[node] = r;
  t.ndepth[node] = (uint8_t)((uint32_t)d | (aq > 0.0 ? kDepthCharged : 0u));
  return j;
}

// All nodes whose first body is i: the leaf at depth ℓ_i and the internal cells above it down to
// depth λ_i + 1.  The leaf is finished here (range end by galloping search, aggregates per
// quadtree.rs:281-306, Σ|q| sums for its ancestors); internal cells only get their depth and first
// body — their skip pointer and sums come from the bottom-up sweep, which reaches the end of a
// node's subtree through its children.  `sink` hands out level-bucket slots and counts diagnostics
// (atomics on the device, plain counters in the emulation).
template <class Sink>
PSIM_HD void emit_nodes_for_body(const uint64_t* keys, uint32_t n, uint32_t i, uint16_t lev,
                                 const uint32_t* nodebase, uint32_t M, const float4* pqr,
                                 const float4* accm, uint32_t leaf_capacity, uint32_t thread_capacity,
                                 float root_size, int dcap, const TreeArrays& t, Sink& sink,
                                 uint32_t body_base = 0, int min_bucket_depth = 0,
                                 int straddle_depth = kMaxLevels + 1, bool internal_ranges = false) {
  // body_base / min_bucket_depth / Sink::kTop serve the sharded build (shard.cuh): `keys`, `pqr` and
  // `nodebase` are then a rank's slice of the sorted order plus a halo, i + body_base is the global body
  // index, cells above the shard depth are reported to the sink instead of the level buckets
  (void)accm;
  const int lam = le_lambda(lev), ell = le_ell(lev);
  if (!(lam < ell)) return;
  const uint32_t base = nodebase[i];
  {
    const int d = ell;
    const uint32_t j = run_end(keys, n, i, i + 1, d);
    const uint32_t node = base + (uint32_t)(d - lam - 1);
    const uint32_t next = (j < n) ? nodebase[j] : M;
    const uint32_t count = j - i;
    const float size = ldexpf(root_size, -d);  // size *= 0.5 per level, exact
    const bool agg = leaf_is_aggregated(count, leaf_capacity, thread_capacity);
    float tq = 0.0f, wx = 0.0f, wy = 0.0f;
    double aq = 0.0, aqx = 0.0, aqy = 0.0;
    for (uint32_t b = i; b < j; ++b) {
      const float4 p = pqr[b];
      const double a = fabs((double)p.z);
      aq += a, aqx += a * (double)p.x, aqy += a * (double)p.y;
      if (agg) {
        wx = f_add(wx, f_mul(p.x, p.z));
        wy = f_add(wy, f_mul(p.y, p.z));
        tq = f_add(tq, p.z);
      }
    }
    if (agg) {
      if (fabsf(tq) > 1e-6f) {
        wx = f_div(wx, tq);
        wy = f_div(wy, tq);
      }
    } else {
      sink.zero_leaf();
    }
    if (count > 1 && d == dcap) sink.cap_leaf();
    t.nodeA[node] = make_float4(wx, wy, tq, size);
    t.nodeB[node] = make_uint4(next, i + body_base, count, (uint32_t)d | kNodeLeaf | (agg ? 0u : kNodeZeroAgg) |
                                                               (aq > 0.0 ? kNodeCharged : 0u));
    const bool last = (j >= n) || lcp_levels(keys[i], keys[j]) < d - 1;
    NodeRec r;
    r.aq = aq, r.aqx = aqx, r.aqy = aqy, r.charge = tq, r.next = next | (last ? kLastSibling : 0u);
    t.rec[node] = r;
    t.ndepth[node] = (uint8_t)((uint32_t)d | (aq > 0.0 ? kDepthCharged : 0u));
    if (Sink::kTop && d <= min_bucket_depth) sink.top_leaf(d, keys[i], node, r);
  }
  // internal_ranges (single-GPU build): an internal cell's body range, and with it its skip pointer, is
  // found here by extending the galloping search outwards from the deeper cell (ranges nest), so the
  // bottom-up sums never read another CTA's nodes: a CTA sums the cells that lie inside its own slab
  // of bodies itself (depth > straddle_depth -> sink.local_node), only cells that straddle a slab
  // boundary go to the level sweeps.
  uint32_t jprev = i + 1;
  if (internal_ranges) jprev = run_end(keys, n, i, i + 1, ell);
  uint32_t big = 0;  // internal nodes of this chain that strict.cuh has to sum (the shallowest `big` ones)
  for (int d = ell - 1; d > lam; --d) {
    const uint32_t node = base + (uint32_t)(d - lam - 1);
    uint32_t nx = 0u, cnt = 0u;
    if (internal_ranges) {
      jprev = run_end(keys, n, i, jprev, d);
      nx = (jprev < n) ? nodebase[jprev] : M;
      cnt = jprev - i;
      if (Sink::kStrict && cnt > sink.strict_direct()) ++big;
    }
    t.nodeB[node] = make_uint4(nx, i + body_base, cnt, (uint32_t)d);
    t.ndepth[node] = (uint8_t)d;
    if (d >= min_bucket_depth) {
      if (d <= straddle_depth) t.level_nodes[sink.level_slot(d)] = node;
      else sink.local_node(d, node);
    }
    if (Sink::kTop && d <= min_bucket_depth) sink.top_internal(d, keys[i], node);
  }
  if (Sink::kStrict && big) sink.strict_chain(i, base, big, jprev - i);
}

// One internal node of the build's bottom-up sweep (quadtree.rs:103-151): charge in the reference's
// ((c0 + c1) + c2) + c3 order (absent children are +0.0 terms), Σ|q| sums in f64 for the centre (the
// reference runs one f32 running sum over the node's whole range), skip pointer = where the last
// child's subtree ends.  It touches one 32-byte record per child and writes one; centres and the
// topology records are finished by the streaming finalize pass.
PSIM_HD void aggregate_node_lean(uint32_t node, int depth, uint32_t M, const TreeArrays& t) {
  double aq = 0.0, aqx = 0.0, aqy = 0.0;
  float charge = 0.0f;
  uint32_t c = node + 1;
  while (true) {
    const NodeRec r = t.rec[c];
    charge = f_add(charge, r.charge);
    aq += r.aq, aqx += r.aqx, aqy += r.aqy;
    c = r.next & kNextMask;
    if (r.next & kLastSibling) break;
  }
  // the node that follows my subtree is my sibling iff it sits at my depth
  const bool last = (c >= M) || ((int)(t.ndepth[c] & kDepthMask8) != depth);
  NodeRec out;
  out.aq = aq, out.aqx = aqx, out.aqy = aqy, out.charge = charge, out.next = c | (last ? kLastSibling : 0u);
  t.rec[node] = out;
  if (aq > 0.0) t.ndepth[node] = (uint8_t)((uint32_t)depth | kDepthCharged);
}

// The same for a node whose skip pointer is already known (emit with internal_ranges): the children are
// the nodes from node + 1 up to the skip pointer, hopping over each child's subtree; no sibling flags and
// no read beyond the node's own subtree.
PSIM_HD void aggregate_node_ranged(uint32_t node, int depth, const TreeArrays& t) {
  const uint32_t end = t.nodeB[node].x;
  double aq = 0.0, aqx = 0.0, aqy = 0.0;
  float charge = 0.0f;
  uint32_t c = node + 1;
  while (c < end) {
    const NodeRec r = t.rec[c];
    charge = f_add(charge, r.charge);
    aq += r.aq, aqx += r.aqx, aqy += r.aqy;
    c = r.next & kNextMask;
  }
  NodeRec out;
  out.aq = aq, out.aqx = aqx, out.aqy = aqy, out.charge = charge, out.next = end;
  t.rec[node] = out;
  if (aq > 0.0) t.ndepth[node] = (uint8_t)((uint32_t)depth | kDepthCharged);
}

struct SubtreeEndCount {  // internal nodes carry their body count (emit with internal_ranges)
  PSIM_HD uint32_t operator()(uint32_t, const uint4& nb) const { return nb.y + nb.z; }
};

// Streaming pass over all nodes in pre-order after the sweep: topology record and centre of every
// internal node from its sums.  A node whose Σ|q| is <= 1e-6 takes the reference's mass / centroid
// fall-back by a direct pass over its bodies; for Σ|q| == 0 (no charge below: the node
// can never contribute to a field sum) that is left to the export sweep.
struct SubtreeEndLocal {  // first body after the subtree that ends where node c starts
  uint32_t M, n_bodies;
  const uint4* nodeB;
  PSIM_HD uint32_t operator()(uint32_t c, const uint4&) const { return c < M ? nodeB[c].y : n_bodies; }
};

// The reference's own loop for one internal node's centre (quadtree.rs:114-139): serial f32 running sums over
// the node's bodies, |q|-weighted, else mass-weighted, else the centroid.
PSIM_HD void strict_centre_direct(uint32_t b0, uint32_t b1, const float4* pqr, const float4* accm, float& px,
                                  float& py) {
  float total_abs = 0.0f, wx = 0.0f, wy = 0.0f;
  for (uint32_t b = b0; b < b1; ++b) {
    const float4 p = pqr[b];
    const float a = fabsf(p.z);
    total_abs = f_add(total_abs, a);
    wx = f_add(wx, f_mul(p.x, a)), wy = f_add(wy, f_mul(p.y, a));
  }
  if (total_abs > 1e-6f) {
    px = f_div(wx, total_abs), py = f_div(wy, total_abs);
    return;
  }
  float total_mass = 0.0f;
  for (uint32_t b = b0; b < b1; ++b) total_mass = f_add(total_mass, accm[b].w);
  wx = 0.0f, wy = 0.0f;
  if (total_mass > 1e-6f) {
    for (uint32_t b = b0; b < b1; ++b) {
      const float4 p = pqr[b];
      const float m = accm[b].w;
      wx = f_add(wx, f_mul(p.x, m)), wy = f_add(wy, f_mul(p.y, m));
    }
    px = f_div(wx, total_mass), py = f_div(wy, total_mass);
  } else if (b1 > b0) {
    for (uint32_t b = b0; b < b1; ++b) wx = f_add(wx, pqr[b].x), wy = f_add(wy, pqr[b].y);
    px = f_div(wx, (float)(b1 - b0)), py = f_div(wy, (float)(b1 - b0));
  } else {
    px = 0.0f, py = 0.0f;
  }
}

// psim_config.strict_centres: nodes of at most `limit` bodies get the reference's serial sums right where the tree
// is finished; larger ones are left to strict.cuh.  A body with q == 0 adds exactly +-0 to the |q|-weighted sums, so
// the loop runs over the node's CHARGED bodies only (cidx = exclusive count of charged bodies, cw = their
// {|q|, x|q|, y|q|} in sorted order, both made before the emit kernel).  A node without any charged body never enters
// a field sum: its (mass-weighted / centroid) centre is left to the export (strict_chargeless_kernel).
struct StrictDirect {
  uint32_t limit = 0;  // 0: off
  const uint32_t* cidx = nullptr;
  const float4* cw = nullptr;
};

// set_count (sharded build, whose emit does not know the ranges of internal cells): the body count goes into B.z.
template <class SubtreeEnd>
PSIM_HD void finalize_node(uint32_t node, float root_size, const float4* pqr, const float4* accm,
                           const TreeArrays& t, const SubtreeEnd& subtree_end, StrictDirect sd = StrictDirect(),
                           bool set_count = false) {
  uint4 nb = t.nodeB[node];
  if (nb.w & kNodeLeaf) return;
  const NodeRec r = t.rec[node];
  const uint32_t c = r.next & kNextMask;
  nb.x = c;
  if (r.aq > 0.0) nb.w |= kNodeCharged;
  if (set_count) nb.z = subtree_end(c, nb) - nb.y;
  t.nodeB[node] = nb;
  float px = 0.0f, py = 0.0f;
  if (sd.limit && nb.z <= sd.limit) {
    const uint32_t c0 = sd.cidx[nb.y], c1 = sd.cidx[nb.y + nb.z];
    float total_abs = 0.0f, wx = 0.0f, wy = 0.0f;
    for (uint32_t k = c0; k < c1; ++k) {
      const float4 w = sd.cw[k];
      total_abs = f_add(total_abs, w.x), wx = f_add(wx, w.y), wy = f_add(wy, w.z);
    }
    if (total_abs > 1e-6f) px = f_div(wx, total_abs), py = f_div(wy, total_abs);
    else if (c1 > c0) strict_centre_direct(nb.y, nb.y + nb.z, pqr, accm, px, py);  // charges too small to weigh
  } else if (r.aq > (double)1e-6f) {
    px = (float)(r.aqx / r.aq), py = (float)(r.aqy / r.aq);
  } else if (r.aq > 0.0) {
    const uint32_t b0 = nb.y, b1 = subtree_end(c, nb);
    double m = 0.0, mx = 0.0, my = 0.0, x = 0.0, y = 0.0;
    for (uint32_t b = b0; b < b1; ++b) {
      const float4 p = pqr[b];
      const double w = (double)accm[b].w;
      m += w, mx += w * (double)p.x, my += w * (double)p.y, x += (double)p.x, y += (double)p.y;
    }
    if (m > (double)1e-6f) px = (float)(mx / m), py = (float)(my / m);
    else if (b1 > b0) px = (float)(x / (double)(b1 - b0)), py = (float)(y / (double)(b1 - b0));
  }
  t.nodeA[node] = make_float4(px, py, r.charge, ldexpf(root_size, -(int)(nb.w & kNodeDepthMask)));
}

// Export sweep (psim_download_nodes only), one internal node, deepest level first: parent links,
// node mass in the reference's child order, body counts, and the mass-weighted / centroid centre of
// nodes without charge (quadtree.rs:127-139), which the build leaves at (0, 0).
PSIM_HD void aggregate_node(uint32_t node, float root_size, const float4* pqr, const float4* accm,
                            const TreeArrays& t, bool write_chargeless_centres = true) {
  (void)root_size;
  const uint4 nb = t.nodeB[node];
  NodeSums s = {0, 0, 0, 0, 0, 0, 0, 0};
  float msum = 0.0f;
  uint32_t count = 0;
  uint32_t c = node + 1;
  while (c < nb.x) {
    const uint4 cb = t.nodeB[c];
    t.parent[c] = node;
    if (cb.w & kNodeLeaf) {
      float lm = 0.0f;
      const bool agg = !(cb.w & kNodeZeroAgg);
      for (uint32_t b = cb.y; b < cb.y + cb.z; ++b) {
        const float4 p = pqr[b];
        const double m = (double)accm[b].w;
        s.m += m, s.mx += m * (double)p.x, s.my += m * (double)p.y;
        s.x += (double)p.x, s.y += (double)p.y;
        if (agg) lm = f_add(lm, accm[b].w);
      }
      t.node_mass[c] = lm;
      msum = f_add(msum, lm);
    } else {
      const NodeSums cs = t.sums[c];
      s.m += cs.m, s.mx += cs.mx, s.my += cs.my;
      s.x += cs.x, s.y += cs.y;
      msum = f_add(msum, t.node_mass[c]);
    }
    count += t.nodeB[c].z;
    c = cb.x;
  }
  t.sums[node] = s;
  t.node_mass[node] = msum;
  t.nodeB[node].z = count;
  if (write_chargeless_centres && !(nb.w & kNodeCharged)) {
    float px = 0.0f, py = 0.0f;
    if (s.m > (double)1e-6f) {
      px = (float)(s.mx / s.m), py = (float)(s.my / s.m);
    } else if (count > 0) {
      px = (float)(s.x / (double)count), py = (float)(s.y / (double)count);
    }
    float4 a = t.nodeA[node];
    a.x = px, a.y = py;
    t.nodeA[node] = a;
  }
}

// ---- export in the reference's shape (node.rs:6-14): ROOT = 0, the 4 children of the r-th
// internal node (pre-order rank r) at 4r+1 .. 4r+4 (quadtree.rs:65-66 hands out groups the same
// way; the reference's raw group order is schedule dependent), `next` = sibling or the
// parent's next, 0 at the end (quadtree.rs:82-87).
PSIM_HD uint64_t export_index(uint32_t node, const uint32_t* parent, const uint32_t* irank,
                              const uint4* nodeB, const uint64_t* keys) {
  if (node == 0) return 0;
  const uint4 nb = nodeB[node];
  const unsigned q = digit_at(keys[nb.y], (int)(nb.w & kNodeDepthMask));
  return 4ull * irank[parent[node]] + 1ull + q;
}

PSIM_HD void export_node(uint32_t node, const uint64_t* keys, RootQuad root, const TreeArrays& t,
                         const uint32_t* irank, PsimNodeOut* out, uint64_t out_cap) {
  const uint4 nb = t.nodeB[node];
  const int depth = (int)(nb.w & kNodeDepthMask);
  const uint64_t key = keys[nb.y];
  const uint64_t me = export_index(node, t.parent, irank, t.nodeB, keys);
  // reference `next`: next sibling slot, or the first ancestor's; 0 when the walk ends
  uint64_t next = 0;
  {
    uint32_t a = node;
    while (a != 0) {
      const uint4 ab = t.nodeB[a];
      const unsigned q = digit_at(keys[ab.y], (int)(ab.w & kNodeDepthMask));
      if (q < 3) {
        next = export_index(a, t.parent, irank, t.nodeB, keys) + 1;
        break;
      }
      a = t.parent[a];
    }
  }
  const RootQuad quad = quad_at(root, key, depth);
  const bool leaf = (nb.w & kNodeLeaf) != 0;
  if (me < out_cap) {
    PsimNodeOut o;
    o.children = leaf ? 0ull : 4ull * irank[node] + 1ull;
    o.next = next;
    const float4 a = t.nodeA[node];
    o.pos[0] = a.x, o.pos[1] = a.y;
    o.mass = t.node_mass[node];
    o.quad_center[0] = quad.cx, o.quad_center[1] = quad.cy;
    o.quad_size = quad.size;
    o.bodies_start = nb.y, o.bodies_end = (uint64_t)nb.y + nb.z;
    o.charge = a.z;
    o._pad = 0;
    out[me] = o;
  }
  if (!leaf) {
    // materialise the empty children: Node::new(next, quad, s..s) with s = the split point
    const uint64_t group = 4ull * irank[node] + 1ull;
    uint32_t c = node + 1;
    unsigned q = 0;
    while (q < 4) {
      unsigned cq = 4;
      uint32_t cstart = nb.y + nb.z;
      uint32_t cnext = 0;
      if (c < nb.x) {
        const uint4 cb = t.nodeB[c];
        cq = digit_at(keys[cb.y], depth + 1);
        cstart = cb.y;
        cnext = cb.x;
      }
      for (; q < cq && q < 4; ++q) {
        if (group + q < out_cap) {
          const RootQuad cqd = quad_child(quad, q);
          PsimNodeOut o;
          o.children = 0;
          o.next = (q < 3) ? group + q + 1 : next;
          o.pos[0] = 0.0f, o.pos[1] = 0.0f;
          o.mass = 0.0f;
          o.quad_center[0] = cqd.cx, o.quad_center[1] = cqd.cy;
          o.quad_size = cqd.size;
          o.bodies_start = cstart, o.bodies_end = cstart;
          o.charge = 0.0f;
          o._pad = 0;
          out[group + q] = o;
        }
      }
      if (cq < 4) {
        q = cq + 1;
        c = cnext;
      }
    }
  }
}

}  // namespace psim

// emu.cpp — TEST-ONLY serial emulation of the device pipeline's logic.
//
// Purpose: check the construction algorithm (psim_core.cuh + tree_logic.cuh, the very same
// functions the sm_100a kernels call per thread) and the warp-lockstep traversal scheme against the
// oracle on the CPU box, before GPU time is spent.  It is NOT a product path: the library has no
// CPU fallback and nothing under particlesim_b200/ references this file.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../particlesim_b200/csrc/tree_logic.cuh"
#include "../../particlesim_b200/csrc/shard_logic.cuh"

using namespace psim;

namespace {
struct HostSink {
  static constexpr bool kTop = false;
  static constexpr bool kStrict = false;
  uint32_t strict_direct() const { return 0xffffffffu; }
  void strict_chain(uint32_t, uint32_t, uint32_t, uint32_t) {}
  void top_leaf(int, uint64_t, uint32_t, const NodeRec&) {}
  void top_internal(int, uint64_t, uint32_t) {}
  std::vector<std::vector<uint32_t>>* local = nullptr;  // per depth: the current slab's own cells
  void local_node(int d, uint32_t node) { (*local)[d].push_back(node); }
  TreeMeta* meta;
  uint32_t level_slot(int d) { return meta->level_start[d] + meta->level_cursor[d]++; }
  void zero_leaf() { meta->num_zero_leaves++; }
  void cap_leaf() { meta->num_cap_leaves++; }
};

struct Emu {
  uint32_t n = 0;
  std::vector<float4> pqr, accm;
  std::vector<uint64_t> keys;
  std::vector<uint32_t> perm;
  TreeMeta meta;
  std::vector<float4> nodeA;
  std::vector<uint4> nodeB;
  std::vector<float> node_mass;
  std::vector<uint32_t> parent, level_nodes, irank;
  std::vector<NodeSums> sums;
  std::vector<NodeRec> rec;
  std::vector<uint8_t> ndepth;
  TreeArrays t;
  float t_sq = 1, e_sq = 4;
  uint32_t slab = 0;  // > 0: replay the emit kernel's slabs (in-slab sums first, level sweeps for the rest)
};
}  // namespace

extern "C" {

void* emu_create() { return new Emu(); }
void emu_destroy(void* h) { delete static_cast<Emu*>(h); }

// returns number of compact nodes; fills perm (pre-build index of the body now at i) and keys
uint32_t emu_build(void* h, uint32_t n, const float* pos_xy, const float* mass, const float* radius,
                   const float* charge, int mode, float hw, float hh, uint32_t leaf_capacity,
                   uint32_t thread_capacity, uint32_t* perm_out, uint64_t* keys_out) {
  Emu& e = *static_cast<Emu*>(h);
  e.n = n;
  memset(&e.meta, 0, sizeof e.meta);
  if (n == 0) return 0;
  std::vector<float4> pqr(n), accm(n);
  for (uint32_t i = 0; i < n; ++i) {
    pqr[i] = make_float4(pos_xy[2 * i], pos_xy[2 * i + 1], charge ? charge[i] : 0.f, radius ? radius[i] : 0.f);
    accm[i] = make_float4(0, 0, 0, mass ? mass[i] : 1.f);
  }
  RootQuad r;
  if (mode == 0) {
    float mnx = 3.402823466e+38f, mny = mnx, mxx = -mnx, mxy = -mnx;
    for (uint32_t i = 0; i < n; ++i) {
      mnx = fminf(mnx, pqr[i].x), mny = fminf(mny, pqr[i].y);
      mxx = fmaxf(mxx, pqr[i].x), mxy = fmaxf(mxy, pqr[i].y);
    }
    r = root_from_bounds(mnx, mny, mxx, mxy);
  } else {
    r = root_for_domain(hw, hh);
  }
  meta_reset(&e.meta, r, n);
  std::vector<uint64_t> k0(n);
  for (uint32_t i = 0; i < n; ++i) k0[i] = morton_key(pqr[i].x, pqr[i].y, r);
  e.perm.resize(n);
  std::iota(e.perm.begin(), e.perm.end(), 0u);
  std::stable_sort(e.perm.begin(), e.perm.end(), [&](uint32_t a, uint32_t b) { return k0[a] < k0[b]; });
  e.keys.resize(n), e.pqr.resize(n), e.accm.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    e.keys[i] = k0[e.perm[i]];
    e.pqr[i] = pqr[e.perm[i]];
    e.accm[i] = accm[e.perm[i]];
  }
  if (perm_out) memcpy(perm_out, e.perm.data(), n * 4);
  if (keys_out) memcpy(keys_out, e.keys.data(), n * 8);

  const uint32_t c_eff = effective_capacity(leaf_capacity, thread_capacity);
  const int dcap = (int)e.meta.dcap;
  std::vector<uint16_t> le(n);
  std::vector<uint32_t> nodebase(n + 1);
  uint32_t run = 0;
  for (uint32_t i = 0; i < n; ++i) {
    le[i] = body_levels(e.keys.data(), e.pqr.data(), n, i, c_eff, dcap);
    nodebase[i] = run;
    run += le_nodes(le[i]);
    const int lam = le_lambda(le[i]), ell = le_ell(le[i]);
    if (lam < ell) {
      if ((uint32_t)ell > e.meta.max_depth) e.meta.max_depth = ell;
      e.meta.internal_total += (uint32_t)(ell - lam - 1);
    }
  }
  // level buckets: every internal cell, or (slab mode, like tree_count_kernel) only those that straddle a
  // slab boundary
  auto straddle_of = [&](uint32_t i) {
    if (!e.slab) return kMaxLevels + 1;
    const uint64_t slab_end = ((uint64_t)(i / e.slab) + 1) * e.slab;
    return slab_end < n ? lcp_levels(e.keys[i], e.keys[slab_end]) : -1;
  };
  for (uint32_t i = 0; i < n; ++i) {
    const int lam = le_lambda(le[i]), ell = le_ell(le[i]), st = straddle_of(i);
    for (int d = lam + 1; d < ell; ++d)
      if (d <= st) e.meta.level_count[d]++;
  }
  nodebase[n] = run;
  e.meta.num_nodes = run;
  const uint32_t M = run;
  level_scan(&e.meta, M);
  e.nodeA.assign(M, make_float4(0, 0, 0, 0));
  e.nodeB.assign(M, make_uint4(0, 0, 0, 0));
  e.node_mass.assign(M, 0.f);
  e.parent.assign(M, 0);
  e.level_nodes.assign(M, 0);
  e.sums.assign(M, NodeSums{0, 0, 0, 0, 0, 0, 0, 0});
  e.rec.assign(M, NodeRec{0, 0, 0, 0.f, 0});
  e.ndepth.assign(M, 0);
  e.t = TreeArrays{e.nodeA.data(), e.nodeB.data(), e.rec.data(), e.ndepth.data(), e.node_mass.data(),
                   e.parent.data(), e.sums.data(), e.level_nodes.data(), nullptr, M};
  std::vector<std::vector<uint32_t>> local(kLevels);
  HostSink sink{&local, &e.meta};
  const uint32_t slab = e.slab ? e.slab : n;
  for (uint32_t lo = 0; lo < n; lo += slab) {
    // one emit CTA: its bodies' nodes, then the sums of the cells inside the slab, deepest level first
    const uint32_t hi = std::min<uint64_t>(n, (uint64_t)lo + slab);
    for (auto& v : local) v.clear();
    for (uint32_t i = lo; i < hi; ++i)
      emit_nodes_for_body(e.keys.data(), n, i, le[i], nodebase.data(), M, e.pqr.data(), e.accm.data(),
                          leaf_capacity, thread_capacity, r.size, dcap, e.t, sink, 0, 0, straddle_of(i),
                          /*internal_ranges=*/true);
    for (int level = kLevels - 1; level >= 0; --level)
      for (uint32_t node : local[level]) {
        aggregate_node_ranged(node, level, e.t);
        finalize_node(node, r.size, e.pqr.data(), e.accm.data(), e.t, SubtreeEndCount{});
      }
  }
  // level sweeps over the bucketed cells, then the export sweep (parents, masses, counts, chargeless centres)
  for (int level = kMaxLevels - 1; level >= 0; --level)
    for (uint32_t k = e.meta.level_start[level]; k < e.meta.level_start[level + 1]; ++k) {
      aggregate_node_ranged(e.level_nodes[k], level, e.t);
      finalize_node(e.level_nodes[k], r.size, e.pqr.data(), e.accm.data(), e.t, SubtreeEndCount{});
    }
  e.meta.num_internal = e.meta.internal_total;
  if (e.slab) {
    // the export sweep visits every internal cell: rebuild the buckets to hold them all (as psim_download_nodes does)
    for (int l = 0; l < kLevels; ++l) e.meta.level_count[l] = 0;
    for (uint32_t node = 0; node < M; ++node)
      if (!(e.nodeB[node].w & kNodeLeaf)) e.meta.level_count[e.nodeB[node].w & kNodeDepthMask]++;
    level_scan(&e.meta, M);
    for (uint32_t node = 0; node < M; ++node)
      if (!(e.nodeB[node].w & kNodeLeaf)) {
        const uint32_t d = e.nodeB[node].w & kNodeDepthMask;
        e.level_nodes[e.meta.level_start[d] + e.meta.level_cursor[d]++] = node;
      }
  }
  if (M && (e.nodeB[0].w & kNodeLeaf)) {
    float lm = 0.f;
    if (!(e.nodeB[0].w & kNodeZeroAgg))
      for (uint32_t b = e.nodeB[0].y; b < e.nodeB[0].y + e.nodeB[0].z; ++b) lm += e.accm[b].w;
    e.node_mass[0] = lm;
  }
  for (int level = kMaxLevels - 1; level >= 0; --level)
    for (uint32_t k = e.meta.level_start[level]; k < e.meta.level_start[level + 1]; ++k)
      aggregate_node(e.level_nodes[k], r.size, e.pqr.data(), e.accm.data(), e.t);
  e.irank.assign(M, 0);
  uint32_t rk = 0;
  for (uint32_t i = 0; i < M; ++i) {
    e.irank[i] = rk;
    if (!(e.nodeB[i].w & kNodeLeaf)) rk++;
  }
  return M;
}

uint64_t emu_reference_node_count(void* h) {
  Emu& e = *static_cast<Emu*>(h);
  return e.n ? 4ull * e.meta.num_internal + 1ull : 0ull;
}
void emu_meta(void* h, uint32_t* out8, float* root3) {
  Emu& e = *static_cast<Emu*>(h);
  out8[0] = e.meta.num_nodes, out8[1] = e.meta.num_internal, out8[2] = e.meta.max_depth;
  out8[3] = e.meta.dcap, out8[4] = e.meta.num_zero_leaves, out8[5] = e.meta.num_cap_leaves;
  out8[6] = e.meta.err, out8[7] = e.meta.n;
  root3[0] = e.meta.root.cx, root3[1] = e.meta.root.cy, root3[2] = e.meta.root.size;
}
void emu_export_nodes(void* h, PsimNodeOut* out, uint64_t cap) {
  Emu& e = *static_cast<Emu*>(h);
  memset(out, 0, cap * sizeof(PsimNodeOut));
  for (uint32_t node = 0; node < e.meta.num_nodes; ++node)
    export_node(node, e.keys.data(), e.meta.root, e.t, e.irank.data(), out, cap);
}

// Serial emulation of bh_walk (traverse.cuh) for groups of 32 consecutive targets: the same
// per-lane skip index and warp-level descend vote, in the reference's arithmetic order.
void emu_set_slab(void* h, uint32_t slab) { static_cast<Emu*>(h)->slab = slab; }
void emu_set_params(void* h, float theta, float epsilon) {
  Emu& e = *static_cast<Emu*>(h);
  e.t_sq = theta * theta;
  e.e_sq = epsilon * epsilon;
}
void emu_walk(void* h, uint32_t m, const float* pts_xy, const float* q, const float* radius, float k_e,
              float* out_xy, uint64_t* warp_steps, uint64_t* pairs) {
  Emu& e = *static_cast<Emu*>(h);
  const uint32_t M = e.meta.num_nodes;
  uint64_t steps = 0, P = 0;
  for (uint32_t g = 0; g < (m + 31) / 32; ++g) {
    float ax[32] = {0}, ay[32] = {0};
    uint32_t skip[32];
    for (int l = 0; l < 32; ++l) skip[l] = (g * 32 + l < m) ? 0u : 0xffffffffu;
    uint32_t nidx = 0;
    while (nidx < M) {
      const float4 na = e.nodeA[nidx];
      const uint4 nb = e.nodeB[nidx];
      ++steps;
      bool descend = false;
      for (int l = 0; l < 32; ++l) {
        const uint32_t i = g * 32 + l;
        if (i >= m) continue;
        const float px = pts_xy[2 * i], py = pts_xy[2 * i + 1];
        const float rad = radius ? radius[i] : 0.f, kq = k_e * (q ? q[i] : 1.f);
        const bool active = nidx >= skip[l];
        const float dx = px - na.x, dy = py - na.y;
        const float d_sq = (dx * dx) + (dy * dy);
        const float dist = sqrtf(d_sq);
        const float dist_adj = fmaxf(dist - rad, 0.0f);
        const bool accept = (na.w * na.w) < ((dist_adj * dist_adj) * e.t_sq);
        const bool leaf = (nb.w & kNodeLeaf) != 0;
        if (active) {
          if (accept) {
            const float r_eff = fmaxf(dist, rad + na.w * 0.5f);
            const float denom = (r_eff * r_eff + e.e_sq) * r_eff;
            const float s = (kq * na.z) / denom;
            ax[l] += dx * s, ay[l] += dy * s;
            skip[l] = nb.x;
          } else if (leaf) {
            for (uint32_t b = nb.y; b < nb.y + nb.z; ++b) {
              const float4 s4 = e.pqr[b];
              const float ex = s4.x - px, ey = s4.y - py;
              if ((ex * ex) + (ey * ey) < 1e-6f) continue;
              ++P;
              const float bx = px - s4.x, by = py - s4.y;
              const float bd = sqrtf((bx * bx) + (by * by));
              const float r_eff = fmaxf(bd, rad + s4.w);
              const float denom = (r_eff * r_eff + e.e_sq) * r_eff;
              const float s = fminf((kq * s4.z) / denom, 3.402823466e+38f);
              ax[l] += bx * s, ay[l] += by * s;
            }
            skip[l] = nb.x;
          } else {
            descend = true;
          }
        }
      }
      nidx = descend ? nidx + 1 : nb.x;
    }
    for (int l = 0; l < 32; ++l) {
      const uint32_t i = g * 32 + l;
      if (i < m) out_xy[2 * i] = ax[l], out_xy[2 * i + 1] = ay[l];
    }
  }
  if (warp_steps) *warp_steps = steps;
  if (pairs) *pairs = P;
}

// ---- serial replay of a `world`-rank sharded build (csrc/shard.cuh) on the sorted bodies
// of the last emu_build: bin-aligned key ranges, virtual halo keys, per-bin tables, top heap, per-rank
// compaction.  The device kernels' glue is restated here; the per-body / per-node logic is the shared
// host+device code.  out[0..7] = mismatches against the single build: nodeA, nodeB, rec, ndepth (per
// node), sentinel depths, heap slots with more than one writer, travA, travB.  Returns the node total.
struct HostShardSink {
  static constexpr bool kTop = true;
  static constexpr bool kStrict = false;
  uint32_t strict_direct() const { return 0xffffffffu; }
  void strict_chain(uint32_t, uint32_t, uint32_t, uint32_t) {}
  TopRec* heap;
  uint32_t* multi_writer;
  uint32_t level_slot(int) { return 0; }  // the level buckets are collected from the (lambda, ell) levels
  void local_node(int, uint32_t) {}
  void zero_leaf() {}
  void cap_leaf() {}
  void top_leaf(int d, uint64_t key, uint32_t node, const NodeRec& r) {
    TopRec& h = heap[top_slot(d, key)];
    if (h.state != kTopAbsent) ++*multi_writer;
    h.r = r, h.node = node, h.state = kTopComplete;
  }
  void top_internal(int d, uint64_t key, uint32_t node) {
    TopRec& h = heap[top_slot(d, key)];
    if (h.state != kTopAbsent) ++*multi_writer;
    h.node = node, h.state = d == kShardDepth ? kTopBin : kTopInternal;
  }
};

uint32_t emu_shard_check(void* hd, uint32_t world, uint32_t leaf_capacity, uint32_t thread_capacity, uint64_t* out) {
  Emu& e = *static_cast<Emu*>(hd);
  for (int k = 0; k < 8; ++k) out[k] = 0;
  const uint32_t n = e.n;
  if (n == 0) return 0;
  const uint32_t M = e.meta.num_nodes;
  const uint32_t c_eff = effective_capacity(leaf_capacity, thread_capacity);
  const int dcap = (int)e.meta.dcap;
  const float root_size = e.meta.root.size;
  // bins, prefix, splitters (bin_split_kernel)
  std::vector<uint32_t> binhist(kBins, 0), binprefix(kBins + 1, 0);
  for (uint32_t i = 0; i < n; ++i) binhist[(uint32_t)(e.keys[i] >> 48)]++;
  for (uint32_t b = 0; b < kBins; ++b) binprefix[b + 1] = binprefix[b] + binhist[b];
  std::vector<uint32_t> bin_lo(world + 1), body_lo(world + 1);
  for (uint32_t r = 0; r <= world; ++r) {
    const uint64_t target = (uint64_t)n * r / world;
    uint32_t lo = (uint32_t)(std::lower_bound(binprefix.begin(), binprefix.begin() + kBins, (uint32_t)target) - binprefix.begin());
    if (r == 0) lo = 0;
    if (r == world) lo = kBins;
    bin_lo[r] = lo, body_lo[r] = binprefix[lo];
  }
  auto bin_of_body = [&](uint32_t g) {
    return (uint32_t)(std::upper_bound(binprefix.begin(), binprefix.begin() + kBins + 1, g) - binprefix.begin()) - 1;
  };
  const uint32_t H = c_eff + 1;
  struct Rank {
    uint32_t s_lo, nl, hl, L, body_base, M_local, node_off, T_local, trav_off;
    std::vector<uint64_t> lkeys;
    std::vector<uint16_t> le;
    std::vector<uint32_t> nodebase;
    std::vector<std::vector<uint32_t>> buckets;
  };
  std::vector<Rank> R(world);
  // phases 0-2: halo keys, levels, local node scan
  for (uint32_t r = 0; r < world; ++r) {
    Rank& k = R[r];
    k.s_lo = body_lo[r], k.nl = body_lo[r + 1] - body_lo[r];
    k.hl = std::min(k.s_lo, H);
    const uint32_t after = n - body_lo[r + 1];
    k.L = k.hl + k.nl + std::min(after, H);
    k.body_base = k.s_lo - k.hl;
    k.lkeys.resize(k.L), k.le.assign(k.L, 1), k.nodebase.assign(k.L + 1, 0);
    for (uint32_t a = 0; a < k.L; ++a) {
      const bool local = a >= k.hl && a < k.hl + k.nl;
      k.lkeys[a] = local ? e.keys[k.body_base + a] : (uint64_t)bin_of_body(k.body_base + a) << 48;
    }
    for (uint32_t a = k.hl; a < k.hl + k.nl; ++a)
      k.le[a] = body_levels(k.lkeys.data(), e.pqr.data() + k.body_base, k.L, a, c_eff, dcap);
    uint32_t run = 0;
    for (uint32_t a = 0; a < k.L; ++a) k.nodebase[a] = run, run += le_nodes(k.le[a]);
    k.nodebase[k.L] = run;
    k.M_local = run;
  }
  // table 1 + offsets (table_nodes_kernel, resolve_table_kernel)
  uint32_t off = 0;
  for (uint32_t r = 0; r < world; ++r) R[r].node_off = off, off += R[r].M_local;
  const uint32_t M_total = off;
  auto owner_of = [&](uint32_t b) {
    uint32_t o = 0;
    while (o + 1 < world && bin_lo[o + 1] <= b) ++o;
    return o;
  };
  std::vector<uint32_t> nb_bin(kBins + 1, M_total);
  for (uint32_t b = 0; b < kBins; ++b) {
    const uint32_t g = binprefix[b];
    if (g >= n) continue;
    const uint32_t bb = binhist[b] ? b : bin_of_body(g);
    const Rank& k = R[owner_of(bb)];
    nb_bin[b] = k.node_off + k.nodebase[binprefix[bb] - k.body_base];
  }
  // global arrays every rank writes its own index range of
  std::vector<float4> nodeA(M_total, make_float4(0, 0, 0, 0));
  std::vector<uint4> nodeB(M_total, make_uint4(0, 0, 0, 0));
  std::vector<NodeRec> rec(M_total, NodeRec{0, 0, 0, 0.f, 0});
  std::vector<uint8_t> ndepth(M_total + 1, 0);
  std::vector<uint32_t> dummy_levels(1);
  TreeArrays t{nodeA.data(), nodeB.data(), rec.data(), ndepth.data(), nullptr, nullptr, nullptr, nullptr, nullptr, M_total};
  std::vector<TopRec> heap(kTopSlots);
  memset(heap.data(), 0, heap.size() * sizeof(TopRec));
  uint32_t multi_writer = 0;
  // phase 3: globalize nodebase, emit
  std::vector<uint8_t> sentinel(world, 0);
  for (uint32_t r = 0; r < world; ++r) {
    Rank& k = R[r];
    for (uint32_t a = k.hl; a < k.hl + k.nl; ++a) k.nodebase[a] += k.node_off;
    for (uint32_t a = k.hl + k.nl; a < k.L; ++a) k.nodebase[a] = nb_bin[(uint32_t)(k.lkeys[a] >> 48)];
    if (k.nl > 0 && k.hl + k.nl < k.L) sentinel[r] = (uint8_t)(lcp_levels(k.lkeys[k.hl + k.nl - 1], k.lkeys[k.hl + k.nl]) + 1);
    k.buckets.assign(kLevels, {});
    for (uint32_t a = k.hl; a < k.hl + k.nl; ++a) {
      const int lam = le_lambda(k.le[a]), ell = le_ell(k.le[a]);
      for (int d = lam + 1; d < ell; ++d)
        if (d >= kShardDepth) k.buckets[d].push_back(k.nodebase[a] + (uint32_t)(d - lam - 1));
    }
    HostShardSink sink{heap.data(), &multi_writer};
    TreeArrays te = t;
    te.level_nodes = dummy_levels.data();
    for (uint32_t a = k.hl; a < k.hl + k.nl; ++a)
      emit_nodes_for_body(k.lkeys.data(), k.L, a, k.le[a], k.nodebase.data(), M_total, e.pqr.data() + k.body_base,
                          (const float4*)nullptr, leaf_capacity, thread_capacity, root_size, dcap, te, sink,
                          k.body_base, kShardDepth);
  }
  out[5] = multi_writer;
  // The sentinel depth must be the real depth of the first node after the rank's piece whenever a local
  // sweep can read it, i.e. when the last local body sits in a leaf of its own rank (a leaf that straddles
  // the boundary lies above the bins: no cell of the local sweeps ends there, the heap handles it).
  for (uint32_t r = 0; r < world; ++r) {
    const Rank& k = R[r];
    const uint32_t first_remote = k.node_off + k.M_local;
    if (k.nl == 0 || first_remote >= M_total || k.hl + k.nl >= k.L) continue;
    const uint32_t last = k.hl + k.nl - 1;
    uint32_t head = last;  // the body that starts the last local body's leaf
    while (head > k.hl && !(le_lambda(k.le[head]) < le_ell(k.le[head]))) --head;
    const bool own_leaf = le_ell(k.le[head]) > lcp_levels(k.lkeys[last], k.lkeys[last + 1]);
    if (own_leaf && sentinel[r] != (ndepth[first_remote] & kDepthMask8)) out[4]++;
  }
  // local sweeps (levels >= kShardDepth), owned internal bins into the heap
  for (uint32_t r = 0; r < world; ++r)
    for (int level = kMaxLevels - 1; level >= kShardDepth; --level)
      for (uint32_t node : R[r].buckets[level]) aggregate_node_lean(node, level, M_total, t);
  for (uint32_t b = 0; b < kBins; ++b) {
    TopRec& hslot = heap[top_base(kShardDepth) + b];
    if (hslot.state == kTopBin) hslot.r = rec[hslot.node], hslot.state = kTopComplete;
  }
  // phase 4: heap sweep, write-back, finalize
  for (int d = kShardDepth - 1; d >= 0; --d)
    for (uint32_t p = 0; p < (1u << (2 * d)); ++p) heap_sweep_cell(heap.data(), d, p);
  for (uint32_t s = 0; s < top_base(kShardDepth); ++s)
    if (heap[s].state == kTopComputed) {
      rec[heap[s].node] = heap[s].r;
      if (heap[s].r.aq > 0.0) ndepth[heap[s].node] |= (uint8_t)kDepthCharged;
    }
  for (uint32_t r = 0; r < world; ++r) {
    const Rank& k = R[r];
    const SubtreeEndShard end{k.node_off, k.node_off + k.M_local, M_total, n, k.body_base, nodeB.data(), k.lkeys.data(),
                              binprefix.data()};
    for (uint32_t node = k.node_off; node < k.node_off + k.M_local; ++node)
      finalize_node(node, root_size, e.pqr.data(), e.accm.data(), t, end);
  }
  // compare with the single build
  if (M_total != M) out[0] = out[1] = out[2] = out[3] = 1u << 30;
  for (uint32_t i = 0; i < M && i < M_total; ++i) {
    const bool leaf = (e.nodeB[i].w & kNodeLeaf) != 0;
    // (the single emulation also ran the export sweep, which gives chargeless internal cells their mass /
    // centroid centre; the build leaves those at (0, 0) and no field sum ever reads them)
    if ((leaf || e.rec[i].aq > 0.0) && memcmp(&nodeA[i], &e.nodeA[i], sizeof(float4)) != 0) out[0]++;
    const uint4 a = nodeB[i], b = e.nodeB[i];
    if (a.x != b.x || a.y != b.y || a.w != b.w || (leaf && a.z != b.z)) out[1]++;
    const NodeRec &ra = rec[i], &rb = e.rec[i];
    if (ra.aq != rb.aq || ra.aqx != rb.aqx || ra.aqy != rb.aqy || memcmp(&ra.charge, &rb.charge, 4) != 0 ||
        (ra.next & kNextMask) != (rb.next & kNextMask))
      out[2]++;
    if (ndepth[i] != e.ndepth[i]) out[3]++;
  }
  // phase 5: per-rank compaction with the traversal table, against a plain compaction of the single tree
  std::vector<uint32_t> rank_single(M + 1, 0);
  for (uint32_t i = 0; i < M; ++i) rank_single[i + 1] = rank_single[i] + (e.ndepth[i] >> 7);
  const uint32_t T = rank_single[M];
  std::vector<float4> sA(T);
  std::vector<uint4> sB(T);
  for (uint32_t i = 0; i < M; ++i)
    if (e.nodeB[i].w & kNodeCharged) {
      uint4 nb = e.nodeB[i];
      nb.x = nb.x < M ? rank_single[nb.x] : T;
      if (!(nb.w & kNodeLeaf)) nb.z = 0;
      sA[rank_single[i]] = e.nodeA[i], sB[rank_single[i]] = nb;
    }
  std::vector<std::vector<uint32_t>> lrank(world);
  uint32_t toff = 0;
  for (uint32_t r = 0; r < world; ++r) {
    Rank& k = R[r];
    lrank[r].assign(k.M_local + 1, 0);
    for (uint32_t j = 0; j < k.M_local; ++j) lrank[r][j + 1] = lrank[r][j] + (ndepth[k.node_off + j] >> 7);
    k.T_local = lrank[r][k.M_local], k.trav_off = toff, toff += k.T_local;
  }
  const uint32_t T_total = toff;
  std::vector<uint32_t> trav_bin(kBins + 1, T_total);
  for (uint32_t b = 0; b < kBins; ++b) {
    const uint32_t g = binprefix[b];
    if (g >= n) continue;
    const uint32_t bb = binhist[b] ? b : bin_of_body(g);
    const uint32_t o = owner_of(bb);
    const uint32_t j = nb_bin[bb] - R[o].node_off;
    trav_bin[b] = R[o].trav_off + (j < R[o].M_local ? lrank[o][j] : R[o].T_local);
  }
  if (T_total != T) out[6] = out[7] = 1u << 30;
  std::vector<float4> dA(T_total);
  std::vector<uint4> dB(T_total);
  for (uint32_t r = 0; r < world; ++r) {
    const Rank& k = R[r];
    for (uint32_t j = 0; j < k.M_local; ++j) {
      uint4 nb = nodeB[k.node_off + j];
      if (!(nb.w & kNodeCharged)) continue;
      const uint32_t x = nb.x;
      if (x >= M_total) nb.x = T_total;
      else if (x - k.node_off < k.M_local) nb.x = k.trav_off + lrank[r][x - k.node_off];
      else nb.x = trav_bin[(uint32_t)(std::lower_bound(nb_bin.begin(), nb_bin.begin() + kBins, x) - nb_bin.begin())];
      const uint32_t rr = k.trav_off + lrank[r][j];
      dA[rr] = nodeA[k.node_off + j], dB[rr] = nb;
    }
  }
  for (uint32_t i = 0; i < T && i < T_total; ++i) {
    if (memcmp(&dA[i], &sA[i], sizeof(float4)) != 0) out[6]++;
    if (memcmp(&dB[i], &sB[i], sizeof(uint4)) != 0) out[7]++;
  }
  return M_total;
}

// ---- per-target interaction signatures -----------------------------------------------------------
// sig[3 i + 0] = sum of hash(node) over the monopoles target i accepted, [1] = sum of hash(body) over its
// direct terms, [2] = number of internal nodes it opened.  Two walks that give equal signatures summed the
// same interaction sets (up to a 2^-64 hash collision), whatever the order.
static inline uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// the reference-order walk (emu_walk's logic, one target at a time: the warp vote only changes which nodes
// the WARP visits, not what a target sums)
void emu_walk_signatures(void* h, uint32_t m, const float* pts_xy, const float* radius, uint64_t* sig) {
  Emu& e = *static_cast<Emu*>(h);
  const uint32_t M = e.meta.num_nodes;
  for (uint32_t i = 0; i < m; ++i) {
    const float px = pts_xy[2 * i], py = pts_xy[2 * i + 1], rad = radius ? radius[i] : 0.f;
    uint64_t sa = 0, sp = 0, so = 0;
    uint32_t n = 0;
    while (n < M) {
      const float4 na = e.nodeA[n];
      const uint4 nb = e.nodeB[n];
      const float dx = px - na.x, dy = py - na.y;
      const float dist = sqrtf((dx * dx) + (dy * dy));
      const float dist_adj = fmaxf(dist - rad, 0.0f);
      if ((na.w * na.w) < ((dist_adj * dist_adj) * e.t_sq)) {
        sa += mix64(n);
        n = nb.x;
      } else if (nb.w & kNodeLeaf) {
        for (uint32_t b = nb.y; b < nb.y + nb.z; ++b) {
          const float ex = e.pqr[b].x - px, ey = e.pqr[b].y - py;
          if ((ex * ex) + (ey * ey) < 1e-6f) continue;
          sp += mix64(b);
        }
        n = nb.x;
      } else {
        ++so;
        n = n + 1;
      }
    }
    sig[3 * i] = sa, sig[3 * i + 1] = sp, sig[3 * i + 2] = so;
  }
}

// Diagnostic (tools/walk_stats.py): interaction-list statistics of the group walk on the CHARGED nodes only (what the
// traversal arrays hold), with the group's targets split into `nsub` sub-boxes for the classification.
// out: [0] walks [1] rounds [2] nodes visited [3] sure entries [4] undecided entries (per-target test) [5] entries decided
// per sub-box with mixed outcome [6] nodes opened for every target [7] leaves with direct terms [8] accepted (target, node)
// pairs [9] targets reaching sure entries [10] targets reaching undecided entries [11] direct (target, body) terms
void emu_group_stats(void* h, uint32_t m, const float* pts_xy, const float* radius, float theta, int nsub, uint64_t* out) {
  Emu& e = *static_cast<Emu*>(h);
  const uint32_t M = e.meta.num_nodes;
  const float inv_theta = 1.0f / theta;
  const float INF = INFINITY;
  for (int k = 0; k < 14; ++k) out[k] = 0;
  const int per = 32 / nsub;
  for (uint32_t g = 0; g < (m + 31) / 32; ++g) {
    float px[32], py[32], rad[32];
    uint32_t live_mask = 0;
    float bx0[8], bx1[8], by0[8], by1[8], rmin[8], rmax[8];
    uint32_t submask[8];
    for (int s2 = 0; s2 < nsub; ++s2) bx0[s2] = by0[s2] = rmin[s2] = INF, bx1[s2] = by1[s2] = rmax[s2] = -INF, submask[s2] = 0;
    for (int l = 0; l < 32; ++l) {
      const uint32_t i = g * 32 + l;
      if (i >= m) continue;
      live_mask |= 1u << l;
      px[l] = pts_xy[2 * i], py[l] = pts_xy[2 * i + 1], rad[l] = radius ? radius[i] : 0.f;
      const int s2 = l / per;
      submask[s2] |= 1u << l;
      bx0[s2] = fminf(bx0[s2], px[l]), bx1[s2] = fmaxf(bx1[s2], px[l]), by0[s2] = fminf(by0[s2], py[l]), by1[s2] = fmaxf(by1[s2], py[l]);
      rmin[s2] = fminf(rmin[s2], rad[l]), rmax[s2] = fmaxf(rmax[s2], rad[l]);
    }
    if (!live_mask || !M) continue;
    out[0]++;
    std::vector<std::pair<uint32_t, uint32_t>> cur, nxt;
    cur.emplace_back(0u, live_mask);
    size_t pos = 0;
    while (pos < cur.size() || !nxt.empty()) {
      if (pos >= cur.size()) cur.swap(nxt), nxt.clear(), pos = 0;
      const size_t k = std::min<size_t>(32, cur.size() - pos);
      out[1]++;
      for (size_t l = 0; l < k; ++l) {
        const uint32_t node = cur[pos + l].first, mask = cur[pos + l].second;
        out[2]++;
        const float4 na = e.nodeA[node];
        const uint4 nb = e.nodeB[node];
        const bool leaf = (nb.w & kNodeLeaf) != 0;
        uint32_t sure_m = 0, rej_m = 0, und_m = 0;
        for (int s2 = 0; s2 < nsub; ++s2) {
          const uint32_t sm = submask[s2] & mask;
          if (!sm) continue;
          const float s_t = na.w * inv_theta;
          const float ddx = fmaxf(fmaxf(bx0[s2] - na.x, na.x - bx1[s2]), 0.0f), ddy = fmaxf(fmaxf(by0[s2] - na.y, na.y - by1[s2]), 0.0f);
          const float fx = fmaxf(na.x - bx0[s2], bx1[s2] - na.x), fy = fmaxf(na.y - by0[s2], by1[s2] - na.y);
          const float dmin2 = ddx * ddx + ddy * ddy, dmax2 = fx * fx + fy * fy;
          const float la = s_t + rmax[s2], lr = s_t + rmin[s2];
          if (dmin2 > la * la * 1.00002f) sure_m |= sm;
          else if (dmax2 < lr * lr * 0.99998f) rej_m |= sm;
          else und_m |= sm;
        }
        uint32_t acc_mask = sure_m;
        for (int t = 0; t < 32; ++t) {
          if (!((und_m >> t) & 1u)) continue;
          const float dx = px[t] - na.x, dy = py[t] - na.y;
          const float d_sq = (dx * dx) + (dy * dy);
          const float dist_adj = fmaxf(sqrtf(d_sq) - rad[t], 0.0f);
          if ((na.w * na.w) < ((dist_adj * dist_adj) * e.t_sq)) acc_mask |= 1u << t;
        }
        if (und_m) {
          out[4]++, out[10] += __builtin_popcount(mask);
          if ((acc_mask & mask) == mask) out[12]++;       // undecided, yet every reaching target accepts
          else if ((acc_mask & mask) == 0) out[13]++;     // ... or none does
        } else if (sure_m && rej_m) out[5]++;
        else if (sure_m) out[3]++, out[9] += __builtin_popcount(mask);
        out[8] += __builtin_popcount(acc_mask);
        const uint32_t rem = mask & ~acc_mask;
        if (!rem) continue;
        if (!leaf) {
          if (!und_m && !sure_m) out[6]++;
          for (uint32_t c = node + 1; c != nb.x; c = e.nodeB[c].x)
            if (e.nodeB[c].w & kNodeCharged) nxt.emplace_back(c, rem);
        } else {
          out[7]++;
          out[11] += (uint64_t)__builtin_popcount(rem) * nb.z;
        }
      }
      pos += k;
    }
  }
}

// Serial port of bh_group_walk (traverse.cuh): 32 consecutive targets share one walk; nodes are classified
// against the group's bounding box with the same conservative margins, undecided nodes take the reference's
// test per target, the ring buffer switches to last-in-first-out above kLifoAbove like the device's.
// Fields in the reference's per-term arithmetic (the order of the additions differs from acc_pos).
void emu_group_walk(void* h, uint32_t m, const float* pts_xy, const float* q, const float* radius, float k_e,
                    float theta, float* out_xy, uint64_t* sig, uint64_t* stats /* [0] nodes visited, [1] lifo rounds */) {
  Emu& e = *static_cast<Emu*>(h);
  const uint32_t M = e.meta.num_nodes;
  constexpr int kCap = 512, kLifo = kCap - 192;
  const float inv_theta = 1.0f / theta;
  const float INF = INFINITY;
  uint64_t visited = 0, lifo_rounds = 0;
  for (uint32_t g = 0; g < (m + 31) / 32; ++g) {
    float px[32], py[32], rad[32], kq[32], ax[32] = {0}, ay[32] = {0};
    uint64_t sa[32] = {0}, sp[32] = {0}, so[32] = {0};
    uint32_t live_mask = 0;
    float bx0 = INF, bx1 = -INF, by0 = INF, by1 = -INF, rmin = INF, rmax = -INF;
    bool box_ok = true;
    for (int l = 0; l < 32; ++l) {
      const uint32_t i = g * 32 + l;
      if (i >= m) continue;
      live_mask |= 1u << l;
      px[l] = pts_xy[2 * i], py[l] = pts_xy[2 * i + 1], rad[l] = radius ? radius[i] : 0.f, kq[l] = k_e * (q ? q[i] : 1.f);
      bx0 = fminf(bx0, px[l]), bx1 = fmaxf(bx1, px[l]), by0 = fminf(by0, py[l]), by1 = fmaxf(by1, py[l]);
      rmin = fminf(rmin, rad[l]), rmax = fmaxf(rmax, rad[l]);
      if (!(fabsf(px[l]) < INF && fabsf(py[l]) < INF && fabsf(rad[l]) < INF)) box_ok = false;
    }
    if (!live_mask || !M) continue;
    std::vector<uint32_t> st_node(kCap), st_mask(kCap);
    st_node[0] = 0, st_mask[0] = live_mask;
    int head = 0, size = 1;
    while (size > 0) {
      const bool lifo = size > kLifo;
      const int k = lifo ? 1 : std::min(size, 32);
      lifo_rounds += lifo;
      uint32_t node[32], mask[32];
      for (int l = 0; l < k; ++l) {
        int idx = lifo ? head + size - 1 : head + l;
        if (idx >= kCap) idx -= kCap;
        node[l] = st_node[idx], mask[l] = st_mask[idx];
      }
      if (!lifo) head = (head + k) % kCap;
      size -= k;
      visited += k;
      std::vector<std::pair<uint32_t, uint32_t>> push;  // (child, mask), in lane order like the device's prefix scan
      for (int l = 0; l < k; ++l) {
        const float4 na = e.nodeA[node[l]];
        const uint4 nb = e.nodeB[node[l]];
        const bool leaf = (nb.w & kNodeLeaf) != 0;
        int cls = 0;
        if (box_ok) {
          const float s_t = na.w * inv_theta;
          const float ddx = fmaxf(fmaxf(bx0 - na.x, na.x - bx1), 0.0f), ddy = fmaxf(fmaxf(by0 - na.y, na.y - by1), 0.0f);
          const float fx = fmaxf(na.x - bx0, bx1 - na.x), fy = fmaxf(na.y - by0, by1 - na.y);
          const float dmin2 = ddx * ddx + ddy * ddy, dmax2 = fx * fx + fy * fy;
          const float la = s_t + rmax, lr = s_t + rmin;
          if (dmin2 > la * la * 1.00002f) cls = 1;
          else if (dmax2 < lr * lr * 0.99998f) cls = 2;
        }
        uint32_t acc_mask = 0;
        for (int t = 0; t < 32; ++t) {
          if (!((mask[l] >> t) & 1u)) continue;
          const float dx = px[t] - na.x, dy = py[t] - na.y;
          const float d_sq = (dx * dx) + (dy * dy);
          bool acc;
          if (cls == 1) acc = true;          // the "sure" list: no test
          else if (cls == 2) acc = false;    // every target rejects: no test
          else {
            const float lim = fmaf(na.w, inv_theta, rad[t]), lim2 = lim * lim;
            acc = d_sq > lim2 * 1.00002f;
            if (!acc && !(d_sq < lim2 * 0.99998f)) {
              const float dist_adj = fmaxf(sqrtf(d_sq) - rad[t], 0.0f);
              acc = (na.w * na.w) < ((dist_adj * dist_adj) * e.t_sq);
            }
          }
          if (!acc) continue;
          acc_mask |= 1u << t;
          sa[t] += mix64(node[l]);
          const float dist = sqrtf(d_sq);
          const float r_eff = fmaxf(dist, rad[t] + na.w * 0.5f);
          const float denom = (r_eff * r_eff + e.e_sq) * r_eff;
          const float sc = (kq[t] * na.z) / denom;
          ax[t] += dx * sc, ay[t] += dy * sc;
        }
        const uint32_t rem = mask[l] & ~acc_mask;
        if (!rem) continue;
        if (!leaf) {
          for (int t = 0; t < 32; ++t) so[t] += (rem >> t) & 1u;
          for (uint32_t c = node[l] + 1; c != nb.x; c = e.nodeB[c].x) push.emplace_back(c, rem);
        } else {
          for (int t = 0; t < 32; ++t) {
            if (!((rem >> t) & 1u)) continue;
            for (uint32_t b = nb.y; b < nb.y + nb.z; ++b) {
              const float4 s4 = e.pqr[b];
              const float ex = s4.x - px[t], ey = s4.y - py[t];
              if ((ex * ex) + (ey * ey) < 1e-6f) continue;
              sp[t] += mix64(b);
              const float bx = px[t] - s4.x, by = py[t] - s4.y;
              const float bd = sqrtf((bx * bx) + (by * by));
              const float r_eff = fmaxf(bd, rad[t] + s4.w);
              const float denom = (r_eff * r_eff + e.e_sq) * r_eff;
              const float sc = fminf((kq[t] * s4.z) / denom, 3.402823466e+38f);
              ax[t] += bx * sc, ay[t] += by * sc;
            }
          }
        }
      }
      for (auto& pm : push) {
        int o2 = (head + size) % kCap;
        if (size >= kCap) {  // the device's capacity argument says this cannot happen
          stats[0] = ~0ull;
          return;
        }
        st_node[o2] = pm.first, st_mask[o2] = pm.second;
        ++size;
      }
    }
    for (int l = 0; l < 32; ++l) {
      const uint32_t i = g * 32 + l;
      if (i >= m) continue;
      out_xy[2 * i] = ax[l], out_xy[2 * i + 1] = ay[l];
      sig[3 * i] = sa[l], sig[3 * i + 1] = sp[l], sig[3 * i + 2] = so[l];
    }
  }
  stats[0] = visited, stats[1] = lifo_rounds;
}

void emu_sorted_bodies(void* h, float* pqr_out) {
  Emu& e = *static_cast<Emu*>(h);
  memcpy(pqr_out, e.pqr.data(), (size_t)e.n * 16);
}

}  // extern "C"

// ---- strict_logic.cuh: the reference's serial f32 running sum, evaluated block-wise (test only) ----
#include "../../particlesim_b200/csrc/strict_logic.cuh"

extern "C" {
// Runs the sum of a[0..n) from s0 three ways: (0) the plain serial loop; (1) per-block functions under a
// binade speculated from an f64 prefix (middle of the block), applied in order with the serial fall-back for
// blocks that are not valid; (2) like (1) but runs of valid blocks are first composed pairwise (the scan
// operator) and applied once.  out[0..2] = the three results, stats[0] = blocks, [1] = serial fall-backs.
void emu_strict_sum(const float* a, uint32_t n, uint32_t block, float s0, float* out, uint64_t* stats) {
  float serial = s0;
  for (uint32_t i = 0; i < n; ++i) serial = serial + a[i];
  out[0] = serial;
  const uint32_t nb = (n + block - 1) / block;
  std::vector<double> pre(nb + 1);
  pre[0] = (double)s0;
  for (uint32_t b = 0; b < nb; ++b) {
    double t = 0;
    for (uint32_t i = b * block; i < std::min(n, (b + 1) * block); ++i) t += (double)a[i];
    pre[b + 1] = pre[b] + t;
  }
  std::vector<BlockFn> fn(nb);
  for (uint32_t b = 0; b < nb; ++b) {
    const int e = spec_exponent(0.5 * (pre[b] + pre[b + 1]));
    blockfn_init(fn[b], e);
    if (e == kBadExp) continue;
    const float iu = inv_ulp(e);
    for (uint32_t i = b * block; i < std::min(n, (b + 1) * block) && fn[b].e != kBadExp; ++i) blockfn_step(fn[b], a[i], iu);
  }
  uint64_t fallbacks = 0;
  float s = s0;
  for (uint32_t b = 0; b < nb; ++b) {
    int e;
    int32_t M;
    if (f32_split(s, e, M) && blockfn_valid(fn[b], e, M)) {
      s = f32_join(e, blockfn_apply(fn[b], M));
    } else {
      ++fallbacks;
      for (uint32_t i = b * block; i < std::min(n, (b + 1) * block); ++i) s = s + a[i];
    }
  }
  out[1] = s;
  // composed application
  s = s0;
  for (uint32_t b = 0; b < nb;) {
    int e;
    int32_t M;
    if (f32_split(s, e, M) && blockfn_valid(fn[b], e, M)) {
      // extend the run while the blocks stay valid given the exact entering mantissa
      int32_t c0 = fn[b].o[0], c1 = fn[b].o[1];
      uint32_t r = b + 1;
      while (r < nb) {
        const int32_t Mr = M + ((M & 1) ? c1 : c0);
        if (!blockfn_valid(fn[r], e, Mr)) break;
        int32_t n0, n1;
        blockfn_compose(n0, n1, c0, c1, fn[r].o[0], fn[r].o[1]);
        c0 = n0, c1 = n1;
        ++r;
      }
      s = f32_join(e, M + ((M & 1) ? c1 : c0));
      b = r;
    } else {
      for (uint32_t i = b * block; i < std::min(n, (b + 1) * block); ++i) s = s + a[i];
      ++b;
    }
  }
  out[2] = s;
  stats[0] = nb, stats[1] = fallbacks;
}
}  // extern "C"

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

using namespace psim;

namespace {
struct HostSink {
  static constexpr bool kTop = false;
  void top_leaf(int, uint64_t, uint32_t, const NodeRec&) {}
  void top_internal(int, uint64_t, uint32_t) {}
  void local_node(int, uint32_t) {}
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
      for (int d = lam + 1; d < ell; ++d) e.meta.level_count[d]++;
      if ((uint32_t)ell > e.meta.max_depth) e.meta.max_depth = ell;
    }
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
  HostSink sink{&e.meta};
  for (uint32_t i = 0; i < n; ++i)
    emit_nodes_for_body(e.keys.data(), n, i, le[i], nodebase.data(), M, e.pqr.data(), e.accm.data(),
                        leaf_capacity, thread_capacity, r.size, dcap, e.t, sink, 0, 0, kMaxLevels + 1,
                        /*internal_ranges=*/true);
  // build sweep, then the export sweep (parents, masses, counts, chargeless centres)
  for (int level = kMaxLevels - 1; level >= 0; --level)
    for (uint32_t k = e.meta.level_start[level]; k < e.meta.level_start[level + 1]; ++k)
      aggregate_node_ranged(e.level_nodes[k], level, e.t);
  for (uint32_t node = 0; node < M; ++node) finalize_node(node, r.size, e.pqr.data(), e.accm.data(), e.t, SubtreeEndCount{});
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

void emu_sorted_bodies(void* h, float* pqr_out) {
  Emu& e = *static_cast<Emu*>(h);
  memcpy(pqr_out, e.pqr.data(), (size_t)e.n * 16);
}

}  // extern "C"

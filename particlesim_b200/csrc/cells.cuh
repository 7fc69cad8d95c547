// cells.cuh — uniform cell list, short-range LJ / repulsion / stack pressure, neighbour queries,
// integrator.
//
// Replaces src/cell_list.rs (rebuild :27-39, coord :47-55, find_neighbors_within :57-85,
// metal_neighbor_count :92-127), src/simulation/forces.rs:182-321 (apply_lj_forces,
// compute_repulsive_force, apply_repulsive_forces, apply_stack_pressure) and Simulation::iterate
// (src/simulation/simulation.rs:1437-1486).
//
// Cell list layout: bodies are counting-sorted by cell id (stable, so each cell lists its bodies in
// body-index order like the reference's per-cell Vec) into `order`; cell c owns
// order[cell_start[c] .. cell_end[c]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sort.cuh"
#include "traverse.cuh"

namespace psim {

struct GridDims {
  uint32_t gx, gy;
  float cell_size, hw, hh;
};

// Rust `f as isize` then clamp(0, g-1) (cell_list.rs:47-55); NaN -> 0
__device__ __forceinline__ uint32_t cell_axis(float p, float min_v, float cell_size, uint32_t g) {
  const float f = floorf(__fdiv_rn(__fsub_rn(p, min_v), cell_size));
  if (!(f == f)) return 0u;
  if (f <= 0.0f) return 0u;
  const float gm1 = (float)(g - 1);
  if (f >= gm1) return g - 1;  // also covers +inf; exact integer compare below for the rest
  const uint32_t v = (uint32_t)f;
  return v > g - 1 ? g - 1 : v;
}

__global__ void __launch_bounds__(256)
    cell_id_kernel(const float4* __restrict__ pqr, uint32_t n, GridDims g,
                   uint32_t* __restrict__ cell, uint32_t* __restrict__ idx) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 p = pqr[i];
    const uint32_t cx = cell_axis(p.x, -g.hw, g.cell_size, g.gx);
    const uint32_t cy = cell_axis(p.y, -g.hh, g.cell_size, g.gy);
    cell[i] = cx + cy * g.gx;
    idx[i] = i;
  }
}

// sorted cell ids -> [start, end) per cell; also publishes the sorted order into a fixed buffer
__global__ void __launch_bounds__(256)
    cell_ranges_kernel(const uint32_t* __restrict__ cell0, const uint32_t* __restrict__ cell1,
                       const uint32_t* __restrict__ idx0, const uint32_t* __restrict__ idx1,
                       const SortPlan* __restrict__ plan, int npass, uint32_t n,
                       uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end,
                       uint32_t* __restrict__ order, uint32_t* __restrict__ body_cell,
                       const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                       float4* __restrict__ cpos) {
  const uint32_t* __restrict__ cell = plan->src[npass] ? cell1 : cell0;
  const uint32_t* __restrict__ idx = plan->src[npass] ? idx1 : idx0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t c = cell[k];
    const uint32_t b = idx[k];
    order[k] = b;
    body_cell[b] = c;
    // neighbour data laid out in cell order: {x, y, species, body index}, so a cell's bodies are one
    // contiguous run of 16-byte records
    const float4 p = pqr[b];
    cpos[k] = make_float4(p.x, p.y, __uint_as_float((uint32_t)species[b]), __uint_as_float(b));
    if (k == 0 || cell[k - 1] != c) cell_start[c] = k;
    if (k + 1 == n || cell[k + 1] != c) cell_end[c] = k + 1;
  }
}

// ------------------------------------------------------------------------------------------------
struct ShortRangeParams {
  GridDims g;
  int do_lj, do_rep, do_stack;
  float max_lj_cutoff;   // species.rs:412-445
  float max_lj_force;    // COLLISION_PASSES as f32 * LJ_FORCE_MAX (forces.rs:221)
  float stack_pressure, stack_decay;
  int range;             // cells to scan on each side: ceil(max cutoff / cell_size)
  float max_rep_cutoff;  // species.rs:447-479 (0 when the repulsion pass is off)
};

// Vec2::normalized() of ultraviolet 0.9.2: multiply by 1/mag
__device__ __forceinline__ float2 normalized_rn(float x, float y, float mag) {
  const float inv = __fdiv_rn(1.0f, mag);
  return make_float2(__fmul_rn(x, inv), __fmul_rn(y, inv));
}

// LJ pair term seen from body `me` (forces.rs:203-229).  (a, b) = (lower, higher) body index.
__device__ __forceinline__ void lj_pair(const SpeciesRow& sa, const SpeciesRow& sb, float ax, float ay,
                                        float bx, float by, bool me_is_a, float my_mass,
                                        float max_lj_force, float& accx, float& accy) {
  const float sigma = __fmul_rn(__fadd_rn(sa.lj_sigma, sb.lj_sigma), 0.5f);
  const float epsilon = __fsqrt_rn(__fmul_rn(sa.lj_epsilon, sb.lj_epsilon));
  const float cutoff = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(sa.lj_cutoff, sa.lj_sigma),
                                                 __fmul_rn(sb.lj_cutoff, sb.lj_sigma)));
  const float rx = __fsub_rn(bx, ax), ry = __fsub_rn(by, ay);
  const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)));
  if (r < cutoff && r > 1e-6f) {
    const float x = __fdiv_rn(sigma, r);
    const float x2 = __fmul_rn(x, x), x4 = __fmul_rn(x2, x2);
    const float sr6 = __fmul_rn(x2, x4);  // powi(6) = x^2 * x^4
    const float t = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, sr6), sr6), sr6);
    const float unclamped = __fdiv_rn(__fmul_rn(__fmul_rn(24.0f, epsilon), t), r);
    float fm = unclamped;
    if (fm < -max_lj_force) fm = -max_lj_force;
    if (fm > max_lj_force) fm = max_lj_force;
    const float2 nrm = normalized_rn(rx, ry, r);
    const float fx = __fmul_rn(fm, nrm.x), fy = __fmul_rn(fm, nrm.y);
    const float dx = __fdiv_rn(fx, my_mass), dy = __fdiv_rn(fy, my_mass);
    if (me_is_a) {
      accx = __fsub_rn(accx, dx), accy = __fsub_rn(accy, dy);
    } else {
      accx = __fadd_rn(accx, dx), accy = __fadd_rn(accy, dy);
    }
  }
}

// repulsion pair term (forces.rs:234-247,259-287); `a` is the lower index and owns the query cutoff
__device__ __forceinline__ void rep_pair(const SpeciesRow& sa, const SpeciesRow& sb, float ax, float ay,
                                         float bx, float by, bool me_is_a, float my_mass,
                                         float& accx, float& accy) {
  const float rx = __fsub_rn(bx, ax), ry = __fsub_rn(by, ay);
  const float r2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
  const float qcut = sa.repulsion_cutoff;
  if (!(r2 < __fmul_rn(qcut, qcut))) return;  // neighbour query of body a
  const float r = __fsqrt_rn(r2);
  const float r0 = __fmul_rn(0.5f, __fadd_rn(sa.repulsion_cutoff, sb.repulsion_cutoff));
  if (r >= r0 || r <= 0.0f) return;
  const float k = __fmul_rn(0.5f, __fadd_rn(sa.repulsion_strength, sb.repulsion_strength));
  const float mag = __fdiv_rn(__fmul_rn(k, __fsub_rn(1.0f, __fdiv_rn(r, r0))), r);
  const float fx = __fmul_rn(rx, mag), fy = __fmul_rn(ry, mag);
  if (fx == 0.0f && fy == 0.0f) return;
  const float dx = __fdiv_rn(fx, my_mass), dy = __fdiv_rn(fy, my_mass);
  if (me_is_a) {
    accx = __fsub_rn(accx, dx), accy = __fsub_rn(accy, dy);
  } else {
    accx = __fadd_rn(accx, dx), accy = __fadd_rn(accy, dy);
  }
}

// per-cell body counts, summed into the monotone offsets array `cell_off` (ncells + 1 entries) that
// makes every row segment of the grid one contiguous run of the cell-ordered records
struct CellCountFn {
  const uint32_t* cell_start;
  const uint32_t* cell_end;
  __device__ __forceinline__ uint32_t operator()(uint32_t c) const { return cell_end[c] - cell_start[c]; }
};

// Gather form of the reference's serial pair loops (forces.rs:182-289): each body sums the terms of
// every pair it is in, in the reference's neighbour order (rows dy, cells dx, index order inside a cell).
//
// Shared-memory tile staging: a CTA owns 128 consecutive bodies, a compact patch in Morton order.  It
// takes the bounding box of their cells, grows it by the interaction range, and stages every record of
// that box of cells — {x, y, species, index}, one contiguous global run per grid row — plus the box's
// cell offsets into shared memory; the pair loops then read neighbours from shared memory only, one
// contiguous run per row.  A patch whose box is larger than the tile limits (a strip of far-apart
// bodies, e.g. at a curve discontinuity) falls back to reading the same runs from global memory.
constexpr int kTileMaxW = 32, kTileMaxH = 32, kTileCap = 768;

__global__ void __launch_bounds__(128)
    short_range_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                       const SpeciesRow* __restrict__ table_g, uint32_t first, uint32_t n,
                       const uint32_t* __restrict__ cell_off, const float4* __restrict__ cpos,
                       const uint32_t* __restrict__ body_cell, ShortRangeParams P,
                       float4* __restrict__ acc_mass) {
  __shared__ SpeciesRow table[kMaxSpecies];
  __shared__ float4 s_rec[kTileCap];
  __shared__ uint16_t s_off[(kTileMaxW + 1) * kTileMaxH];  // offsets into s_rec (kTileCap < 65536)
  __shared__ uint32_t s_rowsrc[kTileMaxH], s_rowbase[kTileMaxH + 1];
  __shared__ int s_box[4];  // min x, max x, min y, max y of the cells of the bodies that have pair work
  __shared__ int s_fits;
  for (int k = threadIdx.x; k < kMaxSpecies * (int)(sizeof(SpeciesRow) / 4); k += blockDim.x)
    reinterpret_cast<uint32_t*>(table)[k] = reinterpret_cast<const uint32_t*>(table_g)[k];
  if (threadIdx.x == 0) s_box[0] = 0x7fffffff, s_box[1] = -1, s_box[2] = 0x7fffffff, s_box[3] = -1, s_fits = 0;
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  float4 me = make_float4(0, 0, 0, 0);
  uint32_t my_sp = 0;
  if (live) {
    me = pqr[i];
    my_sp = species[i] < kMaxSpecies ? species[i] : 0;
  }
  __syncthreads();
  const SpeciesRow si = table[my_sp];
  const bool lj_i = live && P.do_lj && si.lj_enabled;
  const bool rep_i = live && P.do_rep && si.repulsion_enabled;
  const bool need = lj_i || rep_i;
  int cx = 0, cy = 0;
  if (need) {
    const uint32_t c = body_cell[i];
    cx = (int)(c % P.g.gx), cy = (int)(c / P.g.gx);
    atomicMin(&s_box[0], cx), atomicMax(&s_box[1], cx), atomicMin(&s_box[2], cy), atomicMax(&s_box[3], cy);
  }
  const int any_need = __syncthreads_or(need ? 1 : 0);
  float4 am = make_float4(0, 0, 0, 0);
  if (live) am = acc_mass[i];
  if (any_need) {
    const int R = P.range;
    const int x0 = max(s_box[0] - R, 0), x1 = min(s_box[1] + R, (int)P.g.gx - 1);
    const int y0 = max(s_box[2] - R, 0), y1 = min(s_box[3] + R, (int)P.g.gy - 1);
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const bool box_ok = bw <= kTileMaxW && bh <= kTileMaxH;
    if (box_ok) {
      for (int r = threadIdx.x; r < bh; r += blockDim.x)
        s_rowsrc[r] = cell_off[(uint32_t)x0 + (uint32_t)(y0 + r) * P.g.gx];
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int r = 0; r < bh; ++r) {
          s_rowbase[r] = run;
          run += cell_off[(uint32_t)x1 + (uint32_t)(y0 + r) * P.g.gx + 1] - s_rowsrc[r];
        }
        s_rowbase[bh] = run;
        s_fits = run <= (uint32_t)kTileCap;
      }
      __syncthreads();
    }
    const bool staged = box_ok && s_fits;
    if (staged) {
      // the box's cell offsets (re-based to the staged records) and the records themselves
      for (int idx = threadIdx.x; idx < (bw + 1) * bh; idx += blockDim.x) {
        const int r = idx / (bw + 1), xx = idx - r * (bw + 1);
        s_off[idx] = (uint16_t)(cell_off[(uint32_t)(x0 + xx) + (uint32_t)(y0 + r) * P.g.gx] - s_rowsrc[r] + s_rowbase[r]);
      }
      for (int r = 0; r < bh; ++r) {
        const uint32_t len = s_rowbase[r + 1] - s_rowbase[r], src = s_rowsrc[r], dst = s_rowbase[r];
        for (uint32_t t = threadIdx.x; t < len; t += blockDim.x) s_rec[dst + t] = cpos[src + t];
      }
      __syncthreads();
    }
    if (need) {
      const float lj_cut_sq = __fmul_rn(P.max_lj_cutoff, P.max_lj_cutoff);
      float ljx = 0.0f, ljy = 0.0f, rpx = 0.0f, rpy = 0.0f;
      const int xa = max(cx - R, 0), xb = min(cx + R, (int)P.g.gx - 1);
      const int yb = min(cy + R, (int)P.g.gy - 1);
      // no pair term exists at or beyond this squared distance (LJ: forces.rs:203; repulsion: the
      // neighbour query of the lower-index body, forces.rs:262)
      const float reach_sq = fmaxf(P.do_lj ? lj_cut_sq : 0.0f, __fmul_rn(P.max_rep_cutoff, P.max_rep_cutoff));
      int y = max(cy - R, 0) - 1;
      uint32_t k = 0, ke = 0;
      for (;;) {
        // cheap scan, lanes diverge here: advance to this body's next candidate inside the reach ...
        float4 cj;
        bool found = false;
        for (;;) {
          if (k >= ke) {
            if (++y > yb) break;
            if (staged) {
              const int r = y - y0;
              k = s_off[r * (bw + 1) + (xa - x0)], ke = s_off[r * (bw + 1) + (xb - x0) + 1];
            } else {
              k = cell_off[(uint32_t)xa + (uint32_t)y * P.g.gx], ke = cell_off[(uint32_t)xb + (uint32_t)y * P.g.gx + 1];
            }
            continue;
          }
          cj = staged ? s_rec[k] : __ldg(&cpos[k]);
          ++k;
          if (__float_as_uint(cj.w) == i) continue;
          const float ddx = __fsub_rn(cj.x, me.x), ddy = __fsub_rn(cj.y, me.y);
          if (__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)) < reach_sq) {
            found = true;
            break;
          }
        }
        if (!found) break;
        // ... and the warp reconverges for the pair arithmetic, same neighbour order per body
        const uint32_t j = __float_as_uint(cj.w);
        const float jx = cj.x, jy = cj.y;
        uint32_t jsp = __float_as_uint(cj.z);
        if (jsp >= kMaxSpecies) jsp = 0;
        const SpeciesRow& sj = table[jsp];
        const bool me_is_a = i < j;
        const float ax = me_is_a ? me.x : jx, ay = me_is_a ? me.y : jy;
        const float bx = me_is_a ? jx : me.x, by = me_is_a ? jy : me.y;
        const SpeciesRow& sa = me_is_a ? si : sj;
        const SpeciesRow& sb = me_is_a ? sj : si;
        if (lj_i && sj.lj_enabled) {
          const float rx = __fsub_rn(bx, ax), ry = __fsub_rn(by, ay);
          if (__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)) < lj_cut_sq)
            lj_pair(sa, sb, ax, ay, bx, by, me_is_a, am.w, P.max_lj_force, ljx, ljy);
        }
        if (rep_i && sj.repulsion_enabled) rep_pair(sa, sb, ax, ay, bx, by, me_is_a, am.w, rpx, rpy);
      }
      // the reference finishes the LJ pass before the repulsion pass (simulation.rs:1008-1009)
      am.x = __fadd_rn(__fadd_rn(am.x, ljx), rpx);
      am.y = __fadd_rn(__fadd_rn(am.y, ljy), rpy);
    }
  }
  if (!live) return;
  if (P.do_stack) {  // forces.rs:294-321
    const float x_min = -P.g.hw, x_max = P.g.hw;
    const float dist_left = __fsub_rn(me.x, x_min);
    if (dist_left < P.stack_decay && dist_left > 0.0f) {
      const float force = __fmul_rn(P.stack_pressure, __fsub_rn(1.0f, __fdiv_rn(dist_left, P.stack_decay)));
      am.x = __fadd_rn(am.x, __fdiv_rn(force, am.w));
    }
    const float dist_right = __fsub_rn(x_max, me.x);
    if (dist_right < P.stack_decay && dist_right > 0.0f) {
      const float force = __fmul_rn(P.stack_pressure, __fsub_rn(1.0f, __fdiv_rn(dist_right, P.stack_decay)));
      am.x = __fsub_rn(am.x, __fdiv_rn(force, am.w));
    }
  }
  if (need || P.do_stack) acc_mass[i] = am;
}

// ------------------------------------------------------------------------------------------------
// CellList::find_neighbors_within / metal_neighbor_count for a batch of bodies.
// pass 0 counts (out_idx == nullptr), pass 1 fills at offsets[q]; order = the reference's
// (dy, dx, per-cell index order).
__global__ void __launch_bounds__(128)
    cell_neighbors_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                          uint32_t n, const uint32_t* __restrict__ cell_start,
                          const uint32_t* __restrict__ cell_end, const uint32_t* __restrict__ order,
                          GridDims g, const uint32_t* __restrict__ query, uint32_t m, float cutoff,
                          int metals_only, uint32_t* __restrict__ counts,
                          const uint32_t* __restrict__ offsets, uint32_t* __restrict__ out_idx) {
  const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= m) return;
  const uint32_t i = query[qi];
  if (i >= n) {
    if (!out_idx) counts[qi] = 0;
    return;
  }
  const float4 me = pqr[i];
  const int cx = (int)cell_axis(me.x, -g.hw, g.cell_size, g.gx);
  const int cy = (int)cell_axis(me.y, -g.hh, g.cell_size, g.gy);
  const float rf = ceilf(__fdiv_rn(cutoff, g.cell_size));
  const int range = (rf == rf) ? (rf > 1.0e6f ? 1000000 : (rf < -1.0e6f ? -1000000 : (int)rf)) : 0;
  const float cutoff_sq = __fmul_rn(cutoff, cutoff);
  uint32_t cnt = 0;
  uint32_t* dst = out_idx ? out_idx + offsets[qi] : nullptr;
  // same visiting order as the reference's dy/dx loops, with the out-of-grid cells clipped up front
  const int y0 = max(cy - range, 0), y1 = min(cy + range, (int)g.gy - 1);
  const int x0 = max(cx - range, 0), x1 = min(cx + range, (int)g.gx - 1);
  for (int y = y0; y <= y1; ++y) {
    for (int x = x0; x <= x1; ++x) {
      const uint32_t cc = (uint32_t)x + (uint32_t)y * g.gx;
      const uint32_t e = cell_end[cc];
      for (uint32_t k = cell_start[cc]; k < e; ++k) {
        const uint32_t j = order[k];
        if (j == i) continue;
        const float4 pj = pqr[j];
        const float rx = __fsub_rn(pj.x, me.x), ry = __fsub_rn(pj.y, me.y);
        if (__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)) < cutoff_sq) {
          if (metals_only && !(species[j] == 1 || species[j] == 2)) continue;
          if (dst) dst[cnt] = j;
          ++cnt;
        }
      }
    }
  }
  if (!out_idx) counts[qi] = cnt;
}

// ------------------------------------------------------------------------------------------------
// Simulation::iterate (simulation.rs:1437-1486)
struct IterateParams {
  float dt, base_damping, hw, hh, hd;
  int enable_z;
};

__global__ void __launch_bounds__(256)
    iterate_kernel(float4* __restrict__ pqr, float4* __restrict__ vel_z /* {vx, vy, z, vz} */,
                   const float4* __restrict__ acc_mass, const uint8_t* __restrict__ species,
                   const SpeciesRow* __restrict__ table, uint32_t n, IterateParams P) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 p = pqr[i];
    float4 v = vel_z[i];
    const float4 a = acc_mass[i];
    const uint8_t sp = species[i] < kMaxSpecies ? species[i] : 0;
    v.x = __fadd_rn(v.x, __fmul_rn(a.x, P.dt));
    v.y = __fadd_rn(v.y, __fmul_rn(a.y, P.dt));
    const float damping = __fmul_rn(P.base_damping, table[sp].damping);
    v.x = __fmul_rn(v.x, damping);
    v.y = __fmul_rn(v.y, damping);
    p.x = __fadd_rn(p.x, __fmul_rn(v.x, P.dt));
    p.y = __fadd_rn(p.y, __fmul_rn(v.y, P.dt));
    if (P.enable_z) {
      v.w = __fadd_rn(v.w, __fmul_rn(a.z, P.dt));
      v.w = __fmul_rn(v.w, damping);
      v.z = __fadd_rn(v.z, __fmul_rn(v.w, P.dt));
      if (v.z < -P.hd) {
        v.z = -P.hd;
        v.w = -v.w;
      } else if (v.z > P.hd) {
        v.z = P.hd;
        v.w = -v.w;
      }
    }
    if (p.x < -P.hw) {
      p.x = -P.hw;
      v.x = -v.x;
    } else if (p.x > P.hw) {
      p.x = P.hw;
      v.x = -v.x;
    }
    if (p.y < -P.hh) {
      p.y = -P.hh;
      v.y = -v.y;
    } else if (p.y > P.hh) {
      p.y = P.hh;
      v.y = -v.y;
    }
    pqr[i] = p;
    vel_z[i] = v;
  }
}

}  // namespace psim

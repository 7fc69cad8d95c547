// neighbors.cuh — the index-based neighbour consumers that follow the force phase:
//   Simulation::update_surrounded_flags  src/simulation/simulation.rs:1893-1918
//     Body::maybe_update_surrounded      src/body/types.rs:243-286
//     CellList::metal_neighbor_count     src/cell_list.rs:92-127
//   enforce_metal_z_boundaries           src/simulation/out_of_plane.rs:140-254
// Both are pure counting / clamping over the cell grid.  One thread per body walks the reference's
// (dy, dx, per-cell index) order, which matters for the "first five metal neighbours" rule of the z clamp.
// The surround state (last position / frame / flag) is kept by ORIGINAL body id, so tree builds do not
// have to permute it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cells.cuh"

namespace psim {

struct SurroundState {
  float2* last_pos;               // by original id
  unsigned long long* last_frame; // by original id
  uint8_t* flag;                  // by original id
};

__global__ void __launch_bounds__(256)
    surround_init_kernel(const float4* __restrict__ pqr, const uint32_t* __restrict__ orig, uint32_t n,
                         SurroundState st) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t id = orig[i];
    st.last_pos[id] = make_float2(pqr[i].x, pqr[i].y);  // Body::new, body/types.rs:111-113
    st.last_frame[id] = 0ull;
    st.flag[id] = 0;
  }
}

__device__ __forceinline__ bool is_metal(uint8_t s) { return s == 1 || s == 2; }  // LithiumMetal | FoilMetal

struct SurroundParams {
  unsigned long long frame, interval, neighbor_threshold;
  float radius_factor, move_threshold;
};

__global__ void __launch_bounds__(128)
    surrounded_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                      const uint32_t* __restrict__ orig, uint32_t n, const uint32_t* __restrict__ cell_start,
                      const uint32_t* __restrict__ cell_end, const uint32_t* __restrict__ order, GridDims g,
                      SurroundParams P, SurroundState st) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 me = pqr[i];
  const uint32_t id = orig[i];
  const float2 lp = st.last_pos[id];
  const unsigned long long lf = st.last_frame[id];
  const float mx = __fsub_rn(me.x, lp.x), my = __fsub_rn(me.y, lp.y);
  const bool moved = __fsqrt_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my))) > __fmul_rn(P.move_threshold, me.w);
  const unsigned long long frame_diff = P.frame >= lf ? P.frame - lf : P.interval;
  if (!(moved || frame_diff >= P.interval)) return;
  const float cutoff = __fmul_rn(me.w, P.radius_factor);
  const int cx = (int)cell_axis(me.x, -g.hw, g.cell_size, g.gx);
  const int cy = (int)cell_axis(me.y, -g.hh, g.cell_size, g.gy);
  const float rf = ceilf(__fdiv_rn(cutoff, g.cell_size));
  const int range = (rf == rf) ? (rf > 1.0e6f ? 1000000 : (rf < -1.0e6f ? -1000000 : (int)rf)) : 0;
  const float cutoff_sq = __fmul_rn(cutoff, cutoff);
  unsigned long long count = 0;
  const int y0 = max(cy - range, 0), y1 = min(cy + range, (int)g.gy - 1);
  const int x0 = max(cx - range, 0), x1 = min(cx + range, (int)g.gx - 1);
  for (int y = y0; y <= y1; ++y)
    for (int x = x0; x <= x1; ++x) {
      const uint32_t cc = (uint32_t)x + (uint32_t)y * g.gx;
      const uint32_t e = cell_end[cc];
      for (uint32_t k = cell_start[cc]; k < e; ++k) {
        const uint32_t j = order[k];
        if (j == i) continue;
        const float4 pj = pqr[j];
        const float rx = __fsub_rn(pj.x, me.x), ry = __fsub_rn(pj.y, me.y);
        if (__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)) < cutoff_sq && is_metal(species[j])) ++count;
      }
    }
  st.flag[id] = count >= P.neighbor_threshold ? 1 : 0;
  st.last_pos[id] = make_float2(me.x, me.y);
  st.last_frame[id] = P.frame;
}

__global__ void __launch_bounds__(256)
    surround_export_kernel(const uint32_t* __restrict__ orig, uint32_t n, SurroundState st,
                           uint8_t* __restrict__ flags, float2* __restrict__ last_pos,
                           unsigned long long* __restrict__ last_frame) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t id = orig[i];
    if (flags) flags[i] = st.flag[id];
    if (last_pos) last_pos[i] = st.last_pos[id];
    if (last_frame) last_frame[i] = st.last_frame[id];
  }
}

// out_of_plane.rs:164-251, one thread per non-metal body
__global__ void __launch_bounds__(128)
    metal_z_kernel(const float4* __restrict__ pqr, float4* __restrict__ vel_z /* {vx, vy, z, vz} */,
                   const uint8_t* __restrict__ species, uint32_t n, const uint32_t* __restrict__ cell_start,
                   const uint32_t* __restrict__ cell_end, const uint32_t* __restrict__ order, GridDims g,
                   float metal_max_r, float max_z) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (is_metal(species[i])) return;
  const float4 me = pqr[i];
  const float body_radius = me.w;
  const float cutoff = __fadd_rn(__fmul_rn(3.0f, body_radius), metal_max_r);
  const int cx = (int)cell_axis(me.x, -g.hw, g.cell_size, g.gx);
  const int cy = (int)cell_axis(me.y, -g.hh, g.cell_size, g.gy);
  const float rf = ceilf(__fdiv_rn(cutoff, g.cell_size));
  const int range = (rf == rf) ? (rf > 1.0e6f ? 1000000 : (rf < -1.0e6f ? -1000000 : (int)rf)) : 0;
  const float cutoff_sq = __fmul_rn(cutoff, cutoff);
  float min_c = -max_z, max_c = max_z;
  int applied = 0;
  bool stop = false;
  const int y0 = max(cy - range, 0), y1 = min(cy + range, (int)g.gy - 1);
  const int x0 = max(cx - range, 0), x1 = min(cx + range, (int)g.gx - 1);
  for (int y = y0; y <= y1 && !stop; ++y)
    for (int x = x0; x <= x1 && !stop; ++x) {
      const uint32_t cc = (uint32_t)x + (uint32_t)y * g.gx;
      const uint32_t e = cell_end[cc];
      for (uint32_t k = cell_start[cc]; k < e; ++k) {
        const uint32_t j = order[k];
        if (j == i) continue;
        const float4 pj = pqr[j];
        // find_neighbors_within: (other - me).mag_sq() < cutoff^2 (cell_list.rs:76)
        const float rx = __fsub_rn(pj.x, me.x), ry = __fsub_rn(pj.y, me.y);
        if (!(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)) < cutoff_sq)) continue;
        if (!is_metal(species[j])) continue;
        if (++applied > 5) {  // only the first five metal neighbours constrain (:199-203)
          stop = true;
          break;
        }
        const float metal_radius = pj.w;
        const float dx = __fsub_rn(me.x, pj.x), dy = __fsub_rn(me.y, pj.y);
        const float distance_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float reach = __fadd_rn(__fadd_rn(body_radius, metal_radius), __fmul_rn(2.0f, body_radius));
        if (distance_sq > __fmul_rn(reach, reach)) continue;
        if (__fsqrt_rn(distance_sq) < reach) {
          const float metal_z = vel_z[j].z;
          const float lower = __fsub_rn(__fsub_rn(metal_z, metal_radius), 0.01f);
          const float upper = __fadd_rn(__fadd_rn(metal_z, metal_radius), 0.01f);
          if (lower < upper) {
            min_c = fmaxf(min_c, lower);
            max_c = fminf(max_c, upper);
          }
        }
      }
    }
  if (min_c > max_c) min_c = -0.1f, max_c = 0.1f;
  float4 v = vel_z[i];
  if (v.z < min_c) {
    v.z = min_c;
    if (v.w < 0.0f) v.w = 0.0f;
  }
  if (v.z > max_c) {
    v.z = max_c;
    if (v.w > 0.0f) v.w = 0.0f;
  }
  if (v.z > max_z) v.z = max_z, v.w = 0.0f;  // Body::clamp_z, body/types.rs:296-304
  else if (v.z < -max_z) v.z = -max_z, v.w = 0.0f;
  vel_z[i] = v;
}

}  // namespace psim

// hopping.cuh — the field part of the electron-hopping candidate predicate, batched.
//
// Replaces, for a whole list of (donor, acceptor) candidates at once, what the serial hopping loop of the
// reference does per candidate (src/simulation/electron_hopping.rs:283-329):
//   local_field = background_e_field + quadtree.field_at_point(&bodies, src.pos, k_e)     (:290-295)
//   field_dir   = local_field.normalized() if |local_field| > 1e-6 else 0                  (:296-300)
//   hop_dir     = (dst.pos - src.pos).normalized() if longer than 1e-6 else 0              (:284-289)
//   alignment   = max(0, -hop_dir . field_dir), 1 when the field vanishes, times max(bias, 0)   (:301-306)
//   alignment   = max(alignment, 0.5) when both ends are metals / electrode materials and one of them is an
//                 electrode material                                                        (:310-328)
// The Barnes-Hut walk itself is the batched point kernel of traverse.cuh (one walk per DISTINCT donor; the
// reference repeats it for every candidate of the donor).  The candidate lists, their shuffling and the rate /
// d_phi tests stay on the host: they need per-body state this library does not hold.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "psim_core.cuh"
#include "tree_logic.cuh"

namespace psim {

// body/types.rs:12-36 (enum order)
__device__ __forceinline__ bool hop_is_electrode_material(uint32_t s) { return s >= 13u && s <= 20u; }
__device__ __forceinline__ bool hop_is_metal_or_electrode(uint32_t s) { return s == 1u || s == 2u || hop_is_electrode_material(s); }

__global__ void __launch_bounds__(256)
    hop_points_kernel(const float4* __restrict__ pqr, const uint32_t* __restrict__ src, uint32_t m, uint32_t n,
                      float2* __restrict__ pts) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const uint32_t s = src[i] < n ? src[i] : 0u;
    const float4 p = pqr[s];
    pts[i] = make_float2(p.x, p.y);
  }
}

// Vec2::normalized() of ultraviolet 0.9.2 (multiply by 1 / mag) behind the reference's `mag > 1e-6` guard
__device__ __forceinline__ float2 hop_dir_of(float x, float y) {
  const float mag = __fsqrt_rn(f_add(f_mul(x, x), f_mul(y, y)));
  if (!(mag > 1e-6f)) return make_float2(0.0f, 0.0f);
  const float r = __frcp_rn(mag);
  return make_float2(f_mul(x, r), f_mul(y, r));
}

// one donor per thread, its candidates one after the other
__global__ void __launch_bounds__(128)
    hop_alignment_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                         const uint32_t* __restrict__ src, const uint32_t* __restrict__ pair_off,
                         const uint32_t* __restrict__ dst, uint32_t m, uint32_t n, const float2* __restrict__ field,
                         float bg_x, float bg_y, float bias, float2* __restrict__ local_field,
                         float* __restrict__ alignment) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const float b = fmaxf(bias, 0.0f);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const uint32_t s = src[i] < n ? src[i] : 0u;
    const float4 ps = pqr[s];
    const uint32_t ss = species[s];
    const float2 f = field[i];
    const float lx = f_add(bg_x, f.x), ly = f_add(bg_y, f.y);
    local_field[i] = make_float2(lx, ly);
    const float2 fd = hop_dir_of(lx, ly);
    const bool no_field = fd.x == 0.0f && fd.y == 0.0f;
    for (uint32_t k = pair_off[i]; k < pair_off[i + 1]; ++k) {
      const uint32_t d = dst[k] < n ? dst[k] : 0u;
      const float4 pd = pqr[d];
      const float2 hd = hop_dir_of(f_add(pd.x, -ps.x), f_add(pd.y, -ps.y));
      float a = fmaxf(-f_add(f_mul(hd.x, fd.x), f_mul(hd.y, fd.y)), 0.0f);
      if (no_field) a = 1.0f;
      a = f_mul(a, b);
      const uint32_t ds = species[d];
      if (hop_is_metal_or_electrode(ss) && hop_is_metal_or_electrode(ds) &&
          (hop_is_electrode_material(ss) || hop_is_electrode_material(ds)))
        a = fmaxf(a, 0.5f);
      alignment[k] = a;
    }
  }
}

}  // namespace psim

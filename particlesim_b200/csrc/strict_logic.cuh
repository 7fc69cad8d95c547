// strict_logic.cuh — the reference's serial f32 running sums (Quadtree::propagate,
// src/quadtree/quadtree.rs:114-139), evaluated in parallel and still bit for bit.
//
// `propagate` computes every internal node's centre as  fold(acc + pos * |q|) / sum(|q|)  over the node's
// body range with ONE f32 accumulator per component.  A rounded sum is not associative, so a reduction
// tree cannot reproduce it.  What can be done exactly:
//
//  * While the accumulator s stays inside one binade [2^e, 2^(e+1)) it is an integer mantissa M times
//    u = 2^(e-23), and  fl(s + a) = (M + rne_M(a / u)) * u  where the rounding of the real number a / u to
//    an integer only depends on M through its PARITY (ties go to the even result).  So, over a block of
//    addends, "what the block does to s" is one of two integers: o[p] = total mantissa offset when the
//    block is entered with parity p.  These pairs compose associatively (BlockFn / compose), hence any
//    number of blocks can be evaluated independently and chained by a scan - as long as the accumulator
//    really stays in the binade the block assumed, which the running minimum / maximum of the offset
//    (lo[p], hi[p]) decides exactly once the entering mantissa is known.
//  * Blocks for which the assumption fails (the sum crosses a power of two, cancels to zero, meets an
//    addend that is not small against it) are re-evaluated with plain serial f32 additions.  A monotone
//    sum crosses ~24 binades however long it is, so the serial part is bounded.
//
// Everything here is host+device so that tests/emu can check it on the CPU against a plain loop.
#pragma once
#include <stdint.h>

#include "psim_core.cuh"

namespace psim {

constexpr int kBadExp = 0x7fff;          // BlockFn::e of a block that must be evaluated serially
constexpr int32_t kMant0 = 1 << 23;      // mantissa range of a normal f32: [2^23, 2^24)
constexpr int32_t kMant1 = 1 << 24;
constexpr int32_t kMaxStep = 1 << 21;    // |a / u| above this: the addend is not small against the sum
constexpr uint32_t kStrictBlock = 512;   // addends per block (512 * 2^21 < 2^31: offsets fit an int32)

struct BlockFn {
  int32_t o[2];   // mantissa offset after the block, entering parity 0 / 1
  int32_t lo[2];  // minimum and maximum of the running offset inside the block
  int32_t hi[2];
  int32_t e;      // binade the block was evaluated for, or kBadExp
};

PSIM_HD uint32_t f32_bits(float v) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(v);
#else
  union { float f; uint32_t u; } c;
  c.f = v;
  return c.u;
#endif
}
PSIM_HD float f32_from_bits(uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c;
  c.u = b;
  return c.f;
#endif
}

// s = M * 2^(e - 23) with |M| in [2^23, 2^24); false for zero, subnormal, non-finite values and for
// binades whose 1 / u is not a normal f32 (|s| < 2^-104): those go the serial way
PSIM_HD bool f32_split(float s, int& e, int32_t& M) {
  const uint32_t b = f32_bits(s);
  const int ef = (int)((b >> 23) & 0xffu);
  if (ef == 0 || ef == 255 || ef < 127 - 104) return false;
  e = ef - 127;
  const int32_t m = (int32_t)((b & 0x7fffffu) | 0x800000u);
  M = (b >> 31) ? -m : m;
  return true;
}
PSIM_HD float f32_join(int e, int32_t M) {
  const uint32_t m = (uint32_t)(M < 0 ? -M : M);
  return f32_from_bits((M < 0 ? 0x80000000u : 0u) | ((uint32_t)(e + 127) << 23) | (m & 0x7fffffu));
}
// 2^(23 - e): multiplying an addend by it gives a / u exactly (a power of two; an underflowing product is
// far below 1/2 and rounds to "adds nothing" either way)
PSIM_HD float inv_ulp(int e) { return f32_from_bits((uint32_t)(150 - e) << 23); }
// binade of a (double) estimate of the accumulator, kBadExp when f32_split would refuse it
PSIM_HD int spec_exponent(double est) {
  int e;
  int32_t M;
  const float f = (float)est;
  return f32_split(f, e, M) ? e : kBadExp;
}

PSIM_HD void blockfn_init(BlockFn& f, int e) {
  f.o[0] = f.o[1] = 0;
  f.lo[0] = f.lo[1] = 0;
  f.hi[0] = f.hi[1] = 0;
  f.e = e;
}

PSIM_HD float f_floor(float t) {
#if defined(__CUDA_ARCH__)
  return floorf(t);
#else
  return __builtin_floorf(t);
#endif
}

// one more addend; inv_u = inv_ulp(f.e).  t = a / u is exact (scaling by a power of two); the integer and
// fractional parts are taken of |t|, where both are exact (1 + t for a negative t would round).
PSIM_HD void blockfn_step(BlockFn& f, float a, float inv_u) {
  const float t = f_mul(a, inv_u);
  const float at = t < 0.0f ? -t : t;
  if (!(at < (float)kMaxStep)) {  // also catches NaN / inf
    f.e = kBadExp;
    return;
  }
  const float fl = f_floor(at);
  const float frac = f_add(at, -fl);
  const int32_t sgn = t < 0.0f ? -1 : 1;
  const int32_t k = sgn * (int32_t)fl;          // t truncated towards zero
  const int32_t up = frac > 0.5f ? sgn : 0;     // nearest integer is the one away from zero
  const bool tie = frac == 0.5f;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int32_t before = f.o[p] + k;
    // tie: of the two candidates the one that makes the result mantissa M + o even (M has parity p)
    const int32_t o = before + (tie ? sgn * ((p + before) & 1) : up);
    f.o[p] = o;
    f.lo[p] = o < f.lo[p] ? o : f.lo[p];
    f.hi[p] = o > f.hi[p] ? o : f.hi[p];
  }
}

// g after f (only the offsets: validity is checked per block with its own entering mantissa)
PSIM_HD void blockfn_compose(int32_t& o0, int32_t& o1, int32_t f0, int32_t f1, int32_t g0, int32_t g1) {
  o0 = f0 + ((f0 & 1) ? g1 : g0);
  o1 = f1 + (((1 + f1) & 1) ? g1 : g0);
}

// may the block be applied to an accumulator with exponent e and mantissa M?  Every intermediate sum must
// stay strictly inside the binade (a result of exactly +-2^23 or +-2^24 may have been rounded on the
// neighbouring binade's grid).
PSIM_HD bool blockfn_valid(const BlockFn& f, int e, int32_t M) {
  if (f.e != e) return false;
  const int p = M & 1;
  const int32_t a = M + f.lo[p], b = M + f.hi[p];
  return M > 0 ? (a > kMant0 && b < kMant1) : (b < -kMant0 && a > -kMant1);
}
PSIM_HD int32_t blockfn_apply(const BlockFn& f, int32_t M) { return M + f.o[M & 1]; }

}  // namespace psim

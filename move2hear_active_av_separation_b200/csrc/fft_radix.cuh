// In-register radix-R DFTs (R = 8, 16, 32) used by the shared-memory FFT passes.
//
// fft_reg<R, INV>(v): v[j], j < R natural order in; on return v[bitrev(k)] holds
//     sum_j v[j] * exp(-/+ 2 pi i j k / R)      (- forward, + for INV).
// All twiddles are compile-time constants once the loops are unrolled.
#pragma once
#include "m2h_common.cuh"

M2H_HD constexpr float tw32_cos(int k) {   // cos(2 pi k / 32), 0 <= k < 16
  return k == 0 ? 1.0f : k == 1 ? 0.9807852804032304f : k == 2 ? 0.9238795325112867f : k == 3 ? 0.8314696123025452f :
         k == 4 ? 0.7071067811865476f : k == 5 ? 0.5555702330196023f : k == 6 ? 0.38268343236508984f :
         k == 7 ? 0.19509032201612833f : k == 8 ? 0.0f : k == 9 ? -0.19509032201612833f : k == 10 ? -0.38268343236508984f :
         k == 11 ? -0.5555702330196023f : k == 12 ? -0.7071067811865476f : k == 13 ? -0.8314696123025452f :
         k == 14 ? -0.9238795325112867f : -0.9807852804032304f;
}
M2H_HD constexpr float tw32_sin(int k) {   // sin(2 pi k / 32), 0 <= k < 16
  return k == 0 ? 0.0f : k == 1 ? 0.19509032201612825f : k == 2 ? 0.3826834323650898f : k == 3 ? 0.5555702330196022f :
         k == 4 ? 0.7071067811865475f : k == 5 ? 0.8314696123025452f : k == 6 ? 0.9238795325112867f :
         k == 7 ? 0.9807852804032304f : k == 8 ? 1.0f : k == 9 ? 0.9807852804032304f : k == 10 ? 0.9238795325112867f :
         k == 11 ? 0.8314696123025452f : k == 12 ? 0.7071067811865475f : k == 13 ? 0.5555702330196022f :
         k == 14 ? 0.3826834323650898f : 0.19509032201612825f;
}

// d * exp(-/+ 2 pi i k / 32) for a compile-time k in [0, 16)
template <bool INV>
M2H_DEV float2 mul_tw32(float2 d, int k) {
  if (k == 0) return d;
  if (k == 8) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  if (k == 4) {      // (1 -/+ i) / sqrt 2
    const float s = 0.7071067811865476f;
    return c_scale(INV ? c_add(d, make_float2(-d.y, d.x)) : c_add(d, make_float2(d.y, -d.x)), s);
  }
  if (k == 12) {     // (-1 -/+ i) / sqrt 2
    const float s = 0.7071067811865476f;
    return c_scale(INV ? c_add(make_float2(-d.y, d.x), make_float2(-d.x, -d.y)) : c_add(make_float2(d.y, -d.x), make_float2(-d.x, -d.y)), s);
  }
  const float c = tw32_cos(k), s = tw32_sin(k);
  return INV ? c_rot_inv(d, c, s) : c_rot_fwd(d, c, s);
}

// ZERO_UPPER: v[R/2 .. R) is zero padding on entry (never read): the first stage degenerates to twiddle multiplies.
template <int R, bool INV, bool ZERO_UPPER = false>
M2H_DEV void fft_reg(float2 (&v)[R]) {
  M2H_UNROLL
  for (int half = R / 2; half >= 1; half >>= 1) {
    M2H_UNROLL
    for (int blk = 0; blk < R; blk += 2 * half) {
      M2H_UNROLL
      for (int j = 0; j < half; ++j) {
        if (ZERO_UPPER && half == R / 2) {
          v[j + half] = mul_tw32<INV>(v[j], j * (16 / half));
          continue;
        }
        const float2 a = v[blk + j], b = v[blk + j + half];
        v[blk + j] = c_add(a, b);
        v[blk + j + half] = mul_tw32<INV>(c_sub(a, b), j * (16 / half));
      }
    }
  }
}

template <int BITS>
M2H_HD constexpr int bitrev(int k) {
  int r = 0;
  for (int i = 0; i < BITS; ++i) r |= ((k >> i) & 1) << (BITS - 1 - i);
  return r;
}

// out(k, v[bitrev(k)] * w1^k)  (times conj(w1)^k when CONJ) for k = 0..R-1.
// The powers form a geometric sequence with a thread-constant base, so they are generated in registers
// (two-level: k = Q*a + b, depth <= 5 multiplies) instead of being fetched from a shared-memory table:
// table lookups with power-of-two lane strides were 40-67 % bank conflicts in the first ncu captures.
//
// Both anchors w1 and wq = w1^Q come correctly rounded from a table (computing wq from the rounded w1 would amplify
// w1's rounding error by Q): the worst amplification of a table rounding error is (Q - 1) + (R/Q - 1).
#define M2H_TW_Q(R) ((R) == 32 ? 8 : 4)
template <int R, int BITS, bool CONJ, typename Out>
M2H_DEV void twiddle_powers_out(const float2 (&v)[R], float2 w1, float2 wq, Out out) {
  constexpr int Q = M2H_TW_Q(R);
  float2 wb[Q];
  wb[0] = make_float2(1.f, 0.f);
  wb[1] = w1;
  M2H_UNROLL for (int b = 2; b < Q; ++b) wb[b] = c_mul(wb[b >> 1], wb[(b + 1) >> 1]);
  float2 wa = make_float2(1.f, 0.f);
  M2H_UNROLL for (int a = 0; a < R / Q; ++a) {
    M2H_UNROLL for (int b = 0; b < Q; ++b) {
      const int k = a * Q + b;
      if (k == 0) { out(0, v[0]); continue; }
      const float2 t = a == 0 ? wb[b] : (b == 0 ? wa : c_mul(wa, wb[b]));
      out(k, c_mul_sel<CONJ>(v[bitrev<BITS>(k)], t));
    }
    if (a + 1 < R / Q) wa = a == 0 ? wq : (a == 1 ? c_mul(wq, wq) : c_mul(wa, wq));
  }
}
// in-place variant for twiddles applied on load: v[k] *= w1^k (natural order)
template <int R, bool CONJ>
M2H_DEV void twiddle_powers_inplace(float2 (&v)[R], float2 w1, float2 wq) {
  constexpr int Q = M2H_TW_Q(R);
  float2 wb[Q];
  wb[0] = make_float2(1.f, 0.f);
  wb[1] = w1;
  M2H_UNROLL for (int b = 2; b < Q; ++b) wb[b] = c_mul(wb[b >> 1], wb[(b + 1) >> 1]);
  float2 wa = make_float2(1.f, 0.f);
  M2H_UNROLL for (int a = 0; a < R / Q; ++a) {
    M2H_UNROLL for (int b = 0; b < Q; ++b) {
      const int k = a * Q + b;
      if (k == 0) continue;
      const float2 t = a == 0 ? wb[b] : (b == 0 ? wa : c_mul(wa, wb[b]));
      v[k] = c_mul_sel<CONJ>(v[k], t);
    }
    if (a + 1 < R / Q) wa = a == 0 ? wq : (a == 1 ? c_mul(wq, wq) : c_mul(wa, wq));
  }
}

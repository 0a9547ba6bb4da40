// RIR convolution + quantise for RIRs of at most 16384 taps: circular length F = 24576 instead of 32768.
//
// Same contract as render_conv.cuh (scipy.signal.fftconvolve(mono, rir[:, ch], "same") -> np.round -> int16,
// habitat_audio/simulator_train.py:417-423, audio_separation/pretrain/datasets/dataset.py:178-185).  Only taps
// [c-N+1, c+N-1] (c = (L-1)/2) reach the "same" window, and with L <= 16384 the window [c_off, c_off+N) of the full
// convolution ends below 24576, so a circular length of 3 * 8192 is alias-free: 25 % fewer butterflies than F = 32768 and,
// more important, a 102 KB working set - TWO independent CTAs per SM, whose load / store / barrier phases overlap the
// other CTA's butterflies (ncu r2c: the F = 32768 kernel keeps the FP32 pipe only ~50 % busy with one CTA per SM).
//
// Packed real transform of M = 12288 complex points = 3 x 4096:
//     Z[k1 + 3 k2] = FFT_4096( u_k1 )[k2],   u_k1[n2] = W_M^(n2 k1) * sum_n1 z[4096 n1 + n2] W_3^(n1 k1)
// - the radix-3 step rides on the global loads (inputs have at most 8192 packed samples: z[8192 + n2] = 0) and, mirrored,
//   on the quantising store of the inverse;
// - each 4096-point transform is three radix-16 passes over a padded [16][16][16] shared layout
//   (position (a, b, c) = a*272 + b*17 + c: every pass's half-warp touches 16 consecutive or stride-17 slots: conflict-free);
//   decimation in frequency forward, mirrored passes inverse, so the spectrum is used in digit-reversed order and no
//   reorder pass exists;
// - after pass A a slab (fixed first digit) is a self-contained 16 x 16 problem: a half-warp owns a slab through passes
//   B, C, the spectrum product, C' and B' with warp-level synchronisation only.  The real-FFT untangle pairs spectrum
//   index f with M - f = a slab of sub-transform 1 with the complementary slab of sub-transform 2 (and slabs q, 16 - q of
//   sub-transform 0): the two half-warps of a warp own such a pair and exchange rows with warp shuffles, each lane keeping
//   ONE row of 16 points in registers (32 registers: three CTAs' worth of threads fit the register file).
#pragma once
#include "fft_radix.cuh"
#include "render_conv.cuh"      // ConvParams, conv_can_vectorise

#define M2H_C24_THREADS 384
#define M2H_C24_CTAS_PER_SM 2
#define M2H_C24_M 12288                  // complex points of the packed transform
#define M2H_C24_SUB 4096
#define M2H_C24_S1 272                   // slab stride (16 * 17)
#define M2H_C24_S2 17
#define M2H_C24_SUBBUF (16 * M2H_C24_S1) // complex slots of one sub-transform
#define M2H_C24_MAX_TAPS 16384
#define M2H_C24_SCRATCH_F2 12288         // float2 per CTA of spectrum scratch (96 KB, L2-resident)
#define M2H_C24_ROUNDS 2                 // 24 slab pairs / 12 warps

// tables in device global memory
struct M2hConv24Tables {
  float2 a1[256], a4[256];     // W_4096^r, W_4096^(4r): anchors of the pass A / A' twiddles W_4096^(r q1)
  float2 t256[256];            // W_256^(i*16 + j -> i*j): pass B / C' twiddles (copied to shared memory; symmetric in i, j)
  float2 m1[384];              // W_12288^tid: the radix-3 step's twiddle of n2 = tid + 384 i is m1[tid] * W_32^i
  float2 g[768];               // W_24576^(k1 + 3 q1 + 48 q2) at [k1*256 + q1*16 + q2]: untangle twiddle of a row's first element
};

struct Conv24Smem {
  float2 buf[3 * M2H_C24_SUBBUF];
  float2 t256[256];
  float2 sp[3][16];            // the one self-paired row (sub 0, slab 0, q2 = 0): Z, Y and Z' parked for 16 helper lanes
};

M2H_DEV int c24_pos(int n2) { return (n2 >> 8) * M2H_C24_S1 + ((n2 >> 4) & 15) * M2H_C24_S2 + (n2 & 15); }

// ---- radix-3 step on the loads --------------------------------------------------------------------------------------
// LOAD(j) -> packed sample j (0 beyond the valid length).  u0 = x0 + x1, u1 = (x0 + w3 x1) W_M^n2, u2 = (x0 + w3^2 x1) W_M^(2 n2)
template <typename Load>
M2H_DEV void c24_load_radix3(Conv24Smem& s, int tid, const M2hConv24Tables* M2H_RESTRICT t, Load load) {
  const float2 wt = m2h_ldg2(&t->m1[tid]);
  const float hs = 0.8660254037844386f;
  // every global load of the thread is issued before the first use (22 independent loads in flight: the step was bound by
  // their latency when each iteration waited for its own pair - ncu r2e: 20 % of the kernel's samples, long scoreboard)
  M2H_UNROLL for (int half = 0; half < 2; ++half) {          // two batches of 12 / 10 loads: all in flight, no spills at 80 registers
    const int i0 = half * 6, cnt = half ? 5 : 6;
    float2 x0[6], x1[6];
    M2H_UNROLL for (int k = 0; k < 6; ++k) {
      const int n2 = tid + M2H_C24_THREADS * (i0 + k);
      x0[k] = make_float2(0.f, 0.f); x1[k] = make_float2(0.f, 0.f);
      if (k < cnt && n2 < M2H_C24_SUB) { x0[k] = load(n2); x1[k] = load(n2 + M2H_C24_SUB); }
    }
    M2H_UNROLL for (int k = 0; k < 6; ++k) {
      const int i = i0 + k, n2 = tid + M2H_C24_THREADS * i;
      if (k < cnt && n2 < M2H_C24_SUB) {
        const float2 a = c_add(x0[k], c_scale(x1[k], -0.5f));
        const float2 b = c_scale(make_float2(x1[k].y, -x1[k].x), hs);   // -i (sqrt3/2) x1
        const float2 w1 = mul_tw32<false>(wt, i);                          // W_M^n2 (384 / 12288 = 1 / 32)
        const float2 w2 = c_mul(w1, w1);
        const int p = c24_pos(n2);
        s.buf[p] = c_add(x0[k], x1[k]);
        s.buf[M2H_C24_SUBBUF + p] = c_mul(c_add(a, b), w1);
        s.buf[2 * M2H_C24_SUBBUF + p] = c_mul(c_sub(a, b), w2);
      }
    }
  }
}

// ---- pass A / A' of the three 4096-point transforms: 768 (sub, column) tasks, two per thread ---------------------------
M2H_DEV void c24_fwd_pass_a(Conv24Smem& s, int tid, const M2hConv24Tables* M2H_RESTRICT t) {
  M2H_UNROLL for (int j = 0; j < 2; ++j) {
    const int task = tid + M2H_C24_THREADS * j, sub = task >> 8, r = task & 255;
    float2* p = s.buf + sub * M2H_C24_SUBBUF + (r >> 4) * M2H_C24_S2 + (r & 15);
    float2 v[16];
    M2H_UNROLL for (int i1 = 0; i1 < 16; ++i1) v[i1] = p[i1 * M2H_C24_S1];
    fft_reg<16, false>(v);
    twiddle_powers_out<16, 4, false>(v, m2h_ldg2(&t->a1[r]), m2h_ldg2(&t->a4[r]), [&](int q1, float2 x) { p[q1 * M2H_C24_S1] = x; });
  }
}
M2H_DEV void c24_inv_pass_a(Conv24Smem& s, int tid, const M2hConv24Tables* M2H_RESTRICT t) {
  M2H_UNROLL for (int j = 0; j < 2; ++j) {
    const int task = tid + M2H_C24_THREADS * j, sub = task >> 8, r = task & 255;
    float2* p = s.buf + sub * M2H_C24_SUBBUF + (r >> 4) * M2H_C24_S2 + (r & 15);
    float2 v[16];
    M2H_UNROLL for (int q1 = 0; q1 < 16; ++q1) v[q1] = p[q1 * M2H_C24_S1];
    twiddle_powers_inplace<16, true>(v, m2h_ldg2(&t->a1[r]), m2h_ldg2(&t->a4[r]));
    fft_reg<16, true>(v);
    M2H_UNROLL for (int i1 = 0; i1 < 16; ++i1) p[i1 * M2H_C24_S1] = v[bitrev<4>(i1)];
  }
}

// ---- warp-local middle ------------------------------------------------------------------------------------------------
// Slab-pair unit u of (warp w, round rho) = w + 12 rho; lanes 0..15 own the unit's first slab, lanes 16..31 its second:
//   u < 16:      (sub 1, slab u)       | (sub 2, slab 15 - u)        f = 1 + 3 k2  <->  M - f = 2 + 3 (4095 - k2)
//   16 <= u < 23 (q = u - 15):  (sub 0, slab q)  | (sub 0, slab 16 - q)     f = 3 k2 <-> 3 (4096 - k2)
//   u = 23:      (sub 0, slab 8)       | (sub 0, slab 0)
// Lane t = lane & 15 holds row (slab, q2 = t).  The row that holds M - f sits in lane `partner`, element 15 - q3 - except
// for row (sub 0, slab 0, q2 = 0), which pairs with itself through 16 - q3 (f = 0 carries the real DC and Nyquist terms).
struct C24Row { int base; int k1, q1; int partner; bool special; };
M2H_DEV C24Row c24_row(int tid, int round) {
  const int w = tid >> 5, lane = tid & 31, h = lane >> 4, t = lane & 15;
  const int u = w + 12 * round;
  C24Row r;
  r.special = false;
  r.partner = 31 - lane;
  if (u < 16) { r.k1 = 1 + h; r.q1 = h ? 15 - u : u; }
  else if (u < 23) { const int q = u - 15; r.k1 = 0; r.q1 = h ? 16 - q : q; }
  else {
    r.k1 = 0; r.q1 = h ? 0 : 8;
    if (!h) r.partner = 15 - lane;
    else { r.partner = (t == 0 || t == 8) ? lane : 16 + (16 - t); r.special = (t == 0); }
  }
  r.base = r.k1 * M2H_C24_SUBBUF + r.q1 * M2H_C24_S1;
  return r;
}

// pass B (forward) on the lane's slab: column i3 = t, radix 16 along i2, twiddle W_256^(t q2) from the shared table
M2H_DEV void c24_fwd_pass_b(Conv24Smem& s, int tid, int round) {
  const C24Row R = c24_row(tid, round);
  const int t = tid & 15;
  float2* p = s.buf + R.base + t;
  float2 v[16];
  M2H_UNROLL for (int i2 = 0; i2 < 16; ++i2) v[i2] = p[i2 * M2H_C24_S2];
  fft_reg<16, false>(v);
  p[0] = v[0];
  M2H_UNROLL for (int q2 = 1; q2 < 16; ++q2) p[q2 * M2H_C24_S2] = c_mul(v[bitrev<4>(q2)], s.t256[q2 * 16 + t]);
}
// pass B' (inverse): column i3 = t, radix 16 along q2 (its twiddle was applied by pass C')
M2H_DEV void c24_inv_pass_b(Conv24Smem& s, int tid, int round) {
  const C24Row R = c24_row(tid, round);
  const int t = tid & 15;
  float2* p = s.buf + R.base + t;
  float2 v[16];
  M2H_UNROLL for (int q2 = 0; q2 < 16; ++q2) v[q2] = p[q2 * M2H_C24_S2];
  fft_reg<16, true>(v);
  M2H_UNROLL for (int i2 = 0; i2 < 16; ++i2) p[i2 * M2H_C24_S2] = v[bitrev<4>(i2)];
}
// pass C (forward) of the lane's row q2 = t: on return v[bitrev(q3)] = Z at (slab, t, q3)
M2H_DEV void c24_fwd_pass_c(const Conv24Smem& s, int tid, int round, float2 (&v)[16]) {
  const C24Row R = c24_row(tid, round);
  const float2* p = s.buf + R.base + (tid & 15) * M2H_C24_S2;
  M2H_UNROLL for (int i3 = 0; i3 < 16; ++i3) v[i3] = p[i3];
  fft_reg<16, false>(v);
}
// pass C' (inverse) of the lane's row: u[q3] natural order in, twiddle conj W_256^(i3 t), store
M2H_DEV void c24_inv_pass_c(Conv24Smem& s, int tid, int round, float2 (&u)[16]) {
  const C24Row R = c24_row(tid, round);
  const int t = tid & 15;
  float2* p = s.buf + R.base + t * M2H_C24_S2;
  fft_reg<16, true>(u);
  p[0] = u[0];
  M2H_UNROLL for (int i3 = 1; i3 < 16; ++i3) p[i3] = c_mulc(u[bitrev<4>(i3)], s.t256[i3 * 16 + t]);
}

// 2 X[f] = (Zm + conj Zp) - i w (Zm - conj Zp),  w = W_F^f,  Zp = Z[M - f]   (valid for every 0 < f < M, f = M/2 included)
// (conjugations and rotations by +-i are operand modifiers of the packed adds: five FADD2 / FMUL2 / FFMA2 in all)
M2H_DEV float2 c24_untangle(float2 zm, float2 zp, float2 w) {
  const float2 sm = c_add(zm, make_float2(zp.x, -zp.y));
  const float2 df = c_add(zm, make_float2(-zp.x, zp.y));
  const float2 q = c_mul(w, df);
  return c_add(sm, make_float2(q.y, -q.x));
}
// 8 Z'[f] = (Ym + conj Yp) + i conj(w) (Ym - conj Yp)      (Y = 4 x the true product spectrum)
M2H_DEV float2 c24_retangle(float2 ym, float2 yp, float2 w) {
  const float2 sm = c_add(ym, make_float2(yp.x, -yp.y));
  const float2 df = c_add(ym, make_float2(-yp.x, yp.y));
  const float2 q = c_mulc(df, w);
  return c_add(sm, make_float2(-q.y, q.x));
}

// Spectrum step of a lane's row.  `zp(idx)` / `yp(idx)` fetch register idx of the PARTNER lane's v / y array.
// MONO: stores 2 X[f] of the clip to the scratch slot;  else: multiplies the RIR spectrum by it and re-tangles into u[q3].
// Split in two halves so that the host emulation can run them as separate phases (the exchange crosses threads).
template <bool MONO, typename PartnerZ>
M2H_DEV void c24_spectrum_a(int tid, int round, const M2hConv24Tables* M2H_RESTRICT t, float2* g, const float2 (&v)[16], float2 (&y)[16], PartnerZ zp) {
  const C24Row R = c24_row(tid, round);
  const float2 wU = m2h_ldg2(&t->g[R.k1 * 256 + R.q1 * 16 + (tid & 15)]);
  float2* gx = g + (size_t)round * 16 * M2H_C24_THREADS + tid;
  M2H_UNROLL for (int q3 = 0; q3 < 16; ++q3) {
    const float2 zo = zp(bitrev<4>(15 - q3));                      // every lane takes part in the exchange
    const float2 x2 = c24_untangle(v[bitrev<4>(q3)], zo, mul_tw32<false>(wU, q3));       // w = W_F^(f0 + 768 q3)
    if (MONO) { if (!R.special) m2h_st2_keep(gx + q3 * M2H_C24_THREADS, x2); }
    else y[q3] = c_mul(y[q3], x2);                                   // y arrives holding 2 X[f] of the clip (c24_load_x)
  }
}
// (Scratch accesses carry the L2 evict_last policy, the clip / RIR loads and the int16 stores evict_first: the launch's 460 MB of
// streaming traffic no longer pushes dirty scratch lines out - DRAM traffic 555 -> 438 MB per 1024 mixtures, ncu.)
// The clip's spectrum row comes from the L2-resident scratch: issued BEFORE the row's own radix pass so that the ~300-cycle
// L2 latency runs under the butterflies (ncu r2f: 12.7 % of the kernel's samples waited on these loads at the product)
M2H_DEV void c24_load_x(int tid, int round, const float2* g, float2 (&y)[16]) {
  const float2* gx = g + (size_t)round * 16 * M2H_C24_THREADS + tid;
  M2H_UNROLL for (int q3 = 0; q3 < 16; ++q3) y[q3] = m2h_ld2_keep(gx + q3 * M2H_C24_THREADS);
}
template <typename PartnerY>
M2H_DEV void c24_spectrum_b(int tid, int round, const M2hConv24Tables* M2H_RESTRICT t, const float2 (&y)[16], float2 (&u)[16], PartnerY yp) {
  const C24Row R = c24_row(tid, round);
  const float2 wU = m2h_ldg2(&t->g[R.k1 * 256 + R.q1 * 16 + (tid & 15)]);
  M2H_UNROLL for (int q3 = 0; q3 < 16; ++q3) u[q3] = c24_retangle(y[q3], yp(15 - q3), mul_tw32<false>(wU, q3));
}
// The self-paired row (held by lane 16 of the warp that owns unit 23) pairs element q3 with 16 - q3 of ITSELF, and its
// f = 0 slot carries the real DC and Nyquist terms.  Doing that in the owning lane's registers costs every thread of the
// kernel registers (the compiler keeps both variants live: 700 bytes of spills); instead the lane parks its row in shared
// memory and lanes 0..15 of the same warp handle one element each (f = 768 q3, w = W_32^q3 = t256[8 * 16 + q3]).
M2H_DEV bool c24_is_special_warp(int tid, int round) { return (tid >> 5) + 12 * round == 23; }
M2H_DEV int c24_special_tid(int round) { return (23 - 12 * round) * 32 + 16; }
template <bool MONO>
M2H_DEV void c24_special_a(Conv24Smem& s, int tid, int round, float2* g) {      // lanes 0..15 of the special warp
  const int q3 = tid & 31;
  const float2 zm = s.sp[0][q3];
  float2 x2 = c24_untangle(zm, s.sp[0][(16 - q3) & 15], s.t256[8 * 16 + q3]);
  if (q3 == 0) x2 = make_float2(2.f * (zm.x + zm.y), 2.f * (zm.x - zm.y));             // f = 0 | f = M: both real
  float2* gx = g + (size_t)round * 16 * M2H_C24_THREADS + q3 * M2H_C24_THREADS + c24_special_tid(round);
  if (MONO) m2h_st2_keep(gx, x2);
  else { const float2 x = m2h_ld2_keep(gx); s.sp[1][q3] = q3 == 0 ? make_float2(x.x * x2.x, x.y * x2.y) : c_mul(x, x2); }
}
M2H_DEV void c24_special_b(Conv24Smem& s, int tid) {                                       // lanes 0..15 of the special warp
  const int q3 = tid & 31;
  const float2 y = s.sp[1][q3];
  s.sp[2][q3] = q3 == 0 ? make_float2(y.x + y.y, y.x - y.y) : c24_retangle(y, s.sp[1][(16 - q3) & 15], s.t256[8 * 16 + q3]);
}

#ifdef M2H_HOST_EMU
  #define M2H_C24_REGS_DECL static float2 c24_v[M2H_C24_THREADS][16], c24_y[M2H_C24_THREADS][16], c24_u[M2H_C24_THREADS][16]
#else
  M2H_DEV float2 c24_shfl(float2 a, int lane) {
    return make_float2(__shfl_sync(0xffffffffu, a.x, lane), __shfl_sync(0xffffffffu, a.y, lane));
  }
#endif

// ---- radix-3 step of the inverse fused with scale / round: staged int16 pairs replace the consumed slots ---------------
M2H_DEV uint32_t c24_quantise_pair(float a, float b) {
  return ((uint32_t)m2h_f2i_rn(a) & 0xffffu) | ((uint32_t)m2h_f2i_rn(b) << 16);
}
// Phase 1 computes the thread's (up to) 33 output words; after a CTA barrier phase 2 stores them CONTIGUOUSLY over the
// consumed transform buffer (word j of the packed result at stage[j]), so that the copy-out is plain coalesced reads.
#define M2H_C24_NW 33
M2H_DEV void c24_radix3_words(const Conv24Smem& s, int tid, const M2hConv24Tables* M2H_RESTRICT t, uint32_t (&w)[M2H_C24_NW]) {
  const float2 wt = m2h_ldg2(&t->m1[tid]);
  const float hs = 0.8660254037844386f;
  const float scale = 1.0f / (8.0f * (float)M2H_C24_M);              // 1/M inverse scale, 2*2*2 from the untangle / re-tangle steps
  M2H_UNROLL for (int i = 0; i < 11; ++i) {
    const int n2 = tid + M2H_C24_THREADS * i;
    w[3 * i] = w[3 * i + 1] = w[3 * i + 2] = 0u;
    if (n2 < M2H_C24_SUB) {
      const int p = c24_pos(n2);
      const float2 a0 = s.buf[p];
      const float2 w1 = mul_tw32<false>(wt, i);
      const float2 b1 = c_mulc(s.buf[M2H_C24_SUBBUF + p], w1);
      const float2 b2 = c_mulc(s.buf[2 * M2H_C24_SUBBUF + p], c_mul(w1, w1));
      const float2 sm = c_add(b1, b2), df = c_sub(b1, b2);
      const float2 c = c_add(a0, c_scale(sm, -0.5f));
      const float2 d = c_scale(make_float2(-df.y, df.x), hs);         // i (sqrt3/2) (b1 - b2)
      const float2 z0 = c_scale(c_add(a0, sm), scale), z1 = c_scale(c_add(c, d), scale), z2 = c_scale(c_sub(c, d), scale);
      w[3 * i] = c24_quantise_pair(z0.x, z0.y);
      w[3 * i + 1] = c24_quantise_pair(z1.x, z1.y);
      w[3 * i + 2] = c24_quantise_pair(z2.x, z2.y);
    }
  }
}
M2H_DEV void c24_store_words(Conv24Smem& s, int tid, const uint32_t (&w)[M2H_C24_NW]) {
  uint32_t* stage = reinterpret_cast<uint32_t*>(s.buf);
  M2H_UNROLL for (int i = 0; i < 11; ++i) {
    const int n2 = tid + M2H_C24_THREADS * i;
    if (n2 < M2H_C24_SUB) { stage[n2] = w[3 * i]; stage[n2 + M2H_C24_SUB] = w[3 * i + 1]; stage[n2 + 2 * M2H_C24_SUB] = w[3 * i + 2]; }
  }
}
template <bool VEC>
M2H_DEV void c24_copy_out(const Conv24Smem& s, int tid, int16_t* out, int N, int c_off) {
  const uint32_t* stage = reinterpret_cast<const uint32_t*>(s.buf);
  if (VEC) {
    uint32_t* o = (uint32_t*)out;
    const int sh = (c_off & 1) * 16, k0 = c_off >> 1;
    for (int w = tid; w < (N >> 1); w += M2H_C24_THREADS) m2h_st_u32_stream(o + w, m2h_funnel_r(stage[k0 + w], stage[k0 + w + 1], sh));
  } else {
    for (int m = tid; m < N; m += M2H_C24_THREADS) {
      const int q = m + c_off;
      const uint32_t w = stage[q >> 1];
      out[m] = (int16_t)(uint16_t)((q & 1) ? (w >> 16) : (w & 0xffffu));
    }
  }
}

// One forward transform up to (and including) the spectrum step, then - unless MONO - the inverse up to pass A'.
// LOAD(j) returns packed input sample j.
template <bool MONO, typename Load>
M2H_DEV void c24_transform(Conv24Smem& s, const M2hConv24Tables* M2H_RESTRICT t, float2* g, Load load) {
  M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_load_radix3(s, tid, t, [&](int j) { return load(tid, j); });
  M2H_SYNC();
  M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_fwd_pass_a(s, tid, t);
  M2H_SYNC();
  for (int round = 0; round < M2H_C24_ROUNDS; ++round) {
    M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_fwd_pass_b(s, tid, round);
    M2H_SYNC_WARP();
#ifdef M2H_HOST_EMU
    M2H_C24_REGS_DECL;
    const int sp_tid = c24_special_tid(round), sp_w0 = sp_tid - 16;
    if (!MONO) { M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_load_x(tid, round, g, c24_y[tid]); }
    M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_fwd_pass_c(s, tid, round, c24_v[tid]);
    if (round == 1) for (int q3 = 0; q3 < 16; ++q3) s.sp[0][q3] = c24_v[sp_tid][bitrev<4>(q3)];
    M2H_FOR_THREADS(tid, M2H_C24_THREADS) {
      const int pt = (tid & ~31) | c24_row(tid, round).partner;
      c24_spectrum_a<MONO>(tid, round, t, g, c24_v[tid], c24_y[tid], [&](int idx) { return c24_v[pt][idx]; });
    }
    if (round == 1) for (int l = 0; l < 16; ++l) c24_special_a<MONO>(s, sp_w0 + l, round, g);
    if (!MONO) {
      if (round == 1) for (int q3 = 0; q3 < 16; ++q3) c24_y[sp_tid][q3] = s.sp[1][q3];
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) {
        const int pt = (tid & ~31) | c24_row(tid, round).partner;
        c24_spectrum_b(tid, round, t, c24_y[tid], c24_u[tid], [&](int idx) { return c24_y[pt][idx]; });
      }
      if (round == 1) {
        for (int l = 0; l < 16; ++l) c24_special_b(s, sp_w0 + l);
        for (int q3 = 0; q3 < 16; ++q3) c24_u[sp_tid][q3] = s.sp[2][q3];
      }
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_inv_pass_c(s, tid, round, c24_u[tid]);
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_inv_pass_b(s, tid, round);
    }
#else
    {
      const int tid = threadIdx.x, lane = tid & 31;
      const int partner = c24_row(tid, round).partner;
      const bool spw = c24_is_special_warp(tid, round);              // warp-uniform
      float2 v[16], y[16];
      if (!MONO) c24_load_x(tid, round, g, y);
      c24_fwd_pass_c(s, tid, round, v);
      if (spw && lane == 16) { M2H_UNROLL for (int q3 = 0; q3 < 16; ++q3) s.sp[0][q3] = v[bitrev<4>(q3)]; }
      c24_spectrum_a<MONO>(tid, round, t, g, v, y, [&](int idx) { return c24_shfl(v[idx], partner); });
      if (spw) {
        M2H_SYNC_WARP();
        if (lane < 16) c24_special_a<MONO>(s, tid, round, g);
        M2H_SYNC_WARP();
        if (!MONO && lane == 16) { M2H_UNROLL for (int q3 = 0; q3 < 16; ++q3) y[q3] = s.sp[1][q3]; }
      }
      if (!MONO) {
        c24_spectrum_b(tid, round, t, y, v, [&](int idx) { return c24_shfl(y[idx], partner); });
        if (spw) {
          if (lane < 16) c24_special_b(s, tid);
          M2H_SYNC_WARP();
          if (lane == 16) { M2H_UNROLL for (int q3 = 0; q3 < 16; ++q3) v[q3] = s.sp[2][q3]; }
        }
        c24_inv_pass_c(s, tid, round, v);
        M2H_SYNC_WARP();
        c24_inv_pass_b(s, tid, round);
      }
    }
#endif
  }
  if (!MONO) {
    M2H_SYNC();
    M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_inv_pass_a(s, tid, t);
  }
  M2H_SYNC();
}

// L2 prefetch of a row that the NEXT transform will load (one 128-byte line per thread and step)
M2H_DEV void c24_prefetch_row(int tid, const void* row, int bytes) {
  for (int o = tid * 128; o < bytes; o += M2H_C24_THREADS * 128) m2h_prefetch_l2((const char*)row + o);
}

// The whole per-CTA loop (same ConvParams as render_conv.cuh).  Requires P.Lmax <= 16384 and P.N <= 16384.
template <bool VEC>
M2H_DEV void conv24_cta_body(const ConvParams& P, Conv24Smem& s, const M2hConv24Tables* M2H_RESTRICT gtab, int cta, int n_ctas) {
  float2* g = reinterpret_cast<float2*>(P.scratch) + (size_t)cta * M2H_C24_SCRATCH_F2;
  M2H_FOR_THREADS(tid, M2H_C24_THREADS) { if (tid < 256) s.t256[tid] = gtab->t256[tid]; }
  M2H_PDL_WAIT();
  M2H_SYNC();
  const int N = P.N;
  const int n_units = P.ear_split ? 2 * P.n_items : P.n_items;
  for (int unit = cta; unit < n_units; unit += n_ctas) {
    const int item = P.ear_split ? unit >> 1 : unit;
    const int ear_lo = P.ear_split ? (unit & 1) : 0, ear_hi = P.ear_split ? ear_lo + 1 : 2;
    const int next_item = P.ear_split ? (unit + n_ctas) >> 1 : unit + n_ctas;       // whose clip this CTA transforms next
    const int mi = P.mono_index ? P.mono_index[item] : item;
    const int ri = P.rir_index ? P.rir_index[item] : item;
    const int L = P.rir_len[ri];
    const int c = (L - 1) / 2;
    int jmin = c - N + 1 > 0 ? c - N + 1 : 0;
    const int jmax = c + N - 1 < L - 1 ? c + N - 1 : L - 1;
    if (VEC && (jmin & 1)) jmin -= 1;           // 8-byte aligned first tap (rows are: Lmax is even); the extra tap cannot reach the window
    const int Leff = jmax - jmin + 1;
    const int c_off = c - jmin;
    if (L <= 0) {     // unreadable / empty RIR -> the reference substitutes zeros (simulator_train.py:407-415)
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) {
        int16_t* out = P.waves + ((size_t)item * 2 + ear_lo) * N;
        for (int i = tid; i < (ear_hi - ear_lo) * N; i += M2H_C24_THREADS) out[i] = 0;
      }
      continue;
    }
    const int16_t* x = P.mono + (size_t)mi * N;
    M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_prefetch_row(tid, P.rir + ((size_t)ri * 2 + ear_lo) * P.Lmax + jmin, Leff * 4);      // first ear, loaded one transform from now
    c24_transform<true>(s, gtab, g, [&](int tid, int j) {
      float a = 0.f, b = 0.f;
      if (VEC) {
        uint32_t w = 0;
        if (2 * j < N) w = m2h_ldg_u32_stream((const uint32_t*)x + j);     // N even: a packed sample is one aligned 32-bit load
        a = (float)(int16_t)(w & 0xffffu);
        b = (float)(int16_t)(w >> 16);
      } else {
        if (2 * j < N) a = (float)x[2 * j];
        if (2 * j + 1 < N) b = (float)x[2 * j + 1];
      }
      return make_float2(a, b);
    });
    for (int ear = ear_lo; ear < ear_hi; ++ear) {
      const float* h = P.rir + ((size_t)ri * 2 + ear) * P.Lmax + jmin;
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) {          // rows the next transform will load: the other ear, or the next item's clip
        if (ear + 1 < ear_hi) c24_prefetch_row(tid, h + P.Lmax, Leff * 4);
        else if (unit + n_ctas < n_units)
          c24_prefetch_row(tid, P.mono + (size_t)(P.mono_index ? P.mono_index[next_item] : next_item) * N, N * 2);
      }
      c24_transform<false>(s, gtab, g, [&](int tid, int j) {
        float2 a = make_float2(0.f, 0.f);
        if (VEC) {
          if (2 * j < Leff) a = m2h_ldg2_stream((const float2*)h + j);
          if (2 * j + 1 >= Leff) a.y = 0.f;
        } else {
          if (2 * j < Leff) a.x = m2h_ldg_stream(h + 2 * j);
          if (2 * j + 1 < Leff) a.y = m2h_ldg_stream(h + 2 * j + 1);
        }
        return a;
      });
#ifdef M2H_HOST_EMU
      static uint32_t c24_w[M2H_C24_THREADS][M2H_C24_NW];
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_radix3_words(s, tid, gtab, c24_w[tid]);
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_store_words(s, tid, c24_w[tid]);
#else
      {
        uint32_t w[M2H_C24_NW];
        c24_radix3_words(s, threadIdx.x, gtab, w);
        M2H_SYNC();                       // every thread has read its triples: the buffer may be overwritten
        c24_store_words(s, threadIdx.x, w);
      }
#endif
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_C24_THREADS) c24_copy_out<VEC>(s, tid, P.waves + ((size_t)item * 2 + ear) * N, N, c_off);
      M2H_SYNC();
    }
  }
}

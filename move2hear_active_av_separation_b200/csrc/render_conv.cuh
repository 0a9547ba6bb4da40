// RIR convolution + quantise: the first half of the mixture render.
//
// Replaces scipy.signal.fftconvolve(mono, rir[:, ch], mode="same") followed by
// np.round(..).astype("int16") at habitat_audio/simulator_train.py:417-423,
// habitat_audio/simulator_eval.py:429-435 and audio_separation/pretrain/datasets/dataset.py:178-185.
//
// One CTA (512 threads) owns one (mixture b, source k) work item at a time (persistent grid):
//   1. mono (int16, N samples) is packed two-reals-per-complex, FFT'd (M = 16384 complex points, entirely in
//      shared memory) and untangled into the half spectrum X[0..M]; X goes to a 128 KB per-CTA scratch that
//      stays L2-resident (a second shared-memory spectrum does not fit next to the 136 KB working buffer);
//   2. for each ear: the effective RIR taps are packed and FFT'd the same way; one pointwise pass untangles
//      H, multiplies by X and re-tangles for the inverse; the inverse FFT's last pass scales, rounds half-to-even,
//      wraps to int16 like numpy's cast and stores the N-sample "same" window.
//
// FFT-16384 = 32 x 32 x 16 (three in-register radix passes over a padded [32][32][16] shared layout,
// position(i1,i2,i3) = i1*544 + i2*17 + i3, which keeps every pass's 64-bit accesses conflict-free).
// Forward passes are decimation-in-frequency: the spectrum is left in digit-reversed order
// (position (k1,k2,k3) <-> f = k1 + 32*k2 + 1024*k3); the inverse runs the mirrored passes, so no reorder pass.
#pragma once
#include "fft_radix.cuh"

#define M2H_CONV_THREADS 512
#define M2H_CONV_S1 544                 // stride of the slowest digit (32*17)
#define M2H_CONV_S2 17                  // stride of the middle digit (16 + 1 pad)
#define M2H_CONV_BUF (32 * M2H_CONV_S1) // complex elements in the working buffer
#define M2H_CONV_SCRATCH_F4 8192        // float4 per CTA of spectrum scratch

struct ConvParams {
  const int16_t* mono;     // [B, K, N]          (or the clip pool [n_clips, N] when mono_index is set)
  const float* rir;        // [B, K, 2, Lmax]  (ear-planar; or the RIR pool [n_rirs, 2, Lmax] when rir_index is set)
  const int32_t* rir_len;  // [B, K]            (or [n_rirs])
  const int32_t* mono_index;  // [B, K] row of each (mixture, source) in the clip pool, or null = identity
  const int32_t* rir_index;   // [B, K] row in the RIR pool, or null = identity
  int16_t* waves;          // [B, K, 2, N]   rounded per-source binaural waveform (int16 full-scale units)
  float4* scratch;         // [gridDim.x, 8192]
  int n_items;             // B * K
  int N;
  int Lmax;
  int ear_split;           // F = 24576 kernel only: a work unit is (item, ear) - two CTAs per item, each repeating the clip's
                           // transform - for launches that cannot fill the GPU with whole items (rollout ranks, B = 1 shims):
                           // 3 instead of 5 transforms on an item's critical path
};

struct ConvSmem {
  float2 buf[M2H_CONV_BUF];
  uint32_t stage[M2H_CONV_M];   // quantised output of one inverse transform, natural order: word j = samples (2j, 2j+1)
};

// Every inter-pass twiddle set is a geometric sequence whose base depends only on the thread id: each pass fetches
// its two correctly-rounded anchors (base, base^Q) with one coalesced, L1-resident global load and generates the other
// powers in registers (fft_radix.cuh) - no shared-memory twiddle tables, hence no table bank conflicts.
#define M2H_CONV_TW(t, which, tid) m2h_ldg2(&(t)->thr[which][tid])
M2H_HD constexpr float conv_w16384_cos(int i) { return i == 0 ? 1.0f : i == 1 ? 0.9999999264657179f : i == 2 ? 0.9999997058628822f : i == 3 ? 0.9999993381915255f : i == 4 ? 0.9999988234517019f : i == 5 ? 0.999998161643487f : i == 6 ? 0.9999973527669782f : i == 7 ? 0.9999963968222944f : i == 8 ? 0.9999952938095762f : i == 9 ? 0.9999940437289858f : i == 10 ? 0.9999926465807072f : i == 11 ? 0.9999911023649456f : i == 12 ? 0.9999894110819284f : i == 13 ? 0.9999875727319041f : i == 14 ? 0.9999855873151432f : 0.9999834548319377f; }
M2H_HD constexpr float conv_w16384_sin(int i) { return i == 0 ? -0.0f : i == 1 ? -0.00038349518757139556f : i == 2 ? -0.0007669903187427045f : i == 3 ? -0.0011504853371138485f : i == 4 ? -0.0015339801862847655f : i == 5 ? -0.001917474809855419f : i == 6 ? -0.002300969151425805f : i == 7 ? -0.0026844631545959617f : i == 8 ? -0.003067956762965976f : i == 9 ? -0.003451449920135994f : i == 10 ? -0.0038349425697062275f : i == 11 ? -0.004218434655276963f : i == 12 ? -0.0046019261204485705f : i == 13 ? -0.0049854169088215106f : i == 14 ? -0.0053689069639963425f : -0.005752396229573736f; }

M2H_DEV int conv_pos(int f) {            // position of spectrum index f = k1 + 32*k2 + 1024*k3
  return (f & 31) * M2H_CONV_S1 + ((f >> 5) & 31) * M2H_CONV_S2 + (f >> 10);
}
// pair index -> spectrum index f in [0, 8192): k3 fastest so that half-warps touch consecutive positions
M2H_DEV int conv_pair_f(int idx) { return (idx >> 8) + 32 * ((idx >> 3) & 31) + 1024 * (idx & 7); }

// ---- warp-local middle of the transforms -------------------------------------------------------------------------
// Pass A leaves slab k1 (544 elements) holding a 32 x 16 (n2, n3) matrix; passes B (radix 32 along n2), C (radix 16 along
// n3), the pointwise spectrum product and the mirrored inverse passes C', B' never leave a slab - except that the product
// pairs spectrum index f with M - f, i.e. slab k1 with slab 32 - k1.  Warp w therefore owns the slab pair (w, 32 - w)
// (warp 0: the self-paired slabs 0 and 16) from pass B to pass B': inside that stretch warps only need __syncwarp(),
// and passes C, the product and C' run on a row pair held in registers (no shared-memory round trip between them).
M2H_DEV int conv_slab_of(int tid, int half) { const int w = tid >> 5; return half == 0 ? w : (w ? 32 - w : 16); }

M2H_DEV void conv_fwd_pass_b(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t) {
  const int l = tid & 31, n3 = l & 15;
  float2* p = s.buf + conv_slab_of(tid, l >> 4) * M2H_CONV_S1 + n3;
  float2 v[32];
  M2H_UNROLL for (int j = 0; j < 32; ++j) v[j] = p[j * M2H_CONV_S2];
  fft_reg<32, false>(v);
  twiddle_powers_out<32, 5, false>(v, M2H_CONV_TW(t, 2, tid), M2H_CONV_TW(t, 3, tid), [&](int k2, float2 x) { p[k2 * M2H_CONV_S2] = x; });   // W_512^(n3*k2)
}
// pass B' stores untwiddled; conj W_16384^((16*n2+n3)*k1) = conj(W_16384^r)^k1 is applied by pass A' on load
M2H_DEV void conv_inv_pass_b(ConvSmem& s, int tid) {
  const int l = tid & 31, n3 = l & 15;
  float2* p = s.buf + conv_slab_of(tid, l >> 4) * M2H_CONV_S1 + n3;
  float2 v[32];
  M2H_UNROLL for (int j = 0; j < 32; ++j) v[j] = p[j * M2H_CONV_S2];
  fft_reg<32, true>(v);
  M2H_UNROLL for (int n2 = 0; n2 < 32; ++n2) p[n2 * M2H_CONV_S2] = v[bitrev<5>(n2)];
}

// Row pair of a lane in the fused pass: row A = (slab, k2) at position pa, its mirror row B at pb.  For every lane but
// lane 16 of warp 0, spectrum index f = (row A, k3) pairs with M - f = (row B, 15 - k3):
//   warps 1..15: A = (w, l),  B = (32 - w, 31 - l);   warp 0, lanes 0..15: A = (16, l), B = (16, 31 - l);
//   warp 0, lanes 17..31 (j = l - 16): A = (0, j), B = (0, 32 - j).
// Lane 16 of warp 0 holds the two self-paired rows (0, 0) [k3 <-> 16 - k3, with f = 0 and f = M/2 on their own] and
// (0, 16) [k3 <-> 15 - k3].
struct ConvRows { int pa, pb, k2a, k2b; };
M2H_DEV ConvRows conv_rows(int tid) {
  const int w = tid >> 5, l = tid & 31;
  ConvRows r;
  if (w) { r.pa = w * M2H_CONV_S1 + l * M2H_CONV_S2; r.pb = (32 - w) * M2H_CONV_S1 + (31 - l) * M2H_CONV_S2; r.k2a = l; r.k2b = 31 - l; }
  else if (l < 16) { r.pa = 16 * M2H_CONV_S1 + l * M2H_CONV_S2; r.pb = 16 * M2H_CONV_S1 + (31 - l) * M2H_CONV_S2; r.k2a = l; r.k2b = 31 - l; }
  else { const int j = l - 16; r.k2a = j; r.k2b = j ? 32 - j : 16; r.pa = r.k2a * M2H_CONV_S2; r.pb = r.k2b * M2H_CONV_S2; }
  return r;
}

// Untangle the packed-real spectrum at (f, M-f):  returns 2*X[f] in xa and 2*X[M-f] in xb  (w = W_32768^f, any 0 < f < M).
M2H_DEV void conv_untangle(float2 za, float2 zb, float2 w, float2& xa, float2& xb) {
  const float2 p = make_float2(za.x + zb.x, za.y - zb.y);       // Za + conj(Zb)         = 2 E[f]
  const float2 o = make_float2(za.y + zb.y, zb.x - za.x);       // -i (Za - conj(Zb))    = 2 O[f]
  const float2 q = c_mul(w, o);
  xa = c_add(p, q);
  xb = c_conj(c_sub(p, q));
}
// H pair (za, zb) at (f, M-f) times the X pair x = {2X[f], 2X[M-f]}, re-tangled for the inverse packed transform (in place)
M2H_DEV void conv_pair_product(float2& za, float2& zb, float2 w, float4 x) {
  float2 ha, hb;
  conv_untangle(za, zb, w, ha, hb);
  const float2 ya = c_mul(make_float2(x.x, x.y), ha);
  const float2 yb = c_mul(make_float2(x.z, x.w), hb);
  const float2 e = make_float2(ya.x + yb.x, ya.y - yb.y);            // Ya + conj(Yb)
  const float2 d = make_float2(ya.x - yb.x, ya.y + yb.y);            // Ya - conj(Yb)
  const float2 o = c_mulc(d, w);                                       // conj(w) * d
  za = make_float2(e.x - o.y, e.y + o.x);                              // E + i O
  zb = make_float2(e.x + o.y, o.x - e.y);                              // conj(E) + i conj(O)
}

// Fused pass C (+ pointwise work), in three steps separated by __syncwarp():
//   1. forward radix-16 pass on the lane's two rows (registers); every lane but lane 16 of warp 0 then does its 16 pair
//      operations in registers - MONO: store the untangled spectrum pairs of the clip in the CTA's scratch slot
//      (g[k3*512 + tid] = {2X[f], 2X[M-f]}), else: multiply the RIR spectrum by them.  Lane 16 of warp 0 parks its two
//      self-paired rows in shared memory instead;
//   2. lanes 0..15 of warp 0 do one pair operation each on the parked rows (unit u: u = 0 -> f = 0, M, M/2;
//      u = 1..7 -> row (0,0), k3 = u <-> 16 - u; u = 8..15 -> row (0,16), k3 = u - 8 <-> 23 - u; X slots g[u*512 + 16]);
//   3. (not MONO) lane 16 reloads its rows; inverse pass C' on both rows of every lane.
template <bool MONO>
M2H_DEV void conv_pass_c_step1(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, float4* g, float2 (&va)[16], float2 (&vb)[16]) {
  const ConvRows R = conv_rows(tid);
  float4 xs[16];
  if (!MONO && tid != 16) { M2H_UNROLL for (int i = 0; i < 4; ++i) xs[i] = g[i * M2H_CONV_THREADS + tid]; }      // L2 latency hides behind the transforms
  M2H_UNROLL for (int j = 0; j < 16; ++j) { va[j] = s.buf[R.pa + j]; vb[j] = s.buf[R.pb + j]; }
  fft_reg<16, false>(va);
  fft_reg<16, false>(vb);
  // Z_A[k3] = va[bitrev(k3)], Z_B[k3] = vb[bitrev(k3)]
  if (tid != 16) {
    const float2 wU = M2H_CONV_TW(t, 6, tid);                         // W_32768^(k1 + 32*k2) of row A
    M2H_UNROLL for (int k3 = 0; k3 < 16; ++k3) {
      if (!MONO && (k3 & 3) == 0 && k3 + 4 < 16) { M2H_UNROLL for (int i = 0; i < 4; ++i) xs[k3 + 4 + i] = g[(k3 + 4 + i) * M2H_CONV_THREADS + tid]; }
      float2& za = va[bitrev<4>(k3)];
      float2& zb = vb[bitrev<4>(15 - k3)];
      const float2 w = mul_tw32<false>(wU, k3);                       // W_32768^(f), f = k1 + 32*k2 + 1024*k3
      if (MONO) {
        float2 xa, xb;
        conv_untangle(za, zb, w, xa, xb);
        g[k3 * M2H_CONV_THREADS + tid] = make_float4(xa.x, xa.y, xb.x, xb.y);
      } else {
        conv_pair_product(za, zb, w, xs[k3]);
      }
    }
  } else {
    M2H_UNROLL for (int k3 = 0; k3 < 16; ++k3) { s.buf[R.pa + k3] = va[bitrev<4>(k3)]; s.buf[R.pb + k3] = vb[bitrev<4>(k3)]; }
  }
}
template <bool MONO>
M2H_DEV void conv_pass_c_step2(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, float4* g) {
  if (tid >= 16) return;
  const int u = tid;
  float4* gx = g + u * M2H_CONV_THREADS + 16;
  if (u == 0) {           // f = 0 and f = M (real, packed in one complex) and f = M/2
    const int ph = 8;
    const float2 z0 = s.buf[0], zh = s.buf[ph];
    if (MONO) {
      *gx = make_float4(2.f * (z0.x + z0.y), 2.f * (z0.x - z0.y), 2.f * zh.x, -2.f * zh.y);
    } else {
      const float4 x = *gx;
      const float y0 = x.x * (2.f * (z0.x + z0.y));
      const float ym = x.y * (2.f * (z0.x - z0.y));
      s.buf[0] = make_float2(y0 + ym, y0 - ym);
      const float2 y = c_mul(make_float2(x.z, x.w), make_float2(2.f * zh.x, -2.f * zh.y));
      s.buf[ph] = make_float2(2.f * y.x, -2.f * y.y);
    }
    return;
  }
  const int pa = u < 8 ? u : 16 * M2H_CONV_S2 + (u - 8);
  const int pb = u < 8 ? 16 - u : 16 * M2H_CONV_S2 + (23 - u);
  const float2 w = m2h_ldg2(&t->unit_tw[u]);
  float2 za = s.buf[pa], zb = s.buf[pb];
  if (MONO) {
    float2 xa, xb;
    conv_untangle(za, zb, w, xa, xb);
    *gx = make_float4(xa.x, xa.y, xb.x, xb.y);
  } else {
    conv_pair_product(za, zb, w, *gx);
    s.buf[pa] = za;
    s.buf[pb] = zb;
  }
}
M2H_DEV void conv_pass_c_step3(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, float2 (&va)[16], float2 (&vb)[16]) {
  const ConvRows R = conv_rows(tid);
  // inverse pass C' on both rows: natural-order input = the products, output n3 with the twiddle conj W_512^(n3*k2)
  float2 ua[16], ub[16];
  if (tid != 16) {
    M2H_UNROLL for (int k3 = 0; k3 < 16; ++k3) { ua[k3] = va[bitrev<4>(k3)]; ub[k3] = vb[bitrev<4>(k3)]; }
  } else {
    M2H_UNROLL for (int k3 = 0; k3 < 16; ++k3) { ua[k3] = s.buf[R.pa + k3]; ub[k3] = s.buf[R.pb + k3]; }
  }
  fft_reg<16, true>(ua);
  twiddle_powers_out<16, 4, true>(ua, M2H_CONV_TW(t, 4, R.k2a), M2H_CONV_TW(t, 5, R.k2a), [&](int n3, float2 x) { s.buf[R.pa + n3] = x; });
  fft_reg<16, true>(ub);
  twiddle_powers_out<16, 4, true>(ub, M2H_CONV_TW(t, 4, R.k2b), M2H_CONV_TW(t, 5, R.k2b), [&](int n3, float2 x) { s.buf[R.pb + n3] = x; });
}
// The two rows of a thread live in registers across the three steps (host emulation: one static copy per thread id).
#ifdef M2H_HOST_EMU
  #define M2H_CONV_ROWS_DECL static float2 m2h_va[M2H_CONV_THREADS][16], m2h_vb[M2H_CONV_THREADS][16]
  #define M2H_CONV_VA(tid) m2h_va[tid]
  #define M2H_CONV_VB(tid) m2h_vb[tid]
#else
  #define M2H_CONV_ROWS_DECL float2 m2h_va[16], m2h_vb[16]
  #define M2H_CONV_VA(tid) m2h_va
  #define M2H_CONV_VB(tid) m2h_vb
#endif

// Pass A of the forward transform, split into its global loads and the radix pass so that the caller can put
// independent work (the previous window's copy-out) under the load latency.  Thread `tid` owns the residue r = tid of
// the packed index j = n1*512 + r and reads packed samples (x[2j], x[2j+1]); `len` = number of valid real samples.
// VEC: the row is 4-byte (int16) / 8-byte (float) aligned and, for int16, `len` is even, so a packed sample is one load.
// ZU: len <= 16384, i.e. n1 >= 16 is zero padding (not loaded) and the first butterfly stage is pruned.
template <bool VEC, bool ZU>
M2H_DEV void conv_load_mono(float2 (&v)[32], int tid, const int16_t* x, int len) {
  const int r = tid;
  const int lim = len - 2 * r;               // packed sample n1 is valid iff n1*1024 < lim (second half iff n1*1024 + 1 < lim)
  M2H_UNROLL for (int n1 = 0; n1 < (ZU ? 16 : 32); ++n1) {
    float a = 0.f, b = 0.f;
    if (VEC) {
      uint32_t w = 0;
      if (n1 * 1024 < lim) w = m2h_ldg_u32((const uint32_t*)x + n1 * 512 + r);
      a = (float)(int16_t)(w & 0xffffu);
      b = (float)(int16_t)(w >> 16);
    } else {
      if (n1 * 1024 < lim) a = (float)x[n1 * 1024 + 2 * r];
      if (n1 * 1024 + 1 < lim) b = (float)x[n1 * 1024 + 2 * r + 1];
    }
    v[n1] = make_float2(a, b);
  }
}
template <bool VEC, bool ZU>
M2H_DEV void conv_load_rir(float2 (&v)[32], int tid, const float* x, int len) {
  const int r = tid;
  const int lim = len - 2 * r;
  M2H_UNROLL for (int n1 = 0; n1 < (ZU ? 16 : 32); ++n1) {
    float2 a = make_float2(0.f, 0.f);
    if (VEC) {
      if (n1 * 1024 < lim) a = m2h_ldg2((const float2*)x + n1 * 512 + r);
      if (n1 * 1024 + 1 >= lim) a.y = 0.f;
    } else {
      if (n1 * 1024 < lim) a.x = m2h_ldg(x + n1 * 1024 + 2 * r);
      if (n1 * 1024 + 1 < lim) a.y = m2h_ldg(x + n1 * 1024 + 2 * r + 1);
    }
    v[n1] = a;
  }
}
template <bool ZU>
M2H_DEV void conv_fwd_pass_a(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, float2 (&v)[32]) {
  const int r = tid;
  fft_reg<32, false, ZU>(v);
  float2* p = s.buf + (r >> 4) * M2H_CONV_S2 + (r & 15);
  twiddle_powers_out<32, 5, false>(v, M2H_CONV_TW(t, 0, tid), M2H_CONV_TW(t, 1, tid), [&](int k1, float2 x) { p[k1 * M2H_CONV_S1] = x; });    // W_16384^(r*k1)
}
// L2 prefetch of a row that a later phase will load (one 128-byte line per thread and step)
M2H_DEV void conv_prefetch_row(int tid, const void* row, int bytes) {
  for (int o = tid * 128; o < bytes; o += M2H_CONV_THREADS * 128) m2h_prefetch_l2((const char*)row + o);
}

// numpy: np.round (half to even) then astype(int16) wraps modulo 2^16: round-to-nearest-even conversion to int32
// (saturating far outside any reachable amplitude), low 16 bits kept
M2H_DEV uint32_t conv_quantise_pair(float a, float b) {
  return ((uint32_t)m2h_f2i_rn(a) & 0xffffu) | ((uint32_t)m2h_f2i_rn(b) << 16);
}

// Pass A' of the inverse transform fused with scale / round: the packed results j = n1*512 + r that overlap the output
// window (sample indices [c_off, c_off + N)) go to the staging buffer as int16 pairs.
M2H_DEV void conv_inv_pass_a_stage(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, int N, int c_off) {
  const int r = tid;
  const float2* p = s.buf + (r >> 4) * M2H_CONV_S2 + (r & 15);
  float2 v[32];
  M2H_UNROLL for (int k1 = 0; k1 < 32; ++k1) v[k1] = p[k1 * M2H_CONV_S1];
  twiddle_powers_inplace<32, true>(v, M2H_CONV_TW(t, 0, tid), M2H_CONV_TW(t, 1, tid));                                                        // conj W_16384^(r*k1)
  fft_reg<32, true>(v);
  const float scale = 1.0f / (8.0f * 16384.0f);      // 1/M inverse scale, 2*2*2 from the untangle/re-tangle steps
  const int n1_lo = c_off >> 10, n1_hi = (c_off + N - 1) >> 10;
  M2H_UNROLL for (int n1 = 0; n1 < 32; ++n1) {
    if (n1 >= n1_lo && n1 <= n1_hi) {
      const float2 z = v[bitrev<5>(n1)];
      s.stage[n1 * 512 + r] = conv_quantise_pair(z.x * scale, z.y * scale);
    }
  }
}
// Window copy-out: out[m] = staged sample (m + c_off), m < N.  Runs after the barrier that ends pass A' and needs no
// barrier of its own (the staging buffer is next written several phases later).
template <bool VEC>
M2H_DEV void conv_copy_out(const ConvSmem& s, int tid, int16_t* out, int N, int c_off) {
  if (VEC) {                    // N even, row 4-byte aligned: one 32-bit store per sample pair, funnel-shifted when c_off is odd
    uint32_t* o = (uint32_t*)out;
    const int sh = (c_off & 1) * 16;
    for (int w = tid; w < (N >> 1); w += M2H_CONV_THREADS) {
      const int k = (2 * w + c_off) >> 1;
      o[w] = m2h_funnel_r(s.stage[k], s.stage[(k + 1) & (M2H_CONV_M - 1)], sh);
    }
  } else {
    for (int m = tid; m < N; m += M2H_CONV_THREADS) {
      const int q = m + c_off;
      const uint32_t w = s.stage[q >> 1];
      out[m] = (int16_t)(uint16_t)((q & 1) ? (w >> 16) : (w & 0xffffu));
    }
  }
}

// Rows can be read / written with packed accesses when every row start is suitably aligned.
M2H_HD bool conv_can_vectorise(const ConvParams& P) {
  return (P.N % 2 == 0) && (P.Lmax % 2 == 0) && ((uintptr_t)P.mono % 4 == 0) && ((uintptr_t)P.rir % 8 == 0) && ((uintptr_t)P.waves % 4 == 0);
}

// The whole per-CTA loop.  `cta`/`n_ctas` = blockIdx.x / gridDim.x.
template <bool VEC>
M2H_DEV void conv_cta_body(const ConvParams& P, ConvSmem& s, const M2hConvTables* M2H_RESTRICT gtab, int cta, int n_ctas) {
  float4* g = P.scratch + (size_t)cta * M2H_CONV_SCRATCH_F4;
  M2H_PDL_WAIT();
  M2H_CONV_ROWS_DECL;
  const int N = P.N;
  int16_t* pend_out = nullptr;    // window of the previous inverse transform, still in the staging buffer
  int pend_off = 0;
  for (int item = cta; item < P.n_items; item += n_ctas) {
    // device-side caches (ingest.py / m2h_rir_convolve_indexed): items reference pool rows instead of carrying their inputs
    const int mi = P.mono_index ? P.mono_index[item] : item;
    const int ri = P.rir_index ? P.rir_index[item] : item;
    const int L = P.rir_len[ri];
    // taps that can reach the "same" window [c, c+N) of the full convolution: j in [c-N+1, c+N-1]
    const int c = (L - 1) / 2;
    int jmin = c - N + 1 > 0 ? c - N + 1 : 0;
    const int jmax = c + N - 1 < L - 1 ? c + N - 1 : L - 1;
    // VEC: start one tap earlier when that makes the first tap 8-byte aligned (rows are: Lmax is even).  The extra tap
    // j = c - N cannot reach the window, and the window stays alias-free: c_off + N - 1 <= 2N - 1 < F.
    if (VEC && (jmin & 1)) jmin -= 1;
    const int Leff = jmax - jmin + 1;
    const int c_off = c - jmin;
    if (L <= 0) {     // unreadable / empty RIR -> the reference substitutes zeros (simulator_train.py:407-415)
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) {
        int16_t* out = P.waves + (size_t)item * 2 * N;
        for (int i = tid; i < 2 * N; i += M2H_CONV_THREADS) out[i] = 0;
      }
      continue;
    }
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) {
      float2 v[32];
      conv_load_mono<VEC, true>(v, tid, P.mono + (size_t)mi * N, N);                    // N <= 16384 always
      conv_prefetch_row(tid, P.rir + (size_t)ri * 2 * P.Lmax + jmin, Leff * 4);         // ear 0, loaded four phases from now
      if (pend_out) conv_copy_out<VEC>(s, tid, pend_out, N, pend_off);                  // under the load latency
      conv_fwd_pass_a<true>(s, tid, gtab, v);
    }
    pend_out = nullptr;
    M2H_SYNC();
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_b(s, tid, gtab);
    M2H_SYNC_WARP();
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_pass_c_step1<true>(s, tid, gtab, g, M2H_CONV_VA(tid), M2H_CONV_VB(tid));
    M2H_SYNC_WARP();
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_pass_c_step2<true>(s, tid, gtab, g);
    M2H_SYNC();
    for (int ear = 0; ear < 2; ++ear) {
      const float* h = P.rir + ((size_t)ri * 2 + ear) * P.Lmax + jmin;
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) {
        float2 v[32];
        if (Leff <= 16384) {
          conv_load_rir<VEC, true>(v, tid, h, Leff);
          if (pend_out) conv_copy_out<VEC>(s, tid, pend_out, N, pend_off);
          conv_fwd_pass_a<true>(s, tid, gtab, v);
        } else {
          conv_load_rir<VEC, false>(v, tid, h, Leff);
          if (pend_out) conv_copy_out<VEC>(s, tid, pend_out, N, pend_off);
          conv_fwd_pass_a<false>(s, tid, gtab, v);
        }
        // rows the next transforms will load: the other ear, or the next item's clip
        if (ear == 0) conv_prefetch_row(tid, h + P.Lmax, Leff * 4);
        else if (item + n_ctas < P.n_items)
          conv_prefetch_row(tid, P.mono + (size_t)(P.mono_index ? P.mono_index[item + n_ctas] : item + n_ctas) * N, N * 2);
      }
      pend_out = nullptr;
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_b(s, tid, gtab);
      M2H_SYNC_WARP();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_pass_c_step1<false>(s, tid, gtab, g, M2H_CONV_VA(tid), M2H_CONV_VB(tid));
      M2H_SYNC_WARP();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_pass_c_step2<false>(s, tid, gtab, g);
      M2H_SYNC_WARP();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_pass_c_step3(s, tid, gtab, M2H_CONV_VA(tid), M2H_CONV_VB(tid));
      M2H_SYNC_WARP();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_inv_pass_b(s, tid);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_inv_pass_a_stage(s, tid, gtab, N, c_off);
      M2H_SYNC();
      pend_out = P.waves + ((size_t)item * 2 + ear) * N;
      pend_off = c_off;
    }
  }
  if (pend_out) { M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_copy_out<VEC>(s, tid, pend_out, N, pend_off); }
}

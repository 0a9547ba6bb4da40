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
  const int16_t* mono;     // [B, K, N]
  const float* rir;        // [B, K, 2, Lmax]  (ear-planar)
  const int32_t* rir_len;  // [B, K]
  int16_t* waves;          // [B, K, 2, N]   rounded per-source binaural waveform (int16 full-scale units)
  float4* scratch;         // [gridDim.x, 8192]
  int n_items;             // B * K
  int N;
  int Lmax;
};

struct ConvSmem {
  float2 buf[M2H_CONV_BUF];
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

// ---- forward FFT passes B and C (pass A is fused with the loads) ----------------------------
M2H_DEV void conv_fwd_pass_b(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t) {
  const int k1 = tid >> 4, n3 = tid & 15;
  float2* p = s.buf + k1 * M2H_CONV_S1 + n3;
  float2 v[32];
  M2H_UNROLL for (int j = 0; j < 32; ++j) v[j] = p[j * M2H_CONV_S2];
  fft_reg<32, false>(v);
  twiddle_powers_out<32, 5, false>(v, M2H_CONV_TW(t, 2, tid), M2H_CONV_TW(t, 3, tid), [&](int k2, float2 x) { p[k2 * M2H_CONV_S2] = x; });   // W_512^(n3*k2)
}
M2H_DEV void conv_fwd_pass_c(ConvSmem& s, int tid) {
  M2H_UNROLL for (int h = 0; h < 2; ++h) {
    const int q = tid + h * M2H_CONV_THREADS;
    float2* p = s.buf + (q >> 5) * M2H_CONV_S1 + (q & 31) * M2H_CONV_S2;
    float2 v[16];
    M2H_UNROLL for (int j = 0; j < 16; ++j) v[j] = p[j];
    fft_reg<16, false>(v);
    M2H_UNROLL for (int k3 = 0; k3 < 16; ++k3) p[k3] = v[bitrev<4>(k3)];
  }
}
// ---- inverse FFT passes C' and B' (pass A' is fused with the output) ---------------------------
M2H_DEV void conv_inv_pass_c(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t) {
  M2H_UNROLL for (int h = 0; h < 2; ++h) {
    const int q = tid + h * M2H_CONV_THREADS;
    const int k2 = q & 31;
    float2* p = s.buf + (q >> 5) * M2H_CONV_S1 + k2 * M2H_CONV_S2;
    float2 v[16];
    M2H_UNROLL for (int j = 0; j < 16; ++j) v[j] = p[j];
    fft_reg<16, true>(v);
    twiddle_powers_out<16, 4, true>(v, M2H_CONV_TW(t, 4, tid), M2H_CONV_TW(t, 5, tid), [&](int n3, float2 x) { p[n3] = x; });                 // conj W_512^(n3*k2)
  }
}
// pass B' stores untwiddled; conj W_16384^((16*n2+n3)*k1) = conj(W_16384^r)^k1 is applied by pass A' on load
M2H_DEV void conv_inv_pass_b(ConvSmem& s, int tid) {
  const int k1 = tid >> 4, n3 = tid & 15;
  float2* p = s.buf + k1 * M2H_CONV_S1 + n3;
  float2 v[32];
  M2H_UNROLL for (int j = 0; j < 32; ++j) v[j] = p[j * M2H_CONV_S2];
  fft_reg<32, true>(v);
  M2H_UNROLL for (int n2 = 0; n2 < 32; ++n2) p[n2 * M2H_CONV_S2] = v[bitrev<5>(n2)];
}

// Pass A of the forward transform with the global loads fused in.  Thread `tid` owns the residue
// r = tid of the packed index j = n1*512 + r and reads packed samples (x[2j], x[2j+1]).
//   SRC = 0: int16 mono, SRC = 1: float RIR taps.  `len` = number of valid real samples.
template <int SRC>
M2H_DEV void conv_fwd_pass_a(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, const void* src, int len) {
  const int r = tid;
  float2 v[32];
  M2H_UNROLL for (int n1 = 0; n1 < 32; ++n1) {
    const int i0 = 2 * (n1 * 512 + r);
    float a = 0.f, b = 0.f;
    if (SRC == 0) {
      const int16_t* x = (const int16_t*)src;
      if (i0 < len) a = (float)x[i0];
      if (i0 + 1 < len) b = (float)x[i0 + 1];
    } else {
      const float* x = (const float*)src;
      if (i0 < len) a = m2h_ldg(x + i0);
      if (i0 + 1 < len) b = m2h_ldg(x + i0 + 1);
    }
    v[n1] = make_float2(a, b);
  }
  fft_reg<32, false>(v);
  float2* p = s.buf + (r >> 4) * M2H_CONV_S2 + (r & 15);
  twiddle_powers_out<32, 5, false>(v, M2H_CONV_TW(t, 0, tid), M2H_CONV_TW(t, 1, tid), [&](int k1, float2 x) { p[k1 * M2H_CONV_S1] = x; });    // W_16384^(r*k1)
}

// Untangle the packed-real spectrum at (f, M-f):  returns 2*X[f] in xa and 2*X[M-f] in xb.
M2H_DEV void conv_untangle(float2 za, float2 zb, float2 w, float2& xa, float2& xb) {
  const float2 p = make_float2(za.x + zb.x, za.y - zb.y);       // Za + conj(Zb)         = 2 E[f]
  const float2 o = make_float2(za.y + zb.y, zb.x - za.x);       // -i (Za - conj(Zb))    = 2 O[f]
  const float2 q = c_mul(w, o);
  xa = c_add(p, q);
  xb = c_conj(c_sub(p, q));
}

// Phase: spectrum of the mono signal -> per-CTA scratch (float4 = {2X[f], 2X[M-f]}).
M2H_DEV void conv_store_mono_spectrum(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, float4* g) {
  const float2 wU = M2H_CONV_TW(t, 6, tid);
  M2H_UNROLL for (int i = 0; i < 8192 / M2H_CONV_THREADS; ++i) {
    const int idx = tid + i * M2H_CONV_THREADS;
    if (idx == 0) {
      const float2 z0 = s.buf[0], zh = s.buf[conv_pos(8192)];
      g[0] = make_float4(2.f * (z0.x + z0.y), 2.f * (z0.x - z0.y), 2.f * zh.x, -2.f * zh.y);   // X[0], X[M] (real), X[M/2]
    } else {
      const int f = conv_pair_f(idx);
      float2 xa, xb;
      const float2 w = c_mul(wU, make_float2(conv_w16384_cos(i), conv_w16384_sin(i)));   // W_32768^(f0 + 2i)
      conv_untangle(s.buf[conv_pos(f)], s.buf[conv_pos(16384 - f)], w, xa, xb);
      g[idx] = make_float4(xa.x, xa.y, xb.x, xb.y);
    }
  }
}

// Phase: untangle H, multiply by X, re-tangle for the inverse transform (in place).
M2H_DEV void conv_multiply_spectrum(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, const float4* g) {
  const float2 wU = M2H_CONV_TW(t, 6, tid);
  M2H_UNROLL for (int i = 0; i < 8192 / M2H_CONV_THREADS; ++i) {
    const int idx = tid + i * M2H_CONV_THREADS;
    const float4 x = g[idx];
    if (idx == 0) {
      const int ph = conv_pos(8192);
      const float2 z0 = s.buf[0], zh = s.buf[ph];
      const float y0 = x.x * (2.f * (z0.x + z0.y));
      const float ym = x.y * (2.f * (z0.x - z0.y));
      s.buf[0] = make_float2(y0 + ym, y0 - ym);
      const float2 y = c_mul(make_float2(x.z, x.w), make_float2(2.f * zh.x, -2.f * zh.y));
      s.buf[ph] = make_float2(2.f * y.x, -2.f * y.y);
    } else {
      const int f = conv_pair_f(idx);
      const int pa = conv_pos(f), pb = conv_pos(16384 - f);
      const float2 w = c_mul(wU, make_float2(conv_w16384_cos(i), conv_w16384_sin(i)));
      float2 ha, hb;
      conv_untangle(s.buf[pa], s.buf[pb], w, ha, hb);
      const float2 ya = c_mul(make_float2(x.x, x.y), ha);
      const float2 yb = c_mul(make_float2(x.z, x.w), hb);
      const float2 e = make_float2(ya.x + yb.x, ya.y - yb.y);            // Ya + conj(Yb)
      const float2 t = make_float2(ya.x - yb.x, ya.y + yb.y);            // Ya - conj(Yb)
      const float2 o = c_mulc(t, w);                                       // conj(w) * t
      s.buf[pa] = make_float2(e.x - o.y, e.y + o.x);                       // E + i O
      s.buf[pb] = make_float2(e.x + o.y, o.x - e.y);                       // conj(E) + i conj(O)
    }
  }
}

// numpy: np.round (half to even) then astype(int16) wraps modulo 2^16
M2H_DEV int16_t conv_quantise(float v) {
  const float r = m2h_rintf(v);
  const float lim = 2147483520.f;
  const int32_t i = (int32_t)(r > lim ? lim : (r < -lim ? -lim : r));
  return (int16_t)(uint16_t)(uint32_t)i;
}

// Pass A' of the inverse transform fused with scale / round / window / store.
M2H_DEV void conv_inv_pass_a_store(ConvSmem& s, int tid, const M2hConvTables* M2H_RESTRICT t, int16_t* out, int N, int c_off) {
  const int r = tid;
  const float2* p = s.buf + (r >> 4) * M2H_CONV_S2 + (r & 15);
  float2 v[32];
  M2H_UNROLL for (int k1 = 0; k1 < 32; ++k1) v[k1] = p[k1 * M2H_CONV_S1];
  twiddle_powers_inplace<32, true>(v, M2H_CONV_TW(t, 0, tid), M2H_CONV_TW(t, 1, tid));                                                        // conj W_16384^(r*k1)
  fft_reg<32, true>(v);
  const float scale = 1.0f / (8.0f * 16384.0f);      // 1/M inverse scale, 2*2*2 from the untangle/re-tangle steps
  M2H_UNROLL for (int n1 = 0; n1 < 32; ++n1) {
    const float2 z = v[bitrev<5>(n1)];
    const int m = 2 * (n1 * 512 + r) - c_off;        // output sample index of the real part
    if (m >= 0 && m < N) out[m] = conv_quantise(z.x * scale);
    if (m + 1 >= 0 && m + 1 < N) out[m + 1] = conv_quantise(z.y * scale);
  }
}

// The whole per-CTA loop.  `cta`/`n_ctas` = blockIdx.x / gridDim.x.
M2H_DEV void conv_cta_body(const ConvParams& P, ConvSmem& s, const M2hConvTables* M2H_RESTRICT gtab, int cta, int n_ctas) {
  float4* g = P.scratch + (size_t)cta * M2H_CONV_SCRATCH_F4;
  for (int item = cta; item < P.n_items; item += n_ctas) {
    const int N = P.N;
    const int L = P.rir_len[item];
    // taps that can reach the "same" window [c, c+N) of the full convolution: j in [c-N+1, c+N-1]
    const int c = (L - 1) / 2;
    const int jmin = c - N + 1 > 0 ? c - N + 1 : 0;
    const int jmax = c + N - 1 < L - 1 ? c + N - 1 : L - 1;
    const int Leff = jmax - jmin + 1;
    const int c_off = c - jmin;
    if (L <= 0) {     // unreadable / empty RIR -> the reference substitutes zeros (simulator_train.py:407-415)
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) {
        int16_t* out = P.waves + (size_t)item * 2 * N;
        for (int i = tid; i < 2 * N; i += M2H_CONV_THREADS) out[i] = 0;
      }
      continue;
    }
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_a<0>(s, tid, gtab, P.mono + (size_t)item * N, N);
    M2H_SYNC();
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_b(s, tid, gtab);
    M2H_SYNC();
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_c(s, tid);
    M2H_SYNC();
    M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_store_mono_spectrum(s, tid, gtab, g);
    M2H_SYNC();
    for (int ear = 0; ear < 2; ++ear) {
      const float* h = P.rir + ((size_t)item * 2 + ear) * P.Lmax + jmin;
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_a<1>(s, tid, gtab, h, Leff);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_b(s, tid, gtab);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_fwd_pass_c(s, tid);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_multiply_spectrum(s, tid, gtab, g);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_inv_pass_c(s, tid, gtab);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS) conv_inv_pass_b(s, tid);
      M2H_SYNC();
      M2H_FOR_THREADS(tid, M2H_CONV_THREADS)
        conv_inv_pass_a_store(s, tid, gtab, P.waves + ((size_t)item * 2 + ear) * N, N, c_off);
      M2H_SYNC();
    }
  }
}

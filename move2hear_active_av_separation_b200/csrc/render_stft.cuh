// Framed STFT (n_fft = 1023, hop = 512, periodic Hann, centre/reflect) -> magnitude / phase.
//
// Replaces librosa.stft + librosa.magphase (+ np.angle, np.log1p, fp16 casts) at
// habitat_audio/simulator_train.py:425-441, 473-486, habitat_audio/simulator_eval.py:437-500 and
// audio_separation/pretrain/datasets/dataset.py:190-228.
//
// 1023 = 3*11*31 has no radix-2 factor: each frame's 1023-point DFT is evaluated with Bluestein's chirp-z
// identity as a 2048-point circular convolution.  The two ears of a frame ride together as one complex signal
// (left + i*right) and are separated by conjugate symmetry afterwards, so one frame pair costs two complex
// FFT-2048s.  A CTA = M2H_STFT_GROUPS groups of 128 threads; each group transforms one frame pair at a time in
// a padded [16][16][8] shared buffer (position(i1,i2,i3) = i1*136 + i2*8 + (i2>>1) + i3, conflict-free for all passes).
//
// Signals: s = 0 is the mixture ((sum_k wave_k) * 2^-15 / K, exactly the fp32 arithmetic of
// simulator_train.py:469-471), s >= 1 is source s-1 alone (wave * 2^-15).  A generic fp32 input mode serves the
// cached monaural GT spectrograms (simulator_train.py:443-449, dataset.py:201-207).
#pragma once
#include "fft_radix.cuh"
#ifndef M2H_HOST_EMU
#include <cuda_bf16.h>
#endif

#define M2H_STFT_GROUPS 2
#define M2H_STFT_THREADS (128 * M2H_STFT_GROUPS)
#define M2H_STFT_CTAS_PER_SM 3         // 768 threads per SM at 80 registers: no spills (two 512-thread CTAs at 64 registers spilled)
// padded [16][16][8] layout: position(i1,i2,i3) = i1*136 + i2*8 + (i2 >> 1) + i3  (one pad slot per 16 elements and
// 8 per i1 block): every pass's half-warp touches 16 distinct bank pairs
#define M2H_STFT_S1 136
#define M2H_STFT_BUF (16 * M2H_STFT_S1)
#define M2H_STFT_POS2(i2) ((i2) * 8 + ((i2) >> 1))
#define M2H_STFT_X0_LD 34             // halfwords per row of the x0 staging tile: 17 words, so a warp's 32 rows hit 32 banks

enum {
  M2H_OUT_MIX_LOG1P = 1,   // mixture magnitude is written as log1p(mag)
  M2H_OUT_F16 = 2,         // simulator dtype: magnitude cast to fp16 BEFORE log1p, everything stored as fp16
};

struct StftParams {
  // input A: quantised per-source waves
  const int16_t* waves;    // [B, K, 2, N] or null
  // input B: generic fp32 waveforms [B, C, N] (C = 1 or 2), used when waves == null
  const float* fwave;
  int fwave_channels;
  int B, K, N, T;          // T = 1 + (N - 1) / 512 frames (librosa centre framing with the odd n_fft)
  int n_signals;           // waves mode: 1 (mixture only) .. 1 + n_gt ; fwave mode: 1
  int sig0;                // first signal of the launch: 0, or 1 = the GT spectrograms alone (the mixture comes from another launch)
  int n_gt;                // number of leading sources that get their own GT spectrogram (0, 1 or K)
  int flags;
  // outputs (null = not wanted).  All [.., 512, T, C] with C innermost like the reference's np.stack(axis=-1).
  void* mix_mag;           // [B, 512, T, 2]   fp32 or fp16
  void* mix_phase;         // [B, 512, T, 2]
  void* gt_mag;            // dataset layout: [B, n_gt, 512, T, 2] fp32 ; simulator layout: [B, n_gt, 512, T, 4] fp16
  int gt_interleaved_phase;// 1: simulator layout (mag_l, phase_l, mag_r, phase_r)
  // the mixture's log1p magnitude as the separator's first-layer operand: bf16 [B][18][18][128], the space-to-depth, haloed
  // NHWC form of the [B,512,32,2] observation that pack_input_kernel (m2h_unet.cu) otherwise builds from mix_mag - cell
  // (h/2 + 1, t/2 + 1), channel ((h&1)*2 + (t&1))*32 + ear*16 + slice with h = bin % 32, slice = bin / 32
  // (separator_cnn.py:84-90).  Needs T == 32; interior cells only (the zero border is the buffer owner's).
  void* x0;
  void* f_mag;             // fwave mode: [B, 512, T, C] fp32
  void* f_phase;           // fwave mode: [B, 512, T, C] fp32 or null
  // inverse transform (m2h_istft_kernel): magnitude / phase [B, 512, T, C] fp32 -> waveforms [B, C, i_length] (pre-zeroed)
  const float* i_mag;
  const float* i_phase;
  float* i_out;
  int i_channels, i_length;
};

struct StftSmem {
  float2 buf[M2H_STFT_GROUPS][M2H_STFT_BUF];
  float2 bf[2048];          // chirp spectrum, [k3][k1*16+k2] so that consecutive threads read consecutive entries
  uint16_t x0st[M2H_STFT_GROUPS][32 * M2H_STFT_X0_LD];     // one frame's bf16 cells [h][ear*16 + slice] on their way to P.x0
};
// The samples [512*t - 512, 512*t + 512) of a frame (reflection at either end of the signal maps back inside this window) are
// staged in the group's own FFT buffer before the first pass: sample j of the window lives at STFT_STAGE_POS(j); the
// two pad slots per 8 samples keep the 16-byte staging stores of neighbouring threads on distinct banks.
#define M2H_STFT_STAGE_POS(j) ((j) + (((j) >> 3) << 1))
#define M2H_STFT_TW(t, which, tid) m2h_ldg2(&(t)->thr[which][(tid) & 127])

M2H_DEV int stft_reflect(int i, int N) { return i < 0 ? -i : (i >= N ? 2 * (N - 1) - i : i); }

// fp32 value of an integer sample sum: the fp32 sum of x*2^-15 terms is exact for K <= 256; the divide is the reference's
// float32 divide by K (simulator_train.py:471) - a multiplication by the exact reciprocal when K is a power of two
M2H_DEV float stft_scale(int v, int K, bool mix) {
  const float x = (float)v * (1.0f / 32768.0f);
  if (!mix) return x;
  return (K & (K - 1)) == 0 ? x * (1.0f / (float)K) : x / (float)K;
}
// one sample of signal `sig` for item b, both ears
M2H_DEV float2 stft_sample(const StftParams& P, int b, int sig, int i) {
  if (P.waves) {
    const int16_t* w = P.waves + (size_t)b * P.K * 2 * P.N;
    if (sig == 0) {
      int sl = 0, sr = 0;
      for (int k = 0; k < P.K; ++k) {
        sl += w[(size_t)(2 * k) * P.N + i];
        sr += w[(size_t)(2 * k + 1) * P.N + i];
      }
      return make_float2(stft_scale(sl, P.K, true), stft_scale(sr, P.K, true));
    }
    const int k = sig - 1;
    return make_float2(stft_scale(w[(size_t)(2 * k) * P.N + i], 1, false), stft_scale(w[(size_t)(2 * k + 1) * P.N + i], 1, false));
  }
  const float* f = P.fwave + (size_t)b * P.fwave_channels * P.N;
  return make_float2(f[i], P.fwave_channels > 1 ? f[P.N + i] : 0.f);
}

struct StftI16x8 { uint32_t w[4]; };
M2H_DEV void stft_acc8(int (&acc)[8], const int16_t* p) {
  const StftI16x8 q = *reinterpret_cast<const StftI16x8*>(p);       // 16-byte aligned by construction
  M2H_UNROLL for (int u = 0; u < 4; ++u) {
    acc[2 * u] += (int)(int16_t)(q.w[u] & 0xffffu);
    acc[2 * u + 1] += (int)(int16_t)(q.w[u] >> 16);
  }
}
// P0: each 128-thread group stages the 1024-sample window of its own frame (the mixture is formed here, in exactly the
// reference's fp32 arithmetic).  Rows that allow it are read as 8 x int16 vectors: one vector per thread and row.
M2H_DEV void stft_stage(const StftParams& P, StftSmem& s, int tid, int b, int sig, int t0) {
  const int grp = tid >> 7, lt = tid & 127, t = t0 + grp;
  if (t >= P.T) return;
  const int i0 = M2H_HOP * t - M2H_HOP;
  float2* st = s.buf[grp];
  if (P.waves && (P.N & 7) == 0 && ((uintptr_t)P.waves & 15) == 0) {
    const int i = i0 + 8 * lt;
    if (i >= 0 && i < P.N) {
      const int16_t* w = P.waves + (size_t)b * P.K * 2 * P.N + i;
      int al[8], ar[8];
      M2H_UNROLL for (int u = 0; u < 8; ++u) { al[u] = 0; ar[u] = 0; }
      const int k_lo = sig == 0 ? 0 : sig - 1, k_hi = sig == 0 ? P.K : sig;
      for (int k = k_lo; k < k_hi; ++k) {
        stft_acc8(al, w + (size_t)(2 * k) * P.N);
        stft_acc8(ar, w + (size_t)(2 * k + 1) * P.N);
      }
      float4* dst = reinterpret_cast<float4*>(st + M2H_STFT_STAGE_POS(8 * lt));      // 10 float2 per 8 samples: 16-byte aligned
      M2H_UNROLL for (int u = 0; u < 8; u += 2)
        dst[u >> 1] = make_float4(stft_scale(al[u], P.K, sig == 0), stft_scale(ar[u], P.K, sig == 0),
                                  stft_scale(al[u + 1], P.K, sig == 0), stft_scale(ar[u + 1], P.K, sig == 0));
    }
  } else {
    for (int j = lt; j < 2 * M2H_HOP; j += 128) {
      const int i = i0 + j;
      if (i >= 0 && i < P.N) st[M2H_STFT_STAGE_POS(j)] = stft_sample(P, b, sig, i);
    }
  }
}

#ifndef M2H_HOST_EMU
M2H_DEV float stft_log1p_f16(float mag) {    // np.log1p on a float16 array: computed in fp32, rounded to fp16
  return __half2float(__float2half_rn(log1pf(__half2float(__float2half_rn(mag)))));
}
// two adjacent outputs (left, right ear) as one store; idx is even and the buffers are 8-byte aligned (checked on the host)
template <bool F16> M2H_DEV void stft_store2(void* base, size_t idx, float a, float b) {
  if (F16) *reinterpret_cast<__half2*>((__half*)base + idx) = __floats2half2_rn(a, b);
  else *reinterpret_cast<float2*>((float*)base + idx) = make_float2(a, b);
}
// fp32 outputs (tolerance 1e-4 normwise): |x| from the approximate square root (2 ulp) and log1p(m) = log(u) - ((u-1) - m)/u,
// u = fl(1+m), with the hardware logarithm (absolute error ~5e-7).  The fp16 observation keeps the exact library functions:
// it is compared bit for bit.
template <bool F16> M2H_DEV float stft_abs(float2 x) {
  const float q = x.x * x.x + x.y * x.y;
  if (F16) return sqrtf(q);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
  return r;
}
M2H_DEV float stft_log1p_fast(float m) {
  const float u = 1.0f + m;
  return __logf(u) - __fdividef((u - 1.0f) - m, u);
}
#else
float m2h_emu_round_f16(float v);            // provided by the emulation harness (numpy-free fp16 rounding)
M2H_DEV float stft_log1p_f16(float mag) { return m2h_emu_round_f16(log1pf(m2h_emu_round_f16(mag))); }
void m2h_emu_store(void* base, size_t idx, float v, bool f16);
template <bool F16> M2H_DEV void stft_store2(void* base, size_t idx, float a, float b) { m2h_emu_store(base, idx, a, F16); m2h_emu_store(base, idx + 1, b, F16); }
template <bool F16> M2H_DEV float stft_abs(float2 x) { return sqrtf(x.x * x.x + x.y * x.y); }
M2H_DEV float stft_log1p_fast(float m) { return log1pf(m); }
#endif

template <bool F16>
M2H_DEV void stft_emit(const StftParams& P, int b, int sig, int k, int t, float2 xl, float2 xr) {
  const float ml = stft_abs<F16>(xl), mr = stft_abs<F16>(xr);
  const size_t bin = (size_t)k * P.T + t;
  if (!P.waves) {
    const int C = P.fwave_channels;
    const size_t o = ((size_t)b * M2H_NBINS * P.T + bin) * C;
    ((float*)P.f_mag)[o] = ml;
    if (C > 1) ((float*)P.f_mag)[o + 1] = mr;
    if (P.f_phase) {
      ((float*)P.f_phase)[o] = atan2f(xl.y, xl.x);
      if (C > 1) ((float*)P.f_phase)[o + 1] = atan2f(xr.y, xr.x);
    }
    return;
  }
  if (sig == 0) {
    const size_t o = ((size_t)b * M2H_NBINS * P.T + bin) * 2;
    if (P.mix_mag) {
      float vl = ml, vr = mr;
      if (P.flags & M2H_OUT_MIX_LOG1P) {
        vl = F16 ? stft_log1p_f16(ml) : stft_log1p_fast(ml);
        vr = F16 ? stft_log1p_f16(mr) : stft_log1p_fast(mr);
      }
      stft_store2<F16>(P.mix_mag, o, vl, vr);
    }
    if (P.mix_phase) stft_store2<F16>(P.mix_phase, o, atan2f(xl.y, xl.x), atan2f(xr.y, xr.x));
  } else if (P.gt_mag) {
    const int g = sig - 1;
    if (P.gt_interleaved_phase) {
      const size_t o = (((size_t)b * P.n_gt + g) * M2H_NBINS * P.T + bin) * 4;
      stft_store2<F16>(P.gt_mag, o, ml, atan2f(xl.y, xl.x));
      stft_store2<F16>(P.gt_mag, o + 2, mr, atan2f(xr.y, xr.x));
    } else {
      const size_t o = (((size_t)b * P.n_gt + g) * M2H_NBINS * P.T + bin) * 2;
      stft_store2<F16>(P.gt_mag, o, ml, mr);
    }
  }
}

// ---- phases (per thread; grp = tid >> 7 owns frame t, lt = tid & 127) ------------------------------
// P1: load 8 x 128 samples of the frame pair from the staged window, window * conj(chirp) (P1a); after a group barrier -
// the window lives in the buffer the pass writes - radix-16 pass along n1 (upper 8 inputs are the zero padding of the
// 2048-point transform), twiddle W_2048^(r*k1) (P1b).
M2H_DEV void stft_p1_load(const StftParams& P, const StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int tid, int t0, float2 (&v)[16]) {
  const int grp = tid >> 7, r = tid & 127, t = t0 + grp;
  if (t >= P.T) return;
  M2H_UNROLL for (int n1 = 0; n1 < 8; ++n1) {
    const int n = n1 * 128 + r;
    float2 a = make_float2(0.f, 0.f);
    if (n < M2H_NFFT) {
      const int j = stft_reflect(M2H_HOP * t - (M2H_NFFT / 2) + n, P.N) - (M2H_HOP * t - M2H_HOP);
      a = c_mul(s.buf[grp][M2H_STFT_STAGE_POS(j)], gt->wc[n]);
    }
    v[n1] = a;
  }
}
M2H_DEV void stft_p1_pass(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int tid, int t0, float2 (&v)[16]) {
  const int grp = tid >> 7, r = tid & 127;
  if (t0 + grp >= P.T) return;
  fft_reg<16, false, true>(v);
  float2* p = s.buf[grp] + M2H_STFT_POS2(r >> 3) + (r & 7);
  twiddle_powers_out<16, 4, false>(v, M2H_STFT_TW(gt, 0, tid), M2H_STFT_TW(gt, 1, tid), [&](int k1, float2 x) { p[k1 * M2H_STFT_S1] = x; });     // W_2048^(r*k1)
}
// P2: radix-16 pass along n2, twiddle W_128^(n3*k2) = W_2048^(16*n3*k2)
M2H_DEV void stft_p2(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int tid, int t0) {
  const int grp = tid >> 7, lt = tid & 127;
  if (t0 + grp >= P.T) return;
  const int k1 = lt >> 3, n3 = lt & 7;
  float2* p = s.buf[grp] + k1 * M2H_STFT_S1 + n3;
  float2 v[16];
  M2H_UNROLL for (int j = 0; j < 16; ++j) v[j] = p[M2H_STFT_POS2(j)];
  fft_reg<16, false>(v);
  twiddle_powers_out<16, 4, false>(v, M2H_STFT_TW(gt, 2, tid), M2H_STFT_TW(gt, 3, tid), [&](int k2, float2 x) { p[M2H_STFT_POS2(k2)] = x; });    // W_128^(n3*k2)
}
// P3: radix-8 pass along n3, multiply by the chirp spectrum, inverse radix-8 pass, twiddle conj(W_128^(n3*k2))
M2H_DEV void stft_p3(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int tid, int t0) {
  const int grp = tid >> 7, lt = tid & 127;
  if (t0 + grp >= P.T) return;
  M2H_UNROLL for (int h = 0; h < 2; ++h) {
    const int q = lt + h * 128;                 // q = k1*16 + k2
    const int k2 = q & 15;
    float2* p = s.buf[grp] + (q >> 4) * M2H_STFT_S1 + M2H_STFT_POS2(k2);
    float2 v[8], u[8];
    M2H_UNROLL for (int j = 0; j < 8; ++j) v[j] = p[j];
    fft_reg<8, false>(v);
    M2H_UNROLL for (int k3 = 0; k3 < 8; ++k3) u[k3] = c_mul(v[bitrev<3>(k3)], s.bf[k3 * 256 + q]);
    fft_reg<8, true>(u);
    twiddle_powers_out<8, 3, true>(u, M2H_STFT_TW(gt, 4, tid), M2H_STFT_TW(gt, 5, tid), [&](int n3, float2 x) { p[n3] = x; });                   // conj W_128^(n3*k2)
  }
}
// P4: inverse radix-16 pass along k2 -> n2 (its twiddle conj(W_2048^(r*k1)), r = 8*n2+n3, is applied by P5)
M2H_DEV void stft_p4(const StftParams& P, StftSmem& s, int tid, int t0) {
  const int grp = tid >> 7, lt = tid & 127;
  if (t0 + grp >= P.T) return;
  const int k1 = lt >> 3, n3 = lt & 7;
  float2* p = s.buf[grp] + k1 * M2H_STFT_S1 + n3;
  float2 v[16];
  M2H_UNROLL for (int j = 0; j < 16; ++j) v[j] = p[M2H_STFT_POS2(j)];
  fft_reg<16, true>(v);
  M2H_UNROLL for (int n2 = 0; n2 < 16; ++n2) p[M2H_STFT_POS2(n2)] = v[bitrev<4>(n2)];     // twiddle applied by P5 on load
}
// P5: inverse radix-16 pass along k1 -> n1 (only m = n1*128 + r < 1023 is needed), times conj(chirp)/4096
M2H_DEV void stft_p5(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int tid, int t0) {
  const int grp = tid >> 7, r = tid & 127;
  if (t0 + grp >= P.T) return;
  float2* p = s.buf[grp] + M2H_STFT_POS2(r >> 3) + (r & 7);
  float2 v[16];
  M2H_UNROLL for (int k1 = 0; k1 < 16; ++k1) v[k1] = p[k1 * M2H_STFT_S1];
  twiddle_powers_inplace<16, true>(v, M2H_STFT_TW(gt, 0, tid), M2H_STFT_TW(gt, 1, tid));                                                          // conj W_2048^(r*k1)
  fft_reg<16, true>(v);
  M2H_UNROLL for (int n1 = 0; n1 < 8; ++n1)
    p[n1 * M2H_STFT_S1] = c_mul(v[bitrev<4>(n1)], gt->cc[n1 * 128 + r]);
}
M2H_DEV int stft_pos(int m) { return (m >> 7) * M2H_STFT_S1 + M2H_STFT_POS2((m >> 3) & 15) + (m & 7); }
// P6: separate the ears (X_l = (X[k] + conj X[N-k]) / 2, X_r = (X[k] - conj X[N-k]) / 2i; the 1/2 lives in cc) and emit
template <bool F16>
M2H_DEV void stft_p6(const StftParams& P, StftSmem& s, int tid, int b, int sig, int t0) {
  const int grp = tid >> 7, lt = tid & 127, t = t0 + grp;
  if (t >= P.T) return;
  for (int k = lt; k < M2H_NBINS; k += 128) {
    const float2 a = s.buf[grp][stft_pos(k)];
    const float2 c = s.buf[grp][stft_pos(k == 0 ? 0 : M2H_NFFT - k)];
    const float2 xl = make_float2(a.x + c.x, a.y - c.y);
    const float2 xr = make_float2(a.y + c.y, c.x - a.x);
#ifndef M2H_HOST_EMU
    if (!F16 && P.x0 != nullptr && sig == 0) {       // bf16(log1p |X|): exactly what pack_input_kernel makes of the fp32 observation
      uint16_t* st = s.x0st[grp] + (k & 31) * M2H_STFT_X0_LD + (k >> 5);
      st[0] = __bfloat16_as_ushort(__float2bfloat16_rn(stft_log1p_fast(stft_abs<false>(xl))));
      st[16] = __bfloat16_as_ushort(__float2bfloat16_rn(stft_log1p_fast(stft_abs<false>(xr))));
      if (P.mix_mag == nullptr && P.mix_phase == nullptr) continue;
    }
#endif
    stft_emit<F16>(P, b, sig, k, t, xl, xr);
  }
}
#ifndef M2H_HOST_EMU
// after the group barrier that follows P6: 128 threads move the frame's 32 x 64-byte cells, one 16-byte store each
M2H_DEV void stft_x0_copy_out(const StftParams& P, StftSmem& s, int tid, int b, int t0) {
  const int grp = tid >> 7, lt = tid & 127, t = t0 + grp;
  if (t >= P.T) return;
  const int h = lt >> 2, part = lt & 3;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(s.x0st[grp]) + h * (M2H_STFT_X0_LD / 2) + part * 4;
  const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
  uint16_t* dst = reinterpret_cast<uint16_t*>(P.x0) + ((((size_t)b * 18 + (h >> 1) + 1) * 18 + (t >> 1) + 1) * 128 + ((h & 1) * 2 + (t & 1)) * 32 + part * 8);
  *reinterpret_cast<uint4*>(dst) = v;
}
#endif

template <bool F16>     // F16 = simulator dtype (P.flags & M2H_OUT_F16)
M2H_DEV void stft_cta_body(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int cta, int n_ctas) {
  M2H_FOR_THREADS(tid, M2H_STFT_THREADS) {
    for (int i = tid; i < 2048; i += M2H_STFT_THREADS) s.bf[(i & 7) * 256 + (i >> 3)] = gt->bf[i];
  }
  M2H_PDL_WAIT();       // the table staging above overlaps the tail of the convolution kernel
  M2H_SYNC();
  // work item = (mixture, signal, M2H_STFT_GROUPS frames): small batches (the 14 environments of a rollout rank) still fill the GPU
  const int n_fgroups = (P.T + M2H_STFT_GROUPS - 1) / M2H_STFT_GROUPS;
  const int n_items = P.B * P.n_signals * n_fgroups;
  for (int item = cta; item < n_items; item += n_ctas) {
    const int bs = item / n_fgroups, t0 = (item - bs * n_fgroups) * M2H_STFT_GROUPS;
    const int b = bs / P.n_signals, sig = P.sig0 + bs - b * P.n_signals;
    {
      // each 128-thread group owns one frame pair and its buffer: only the group meets between the passes, the four
      // groups of a CTA drift apart freely
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_stage(P, s, tid, b, sig, t0);
      M2H_SYNC_GROUP128();
#ifdef M2H_HOST_EMU
      static float2 vs[M2H_STFT_THREADS][16];
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p1_load(P, s, gt, tid, t0, vs[tid]);
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p1_pass(P, s, gt, tid, t0, vs[tid]);
#else
      {
        float2 v[16];
        stft_p1_load(P, s, gt, threadIdx.x, t0, v);
        M2H_SYNC_GROUP128();
        stft_p1_pass(P, s, gt, threadIdx.x, t0, v);
      }
#endif
      M2H_SYNC_GROUP128();
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p2(P, s, gt, tid, t0);
      M2H_SYNC_GROUP128();
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p3(P, s, gt, tid, t0);
      M2H_SYNC_GROUP128();
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p4(P, s, tid, t0);
      M2H_SYNC_GROUP128();
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p5(P, s, gt, tid, t0);
      M2H_SYNC_GROUP128();
      M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p6<F16>(P, s, tid, b, sig, t0);
      M2H_SYNC_GROUP128();      // the next window is staged into the buffer P6 reads
#ifndef M2H_HOST_EMU
      if (!F16 && P.x0 != nullptr && sig == 0) stft_x0_copy_out(P, s, threadIdx.x, b, t0);    // the tile is rewritten five group barriers later
#endif
    }
  }
}


// ================================================================================================
// Inverse STFT as the reference's evaluation calls it: librosa.istft(spec, hop_length=512, length=16000)
// (audio_separation/common/eval_metrics.py:232-252).  librosa infers n_fft = 2*(512-1) = 1022 from the 512 bins - not the
// 1023 of the forward transform - so a frame is the 1022-point inverse real DFT of its column (imaginary parts of bins 0
// and 511 ignored), times the periodic Hann window of 1022, overlap-added at 512*t, divided by the window's sum of
// squares, with the first 511 samples dropped.  1022 = 2*7*73: the inverse DFT runs through the same Bluestein machinery
// (2048-point circular convolution, passes P2-P5 above with the tables of N = 1022, inverse); the two ears ride as
// real and imaginary part of one complex signal.
// ================================================================================================
#define M2H_ISTFT_N 1022

// P1 of the inverse: Z[k] = X_l[k] + i X_r[k] with the Hermitian extension X[1022-k] = conj(X[k]), times the chirp, radix-16
// pass along n1 (upper half of the 2048 points is zero padding)
M2H_DEV void istft_p1(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int tid, int b, int t0) {
  const int grp = tid >> 7, r = tid & 127, t = t0 + grp;
  if (t >= P.T) return;
  const int C = P.i_channels;
  float2 v[16];
  M2H_UNROLL for (int n1 = 0; n1 < 8; ++n1) {
    const int k = n1 * 128 + r;
    float2 z = make_float2(0.f, 0.f);
    if (k < M2H_ISTFT_N) {
      const int kk = k < M2H_NBINS ? k : M2H_ISTFT_N - k;              // source bin; mirrored bins are conjugated
      const float sgn = k < M2H_NBINS ? 1.f : -1.f;
      const bool real_only = kk == 0 || kk == M2H_NBINS - 1;            // irfft ignores the imaginary part of DC and Nyquist
      const size_t o = (((size_t)b * M2H_NBINS + kk) * P.T + t) * C;
      float sl, cl;
      sincosf(P.i_phase[o], &sl, &cl);
      const float ml = P.i_mag[o];
      float2 xl = make_float2(ml * cl, real_only ? 0.f : sgn * ml * sl);
      float2 xr = make_float2(0.f, 0.f);
      if (C > 1) {
        float sr, cr;
        sincosf(P.i_phase[o + 1], &sr, &cr);
        const float mr = P.i_mag[o + 1];
        xr = make_float2(mr * cr, real_only ? 0.f : sgn * mr * sr);
      }
      z = c_mul(make_float2(xl.x - xr.y, xl.y + xr.x), gt->wc[k]);
    }
    v[n1] = z;
  }
  fft_reg<16, false, true>(v);
  float2* p = s.buf[grp] + M2H_STFT_POS2(r >> 3) + (r & 7);
  twiddle_powers_out<16, 4, false>(v, M2H_STFT_TW(gt, 0, tid), M2H_STFT_TW(gt, 1, tid), [&](int k1, float2 x) { p[k1 * M2H_STFT_S1] = x; });
}
#ifdef M2H_HOST_EMU
M2H_DEV void istft_add(float* p, float v) { *p += v; }
#else
M2H_DEV void istft_add(float* p, float v) { atomicAdd(p, v); }
#endif
// Periodic Hann window of 1022 points, squared
M2H_DEV float istft_win2(int m) {
  const float sn = m2h_sinpif((float)m * (1.0f / (float)M2H_ISTFT_N));      // hann = sin^2(pi m / N): no cancellation near the window's ends
  const float w = sn * sn;
  return w * w;
}
// P6 of the inverse: the frame arrives windowed and scaled (both folded into the table's output multiplier); divide by the
// window's sum of squares over the frames that cover the sample and overlap-add.  At most two frames meet in a sample, so
// the two fp32 additions commute: the result does not depend on their order.
M2H_DEV void istft_p6(const StftParams& P, StftSmem& s, int tid, int b, int t0) {
  const int grp = tid >> 7, lt = tid & 127, t = t0 + grp;
  if (t >= P.T) return;
  const int C = P.i_channels;
  for (int m = lt; m < M2H_ISTFT_N; m += 128) {
    const int smp = M2H_HOP * t + m - M2H_ISTFT_N / 2;                 // centre=True: the first n_fft/2 samples are dropped
    if (smp < 0 || smp >= P.i_length) continue;
    float wss = istft_win2(m);
    if (m >= M2H_HOP && t + 1 < P.T) wss += istft_win2(m - M2H_HOP);
    if (m + M2H_HOP < M2H_ISTFT_N && t >= 1) wss += istft_win2(m + M2H_HOP);
    const float inv = wss > 1.17549435e-38f ? 1.0f / wss : 1.0f;       // librosa: divide where the sum exceeds tiny
    const float2 z = s.buf[grp][stft_pos(m)];
    float* o = P.i_out + (size_t)b * C * P.i_length + smp;
    istft_add(o, z.x * inv);
    if (C > 1) istft_add(o + P.i_length, z.y * inv);
  }
}

M2H_DEV void istft_cta_body(const StftParams& P, StftSmem& s, const M2hStftTables* M2H_RESTRICT gt, int cta, int n_ctas) {
  M2H_FOR_THREADS(tid, M2H_STFT_THREADS) {
    for (int i = tid; i < 2048; i += M2H_STFT_THREADS) s.bf[(i & 7) * 256 + (i >> 3)] = gt->bf[i];
  }
  M2H_SYNC();
  const int n_fgroups = (P.T + M2H_STFT_GROUPS - 1) / M2H_STFT_GROUPS;
  const int n_items = P.B * n_fgroups;
  for (int item = cta; item < n_items; item += n_ctas) {
    const int b = item / n_fgroups, t0 = (item - b * n_fgroups) * M2H_STFT_GROUPS;
    M2H_FOR_THREADS(tid, M2H_STFT_THREADS) istft_p1(P, s, gt, tid, b, t0);
    M2H_SYNC_GROUP128();
    M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p2(P, s, gt, tid, t0);
    M2H_SYNC_GROUP128();
    M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p3(P, s, gt, tid, t0);
    M2H_SYNC_GROUP128();
    M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p4(P, s, tid, t0);
    M2H_SYNC_GROUP128();
    M2H_FOR_THREADS(tid, M2H_STFT_THREADS) stft_p5(P, s, gt, tid, t0);
    M2H_SYNC_GROUP128();
    M2H_FOR_THREADS(tid, M2H_STFT_THREADS) istft_p6(P, s, tid, b, t0);
    M2H_SYNC_GROUP128();
  }
}

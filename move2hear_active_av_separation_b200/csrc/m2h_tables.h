// Host-side construction of the twiddle / chirp tables (double precision, rounded once to fp32).
// Shared by the CUDA library (m2h_render.cu) and by the CPU emulation harness used in the tests.
#pragma once
#include <math.h>
#include <vector>
#include "m2h_common.cuh"

static float2 cis(double turns_neg) {     // exp(-2 pi i * turns)
  const double a = -2.0 * M_PI * turns_neg;
  return make_float2((float)cos(a), (float)sin(a));
}

static void build_conv_tables(M2hConvTables& t) {
  for (int i = 0; i < 512; ++i) {
    t.thr[0][i] = cis((double)i / 16384.0);
    t.thr[1][i] = cis((double)(8 * i) / 16384.0);
    t.thr[2][i] = cis((double)(i & 15) / 512.0);
    t.thr[3][i] = cis((double)(8 * (i & 15)) / 512.0);
    t.thr[4][i] = cis((double)(i & 31) / 512.0);
    t.thr[5][i] = cis((double)(4 * (i & 31)) / 512.0);
    const int w = i >> 5, l = i & 31;
    const int k1 = w ? w : (l < 16 ? 16 : 0), k2 = w ? l : (l < 16 ? l : l - 16);
    t.thr[6][i] = cis((double)(k1 + 32 * k2) / 32768.0);
  }
  for (int u = 0; u < 16; ++u) t.unit_tw[u] = cis((double)(u < 8 ? 1024 * u : 512 + 1024 * (u - 8)) / 32768.0);
}

#ifdef M2H_C24_M
static void build_conv24_tables(M2hConv24Tables& t) {
  for (int r = 0; r < 256; ++r) {
    t.a1[r] = cis((double)r / 4096.0);
    t.a4[r] = cis((double)(4 * r) / 4096.0);
    t.t256[r] = cis((double)((r >> 4) * (r & 15)) / 256.0);
  }
  for (int i = 0; i < 384; ++i) t.m1[i] = cis((double)i / 12288.0);
  for (int k1 = 0; k1 < 3; ++k1)
    for (int q1 = 0; q1 < 16; ++q1)
      for (int q2 = 0; q2 < 16; ++q2) t.g[k1 * 256 + q1 * 16 + q2] = cis((double)(k1 + 3 * q1 + 48 * q2) / 24576.0);
}
#endif

// Bluestein tables of an N-point DFT evaluated as a 2048-point circular convolution (c[n] = exp(+i pi n^2 / N)):
//   forward (the STFT, N = 1023):  X[k] = conj(c[k]) sum_n (x[n] w[n] conj(c[n])) c[k-n]
//        wc = hann_N * conj(c),  bf = FFT_2048(c on the circle),  cc = conj(c) / 4096   (1/2048 inverse-FFT scale, 1/2 ear untangle)
//   inverse (librosa.istft's frames, N = 1022):  x[n] = (1/N) c[n] sum_k (X[k] c[k]) conj(c)[n-k]
//        wc = c,  bf = FFT_2048(conj(c) on the circle),  cc = hann_N * c / (N * 2048)     (the synthesis window rides in cc)
static void build_bluestein_tables(M2hStftTables& t, int N, bool inverse) {
  const int F = 2048;
  for (int i = 0; i < 128; ++i) {
    t.thr[0][i] = cis((double)i / 2048.0);
    t.thr[1][i] = cis((double)(4 * i) / 2048.0);
    t.thr[2][i] = cis((double)(i & 7) / 128.0);
    t.thr[3][i] = cis((double)(4 * (i & 7)) / 128.0);
    t.thr[4][i] = cis((double)(i & 15) / 128.0);
    t.thr[5][i] = cis((double)(4 * (i & 15)) / 128.0);
  }
  std::vector<double> cr(N), ci(N);
  for (int n = 0; n < N; ++n) {          // chirp c[n] = exp(+i pi n^2 / N), phase reduced exactly in integers
    const long long m = ((long long)n * n) % (2 * N);
    cr[n] = cos(M_PI * (double)m / N);
    ci[n] = sin(M_PI * (double)m / N);
  }
  const double sgn = inverse ? -1.0 : 1.0;      // the convolution kernel is c (forward) or conj(c) (inverse)
  for (int n = 0; n < 1024; ++n) {
    if (n < N) {
      const double w = 0.5 - 0.5 * cos(2.0 * M_PI * n / N);       // periodic Hann (fftbins=True)
      if (!inverse) {
        t.wc[n] = make_float2((float)(w * cr[n]), (float)(-w * ci[n]));
        t.cc[n] = make_float2((float)(cr[n] / 4096.0), (float)(-ci[n] / 4096.0));
      } else {
        t.wc[n] = make_float2((float)cr[n], (float)ci[n]);
        t.cc[n] = make_float2((float)(w * cr[n] / (2048.0 * N)), (float)(w * ci[n] / (2048.0 * N)));
      }
    } else {
      t.wc[n] = make_float2(0.f, 0.f);
      t.cc[n] = make_float2(0.f, 0.f);
    }
  }
  // b[j] = c[|j|] (or its conjugate) on the circle of 2048, Bf = DFT_2048(b) in double
  std::vector<double> br(F, 0.0), bi(F, 0.0);
  for (int j = 0; j < N; ++j) {
    br[j] = cr[j]; bi[j] = sgn * ci[j];
    if (j > 0) { br[F - j] = cr[j]; bi[F - j] = sgn * ci[j]; }
  }
  std::vector<double> wr(F), wi(F);
  for (int i = 0; i < F; ++i) { wr[i] = cos(-2.0 * M_PI * i / F); wi[i] = sin(-2.0 * M_PI * i / F); }
  for (int f = 0; f < F; ++f) {
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < F; ++j) {
      const int e = (int)(((long long)f * j) & (F - 1));
      sr += br[j] * wr[e] - bi[j] * wi[e];
      si += br[j] * wi[e] + bi[j] * wr[e];
    }
    const int k1 = f & 15, k2 = (f >> 4) & 15, k3 = f >> 8;
    t.bf[(k1 * 16 + k2) * 8 + k3] = make_float2((float)sr, (float)si);
  }
}
static void build_stft_tables(M2hStftTables& t) { build_bluestein_tables(t, M2H_NFFT, false); }
static void build_istft_tables(M2hStftTables& t) { build_bluestein_tables(t, 1022, true); }

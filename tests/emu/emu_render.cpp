// TEST-ONLY: runs the render kernels' bodies (csrc/render_conv.cuh, csrc/render_stft.cuh) on the CPU by
// executing every barrier-separated phase as a loop over thread ids (-DM2H_HOST_EMU).  It exists so that the
// CPU test-suite can check the kernels' index arithmetic and fp32 numerics against the oracle without a GPU.
// It is not linked into, or reachable from, the product library.
#define M2H_HOST_EMU 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../move2hear_active_av_separation_b200/csrc/render_conv.cuh"
#include "../../move2hear_active_av_separation_b200/csrc/render_conv24.cuh"
#include "../../move2hear_active_av_separation_b200/csrc/render_stft.cuh"
#include "../../move2hear_active_av_separation_b200/csrc/m2h_tables.h"

// IEEE fp16 round-to-nearest-even of a float, returned as float
static uint16_t f32_to_f16_bits(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((x > 0x7f800000u) ? 0x200u : 0));
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                 // overflow -> inf (>= 65520)
  if (x < 0x33000001u) return (uint16_t)sign;                              // underflow -> 0
  int e = (int)(x >> 23) - 127;
  uint32_t m = (x & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t he;
  if (e < -14) { shift = 13 + (-14 - e); he = 0; } else { shift = 13; he = (uint32_t)(e + 15); }
  uint32_t hm = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (hm & 1))) hm++;
  uint32_t h = (he == 0) ? hm : ((he << 10) + (hm - 0x400u));              // carries propagate into the exponent
  return (uint16_t)(sign | h);
}
static float f16_bits_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 31, m = h & 0x3ffu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int s = 0; while (!(m & 0x400u)) { m <<= 1; s++; } m &= 0x3ffu; x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | (m << 13); }
  } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &x, 4); return f;
}
float m2h_emu_round_f16(float v) { return f16_bits_to_f32(f32_to_f16_bits(v)); }
void m2h_emu_store(void* base, size_t idx, float v, bool f16) {
  if (f16) ((uint16_t*)base)[idx] = f32_to_f16_bits(v); else ((float*)base)[idx] = v;
}

static M2hConvTables g_ct;
static M2hConv24Tables g_c24;
static M2hStftTables g_st;
static bool g_ready = false;
static void ensure() { if (!g_ready) { build_conv_tables(g_ct); build_conv24_tables(g_c24); build_stft_tables(g_st); g_ready = true; } }

extern "C" int emu_rir_convolve(const int16_t* mono, const float* rir, const int32_t* rir_len,
                                int B, int K, int N, int Lmax, int16_t* waves, int n_ctas) {
  ensure();
  std::vector<float4> scratch((size_t)n_ctas * M2H_CONV_SCRATCH_F4);
  ConvSmem* s = new ConvSmem;
  ConvParams P;
  P.mono = mono; P.rir = rir; P.rir_len = rir_len; P.waves = waves; P.scratch = scratch.data();
  P.n_items = B * K; P.N = N; P.Lmax = Lmax; P.mono_index = nullptr; P.rir_index = nullptr; P.ear_split = 0;
  const bool vec = conv_can_vectorise(P);      // same dispatch rules as m2h_rir_convolve
  if (Lmax <= M2H_C24_MAX_TAPS) {              // F = 24576 kernel (render_conv24.cuh)
    Conv24Smem* s24 = new Conv24Smem;
    for (int cta = 0; cta < n_ctas; ++cta) {
      if (vec) conv24_cta_body<true>(P, *s24, &g_c24, cta, n_ctas); else conv24_cta_body<false>(P, *s24, &g_c24, cta, n_ctas);
    }
    delete s24;
    delete s;
    return vec ? 3 : 2;
  }
  for (int cta = 0; cta < n_ctas; ++cta) {
    if (vec) conv_cta_body<true>(P, *s, &g_ct, cta, n_ctas); else conv_cta_body<false>(P, *s, &g_ct, cta, n_ctas);
  }
  delete s;
  return vec ? 1 : 0;
}

extern "C" int emu_stft_waves(const int16_t* waves, int B, int K, int N, int n_gt, int gt_with_phase, int f16, int raw,
                              void* mix_mag, void* mix_phase, void* gt, int n_ctas) {
  ensure();
  StftSmem* s = new StftSmem;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.waves = waves; P.B = B; P.K = K; P.N = N; P.T = 1 + (N - 1) / M2H_HOP; P.n_gt = n_gt; P.n_signals = 1 + n_gt;
  P.flags = (raw ? 0 : M2H_OUT_MIX_LOG1P) | (f16 ? M2H_OUT_F16 : 0);
  P.mix_mag = mix_mag; P.mix_phase = mix_phase; P.gt_mag = gt; P.gt_interleaved_phase = gt_with_phase;
  for (int cta = 0; cta < n_ctas; ++cta) { if (f16) stft_cta_body<true>(P, *s, &g_st, cta, n_ctas); else stft_cta_body<false>(P, *s, &g_st, cta, n_ctas); }
  delete s;
  return 0;
}

extern "C" int emu_stft_f32(const float* wave, int B, int C, int N, float* mag, float* phase) {
  ensure();
  StftSmem* s = new StftSmem;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.fwave = wave; P.fwave_channels = C; P.B = B; P.N = N; P.T = 1 + (N - 1) / M2H_HOP; P.n_signals = 1;
  P.f_mag = mag; P.f_phase = phase;
  stft_cta_body<false>(P, *s, &g_st, 0, 1);
  delete s;
  return 0;
}

extern "C" int emu_istft(const float* mag, const float* phase, int B, int C, int T, int length, float* out, int n_ctas) {
  static M2hStftTables it;
  static bool ready = false;
  if (!ready) { build_istft_tables(it); ready = true; }
  StftSmem* s = new StftSmem;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.T = T; P.i_mag = mag; P.i_phase = phase; P.i_out = out; P.i_channels = C; P.i_length = length;
  memset(out, 0, sizeof(float) * (size_t)B * C * length);
  for (int cta = 0; cta < n_ctas; ++cta) istft_cta_body(P, *s, &it, cta, n_ctas);
  delete s;
  return 0;
}

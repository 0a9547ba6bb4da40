// Consumers of the separator's masks on the rollout / evaluation step:
//   * STFT_L2_distance            audio_separation/common/eval_metrics.py:306-366   (rl/ppo/ppo_trainer.py:407-420)
//   * AcousticMem.forward         audio_separation/rl/models/memory_nets.py:5-69    (rl/ppo/ppo_trainer.py:310-319)
//   * waveform moments for SI-SDR audio_separation/common/eval_metrics.py:12-57, 167-230 (scale_bss_eval_helper / preprocess)
// All three are small, latency-bound fp32 jobs at the rollout batch (B = 14): one launch each, no intermediate tensors
// (the reference builds ~30 elementwise temporaries for STFT_L2_distance and 8 layout ops + 2 cuDNN calls for the refiner).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/m2h.h"
#include "m2h_internal.h"

// ---------------------------------------------------------------------------------------------
// block reduction helpers
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_add(double (&v)[NV], double* smem /* [NV * 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if (lane == 0) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < n_warps ? smem[i * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[i] = x;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// STFT_L2_distance
// ---------------------------------------------------------------------------------------------
struct StftL2Params {
  const float* mixed;      // [B, 512, T, 2] contiguous
  const float* masks;      // [B, 512, T, 2] via strides
  int64_t ms[4];
  const float* gt_bin;     // [B, 512, T, c_bin] contiguous, components 0..3 = mag_l, phase_l, mag_r, phase_r
  int c_bin;
  const float* pred_mono;  // [B, 512, T, 1] via strides
  int64_t ps[4];
  const float* gt_mono;    // [B, 512, T, c_mono] contiguous, components 0..1 = mag, phase
  int c_mono;
  int T;
  float* bin_out;          // [B]
  float* mono_out;         // [B]
};

// One CTA per sample.  The arithmetic follows the reference term by term in fp32 (real / imaginary parts of the GT and of
// the prediction carried by the GT phase, squared differences); only the mean is accumulated in fp64.
__global__ void __launch_bounds__(512) stft_l2_kernel(const StftL2Params P) {
  __shared__ double red[3 * 32];
  const int b = blockIdx.x, T = P.T, n = 512 * T;
  double acc[3] = {0.0, 0.0, 0.0};            // left ear, right ear, mono
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int f = i / T, t = i - f * T;
    const size_t e = (size_t)b * n + i;
    const float* gb = P.gt_bin + e * P.c_bin;
    const float2 mix = *reinterpret_cast<const float2*>(P.mixed + e * 2);
    const float* mk = P.masks + b * P.ms[0] + f * P.ms[1] + t * P.ms[2];
    const float pl = (expf(mix.x) - 1.0f) * mk[0];
    const float pr = (expf(mix.y) - 1.0f) * mk[P.ms[3]];
    float s, c;
    sincosf(gb[1], &s, &c);
    float dr = gb[0] * c - pl * c, di = gb[0] * s - pl * s;
    acc[0] += (double)(dr * dr) + (double)(di * di);
    sincosf(gb[3], &s, &c);
    dr = gb[2] * c - pr * c; di = gb[2] * s - pr * s;
    acc[1] += (double)(dr * dr) + (double)(di * di);
    const float* gm = P.gt_mono + e * P.c_mono;
    const float pm = P.pred_mono[b * P.ps[0] + f * P.ps[1] + t * P.ps[2]];
    sincosf(gm[1], &s, &c);
    dr = gm[0] * c - pm * c; di = gm[0] * s - pm * s;
    acc[2] += (double)(dr * dr) + (double)(di * di);
  }
  block_reduce_add<3>(acc, red);
  if (threadIdx.x == 0) {
    const double inv = 1.0 / (2.0 * n);        // mean over the [2 (re, im) x 512 x T] view
    P.bin_out[b] = (float)(acc[0] * inv) + (float)(acc[1] * inv);
    P.mono_out[b] = (float)(acc[2] * inv);
  }
}

extern "C" int m2h_stft_l2_distance(const float* mixed_audio, const float* pred_masks, const int64_t* mask_strides,
                                    const float* gt_bin_comps, int c_bin, const float* pred_mono, const int64_t* mono_strides,
                                    const float* gt_mono_comps, int c_mono, int B, int T, float* bin_dist, float* mono_dist,
                                    void* stream) {
  if (B < 0 || T < 1 || c_bin < 4 || c_mono < 2) { m2h_set_error("m2h_stft_l2_distance: bad sizes B=%d T=%d c_bin=%d c_mono=%d", B, T, c_bin, c_mono); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!mixed_audio || !pred_masks || !mask_strides || !gt_bin_comps || !pred_mono || !mono_strides || !gt_mono_comps || !bin_dist || !mono_dist) {
    m2h_set_error("m2h_stft_l2_distance: null pointer"); return M2H_ERR_ARG;
  }
  if (((uintptr_t)mixed_audio & 7) != 0) { m2h_set_error("m2h_stft_l2_distance: mixed_audio must be 8-byte aligned"); return M2H_ERR_ARG; }
  StftL2Params P;
  P.mixed = mixed_audio; P.masks = pred_masks; P.gt_bin = gt_bin_comps; P.c_bin = c_bin; P.pred_mono = pred_mono;
  P.gt_mono = gt_mono_comps; P.c_mono = c_mono; P.T = T; P.bin_out = bin_dist; P.mono_out = mono_dist;
  for (int i = 0; i < 4; ++i) { P.ms[i] = mask_strides[i]; P.ps[i] = mono_strides[i]; }
  stft_l2_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(P);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

// ---------------------------------------------------------------------------------------------
// waveform moments for the scale-invariant BSS metrics
// ---------------------------------------------------------------------------------------------
// out[b] = { n, S r, S e, S m, S rr, S ee, S mm, S re, S rm }  with m = mean of the mixture's channels
__global__ void __launch_bounds__(512) wave_moments_kernel(const float* __restrict__ ref, const float* __restrict__ est,
                                                           const float* __restrict__ mix, int mix_channels, int L, double* __restrict__ out) {
  __shared__ double red[8 * 32];
  const int b = blockIdx.x;
  const float* r = ref + (size_t)b * L;
  const float* e = est + (size_t)b * L;
  const float* m = mix + (size_t)b * mix_channels * L;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float inv_c = 1.0f / (float)mix_channels;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const double rv = r[i], ev = e[i];
    float ms = 0.f;
    for (int c = 0; c < mix_channels; ++c) ms += m[(size_t)c * L + i];
    const double mv = (double)(ms * inv_c);
    a[0] += rv; a[1] += ev; a[2] += mv; a[3] += rv * rv; a[4] += ev * ev; a[5] += mv * mv; a[6] += rv * ev; a[7] += rv * mv;
  }
  block_reduce_add<8>(a, red);
  if (threadIdx.x == 0) {
    out[(size_t)b * 9] = (double)L;
    for (int i = 0; i < 8; ++i) out[(size_t)b * 9 + 1 + i] = a[i];
  }
}

extern "C" int m2h_wave_moments(const float* reference, const float* estimate, const float* mixture, int mix_channels,
                                int B, int L, double* moments, void* stream) {
  if (B < 0 || L < 1 || mix_channels < 1) { m2h_set_error("m2h_wave_moments: bad sizes B=%d L=%d channels=%d", B, L, mix_channels); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!reference || !estimate || !mixture || !moments) { m2h_set_error("m2h_wave_moments: null pointer"); return M2H_ERR_ARG; }
  wave_moments_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(reference, estimate, mixture, mix_channels, L, moments);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

// ---------------------------------------------------------------------------------------------
// AcousticMem refiner: conv3x3(32 -> 32) [+ BatchNorm(eval)] + ReLU + conv3x3(32 -> 16), one launch
// ---------------------------------------------------------------------------------------------
// The reference slices the [B, 512, 32, 1] spectrograms along frequency into 16 chunks of 32 bins: channel s of the
// [B, 16, 32, 32] map is bins [32 s, 32 s + 32), spatial = (bin within chunk, frame).  Channels 0..15 = pred_mono,
// 16..31 = the (masked) previous memory prediction.  A CTA produces an 8-row x 32-column output tile of all 16 output
// channels: the 12 x 36 x 32 input tile and the 10 x 34 x 32 hidden tile stay in shared memory.
#define AM_TILE_H 8
#define AM_W 32
#define AM_IN_H (AM_TILE_H + 4)
#define AM_IN_W (AM_W + 4)
#define AM_MID_H (AM_TILE_H + 2)
#define AM_MID_W (AM_W + 2)
#define AM_THREADS 256

struct AcousticMemParams {
  const float* pred_mono; int64_t s0[4];
  const float* prev;      int64_t s1[4];
  const float* w1;        // [32, 32, 3, 3]
  const float* w2;        // [16, 32, 3, 3]
  const float* bn_w; const float* bn_b; const float* bn_mean; const float* bn_var;   // [32] or null (use_ddppo variant)
  float bn_eps;
  float* out;             // [B, 512, 32, 1] contiguous
};

struct AcousticMemSmem {
  float in[32][AM_IN_H][AM_IN_W + 1];
  float mid[32][AM_MID_H][AM_MID_W + 1];
  float w1[9][32][32];     // [tap][cin][cout]
  float w2[9][32][16];
  float scale[32], shift[32];
};

__global__ void __launch_bounds__(AM_THREADS) acoustic_mem_kernel(const AcousticMemParams P) {
  extern __shared__ __align__(16) unsigned char am_raw[];
  AcousticMemSmem& s = *reinterpret_cast<AcousticMemSmem*>(am_raw);
  const int b = blockIdx.y, h0 = blockIdx.x * AM_TILE_H, tid = threadIdx.x;
  for (int i = tid; i < 9 * 32 * 32; i += AM_THREADS) {
    const int co = i & 31, ci = (i >> 5) & 31, tap = i >> 10;
    s.w1[tap][ci][co] = P.w1[(co * 32 + ci) * 9 + tap];
  }
  for (int i = tid; i < 9 * 32 * 16; i += AM_THREADS) {
    const int co = i & 15, ci = (i >> 4) & 31, tap = i >> 9;
    s.w2[tap][ci][co] = P.w2[(co * 32 + ci) * 9 + tap];
  }
  if (tid < 32) {
    float sc = 1.f, sh = 0.f;
    if (P.bn_w) { sc = P.bn_w[tid] / sqrtf(P.bn_var[tid] + P.bn_eps); sh = P.bn_b[tid] - P.bn_mean[tid] * sc; }
    s.scale[tid] = sc; s.shift[tid] = sh;
  }
  // input tile with zero padding: rows h0-2 .. h0+9, columns -2 .. 33
  for (int i = tid; i < 32 * AM_IN_H * AM_IN_W; i += AM_THREADS) {
    const int x = i % AM_IN_W, y = (i / AM_IN_W) % AM_IN_H, c = i / (AM_IN_W * AM_IN_H);
    const int h = h0 + y - 2, w = x - 2;
    float v = 0.f;
    if (h >= 0 && h < 32 && w >= 0 && w < AM_W) {
      const int slice = c & 15, f = slice * 32 + h;
      v = c < 16 ? P.pred_mono[b * P.s0[0] + f * P.s0[1] + w * P.s0[2]] : P.prev[b * P.s1[0] + f * P.s1[1] + w * P.s1[2]];
    }
    s.in[c][y][x] = v;
  }
  __syncthreads();
  // hidden tile: rows h0-1 .. h0+8, columns -1 .. 32 (positions outside the image are the second conv's zero padding)
  for (int i = tid; i < AM_MID_H * AM_MID_W * 8; i += AM_THREADS) {
    const int cg = i & 7, p = i >> 3, x = p % AM_MID_W, y = p / AM_MID_W;      // 4 output channels per work item
    const int h = h0 + y - 1, w = x - 1;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (h >= 0 && h < 32 && w >= 0 && w < AM_W) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
#pragma unroll 8
        for (int ci = 0; ci < 32; ++ci) {
          const float v = s.in[ci][y + dy][x + dx];
          const float4 wv = *reinterpret_cast<const float4*>(&s.w1[tap][ci][cg * 4]);
          a0 = fmaf(v, wv.x, a0); a1 = fmaf(v, wv.y, a1); a2 = fmaf(v, wv.z, a2); a3 = fmaf(v, wv.w, a3);
        }
      }
      a0 = fmaxf(a0 * s.scale[cg * 4] + s.shift[cg * 4], 0.f);
      a1 = fmaxf(a1 * s.scale[cg * 4 + 1] + s.shift[cg * 4 + 1], 0.f);
      a2 = fmaxf(a2 * s.scale[cg * 4 + 2] + s.shift[cg * 4 + 2], 0.f);
      a3 = fmaxf(a3 * s.scale[cg * 4 + 3] + s.shift[cg * 4 + 3], 0.f);
    }
    s.mid[cg * 4][y][x] = a0; s.mid[cg * 4 + 1][y][x] = a1; s.mid[cg * 4 + 2][y][x] = a2; s.mid[cg * 4 + 3][y][x] = a3;
  }
  __syncthreads();
  for (int i = tid; i < AM_TILE_H * AM_W * 4; i += AM_THREADS) {
    const int x = i & 31, cg = (i >> 5) & 3, y = i >> 7;                          // x fastest: coalesced stores along the frame axis
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
#pragma unroll 8
      for (int ci = 0; ci < 32; ++ci) {
        const float v = s.mid[ci][y + dy][x + dx];
        const float4 wv = *reinterpret_cast<const float4*>(&s.w2[tap][ci][cg * 4]);
        a0 = fmaf(v, wv.x, a0); a1 = fmaf(v, wv.y, a1); a2 = fmaf(v, wv.z, a2); a3 = fmaf(v, wv.w, a3);
      }
    }
    const int h = h0 + y;
    float* o = P.out + (size_t)b * 512 * AM_W + (size_t)h * AM_W + x;              // out[b, slice*32 + h, x, 0]
    o[(size_t)(cg * 4) * 32 * AM_W] = a0;
    o[(size_t)(cg * 4 + 1) * 32 * AM_W] = a1;
    o[(size_t)(cg * 4 + 2) * 32 * AM_W] = a2;
    o[(size_t)(cg * 4 + 3) * 32 * AM_W] = a3;
  }
}

extern "C" int m2h_acoustic_mem_forward(const float* pred_mono, const int64_t* pred_mono_strides, const float* prev_pred,
                                        const int64_t* prev_pred_strides, const float* conv1_w, const float* conv2_w,
                                        const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var, float bn_eps,
                                        int B, float* out, void* stream) {
  if (B < 0) { m2h_set_error("m2h_acoustic_mem_forward: bad batch %d", B); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!pred_mono || !pred_mono_strides || !prev_pred || !prev_pred_strides || !conv1_w || !conv2_w || !out) { m2h_set_error("m2h_acoustic_mem_forward: null pointer"); return M2H_ERR_ARG; }
  if (bn_w && (!bn_b || !bn_mean || !bn_var)) { m2h_set_error("m2h_acoustic_mem_forward: incomplete BatchNorm parameters"); return M2H_ERR_ARG; }
  static bool attr_set[M2H_MAX_DEVICES] = {false};
  int dev;
  int rc = m2h_current_device_checked(&dev);
  if (rc != M2H_OK) return rc;
  if (!attr_set[dev]) {
    M2H_CUDA_CHECK(cudaFuncSetAttribute(acoustic_mem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AcousticMemSmem)));
    attr_set[dev] = true;
  }
  AcousticMemParams P;
  P.pred_mono = pred_mono; P.prev = prev_pred; P.w1 = conv1_w; P.w2 = conv2_w;
  P.bn_w = bn_w; P.bn_b = bn_b; P.bn_mean = bn_mean; P.bn_var = bn_var; P.bn_eps = bn_eps; P.out = out;
  for (int i = 0; i < 4; ++i) { P.s0[i] = pred_mono_strides[i]; P.s1[i] = prev_pred_strides[i]; }
  acoustic_mem_kernel<<<dim3(32 / AM_TILE_H, B), AM_THREADS, sizeof(AcousticMemSmem), (cudaStream_t)stream>>>(P);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

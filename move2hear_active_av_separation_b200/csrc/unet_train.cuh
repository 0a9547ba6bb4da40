// Training-side kernels of the passive-separator U-Net: train-mode BatchNorm (batch statistics, running-stat update,
// backward), the 1x1 head forward/backward, the L1 mask loss, and the weight-gradient GEMM on tcgen05.
//
// Reference semantics: nn.BatchNorm2d in .train() (audio_separation/rl/models/separator_cnn.py:5-24: momentum 0.1,
// eps 1e-5, biased variance for normalisation, unbiased for the running estimate), LeakyReLU(0.2) / ReLU, the output
// Conv2d 1x1 (:135), and PassiveTrainer.optimize_supervised_loss (audio_separation/pretrain/passive/passive_trainer.py:269-286).
//
// All feature maps are haloed NHWC bf16 ([B][H+2][W+2][ld], zero border), addressed as a flat list of positions.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "unet_gemm.cuh"

// A [positions][ld] bf16 view: channels [coff, coff + C) of every position; Hp == 0 means "no halo" (a plain [P][ld] matrix)
struct FMap {
  __nv_bfloat16* p; int ld, coff, Hp; long long P;     // P = B * Hp * Hp positions (or B rows)
};
// position counts stay below 2^31 (checked on the host: bn_forward / bn_backward): 32-bit unsigned divisions - the 64-bit
// modulo of the first version was a ~150-instruction call per position in kernels that otherwise do 8 FMAs
__device__ __forceinline__ bool fmap_interior(const FMap& m, long long pos) {
  if (m.Hp == 0) return true;
  const unsigned hp = (unsigned)m.Hp, r = (unsigned)pos % (hp * hp), y = r / hp, x = r - y * hp;
  return y >= 1u && y <= hp - 2u && x >= 1u && x <= hp - 2u;
}
__device__ __forceinline__ void ld8(const FMap& m, long long pos, int c8, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(m.p + pos * m.ld + m.coff + c8 * 8);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void st8(const FMap& m, long long pos, int c8, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(m.p + pos * m.ld + m.coff + c8 * 8) = u;
}

// ---- BatchNorm forward (batch statistics) ----------------------------------------------------------------------------------
// per-channel sum / sum of squares of z over all positions (border positions hold zeros and add nothing): stats[0][C], stats[1][C]
// block-level tree reduction of per-thread partial sums laid out [pl][C] in shared memory, then ONE fp64 atomic per channel and
// block (thread-level fp64 atomics on a few hundred addresses serialise in L2: first capture 700 us per call)
__device__ __forceinline__ void block_reduce_2xC(float* sm, int C, int ppi, int c8, int pl, const float (&s)[8], const float (&q)[8],
                                                 double* out0, double* out1) {
  float* sm1 = sm + ppi * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sm[pl * C + c8 * 8 + i] = s[i]; sm1[pl * C + c8 * 8 + i] = q[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const float* src = c < C ? sm + c : sm1 + (c - C);
    float t = 0.f;
    for (int p = 0; p < ppi; ++p) t += src[p * C];
    atomicAdd(c < C ? out0 + c : out1 + (c - C), (double)t);
  }
}
__global__ void bn_stats_kernel(const FMap z, int C, double* __restrict__ stats) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  extern __shared__ float bn_sm[];
  const int chunks = C / 8, ppi = blockDim.x / chunks;            // positions per block iteration
  const int c8 = threadIdx.x % chunks, pl = threadIdx.x / chunks;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long pos = (long long)blockIdx.x * ppi + pl; pos < z.P; pos += (long long)gridDim.x * ppi) {
    float v[8];
    ld8(z, pos, c8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
  }
  block_reduce_2xC(bn_sm, C, ppi, c8, pl, s, q, stats, stats + C);
}
// mean / biased variance -> scale = gamma * invstd, shift = beta - mean * scale; running stats (unbiased variance) updated in place.
// aux = [scale | shift | mean | invstd] (each Cpad); channels >= Creal (zero-padded weight rows) get scale = shift = 0.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double n, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* __restrict__ run_mean, float* __restrict__ run_var,
                                   long long* __restrict__ num_batches, float* __restrict__ aux, int Creal, int Cpad) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches) *num_batches += 1;
  if (c >= Cpad) return;
  float scale = 0.f, shift = 0.f, mean = 0.f, invstd = 0.f;
  if (c < Creal) {
    const double m = stats[c] / n;
    double var = stats[Cpad + c] / n - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    scale = gamma[c] * invstd;
    shift = beta[c] - mean * scale;
    if (run_mean) {
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(n > 1 ? var * n / (n - 1) : var);
    }
  }
  aux[c] = scale; aux[Cpad + c] = shift; aux[2 * Cpad + c] = mean; aux[3 * Cpad + c] = invstd;
}
// a = act(z * scale + shift) at interior positions
__global__ void bn_apply_kernel(const FMap z, const FMap a, int C, const float* __restrict__ aux, float slope) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  // C / 8 is a power of two and divides the grid stride: a thread keeps its eight channels.  (Their constants stay in L1: with
  // ~4 items per thread, hoisting 16-48 loads into a prologue cost more than it saved - measured.)
  const unsigned chunks = (unsigned)C / 8u, cshift = 31u - (unsigned)__clz((int)chunks);
  const unsigned total = (unsigned)z.P * chunks, stride = gridDim.x * blockDim.x;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = (int)(i0 & (chunks - 1u));
  for (unsigned i = i0; i < total; i += stride) {
    const long long pos = (long long)(i >> cshift);
    if (!fmap_interior(z, pos)) continue;
    float v[8];
    ld8(z, pos, c8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float y = fmaf(v[k], aux[c8 * 8 + k], aux[C + c8 * 8 + k]);
      v[k] = y > 0.f ? y : y * slope;
    }
    st8(a, pos, c8, v);
  }
}

// ---- BatchNorm backward -------------------------------------------------------------------------------------------------------
// dy = (da1 + da2) * act'(a);  red[0][C] = sum dy * xhat, red[1][C] = sum dy      (xhat = (z - mean) * invstd)
// The gate a > 0 is re-derived from z (a = act(z * scale + shift) with a slope in [0, 1): a > 0 <=> z * scale + shift > 0, the very
// fmaf the forward pass evaluated) instead of being read back: one 2-byte read per element less in each of the two backward
// passes - they are streaming kernels over 1.3 GB per step.
__global__ void bn_bwd_reduce_kernel(const FMap da1, const FMap da2, const FMap a, const FMap z, int C, const float* __restrict__ aux,
                                     float slope, double* __restrict__ red, int gate_from_a) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  extern __shared__ float bn_sm[];
  const int chunks = C / 8, ppi = blockDim.x / chunks;
  const int c8 = threadIdx.x % chunks, pl = threadIdx.x / chunks;
  float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long pos = (long long)blockIdx.x * ppi + pl; pos < z.P; pos += (long long)gridDim.x * ppi) {
    if (!fmap_interior(z, pos)) continue;
    float g[8], g2[8], av[8], zv[8];
    ld8(da1, pos, c8, g);
    if (da2.p) { ld8(da2, pos, c8, g2);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += g2[k]; }
    ld8(z, pos, c8, zv);
    if (gate_from_a) ld8(a, pos, c8, av);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!gate_from_a) av[k] = fmaf(zv[k], aux[c8 * 8 + k], aux[C + c8 * 8 + k]);
      const float dy = av[k] > 0.f ? g[k] : g[k] * slope;
      const float xh = (zv[k] - aux[2 * C + c8 * 8 + k]) * aux[3 * C + c8 * 8 + k];
      sg[k] = fmaf(dy, xh, sg[k]);
      sb[k] += dy;
    }
  }
  block_reduce_2xC(bn_sm, C, ppi, c8, pl, sg, sb, red, red + C);
}
// dz = scale * (dy - dbeta/n - xhat * dgamma/n), written over z (interior positions); block 0 also emits the fp32 parameter grads
__global__ void bn_bwd_apply_kernel(const FMap da1, const FMap da2, const FMap a, const FMap z, int C, int Creal, const float* __restrict__ aux,
                                    float slope, const double* __restrict__ red, double n, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int gate_from_a) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < Creal; c += blockDim.x) { dgamma[c] = (float)red[c]; dbeta[c] = (float)red[C + c]; }
  const unsigned chunks = (unsigned)C / 8u, cshift = 31u - (unsigned)__clz((int)chunks);
  const unsigned total = (unsigned)z.P * chunks, stride = gridDim.x * blockDim.x;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = (int)(i0 & (chunks - 1u));
  const float inv_n = (float)(1.0 / n);
  for (unsigned i = i0; i < total; i += stride) {
    const long long pos = (long long)(i >> cshift);
    if (!fmap_interior(z, pos)) continue;
    float g[8], g2[8], av[8], zv[8];
    ld8(da1, pos, c8, g);
    if (da2.p) { ld8(da2, pos, c8, g2);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += g2[k]; }
    ld8(z, pos, c8, zv);
    if (gate_from_a) ld8(a, pos, c8, av);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      if (!gate_from_a) av[k] = fmaf(zv[k], aux[c], aux[C + c]);
      const float dy = av[k] > 0.f ? g[k] : g[k] * slope;
      const float xh = (zv[k] - aux[2 * C + c]) * aux[3 * C + c];
      zv[k] = aux[c] * (dy - (float)red[C + c] * inv_n - xh * (float)red[c] * inv_n);
    }
    st8(z, pos, c8, zv);
  }
}

// ---- 1x1 head (separator_cnn.py:135), fp32 NCHW mask [B][Creal][32][32] ---------------------------------------------------------
// a4: haloed [B][34][34][32] bf16; w: fp32 [Creal][Creal]; operands rounded to bf16 like the inference kernels
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ a4, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ mask, int B, int Creal) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  __shared__ float ws[32 * 33];
  for (int i = threadIdx.x; i < Creal * Creal; i += blockDim.x) ws[(i / Creal) * 33 + i % Creal] = __bfloat162float(__float2bfloat16_rn(w[i]));
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 1024) return;
  const int x = i & 31, y = (i >> 5) & 31, b = i >> 10;
  float v[32];
  const uint4* src = reinterpret_cast<const uint4*>(a4 + (((size_t)b * 34 + y + 1) * 34 + x + 1) * 32);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 u = src[k];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[k * 8 + 2 * j] = f.x; v[k * 8 + 2 * j + 1] = f.y; }
  }
  for (int n = 0; n < Creal; ++n) {
    float z = bias[n];
#pragma unroll
    for (int c = 0; c < 32; ++c) z = fmaf(c < Creal ? ws[n * 33 + c] : 0.f, v[c], z);
    mask[(((size_t)b * Creal + n) * 32 + y) * 32 + x] = z;
  }
}
// da4 = W^T dmask (bf16, haloed), dW += dmask a4^T, db += sum dmask   (dW / db fp32, zeroed by the caller)
// Register budget: the first version kept v[32], g[32] and d[32] live per thread (298 local-memory instructions, 0.18 ms per
// launch at B = 256: ncu r1g).  Now the pixel's activations go straight to the shared tile, d is produced eight channels at a
// time, and the dW pass reads its eight v values as two 16-byte loads.
#define M2H_HB_VS 36          // row stride of the shared tiles (floats): 16-byte aligned rows, 4-bank skew between pixels
__global__ void __launch_bounds__(128, 4) head_bwd_kernel(const __nv_bfloat16* __restrict__ a4, const float* __restrict__ w, const float* __restrict__ dmask,
                                __nv_bfloat16* __restrict__ da4, float* __restrict__ dw, float* __restrict__ db, int B, int Creal) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  __shared__ float ws[32 * 33];
  __shared__ __align__(16) float gs[128 * M2H_HB_VS];
  __shared__ __align__(16) float vs[128 * M2H_HB_VS];
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) ws[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < Creal * Creal; i += blockDim.x) ws[(i / Creal) * 33 + i % Creal] = __bfloat162float(__float2bfloat16_rn(w[i]));
  __syncthreads();
  // dW[n][c] = sum_pix g[n] v[c], db[n] = sum_pix g[n]: each thread owns 8 (n, c) pairs and keeps them in registers over all the
  // 128-pixel chunks of its block - one round of atomics per block (one per chunk serialised 2 M atomics on 1056 addresses)
  const int wn = threadIdx.x >> 2, wc0 = (threadIdx.x & 3) * 8;
  float t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tb = 0.f;
  const int n_chunks = (B * 1024 + 127) / 128;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int i = chunk * 128 + threadIdx.x;      // launched with 128 threads per block
    const bool ok = i < B * 1024;
    const int x = i & 31, y = (i >> 5) & 31, b = i >> 10;
    float* vrow = vs + threadIdx.x * M2H_HB_VS;
    float* grow = gs + threadIdx.x * M2H_HB_VS;
    {
      const uint4* src = reinterpret_cast<const uint4*>(a4 + (((size_t)b * 34 + y + 1) * 34 + x + 1) * 32);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (ok) u = src[k];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
        reinterpret_cast<float4*>(vrow)[2 * k] = make_float4(f0.x, f0.y, f1.x, f1.y);
        reinterpret_cast<float4*>(vrow)[2 * k + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
      }
    }
    float g[32];
#pragma unroll
    for (int n = 0; n < 32; ++n) g[n] = (ok && n < Creal) ? dmask[(((size_t)b * Creal + n) * 32 + y) * 32 + x] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) reinterpret_cast<float4*>(grow)[k] = make_float4(g[4 * k], g[4 * k + 1], g[4 * k + 2], g[4 * k + 3]);
    if (ok) {
      uint4* dst = reinterpret_cast<uint4*>(da4 + (((size_t)b * 34 + y + 1) * 34 + x + 1) * 32);
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {               // eight input channels at a time (ws rows beyond Creal are zero)
        float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int n = 0; n < 32; ++n) {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = fmaf(ws[n * 33 + k * 8 + j], g[n], d[j]);
        }
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(d[2 * j], d[2 * j + 1]);
        dst[k] = u;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < 128; ++px) {
      const float gn = gs[px * M2H_HB_VS + wn];
      const float4 va = *reinterpret_cast<const float4*>(vs + px * M2H_HB_VS + wc0);
      const float4 vb = *reinterpret_cast<const float4*>(vs + px * M2H_HB_VS + wc0 + 4);
      tb += gn;
      t[0] = fmaf(gn, va.x, t[0]); t[1] = fmaf(gn, va.y, t[1]); t[2] = fmaf(gn, va.z, t[2]); t[3] = fmaf(gn, va.w, t[3]);
      t[4] = fmaf(gn, vb.x, t[4]); t[5] = fmaf(gn, vb.y, t[5]); t[6] = fmaf(gn, vb.z, t[6]); t[7] = fmaf(gn, vb.w, t[7]);
    }
    __syncthreads();
  }
  if (wn < Creal) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (wc0 + k < Creal) atomicAdd(dw + wn * Creal + wc0 + k, t[k]);
    if (wc0 == 0) atomicAdd(db + wn, tb);
  }
}

// ---- L1 mask loss (passive_trainer.py:269-275) -----------------------------------------------------------------------------------
// mask: fp32 [B][C][512][32] (memory behind the decoder's permuted view); obs, gt: fp32 [B][512][32][C(obs: 2)].
// use_mix = 1: pred = mask * (exp(obs) - 1) (binaural), 0: pred = mask (monaural).  loss_sum += sum |pred - gt|;
// dmask = sign(pred - gt) * mix * grad_scale  (grad_scale = 1 / numel for the mean reduction)
__global__ void l1_mask_loss_kernel(const float* __restrict__ mask, const float* __restrict__ obs, const float* __restrict__ gt,
                                    int B, int C, int use_mix, float grad_scale, double* __restrict__ loss_sum, float* __restrict__ dmask) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const long long total = (long long)B * C * 16384;
  float part = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i & 31), f = (int)((i >> 5) & 511);
    const int c = (int)((i >> 14) % C);
    const long long b = i / ((long long)C * 16384);
    const float mix = use_mix ? expf(obs[((b * 512 + f) * 32 + w) * 2 + c]) - 1.f : 1.f;
    const float d = mask[i] * mix - gt[((b * 512 + f) * 32 + w) * C + c];
    part += fabsf(d);
    if (dmask) dmask[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * mix * grad_scale;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss_sum, (double)part);
}

// ---- class-plane weight gradient of encoder layer 0 -------------------------------------------------------------------------------
// dW[co][32][kh][kw] = sum_b (cls_b + 1) * sum over output pixels whose tap (kh, kw) lies inside the image of dz[b, oy, ox, co]
// dz: haloed [B][18][18][ld] (+coff).  One block per (co), threads over (b, oy, ox); 16 taps reduced in shared memory.
__global__ void wgrad_cls_kernel(const __nv_bfloat16* __restrict__ dz, int ld, int coff, const float* __restrict__ cls, int B,
                                 float* __restrict__ dw, int CinFull) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  // grid = (64 output channels, batch slices); per-tap sums reduced over the block, then one atomic per tap (dw pre-zeroed)
  __shared__ float acc[16];
  if (threadIdx.x < 16) acc[threadIdx.x] = 0.f;
  __syncthreads();
  const int co = blockIdx.x;
  float t[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) t[k] = 0.f;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < B * 256; i += gridDim.y * blockDim.x) {
    const int ox = i & 15, oy = (i >> 4) & 15, b = i >> 8;
    const float v = __bfloat162float(dz[(((size_t)b * 18 + oy + 1) * 18 + ox + 1) * ld + coff + co]) * cls[b];
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int iy = 2 * oy + kh - 1;
      if (iy < 0 || iy >= 32) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ix = 2 * ox + kw - 1;
        if (ix >= 0 && ix < 32) t[kh * 4 + kw] += v;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float s = t[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[k], s);
  }
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(&dw[((size_t)co * CinFull + (CinFull - 1)) * 16 + threadIdx.x], acc[threadIdx.x]);
}

// ---- weight-gradient GEMM on tcgen05 ------------------------------------------------------------------------------------------------
// dW[m][n][kh][kw] = sum over positions (b, y, x) of dense[b, y, x, m] * sampled[b, 2y + kh - 1, 2x + kw - 1, n]
//   Conv2d:          dense = dz (output resolution),  sampled = layer input        -> W[co = m][ci = n]
//   ConvTranspose2d: dense = layer input,            sampled = dz (output res.)   -> W[ci = m][co = n]
// The reduction runs over positions, which are the slow axis of both NHWC operands: both tcgen05 operands are MN-major
// (channels contiguous), read straight from TMA boxes [128 positions][64 channels] (128B swizzle; 32-channel tensors use
// [128][32] boxes with the 64B swizzle).  The sampled operand's stride-2 window is a dense box of the 5-D view
// [B][Hp/2][2][Wp/2][2*ld] of its haloed map (pixel parity = tap parity).
// One CTA = (group of TC consecutive taps, 128 dense channels, BN <= 64 sampled channels, a slice of the positions): the dense
// box of a position chunk is loaded once and multiplied with the TC shifted sampled boxes into TC accumulators (TC * BN = 256
// TMEM columns).  The taps of a group are adjacent in the [..][4][4] weight tensor, so partial sums leave as 16-byte vector
// reductions (red.global.add.v4.f32) - the first version ran one tap per CTA and paid one 4-byte atomic, each on its own
// 32-byte sector, per gradient element and position slice: the L2 atomic rate, not the tensor pipe, set its time.
#define WG_THREADS 192
#define WG_STAGES 2
struct WGLayer {
  int Cd, Cs;               // dense / sampled channel counts (gradient is [Cd][Cs_full][16])
  int Cs_full;              // second dimension of the weight tensor (33 for encoder layer 0)
  int d_coff, s_coff;       // first channel inside the position's channel vector
  int d_pad;                // 1: the dense map is haloed (interior starts at pixel (1, 1)); 0: plain [B][Cd] rows (1x1 maps)
  int s_ld;                 // channel pitch of the sampled map (its 5-D view interleaves the two column parities along channels)
  int Hq;                   // dense resolution (sampled = 2 Hq)
  int nb, rows_y;           // box = [nb images][rows_y rows][Hq columns] = 128 positions (zero-filled past the batch)
  int n_kchunks;            // position chunks in total: ceil(B / nb) * (Hq / rows_y)
  int kchunks_per_cta;
  int BN;                   // sampled channels per CTA (32 | 64)
  int TC;                   // taps per CTA (256 / BN: 8 | 4), taps blockIdx.x * TC ..
  int s_sw64;               // sampled operand uses 32-channel boxes / 64B swizzle
  int tap_mask;             // bit (kh*4+kw): tap can touch the image (1x1 / 2x2 maps only see the centre taps)
  float* dw;
};
namespace wg {
using namespace ug;
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// MN-major operand: 64-channel groups (128B swizzle) lbo bytes apart, 8-position groups 1024 B apart; or one 32-channel
// group (64B swizzle), 8-position groups 512 B apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes, bool sw64) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sw64 ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw64 ? 4 : 2) << 61;
  return d;
}
constexpr int kDenseBytes = 16384 * 2;                   // dense: 2 x [128][64]
constexpr int kStageBytes = kDenseBytes + 65536;         // + TC sampled boxes: 4 x [128][64] or 8 x [128][32]
constexpr int kSmemBytes = WG_STAGES * kStageBytes + 1024 + 256;
}  // namespace wg

__global__ void __launch_bounds__(WG_THREADS, 1)
unet_wgrad_kernel(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_s, const __grid_constant__ WGLayer L) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  extern __shared__ unsigned char wg_smem_raw[];
  const uint32_t raw = ug::smem_u32(wg_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* tiles_ptr = wg_smem_raw + (tiles - raw);
  const uint32_t bars = tiles + WG_STAGES * wg::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
  const uint32_t accum_bar = bars + 8u * (2 * WG_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tiles_ptr + WG_STAGES * wg::kStageBytes + 8 * (2 * WG_STAGES + 1));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // blockIdx.x = tap group, blockIdx.y = (m tile, n tile), blockIdx.z = position slice
  const int tap0 = blockIdx.x * L.TC;
  const int n_tiles_n = (L.Cs + L.BN - 1) / L.BN;
  const int mt = blockIdx.y / n_tiles_n, nt = blockIdx.y % n_tiles_n;
  const int m0 = mt * 128, n0 = nt * L.BN;
  const int kc0 = blockIdx.z * L.kchunks_per_cta;
  int kc1 = kc0 + L.kchunks_per_cta;
  if (kc1 > L.n_kchunks) kc1 = L.n_kchunks;
  const uint32_t grp_mask = ((uint32_t)L.tap_mask >> tap0) & ((1u << L.TC) - 1u);      // live taps of this group
  const bool live = grp_mask != 0 && kc0 < kc1;
  const uint32_t tmem_cols = 256;

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { ug::mbar_init(full_bar(s), 1); ug::mbar_init(empty_bar(s), 1); }
    ug::mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ug::smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ug::tc_fence_before();
  __syncthreads();
  ug::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int chunks_per_img_grp = L.Hq / L.rows_y;          // position chunks per group of nb images
  const uint32_t s_box_bytes = L.s_sw64 ? 128u * 64u : 128u * 128u;
  const int n_dboxes = (L.Cd - m0) >= 128 ? 2 : 1;
  const uint32_t stage_tx = (uint32_t)n_dboxes * 16384u + (uint32_t)__popc(grp_mask) * s_box_bytes;

  if (live) {
    if (warp == 5) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        int it = 0;
        for (int kc = kc0; kc < kc1; ++kc, ++it) {
          const int s = it % WG_STAGES;
          if (it >= WG_STAGES) ug::mbar_wait(empty_bar(s), (uint32_t)((it / WG_STAGES - 1) & 1));
          const int b0 = (kc / chunks_per_img_grp) * L.nb, y0 = (kc % chunks_per_img_grp) * L.rows_y;
          const uint32_t st = tiles + s * wg::kStageBytes;
          wg::mbar_expect_tx(full_bar(s), stage_tx);
          for (int j = 0; j < n_dboxes; ++j)     // dense: interior pixels (1 + y0 .., 1 ..) of images b0 ..
            wg::tma_load_4d(st + j * 16384, &tm_d, L.d_coff + m0 + j * 64, L.d_pad, L.d_pad + y0, b0, full_bar(s));
          // sampled: pixel (2y + kh, 2x + kw) of the haloed map = cell (y + (kh >> 1), x + (kw >> 1)), parity (kh & 1, kw & 1)
          for (int j = 0; j < L.TC; ++j) {
            if (!((grp_mask >> j) & 1u)) continue;
            const int kh = (tap0 + j) >> 2, kw = (tap0 + j) & 3;
            wg::tma_load_5d(st + wg::kDenseBytes + j * s_box_bytes, &tm_s, (kw & 1) * L.s_ld + L.s_coff + n0, (kw >> 1), (kh & 1), y0 + (kh >> 1), b0, full_bar(s));
          }
        }
      }
    } else if (warp == 4) {
      // ------------------------------ MMA issuer: D_j[m, n] += dense[k, m] * sampled_j[k, n], both MN-major ------------------------------
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(L.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int it = 0;
      for (int kc = kc0; kc < kc1; ++kc, ++it) {
        const int s = it % WG_STAGES;
        ug::mbar_wait(full_bar(s), (uint32_t)((it / WG_STAGES) & 1));
        ug::tc_fence_after();
        const uint32_t st = tiles + s * wg::kStageBytes;
        for (int j = 0; j < L.TC; ++j) {
          if (!((grp_mask >> j) & 1u)) continue;
          const uint32_t sb = st + wg::kDenseBytes + j * s_box_bytes, d = tmem_base + (uint32_t)(j * L.BN);
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // 128 positions = 8 steps of K = 16
            const uint64_t adesc = wg::desc_mn(st + k * 2048, 16384, false);
            const uint64_t bdesc = L.s_sw64 ? wg::desc_mn(sb + k * 1024, 0, true) : wg::desc_mn(sb + k * 2048, 16384, false);
            us::umma_bf16_elect(d, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        us::umma_commit_elect(empty_bar(s));
      }
      us::umma_commit_elect(accum_bar);
    } else {
      // ------------------------------ epilogue: accumulator row m, taps 4g .. 4g + 3 of column n -> one 16-byte reduction ------------------------------
      ug::mbar_wait(accum_bar, 0);
      ug::tc_fence_after();
      const int m = m0 + warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int g = 0; g < L.TC; g += 4) {
        const uint32_t gm = (grp_mask >> g) & 15u;
        if (gm == 0) continue;
        for (int cb = 0; cb < L.BN; cb += 16) {
          uint32_t v[4][16];
#pragma unroll
          for (int j = 0; j < 4; ++j) ug::tmem_ld16_nowait(trow + (uint32_t)((g + j) * L.BN + cb), v[j]);
          ug::tmem_ld_wait();
          if (m < L.Cd) {
            float* o = L.dw + ((size_t)m * L.Cs_full + n0 + cb) * 16 + tap0 + g;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + cb + i < L.Cs)       // accumulators of taps that cannot touch the image were never written: they add zero
                wg::red_add_v4(o + (size_t)i * 16, (gm & 1u) ? __uint_as_float(v[0][i]) : 0.f, (gm & 2u) ? __uint_as_float(v[1][i]) : 0.f,
                               (gm & 4u) ? __uint_as_float(v[2][i]) : 0.f, (gm & 8u) ? __uint_as_float(v[3][i]) : 0.f);
          }
        }
      }
      ug::tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    ug::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

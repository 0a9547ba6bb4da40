// C-ABI entry points of the passive-separator U-Net (include/m2h.h): weight packing, input packing, layer launches.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <mutex>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/m2h.h"
#include "m2h_internal.h"
#include "unet_gemm.cuh"
#include "unet_slab.cuh"
#include "unet_gemm_tma.cuh"

// ------------------------------------------------------------------------------------------------
// static network description (separator_cnn.py:44-50, 128-135)
// ------------------------------------------------------------------------------------------------
static const int kEncCin[5] = {32, 64, 128, 256, 512};       // GEMM input channels (the class plane of layer 0 is an epilogue term)
static const int kEncCout[5] = {64, 128, 256, 512, 512};
static const int kEncTaps[5] = {16, 16, 16, 16, 4};          // layer 4 sees a 2x2 input: only kh, kw in {1, 2} touch it
static const int kDecCin[5] = {512, 1024, 512, 256, 128};
static const int kDecCoutBase[5] = {512, 256, 128, 64, 32};  // last entry: 32 (binaural) / 16 (bin2mono, padded to 32)

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct EncLayout {
  size_t w[5], scale[5], bias[5], cls_w, total;
};
struct DecLayout {
  size_t w[5], scale[5], bias[5], head_w, head_wf, head_b, total;
};

static EncLayout enc_layout() {
  EncLayout L;
  size_t o = 0;
  // layer 0 runs on the slab kernel: 12 weight tiles of [64 co][64 k] (unet_slab.cuh)
  for (int i = 0; i < 5; ++i) { L.w[i] = o; o = align256(o + (i == 0 ? (size_t)12 * 64 * 64 * 2 : (size_t)kEncCout[i] * kEncTaps[i] * kEncCin[i] * 2)); }
  for (int i = 0; i < 5; ++i) { L.scale[i] = o; o = align256(o + kEncCout[i] * 4); L.bias[i] = o; o = align256(o + kEncCout[i] * 4); }
  L.cls_w = o; o = align256(o + 9 * 64 * 4);
  L.total = o;
  return L;
}
static DecLayout dec_layout() {
  DecLayout L;
  size_t o = 0;
  // layer 0: [4 pixels * 512][512]; layers 1-4: [4 phases][Cout][4 * Cin]; Cout of layer 4 padded to 32
  for (int i = 0; i < 5; ++i) { L.w[i] = o; o = align256(o + (size_t)4 * kDecCoutBase[i] * (i == 0 ? 1 : 4) * kDecCin[i] * 2); }
  for (int i = 0; i < 5; ++i) { L.scale[i] = o; o = align256(o + kDecCoutBase[i] * 4); L.bias[i] = o; o = align256(o + kDecCoutBase[i] * 4); }
  L.head_w = o; o = align256(o + 32 * 64 * 2);
  L.head_wf = o; o = align256(o + 32 * 32 * 4);
  L.head_b = o; o = align256(o + 32 * 4);
  L.total = o;
  return L;
}

extern "C" size_t m2h_unet_enc_packed_bytes(int) { return enc_layout().total; }
extern "C" size_t m2h_unet_dec_packed_bytes(int) { return dec_layout().total; }

// workspace (bf16 elements per sample)
// every feature map carries a zero border of one pixel ("haloed" NHWC); x0 is stored space-to-depth (16x16 cells of 2x2 pixels)
static const int kWsHp[5] = {18, 18, 10, 6, 4};
static const int kWsLd[5] = {128, 128, 256, 512, 1024};
static const size_t kWsElems[7] = {18 * 18 * 128, 18 * 18 * 128, 10 * 10 * 256, 6 * 6 * 512, 4 * 4 * 1024, 512, 0};
extern "C" int m2h_unet_workspace_layout(int B, size_t offsets[7]) {
  if (B < 0 || !offsets) { m2h_set_error("m2h_unet_workspace_layout: bad arguments"); return M2H_ERR_ARG; }
  size_t o = 0;
  for (int i = 0; i < 7; ++i) { offsets[i] = o; o = align256(o + (size_t)B * kWsElems[i] * 2); }
  return M2H_OK;
}
extern "C" size_t m2h_unet_workspace_bytes(int B) {
  size_t o = 0;
  for (int i = 0; i < 7; ++i) o = align256(o + (size_t)(B > 0 ? B : 0) * kWsElems[i] * 2);
  return o;
}

// ------------------------------------------------------------------------------------------------
// packing kernels
// ------------------------------------------------------------------------------------------------
// Conv2d weight [Cout][CinFull][4][4] -> bf16 [Cout][n_taps*Cin], taps = all 16 (kh*4+kw) or the 4 central ones.
// One thread per (n, c): its 4x4 kernel is one contiguous 64-byte read (a warp reads 2 KB in a row), and for each tap the
// warp writes 32 consecutive bf16.  (The first version walked the OUTPUT index: 64-byte-strided 4-byte reads, 16x read
// amplification - 0.6 ms per training step went into re-packing the weights.)
__global__ void pack_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int CinFull, int Cin, int n_taps) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const size_t total = (size_t)Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int n = (int)(i / Cin);
    float v[16];
    if (c < CinFull) {
      const float* src = w + ((size_t)n * CinFull + c) * 16;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
          v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __ldg(src + q);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = 0.f;
    }
    __nv_bfloat16* o = out + (size_t)n * n_taps * Cin + c;
    if (n_taps == 16) {
#pragma unroll
      for (int t = 0; t < 16; ++t) o[(size_t)t * Cin] = __float2bfloat16_rn(v[t]);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) o[(size_t)t * Cin] = __float2bfloat16_rn(v[(1 + (t >> 1)) * 4 + 1 + (t & 1)]);
    }
  }
}
// class plane of encoder layer 0: cls_w[cy][cx][n] = sum over taps that stay inside the image of w[n][32][kh][kw]
__global__ void pack_cls_w_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int CinFull) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * Cout) return;
  const int n = i % Cout, cx = (i / Cout) % 3, cy = i / (3 * Cout);
  float s = 0.f;
  for (int kh = 0; kh < 4; ++kh) {
    if ((cy == 0 && kh == 0) || (cy == 2 && kh == 3)) continue;
    for (int kw = 0; kw < 4; ++kw) {
      if ((cx == 0 && kw == 0) || (cx == 2 && kw == 3)) continue;
      s += w[(((size_t)n * CinFull + (CinFull - 1)) * 4 + kh) * 4 + kw];
    }
  }
  out[i] = s;
}
// sub-pixel phase decomposition of ConvTranspose2d(k4, s2, p1): output row 2q+p reads input rows q (tap 0) and
// q-1 (p = 0) / q+1 (p = 1) (tap 1) with kernel rows {1, 3} / {2, 0}.
__host__ __device__ inline int deconv_k(int p, int t) { return t == 0 ? (p == 0 ? 1 : 2) : (p == 0 ? 3 : 0); }
__host__ __device__ inline int deconv_d(int p, int t) { return t == 0 ? 0 : (p == 0 ? -1 : 1); }
// ConvTranspose2d weight [Cin][CoutReal][4][4] -> bf16 [4 phases][CoutPad][4 taps * Cin].
// Reads want n fastest (the 4x4 kernels of consecutive output channels are contiguous), writes want c fastest: a block
// transposes a 32 (c) x 32 (n) tile of kernels through shared memory; every (phase, tap) pair is exactly one kernel tap.
__global__ void __launch_bounds__(256) pack_deconv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int CoutReal, int CoutPad) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  __shared__ __nv_bfloat16 tile[16][32][33];            // [kh*4+kw][c][n], one pad column
  const int tiles_n = (CoutPad + 31) / 32, tiles_c = (Cin + 31) / 32;
  for (int tile_id = blockIdx.x; tile_id < tiles_n * tiles_c; tile_id += gridDim.x) {
    const int c0 = (tile_id / tiles_n) * 32, n0 = (tile_id % tiles_n) * 32;
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int cl = e >> 5, nl = e & 31, c = c0 + cl, n = n0 + nl;
      float v[16];
      if (c < Cin && n < CoutReal) {
        const float* src = w + ((size_t)c * CoutReal + n) * 16;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __ldg(src + q);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) tile[q][cl][nl] = __float2bfloat16_rn(v[q]);
    }
    __syncthreads();
    // 16 (phase, tap) x 32 n rows of 32 consecutive c each
    for (int e = threadIdx.x; e < 16 * 32 * 32; e += 256) {
      const int cl = e & 31, nl = (e >> 5) & 31, pt = e >> 10;
      const int ph = pt >> 2, t = pt & 3;
      const int kh = deconv_k(ph >> 1, t >> 1), kw = deconv_k(ph & 1, t & 1);
      const int c = c0 + cl, n = n0 + nl;
      if (c < Cin && n < CoutPad) out[(((size_t)ph * CoutPad + n) * 4 + t) * Cin + c] = tile[kh * 4 + kw][cl][nl];
    }
    __syncthreads();
  }
}
// decoder layer 0 (1x1 input): output pixel (oy, ox) = weight tap (oy+1, ox+1): bf16 [4 pixels * Cout][Cin]
__global__ void pack_deconv0_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cout) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const size_t total = (size_t)4 * Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int n = (int)((i / Cin) % Cout);
    const int pix = (int)(i / ((size_t)Cin * Cout));
    out[i] = __float2bfloat16_rn(w[(((size_t)c * Cout + n) * 4 + (pix >> 1) + 1) * 4 + (pix & 1) + 1]);
  }
}
// head Conv2d 1x1 [C][C] -> bf16 [32][64] zero padded; bias -> fp32 [32]
__global__ void pack_head_kernel(const float* __restrict__ w, const float* __restrict__ b, __nv_bfloat16* __restrict__ wo, float* __restrict__ wf, float* __restrict__ bo, int C) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 32 * 64) {
    const int n = i / 64, c = i % 64;
    wo[i] = __float2bfloat16_rn(n < C && c < C ? w[n * C + c] : 0.f);
    if (c < 32) wf[n * 32 + c] = __bfloat162float(wo[i]);       // bf16-rounded copy for the fused (epilogue) head
  }
  if (i < 32) bo[i] = i < C ? b[i] : 0.f;
}
// BatchNorm (inference): y = x * scale + bias
__global__ void fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ scale, float* __restrict__ bias, int C, int Cpad) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cpad) return;
  if (i < C) {
    const float s = g[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    bias[i] = b[i] - mean[i] * s;
  } else { scale[i] = 0.f; bias[i] = 0.f; }
}

// ------------------------------------------------------------------------------------------------
// slab-kernel weight tiles (unet_slab.cuh): bf16 [rows][64 k], rows in the order the kernel's items consume them
// ------------------------------------------------------------------------------------------------
struct SlabPackBlock { int16_t kh, kw, kw2, chunk; };      // one block of Cout rows
struct SlabPackDesc { int n_blocks, Cout, CoutReal, CinFull; SlabPackBlock blk[64]; };
// ConvTranspose2d weight [Cin][CoutReal][4][4]: row block = (phase, tap) -> kernel position (kh, kw); k = channel inside the chunk
__host__ __device__ inline float slab_deconv_w_value(const float* w, const SlabPackDesc& D, size_t i) {
  const int k = (int)(i & 63);
  const int row = (int)(i >> 6);
  const int n = row % D.Cout;
  const SlabPackBlock b = D.blk[row / D.Cout];
  const int c = b.chunk * 64 + k;
  return n < D.CoutReal ? w[(((size_t)c * D.CoutReal + n) * 4 + b.kh) * 4 + b.kw] : 0.f;
}
__global__ void pack_slab_deconv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, const SlabPackDesc D) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const size_t total = (size_t)D.n_blocks * D.Cout * 64;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(slab_deconv_w_value(w, D, i));
}
// Conv2d weight [Cout][CinFull][4][4], Cin = 32, space-to-depth input: k < 32 -> kernel column kw, k >= 32 -> kw2 (or zero)
__host__ __device__ inline float slab_conv_w_value(const float* w, const SlabPackDesc& D, size_t i) {
  const int k = (int)(i & 63);
  const int row = (int)(i >> 6);
  const int n = row % D.Cout;
  const SlabPackBlock b = D.blk[row / D.Cout];
  const int kw = k < 32 ? b.kw : b.kw2;
  return kw >= 0 ? w[(((size_t)n * D.CinFull + (k & 31)) * 4 + b.kh) * 4 + kw] : 0.f;
}
__global__ void pack_slab_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, const SlabPackDesc D) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const size_t total = (size_t)D.n_blocks * D.Cout * 64;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(slab_conv_w_value(w, D, i));
}

// Items of a transposed-conv layer: per 64-channel chunk, one weight tile per distinct tap shift (dy, dx); the phases that
// use the shift are stacked along N.  Fills the kernel's item list and the matching packing descriptor.
static void build_deconv_items(int Cin, int Cout, int CoutReal, int Wp, USLayer& L, SlabPackDesc& D) {
  static const int order[9][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
  L.n_items = 0; L.n_chunks = Cin / 64; L.n_phase = 4; L.Cout = Cout;
  D.n_blocks = 0; D.Cout = Cout; D.CoutReal = CoutReal; D.CinFull = 0;
  bool written[4] = {false, false, false, false};
  for (int c = 0; c < L.n_chunks; ++c) {
    for (int s = 0; s < 9; ++s) {
      const int dy = order[s][0], dx = order[s][1];
      int ph[4], n_ph = 0;
      for (int py = 0; py < 2; ++py) {
        if (!(dy == 0 || dy == (py == 0 ? -1 : 1))) continue;
        for (int px = 0; px < 2; ++px)
          if (dx == 0 || dx == (px == 0 ? -1 : 1)) ph[n_ph++] = py * 2 + px;
      }
      int p = 0;
      while (p < n_ph) {               // split into items of <= 128 weight rows and <= 2 runs of consecutive phases
        USItem& it = L.items[L.n_items++];
        memset(&it, 0, sizeof(it));
        it.a_shift = (int16_t)((Wp + 1) + dy * Wp + dx); it.chunk = (int8_t)c; it.a_k16 = 0; it.n_k16 = 4;
        int rows = 0;
        while (p < n_ph && rows + Cout <= 128) {
          const bool extend = it.n_parts > 0 && ph[p] == ph[p - 1] + 1 && written[ph[p]] == written[ph[p - 1]] &&
                              it.part[it.n_parts - 1].b_row + it.part[it.n_parts - 1].n == rows;
          if (!extend) {
            if (it.n_parts == 2) break;
            USPart& pt = it.part[it.n_parts++];
            pt.b_row = (int16_t)rows; pt.n = 0; pt.tmem_col = (int16_t)(ph[p] * Cout); pt.init = written[ph[p]] ? 0 : 1;
          }
          it.part[it.n_parts - 1].n = (int16_t)(it.part[it.n_parts - 1].n + Cout);
          SlabPackBlock& b = D.blk[D.n_blocks++];
          b.kh = (int16_t)deconv_k(ph[p] >> 1, dy == 0 ? 0 : 1); b.kw = (int16_t)deconv_k(ph[p] & 1, dx == 0 ? 0 : 1); b.kw2 = -1; b.chunk = (int16_t)c;
          rows += Cout;
          ++p;
        }
        it.b_rows = (int16_t)rows;
        for (int q = 0; q < it.n_parts; ++q)
          for (int col = it.part[q].tmem_col; col < it.part[q].tmem_col + it.part[q].n; col += Cout) written[col / Cout] = true;
      }
    }
  }
}
// Items of encoder layer 0 (Conv2d k4 s2 p1, Cin = 32) on the space-to-depth input: chunk = row parity py,
// channel inside the chunk = px*32 + c.  Kernel rows kh with that parity: py = 0 -> kh = 1 (dy 0), 3 (dy +1); py = 1 -> kh = 0 (dy -1), 2 (dy 0).
static void build_conv0_items(int Wp, int CinFull, USLayer& L, SlabPackDesc& D) {
  L.n_items = 0; L.n_chunks = 2; L.n_phase = 1; L.Cout = 64;
  D.n_blocks = 0; D.Cout = 64; D.CoutReal = 64; D.CinFull = CinFull;
  bool first = true;
  for (int py = 0; py < 2; ++py) {
    for (int j = 0; j < 2; ++j) {
      const int kh = py == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 0 : 2);
      const int dy = py == 0 ? (j == 0 ? 0 : 1) : (j == 0 ? -1 : 0);
      // (kw = 1 | kw = 2): dx = 0, both column parities = the whole 64-channel chunk; kw = 0: px = 1, dx = -1; kw = 3: px = 0, dx = +1
      const int dxs[3] = {0, -1, 1}, ak[3] = {0, 2, 0}, nk[3] = {4, 2, 2}, kw1[3] = {1, 0, 3}, kw2[3] = {2, -1, -1};
      for (int t = 0; t < 3; ++t) {
        USItem& it = L.items[L.n_items++];
        memset(&it, 0, sizeof(it));
        it.a_shift = (int16_t)((Wp + 1) + dy * Wp + dxs[t]); it.chunk = (int8_t)py; it.a_k16 = (int8_t)ak[t]; it.n_k16 = (int8_t)nk[t];
        it.n_parts = 1; it.b_rows = 64;
        it.part[0].b_row = 0; it.part[0].n = 64; it.part[0].tmem_col = 0; it.part[0].init = first ? 1 : 0;
        first = false;
        SlabPackBlock& b = D.blk[D.n_blocks++];
        b.kh = (int16_t)kh; b.kw = (int16_t)kw1[t]; b.kw2 = (int16_t)kw2[t]; b.chunk = (int16_t)py;
      }
    }
  }
}

// obs fp32 [B,512,32,2] -> x0 bf16, space-to-depth + halo: [B][18][18][128], cell (h/2+1, w/2+1), channel = ((h&1)*2 + (w&1))*32 + ear*16 + slice,
// where h = freq % 32, slice = freq / 32 (separator_cnn.py:84-90);
// bin2mono: x = log1p(max(mask * (exp(obs) - 1), 0)) first (separator_cnn.py:73-79)
struct PackInParams {
  const float* obs; const float* masks; long long ms[4]; __nv_bfloat16* x0; int B;
  __nv_bfloat16* x0_plain;     // optional second copy, plain haloed NHWC [B][34][34][32] (weight-gradient operand)
};
__global__ void pack_input_kernel(const PackInParams P) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (b, h, w)
  if (i >= P.B * 1024) return;
  const int w = i & 31, h = (i >> 5) & 31, b = i >> 10;
  __align__(16) __nv_bfloat16 v[32];
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int f = s * 32 + h;
    const float2 o = *reinterpret_cast<const float2*>(P.obs + (((size_t)b * 512 + f) * 32 + w) * 2);
    float l = o.x, r = o.y;
    if (P.masks) {
      const float* m = P.masks + (long long)b * P.ms[0] + (long long)f * P.ms[1] + (long long)w * P.ms[2];
      l = log1pf(fmaxf(m[0] * (expf(o.x) - 1.f), 0.f));
      r = log1pf(fmaxf(m[P.ms[3]] * (expf(o.y) - 1.f), 0.f));
    }
    v[s] = __float2bfloat16_rn(l);
    v[16 + s] = __float2bfloat16_rn(r);
  }
  uint4* dst = reinterpret_cast<uint4*>(P.x0 + (((size_t)b * 18 + (h >> 1) + 1) * 18 + (w >> 1) + 1) * 128 + ((h & 1) * 2 + (w & 1)) * 32);
  const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = src[k];
  if (P.x0_plain) {
    uint4* d2 = reinterpret_cast<uint4*>(P.x0_plain + (((size_t)b * 34 + h + 1) * 34 + w + 1) * 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) d2[k] = src[k];
  }
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __bfloat162float(in[i]);
}
// zero the one-pixel border of the five haloed feature maps (the kernels only ever write interiors): one thread per
// (border position, 16-byte chunk)
struct HaloParams { __nv_bfloat16* buf[5]; int Hp[5], ld8[5]; long long start[6]; };
__global__ void clear_halo_kernel(const HaloParams P) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P.start[5]; i += (long long)gridDim.x * blockDim.x) {
    int t = 0;
    while (i >= P.start[t + 1]) ++t;
    const long long j = i - P.start[t];
    const int hp = P.Hp[t], ld8 = P.ld8[t], hc = 4 * hp - 4;
    const int ch = (int)(j % ld8);
    const long long bh = j / ld8;
    const int h = (int)(bh % hc);
    const long long b = bh / hc;
    int y, x;
    if (h < hp) { y = 0; x = h; }
    else if (h < 2 * hp) { y = hp - 1; x = h - hp; }
    else { const int k = h - 2 * hp; y = 1 + (k >> 1); x = (k & 1) ? hp - 1 : 0; }
    reinterpret_cast<uint4*>(P.buf[t])[((b * hp + y) * hp + x) * ld8 + ch] = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
// Launch with programmatic stream serialisation (see ug::pdl_wait): the forward pass is a chain of 14 short kernels, and
// each one's prologue now overlaps the tail of its predecessor.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("M2H_NO_PDL"); enabled = (e && atoi(e)) ? 0 : 1; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static bool g_unet_attr_set[M2H_MAX_DEVICES];
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

static const int kSlabMaxSmem = 227 * 1024;
static int unet_prepare_device() {
  int dev;
  int rc = m2h_current_device_checked(&dev);
  if (rc != M2H_OK) return rc;
  if (!g_unet_attr_set[dev]) {
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlabMaxSmem));
    // the default carve-out left room for one CTA per SM (first ncu capture: 100 KB config, 5 warps/SM): ask for the maximum
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_slab_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    g_unet_attr_set[dev] = true;
  }
  if (!g_encode_tiled) {
    // the driver's tensor-map encoder, resolved through the runtime (libcuda is not linked)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    M2H_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { m2h_set_error("cuTensorMapEncodeTiled is not available from this driver"); return M2H_ERR_CUDA; }
    g_encode_tiled = (EncodeTiledFn)fn;
  }
  return M2H_OK;
}
static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

// 2D bf16 tensor map [rows][cols] (cols contiguous, row pitch ld elements), box = [box_rows][64], 128-byte swizzle
static int make_tmap(CUtensorMap* tm, const void* base, size_t rows, int cols, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { m2h_set_error("cuTensorMapEncodeTiled failed (%d): rows=%zu cols=%d ld=%d box_rows=%d", (int)r, rows, cols, ld, box_rows); return M2H_ERR_CUDA; }
  return M2H_OK;
}
static int make_tmap_nd(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, bool sw64) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { m2h_set_error("cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r); return M2H_ERR_CUDA; }
  return M2H_OK;
}
// TMA-fed layer launch (unet_gemm_tma.cuh): Cin >= 64 (64-channel boxes, 128B swizzle) or the 32-channel stride-2 conv (64B swizzle), bf16 NHWC output
template <int BN, bool DEEP>
static int launch_tma_cfg(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const UGLayer& L, const UTGeom& G, dim3 grid, cudaStream_t st) {
  static bool attr_set[M2H_MAX_DEVICES];
  int dev = 0;
  M2H_CUDA_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_tma_kernel<BN, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, UTCfg<BN, DEEP>::kSmemBytes));
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_tma_kernel<BN, DEEP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[dev] = true;
  }
  M2H_CUDA_CHECK(launch_pdl(unet_gemm_tma_kernel<BN, DEEP>, grid, dim3(UT_THREADS), (size_t)UTCfg<BN, DEEP>::kSmemBytes, st, tm_a, tm_w, L, G));
  return M2H_OK;
}
template <int BN>
static int launch_tma_persistent(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const UGLayer& L, const UTGeom& G, int n_mtiles, int sms, cudaStream_t st) {
  static bool attr_set[M2H_MAX_DEVICES];
  int dev = 0;
  M2H_CUDA_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_tma_persistent_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, UTPCfg<BN>::kSmemBytes));
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_tma_persistent_kernel<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[dev] = true;
  }
  const int n_mpairs = n_mtiles / 2, n_ntiles = L.Cout / BN, n_tiles = n_mpairs * n_ntiles * L.n_phase;
  const int grid = n_tiles < sms ? n_tiles : sms;
  M2H_CUDA_CHECK(launch_pdl(unet_gemm_tma_persistent_kernel<BN>, dim3(grid), dim3(UT_THREADS), (size_t)UTPCfg<BN>::kSmemBytes, st, tm_a, tm_w, L, G,
                            n_mpairs, n_ntiles, n_tiles));
  return M2H_OK;
}
template <int BN>
static int launch_tma_bn(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const UGLayer& L, const UTGeom& G, int n_mtiles, cudaStream_t st) {
  int dev = 0, sms = 0;
  M2H_CUDA_CHECK(cudaGetDevice(&dev));
  M2H_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static int persistent = -1;
  if (persistent < 0) { const char* e = getenv("M2H_TMA_PERSISTENT"); persistent = e ? atoi(e) : 1; }
  // 256-row super-tiles (shared weight tile) when there is at least one pair per SM (enc1, dec2 at B = 1024: 35 -> 34 us, 48 -> 43 us);
  // with fewer pairs than SMs the one-tile kernels are faster (enc2: 26 us vs 28 us)
  if (persistent && (n_mtiles & 1) == 0 && (long long)(n_mtiles / 2) * (L.Cout / BN) * L.n_phase >= sms)
    return launch_tma_persistent<BN>(tm_a, tm_w, L, G, n_mtiles, sms, st);
  dim3 grid(n_mtiles, L.Cout / BN, L.n_phase);
  // at most one CTA per SM anyway: give it the whole shared memory as ring depth
  if ((long long)grid.x * grid.y * grid.z <= sms) return launch_tma_cfg<BN, true>(tm_a, tm_w, L, G, grid, st);
  return launch_tma_cfg<BN, false>(tm_a, tm_w, L, G, grid, st);
}
static int launch_layer_auto(const UGLayer& L, int bn, int B, cudaStream_t st) {
  const bool k32 = L.Cin == 32 && L.in_stride == 2 && L.in_ld == 32;      // 32-channel stride-2 conv: 64B-swizzled boxes
  if ((L.Cin < 64 && !k32) || L.out_mode != UG_OUT_NHWC_BF16 || L.cls != nullptr || (bn != 64 && bn != 128) || L.Hq != L.Wq || L.Hq > 16 || L.Kvalid != L.Kpad) {
    m2h_set_error("unet layer: no kernel for Cin=%d in_stride=%d in_ld=%d Hq=%dx%d bn=%d out_mode=%d", L.Cin, L.in_stride, L.in_ld, L.Hq, L.Wq, bn, L.out_mode);
    return M2H_ERR_ARG;
  }
  if (L.Cout % bn != 0 || L.M <= 0) { m2h_set_error("unet layer: unsupported shape Cout=%d bn=%d", L.Cout, bn); return M2H_ERR_ARG; }
  UTGeom G;
  const int Hq = L.Hq;
  G.rows_y = Hq >= 8 ? 8 : Hq; G.nb = 128 / (Hq * G.rows_y); G.tiles_per_grp = Hq / G.rows_y; G.conv = L.in_stride == 2 ? 1 : 0; G.B = B; G.k32 = k32 ? 1 : 0;
  const int n_mtiles = ((B + G.nb - 1) / G.nb) * G.tiles_per_grp;
  // 256-wide N tiles (opt-in: M2H_TMA_BN256_MIN_CTAS = smallest grid that may use them; default 0 = never): the tensor core
  // re-reads the 128-row A tile once per MMA whatever N is, so a 128 x 256 tile moves 25 % fewer shared-memory wavefronts per
  // FLOP than two 128 x 128 tiles.  Measured at B = 1024 (enc2, dec1: 128 CTAs of 256 columns, 4 stages, one
  // CTA per SM, against 256 CTAs of 128 columns, two per SM): U-Net forward 0.505 vs 0.494 ms - two co-resident CTAs hide each
  // other's pipeline bubbles better than one wide CTA saves on operand reads.
  {
    static int min_ctas = -1;
    if (min_ctas < 0) { const char* e = getenv("M2H_TMA_BN256_MIN_CTAS"); min_ctas = e ? atoi(e) : 0; }
    int sms = 148;
    if (bn == 128 && min_ctas > 0 && L.Cout % 256 == 0) {
      const long long ctas = (long long)n_mtiles * (L.Cout / 256) * L.n_phase;
      if (ctas >= min_ctas && ctas <= sms) bn = 256;
    }
  }
  alignas(64) CUtensorMap tm_a, tm_w;
  int rc;
  if (G.conv) {
    const int hp = L.Hin + 2;
    cuuint64_t dims[5] = {(cuuint64_t)2 * L.in_ld, (cuuint64_t)hp / 2, 2, (cuuint64_t)hp / 2, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)2 * L.in_ld * 2, (cuuint64_t)hp * L.in_ld * 2, (cuuint64_t)2 * hp * L.in_ld * 2, (cuuint64_t)hp * hp * L.in_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)(k32 ? 32 : 64), (cuuint32_t)Hq, 1, (cuuint32_t)G.rows_y, (cuuint32_t)G.nb};
    rc = make_tmap_nd(&tm_a, L.in, 5, dims, strides, box, k32);
  } else {
    const int hp = L.Hin + 2 * L.in_pad;
    cuuint64_t dims[4] = {(cuuint64_t)L.Cin, (cuuint64_t)hp, (cuuint64_t)hp, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)L.in_ld * 2, (cuuint64_t)hp * L.in_ld * 2, (cuuint64_t)hp * hp * L.in_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Hq, (cuuint32_t)G.rows_y, (cuuint32_t)G.nb};
    rc = make_tmap_nd(&tm_a, L.in, 4, dims, strides, box, false);
  }
  if (rc != M2H_OK) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)L.Kpad, (cuuint64_t)L.n_phase * L.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)L.Kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)(k32 ? 32 : 64), (cuuint32_t)bn};
    rc = make_tmap_nd(&tm_w, L.w, 2, dims, strides, box, k32);
    if (rc != M2H_OK) return rc;
  }
  if (k32) {      // one-tile kernel, two CTAs per SM (the 256-row persistent form assumes 64-channel boxes)
    dim3 grid(n_mtiles, L.Cout / bn, L.n_phase);
    return bn == 128 ? launch_tma_cfg<128, false>(tm_a, tm_w, L, G, grid, st) : launch_tma_cfg<64, false>(tm_a, tm_w, L, G, grid, st);
  }
  if (bn == 256) {
    dim3 grid(n_mtiles, L.Cout / 256, L.n_phase);
    return launch_tma_cfg<256, true>(tm_a, tm_w, L, G, grid, st);
  }
  return bn == 128 ? launch_tma_bn<128>(tm_a, tm_w, L, G, n_mtiles, st) : launch_tma_bn<64>(tm_a, tm_w, L, G, n_mtiles, st);
}
// geometry + tensor maps + launch of one slab layer; items / epilogue fields of L are filled by the caller
static int launch_slab(USLayer& L, const __nv_bfloat16* in, int in_ld, int Hp, int B, const __nv_bfloat16* w, cudaStream_t st) {
  L.Hp = L.Wp = Hp; L.PI = Hp * Hp; L.B = B;
  L.slab_rows = 128 + 2 * (Hp + 1);
  L.slab_bytes = (L.slab_rows * 128 + 1023) & ~1023;
  L.tile_cols = L.n_phase * L.Cout;
  L.n_buf = 4 * L.tile_cols <= 512 ? 4 : 2;
  L.tmem_cols = 32;
  while (L.tmem_cols < L.n_buf * L.tile_cols) L.tmem_cols *= 2;
  int w_rows = 0;
  for (int i = 0; i < L.n_items; ++i) { L.items[i].w_row = (int16_t)w_rows; w_rows += L.items[i].b_rows; }
  L.w_rows = w_rows;
  L.w_box_rows = L.Cout < 128 ? L.Cout : 128;
  const size_t positions = (size_t)B * L.PI;
  L.n_tiles = (int)((positions + 127) / 128);
  // weights stay resident in shared memory when they fit next to >= 2 tiles' worth of slab chunks, else they stream through a ring
  L.head_bytes = L.out_mode == UG_OUT_HEAD_NCHW_F32 ? US_HEAD_BYTES : 0;
  const int budget = kSlabMaxSmem - 1024 - 512 - US_AUX_BYTES - L.head_bytes;
  // at least one PAIR of tiles' worth of chunks in flight (the two issuers work on two tiles at once; with fewer stages every
  // chunk's load latency is exposed - the last decoder layer ran with 2 stages for 4 boxes per pair)
  static const bool old_rule = getenv("M2H_SLAB_OLD_RESIDENT_RULE") != nullptr;
  const int min_stages = old_rule ? (L.n_chunks > 2 ? L.n_chunks : 2) : 2 * L.n_chunks;
  if (w_rows * 128 + min_stages * L.slab_bytes <= budget) {
    L.n_b_stages = 0;
    L.n_slab_stages = (budget - w_rows * 128) / L.slab_bytes;
  } else {
    L.n_b_stages = 6;
    L.n_slab_stages = (budget - L.n_b_stages * US_B_STAGE_BYTES) / L.slab_bytes;
  }
  if (L.n_slab_stages > US_MAX_STAGES) L.n_slab_stages = US_MAX_STAGES;
  // issue records (unet_slab.cuh::USIssue): what the MMA-issuing warp reads, one item ahead of its use
  if (L.n_b_stages) {                 // the ring depth divides the item count: an item's stage and parity bit are launch constants
    while (L.n_b_stages > 2 && L.n_items % L.n_b_stages != 0) --L.n_b_stages;
    if (L.n_items % L.n_b_stages != 0) { m2h_set_error("slab layer: %d items do not fill a weight ring", L.n_items); return M2H_ERR_ARG; }
  }
  const uint32_t desc_lo0 = 1u << 16;         // low word of ug::smem_desc_sw128(0): LBO field = 1
  for (int i = 0; i < L.n_items; ++i) {
    const USItem& it = L.items[i];
    USIssue& r = L.rec[i];
    memset(&r, 0, sizeof(r));
    if (it.n_k16 != 2 && it.n_k16 != 4) { m2h_set_error("slab layer: unsupported K16 count %d", (int)it.n_k16); return M2H_ERR_ARG; }
    const bool first_of_chunk = i == 0 || L.items[i - 1].chunk != it.chunk;
    const bool last_of_chunk = i + 1 == L.n_items || L.items[i + 1].chunk != it.chunk;
    const int stage = L.n_b_stages ? i % L.n_b_stages : 0, turn = L.n_b_stages ? i / L.n_b_stages : 0;
    r.a_off16 = ((uint32_t)(it.a_shift * 128 + it.a_k16 * 32) >> 4) + desc_lo0;
    r.b_off16 = L.n_b_stages ? (uint32_t)stage * (US_B_STAGE_BYTES >> 4) : (uint32_t)(it.w_row * 128) >> 4;
    r.ctl = (uint32_t)it.n_k16 | ((uint32_t)it.n_parts << 8) | (first_of_chunk ? 1u << 16 : 0u) | (last_of_chunk ? 1u << 17 : 0u) | ((uint32_t)(turn & 1) << 18);
    r.b_bar = 8u * (uint32_t)stage;
    for (int p = 0; p < it.n_parts; ++p) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(it.part[p].n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t col = (uint32_t)it.part[p].tmem_col, brow = (uint32_t)(it.part[p].b_row * 8) + desc_lo0, acc = it.part[p].init ? 0u : 1u;
      if (p == 0) { r.idesc0 = idesc; r.col0 = col; r.brow0 = brow; r.acc0 = acc; } else { r.idesc1 = idesc; r.col1 = col; r.brow1 = brow; r.acc1 = acc; }
    }
  }
  if (L.n_items > US_MAX_ITEMS || L.slab_rows > 256 || L.n_slab_stages < 2 || L.tmem_cols > 512 || w_rows > 32767) { m2h_set_error("slab layer: unsupported geometry"); return M2H_ERR_ARG; }
  alignas(64) CUtensorMap tm_in, tm_w;
  int rc = make_tmap(&tm_in, in, positions, in_ld, in_ld, L.slab_rows);
  if (rc != M2H_OK) return rc;
  rc = make_tmap(&tm_w, w, (size_t)w_rows, 64, 64, L.w_box_rows);
  if (rc != M2H_OK) return rc;
  int dev = 0, sms = 0;
  M2H_CUDA_CHECK(cudaGetDevice(&dev));
  M2H_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)(L.n_tiles < sms ? L.n_tiles : sms);
  M2H_CUDA_CHECK(launch_pdl(unet_slab_kernel, dim3(grid), dim3(US_THREADS), (size_t)us::smem_bytes(L.slab_bytes, L.n_slab_stages, L.n_b_stages, w_rows, L.head_bytes), st, tm_in, tm_w, L));
  return M2H_OK;
}

extern "C" int m2h_unet_pack_encoder(const m2h_unet_enc_params* p, int convert_bin2mono, float bn_eps, void* packed, void* stream) {
  if (!p || !packed) { m2h_set_error("m2h_unet_pack_encoder: null pointer"); return M2H_ERR_ARG; }
  int rc = unet_prepare_device();
  if (rc != M2H_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const EncLayout L = enc_layout();
  char* base = (char*)packed;
  for (int i = 0; i < 5; ++i) {
    if (!p->conv_w[i] || !p->bn_w[i] || !p->bn_b[i] || !p->bn_mean[i] || !p->bn_var[i]) { m2h_set_error("m2h_unet_pack_encoder: null parameter for layer %d", i); return M2H_ERR_ARG; }
    const int cin_full = i == 0 ? (convert_bin2mono ? 32 : 33) : kEncCin[i];
    if (i == 0) {
      USLayer sl; SlabPackDesc pd;
      build_conv0_items(18, cin_full, sl, pd);
      M2H_CUDA_CHECK(launch_pdl(pack_slab_conv_w_kernel, dim3(96), dim3(256), (size_t)(0), st, p->conv_w[0], (__nv_bfloat16*)(base + L.w[0]), pd));
    } else {
      M2H_CUDA_CHECK(launch_pdl(pack_conv_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->conv_w[i], (__nv_bfloat16*)(base + L.w[i]), kEncCout[i], cin_full, kEncCin[i], kEncTaps[i]));
    }
    fold_bn_kernel<<<(kEncCout[i] + 127) / 128, 128, 0, st>>>(p->bn_w[i], p->bn_b[i], p->bn_mean[i], p->bn_var[i], bn_eps,
                                                             (float*)(base + L.scale[i]), (float*)(base + L.bias[i]), kEncCout[i], kEncCout[i]);
  }
  if (!convert_bin2mono) M2H_CUDA_CHECK(launch_pdl(pack_cls_w_kernel, dim3((9 * 64 + 127) / 128), dim3(128), (size_t)(0), st, p->conv_w[0], (float*)(base + L.cls_w), 64, 33));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

extern "C" int m2h_unet_pack_decoder(const m2h_unet_dec_params* p, int convert_bin2mono, float bn_eps, void* packed, void* stream) {
  if (!p || !packed || !p->head_w || !p->head_b) { m2h_set_error("m2h_unet_pack_decoder: null pointer"); return M2H_ERR_ARG; }
  int rc = unet_prepare_device();
  if (rc != M2H_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const DecLayout L = dec_layout();
  char* base = (char*)packed;
  const int c_last = convert_bin2mono ? 16 : 32;
  for (int i = 0; i < 5; ++i) {
    if (!p->deconv_w[i] || !p->bn_w[i] || !p->bn_b[i] || !p->bn_mean[i] || !p->bn_var[i]) { m2h_set_error("m2h_unet_pack_decoder: null parameter for layer %d", i); return M2H_ERR_ARG; }
    const int cout_real = i == 4 ? c_last : kDecCoutBase[i];
    if (i == 0) {
      M2H_CUDA_CHECK(launch_pdl(pack_deconv0_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->deconv_w[0], (__nv_bfloat16*)(base + L.w[0]), 512, 512));
    } else if (i >= 3) {     // slab layers
      USLayer sl; SlabPackDesc pd;
      build_deconv_items(kDecCin[i], kDecCoutBase[i], cout_real, i == 3 ? 10 : 18, sl, pd);
      M2H_CUDA_CHECK(launch_pdl(pack_slab_deconv_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->deconv_w[i], (__nv_bfloat16*)(base + L.w[i]), pd));
    } else {
      M2H_CUDA_CHECK(launch_pdl(pack_deconv_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->deconv_w[i], (__nv_bfloat16*)(base + L.w[i]), kDecCin[i], cout_real, kDecCoutBase[i]));
    }
    fold_bn_kernel<<<(kDecCoutBase[i] + 127) / 128, 128, 0, st>>>(p->bn_w[i], p->bn_b[i], p->bn_mean[i], p->bn_var[i], bn_eps,
                                                                 (float*)(base + L.scale[i]), (float*)(base + L.bias[i]), cout_real, kDecCoutBase[i]);
  }
  M2H_CUDA_CHECK(launch_pdl(pack_head_kernel, dim3((32 * 64 + 127) / 128), dim3(128), (size_t)(0), st, p->head_w, p->head_b, (__nv_bfloat16*)(base + L.head_w), (float*)(base + L.head_wf), (float*)(base + L.head_b), c_last));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

static void layer_defaults(UGLayer& g) {
  memset(&g, 0, sizeof(g));
  g.n_phase = 1; g.in_stride = 1; g.out_stride = 1; g.slope = 1.f; g.n_split = 1 << 30;
  g.out_mode = UG_OUT_NHWC_BF16;
}

struct WsPtrs { __nv_bfloat16* x0; __nv_bfloat16* cat[5]; __nv_bfloat16* bott; };
static WsPtrs ws_ptrs(void* workspace, int B) {
  size_t off[7];
  m2h_unet_workspace_layout(B, off);
  char* ws = (char*)workspace;
  WsPtrs P;
  P.x0 = (__nv_bfloat16*)(ws + off[0]);
  P.cat[0] = nullptr;
  for (int i = 1; i < 5; ++i) P.cat[i] = (__nv_bfloat16*)(ws + off[i]);
  P.bott = (__nv_bfloat16*)(ws + off[5]);
  return P;
}

extern "C" int m2h_unet_workspace_clear_halos(int B, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0) { m2h_set_error("m2h_unet_workspace_clear_halos: B < 0"); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!workspace) { m2h_set_error("m2h_unet_workspace_clear_halos: null pointer"); return M2H_ERR_ARG; }
  if (workspace_bytes < m2h_unet_workspace_bytes(B)) { m2h_set_error("m2h_unet_workspace_clear_halos: workspace too small"); return M2H_ERR_WORKSPACE; }
  int dev;
  int rc = m2h_current_device_checked(&dev);
  if (rc != M2H_OK) return rc;
  const WsPtrs W = ws_ptrs(workspace, B);
  HaloParams P;
  P.buf[0] = W.x0;
  for (int i = 1; i < 5; ++i) P.buf[i] = W.cat[i];
  P.start[0] = 0;
  for (int i = 0; i < 5; ++i) { P.Hp[i] = kWsHp[i]; P.ld8[i] = kWsLd[i] / 8; P.start[i + 1] = P.start[i] + (long long)B * (4 * kWsHp[i] - 4) * P.ld8[i]; }
  const long long blocks = (P.start[5] + 255) / 256;
  clear_halo_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(P);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

// x0_ext != null: the first layer's operand (bf16 [B][18][18][128], zero border) was written by the STFT kernel
// (m2h_stft_waves_x0): no fp32 observation, no pack_input_kernel
static int encoder_forward_impl(const void* packed_enc, int convert_bin2mono, const float* obs, const __nv_bfloat16* x0_ext,
                                const float* target_class_plus1, const int64_t* target_class_i64, const float* masks, const int64_t* mask_strides,
                                int B, void* workspace, size_t workspace_bytes, float* bottleneck_f32, void* stream) {
  if (B < 0) { m2h_set_error("m2h_unet_encoder_forward: B < 0"); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!packed_enc || (!obs && !x0_ext) || !workspace) { m2h_set_error("m2h_unet_encoder_forward: null pointer"); return M2H_ERR_ARG; }
  if (convert_bin2mono ? (!masks || !mask_strides) : (!target_class_plus1 && !target_class_i64)) { m2h_set_error("m2h_unet_encoder_forward: %s", convert_bin2mono ? "bin2mono needs masks + strides" : "binaural separator needs target_class"); return M2H_ERR_ARG; }
  if (workspace_bytes < m2h_unet_workspace_bytes(B)) { m2h_set_error("m2h_unet_encoder_forward: workspace too small"); return M2H_ERR_WORKSPACE; }
  int rc = unet_prepare_device();
  if (rc != M2H_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const WsPtrs W = ws_ptrs(workspace, B);
  const EncLayout PL = enc_layout();
  const char* pk = (const char*)packed_enc;

  // the workspace's zero borders are the caller's contract (m2h_unet_workspace_clear_halos once per allocation): every
  // kernel below writes interiors only
  if (!x0_ext) {
    PackInParams pp;
    pp.obs = obs; pp.masks = convert_bin2mono ? masks : nullptr; pp.x0 = W.x0; pp.B = B; pp.x0_plain = nullptr;
    for (int i = 0; i < 4; ++i) pp.ms[i] = convert_bin2mono ? (long long)mask_strides[i] : 0;
    M2H_CUDA_CHECK(launch_pdl(pack_input_kernel, dim3((B * 1024 + 255) / 256), dim3(256), (size_t)0, st, pp));
    M2H_CUDA_CHECK(cudaGetLastError());
  }

  {  // layer 0 on the slab kernel: x0 (space-to-depth, 18x18 cells x 128 ch) -> cat1[..., 64:128]
    USLayer sl; SlabPackDesc pd;
    memset(&sl, 0, sizeof(sl));
    build_conv0_items(18, 0, sl, pd);
    sl.in_coff = 0;
    sl.scale = (const float*)(pk + PL.scale[0]); sl.bias = (const float*)(pk + PL.bias[0]); sl.slope = 0.2f;
    if (!convert_bin2mono) { sl.cls = target_class_plus1; sl.cls_i64 = (const long long*)target_class_i64; sl.cls_w = (const float*)(pk + PL.cls_w); }
    sl.out = W.cat[1]; sl.out_mode = UG_OUT_NHWC_BF16; sl.out_stride = 1; sl.Hop = sl.Wop = 18; sl.out_ld = 128; sl.out_coff = 64;
    rc = launch_slab(sl, x0_ext ? x0_ext : W.x0, 128, 18, B, (const __nv_bfloat16*)(pk + PL.w[0]), st);
    if (rc != M2H_OK) return rc;
  }
  // layers 1-4 (implicit-GEMM gather kernel); skips are written into the upper channel half of the concat buffers
  const __nv_bfloat16* in_ptr[5] = {nullptr, W.cat[1] + 64, W.cat[2] + 128, W.cat[3] + 256, W.cat[4] + 512};
  const int in_ld[5] = {0, 128, 256, 512, 1024};
  __nv_bfloat16* out_ptr[5] = {nullptr, W.cat[2] + 128, W.cat[3] + 256, W.cat[4] + 512, W.bott};
  const int out_ld[5] = {0, 256, 512, 1024, 512};
  const int hin[5] = {32, 16, 8, 4, 2};
  for (int i = 1; i < 5; ++i) {
    UGLayer g;
    layer_defaults(g);
    g.in = in_ptr[i]; g.Hin = g.Win = hin[i]; g.Cin = kEncCin[i]; g.cin_shift = ilog2(kEncCin[i]); g.in_ld = in_ld[i]; g.in_pad = 1;
    g.Hq = g.Wq = hin[i] / 2; g.M = B * g.Hq * g.Wq; g.in_stride = 2; g.n_taps = kEncTaps[i];
    for (int t = 0; t < kEncTaps[i]; ++t) {
      const int kh = kEncTaps[i] == 16 ? t >> 2 : 1 + (t >> 1), kw = kEncTaps[i] == 16 ? t & 3 : 1 + (t & 1);
      g.dy[0][t] = (int8_t)(kh - 1); g.dx[0][t] = (int8_t)(kw - 1);
    }
    g.w = (const __nv_bfloat16*)(pk + PL.w[i]); g.Cout = kEncCout[i]; g.Kvalid = g.Kpad = kEncTaps[i] * kEncCin[i];
    g.scale = (const float*)(pk + PL.scale[i]); g.bias = (const float*)(pk + PL.bias[i]); g.n_mod = kEncCout[i]; g.slope = 0.2f;
    g.out = out_ptr[i]; g.Hout = g.Wout = g.Hq; g.out_ld = out_ld[i]; g.out_pad = i == 4 ? 0 : 1; g.n_valid = g.Cout;
    rc = launch_layer_auto(g, 128, B, st);
    if (rc != M2H_OK) return rc;
  }
  if (bottleneck_f32) {
    M2H_CUDA_CHECK(launch_pdl(bf16_to_f32_kernel, dim3((B * 512 + 255) / 256), dim3(256), (size_t)0, st, (const __nv_bfloat16*)W.bott, bottleneck_f32, (size_t)B * 512));
    M2H_CUDA_CHECK(cudaGetLastError());
  }
  return M2H_OK;
}

extern "C" int m2h_unet_encoder_forward(const void* packed_enc, int convert_bin2mono, const float* obs,
                                        const float* target_class_plus1, const float* masks, const int64_t* mask_strides,
                                        int B, void* workspace, size_t workspace_bytes, float* bottleneck_f32, void* stream) {
  if (B > 0 && !obs) { m2h_set_error("m2h_unet_encoder_forward: null pointer"); return M2H_ERR_ARG; }
  return encoder_forward_impl(packed_enc, convert_bin2mono, obs, nullptr, target_class_plus1, nullptr, masks, mask_strides, B, workspace, workspace_bytes, bottleneck_f32, stream);
}

extern "C" int m2h_unet_encoder_forward_x0(const void* packed_enc, const void* x0, const int64_t* target_class, int B,
                                           void* workspace, size_t workspace_bytes, float* bottleneck_f32, void* stream) {
  if (B > 0 && (!x0 || ((uintptr_t)x0 & 127) != 0)) { m2h_set_error("m2h_unet_encoder_forward_x0: x0 must be a 128-byte aligned device pointer"); return M2H_ERR_ARG; }
  return encoder_forward_impl(packed_enc, 0, nullptr, (const __nv_bfloat16*)x0, nullptr, target_class, nullptr, nullptr, B, workspace, workspace_bytes, bottleneck_f32, stream);
}

extern "C" int m2h_unet_decoder_forward(const void* packed_dec, int convert_bin2mono, int B, void* workspace,
                                        size_t workspace_bytes, float* mask_out, void* stream) {
  if (B < 0) { m2h_set_error("m2h_unet_decoder_forward: B < 0"); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!packed_dec || !workspace || !mask_out) { m2h_set_error("m2h_unet_decoder_forward: null pointer"); return M2H_ERR_ARG; }
  if (workspace_bytes < m2h_unet_workspace_bytes(B)) { m2h_set_error("m2h_unet_decoder_forward: workspace too small"); return M2H_ERR_WORKSPACE; }
  int rc = unet_prepare_device();
  if (rc != M2H_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const WsPtrs W = ws_ptrs(workspace, B);
  const DecLayout PL = dec_layout();
  const char* pk = (const char*)packed_dec;

  {  // layer 0: 1x1 -> 2x2, a plain GEMM [B, 512] x [2048, 512]^T whose column blocks are the 4 output pixels (cat4 is a haloed 4x4 grid)
    UGLayer g;
    layer_defaults(g);
    g.in = W.bott; g.Hin = g.Win = 1; g.Cin = 512; g.cin_shift = 9; g.in_ld = 512;
    g.Hq = g.Wq = 1; g.M = B; g.n_taps = 1;
    g.w = (const __nv_bfloat16*)(pk + PL.w[0]); g.Cout = 2048; g.Kvalid = g.Kpad = 512;
    g.scale = (const float*)(pk + PL.scale[0]); g.bias = (const float*)(pk + PL.bias[0]); g.n_mod = 512; g.slope = 0.f;
    g.out = W.cat[4]; g.Hout = g.Wout = 1; g.out_ld = 16 * 1024; g.n_split = 512; g.n_valid = 2048;
    for (int pix = 0; pix < 4; ++pix) g.pix_off[pix] = (((pix >> 1) + 1) * 4 + (pix & 1) + 1) * 1024;
    rc = launch_layer_auto(g, 128, B, st);
    if (rc != M2H_OK) return rc;
  }
  const __nv_bfloat16* in_ptr[5] = {nullptr, W.cat[4], W.cat[3], W.cat[2], W.cat[1]};
  __nv_bfloat16* out_ptr[5] = {nullptr, W.cat[3], W.cat[2], W.cat[1], nullptr};
  const int out_ld[5] = {0, 512, 256, 128, 0};
  const int hin[5] = {1, 2, 4, 8, 16};
  for (int i = 1; i < 3; ++i) {   // layers 1, 2: gather kernel, 4 sub-pixel phases
    UGLayer g;
    layer_defaults(g);
    g.in = in_ptr[i]; g.Hin = g.Win = hin[i]; g.Cin = kDecCin[i]; g.cin_shift = ilog2(kDecCin[i]); g.in_ld = kDecCin[i]; g.in_pad = 1;
    g.Hq = g.Wq = hin[i]; g.M = B * hin[i] * hin[i]; g.n_taps = 4; g.n_phase = 4;
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < 4; ++t) { g.dy[ph][t] = (int8_t)deconv_d(ph >> 1, t >> 1); g.dx[ph][t] = (int8_t)deconv_d(ph & 1, t & 1); }
    g.w = (const __nv_bfloat16*)(pk + PL.w[i]); g.Cout = kDecCoutBase[i]; g.Kvalid = g.Kpad = 4 * kDecCin[i];
    g.scale = (const float*)(pk + PL.scale[i]); g.bias = (const float*)(pk + PL.bias[i]); g.n_mod = kDecCoutBase[i]; g.slope = 0.f;
    g.out = out_ptr[i]; g.Hout = g.Wout = 2 * hin[i]; g.out_stride = 2; g.out_ld = out_ld[i]; g.out_pad = 1; g.n_valid = g.Cout;
    rc = launch_layer_auto(g, 128, B, st);
    if (rc != M2H_OK) return rc;
  }
  for (int i = 3; i < 5; ++i) {   // layers 3, 4: slab kernel; layer 4 carries the 1x1 head and writes fp32 [B, C, 512, 32]
    USLayer sl; SlabPackDesc pd;
    memset(&sl, 0, sizeof(sl));
    const int hp = hin[i] + 2;
    build_deconv_items(kDecCin[i], kDecCoutBase[i], kDecCoutBase[i], hp, sl, pd);
    sl.in_coff = 0;
    sl.scale = (const float*)(pk + PL.scale[i]); sl.bias = (const float*)(pk + PL.bias[i]); sl.slope = 0.f;
    if (i == 3) {
      sl.out = out_ptr[3]; sl.out_mode = UG_OUT_NHWC_BF16; sl.out_stride = 2; sl.Hop = sl.Wop = 18; sl.out_ld = out_ld[3]; sl.out_coff = 0;
    } else {
      sl.out = mask_out; sl.out_mode = UG_OUT_HEAD_NCHW_F32; sl.n_valid = convert_bin2mono ? 16 : 32;
      sl.head_w = (const __nv_bfloat16*)(pk + PL.head_w); sl.head_b = (const float*)(pk + PL.head_b);
    }
    rc = launch_slab(sl, in_ptr[i], kDecCin[i], hp, B, (const __nv_bfloat16*)(pk + PL.w[i]), st);
    if (rc != M2H_OK) return rc;
  }
  return M2H_OK;
}

#include "m2h_unet_train.inl"

// C-ABI entry points of the passive-separator U-Net (include/m2h.h): weight packing, input packing, layer launches.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "../../include/m2h.h"
#include "m2h_internal.h"
#include "unet_gemm.cuh"

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
  for (int i = 0; i < 5; ++i) { L.w[i] = o; o = align256(o + (size_t)kEncCout[i] * kEncTaps[i] * kEncCin[i] * 2); }
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
static const size_t kWsElems[7] = {32 * 32 * 32, 16 * 16 * 128, 8 * 8 * 256, 4 * 4 * 512, 2 * 2 * 1024, 512, 32 * 32 * 32};
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
// Conv2d weight [Cout][CinFull][4][4] -> bf16 [Cout][n_taps*Cin], taps = all 16 (kh*4+kw) or the 4 central ones
__global__ void pack_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int CinFull, int Cin, int n_taps) {
  const size_t total = (size_t)Cout * n_taps * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int t = (int)((i / Cin) % n_taps);
    const int n = (int)(i / ((size_t)Cin * n_taps));
    const int kh = n_taps == 16 ? t >> 2 : 1 + (t >> 1), kw = n_taps == 16 ? t & 3 : 1 + (t & 1);
    out[i] = __float2bfloat16_rn(w[(((size_t)n * CinFull + c) * 4 + kh) * 4 + kw]);
  }
}
// class plane of encoder layer 0: cls_w[cy][cx][n] = sum over taps that stay inside the image of w[n][32][kh][kw]
__global__ void pack_cls_w_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int CinFull) {
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
// ConvTranspose2d weight [Cin][CoutReal][4][4] -> bf16 [4 phases][CoutPad][4 taps * Cin]
__global__ void pack_deconv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int CoutReal, int CoutPad) {
  const size_t total = (size_t)4 * CoutPad * 4 * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int t = (int)((i / Cin) % 4);
    const int n = (int)((i / ((size_t)4 * Cin)) % CoutPad);
    const int ph = (int)(i / ((size_t)4 * Cin * CoutPad));
    const int kh = deconv_k(ph >> 1, t >> 1), kw = deconv_k(ph & 1, t & 1);
    out[i] = n < CoutReal ? __float2bfloat16_rn(w[(((size_t)c * CoutReal + n) * 4 + kh) * 4 + kw]) : __float2bfloat16_rn(0.f);
  }
}
// decoder layer 0 (1x1 input): output pixel (oy, ox) = weight tap (oy+1, ox+1): bf16 [4 pixels * Cout][Cin]
__global__ void pack_deconv0_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cout) {
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
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cpad) return;
  if (i < C) {
    const float s = g[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    bias[i] = b[i] - mean[i] * s;
  } else { scale[i] = 0.f; bias[i] = 0.f; }
}

// obs fp32 [B,512,32,2] -> x0 bf16 NHWC [B,32,32,32], channel = ear*16 + slice, row = freq % 32 (separator_cnn.py:84-90);
// bin2mono: x = log1p(max(mask * (exp(obs) - 1), 0)) first (separator_cnn.py:73-79)
struct PackInParams {
  const float* obs; const float* masks; long long ms[4]; __nv_bfloat16* x0; int B;
};
__global__ void pack_input_kernel(const PackInParams P) {
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
  uint4* dst = reinterpret_cast<uint4*>(P.x0 + (size_t)i * 32);
  const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = src[k];
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __bfloat162float(in[i]);
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
static bool g_unet_attr_set[M2H_MAX_DEVICES];
#ifndef M2H_UNFUSED_HEAD
#define M2H_UNFUSED_HEAD 0      // 1: run the 1x1 head as its own GEMM launch (kept for A/B testing)
#endif

template <int BN>
static int launch_bn(const UGLayer& L, cudaStream_t st) {
  dim3 grid((L.M + UG_BM - 1) / UG_BM, L.Cout / BN, L.n_phase);
  unet_gemm_kernel<BN><<<grid, UG_THREADS, ug::Cfg<BN>::kSmemBytes, st>>>(L);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}
static int launch_layer(const UGLayer& L, int bn, cudaStream_t st) {
  if (L.Cout % bn != 0 || L.Kpad % UG_BK != 0 || (1 << L.cin_shift) != L.Cin || L.M <= 0) { m2h_set_error("unet layer: unsupported shape Cout=%d bn=%d Kpad=%d Cin=%d M=%d", L.Cout, bn, L.Kpad, L.Cin, L.M); return M2H_ERR_ARG; }
  switch (bn) {
    case 32: return launch_bn<32>(L, st);
    case 64: return launch_bn<64>(L, st);
    case 128: return launch_bn<128>(L, st);
    default: m2h_set_error("unet layer: no kernel for BN=%d", bn); return M2H_ERR_ARG;
  }
}
static int unet_prepare_device() {
  int dev;
  int rc = m2h_current_device_checked(&dev);
  if (rc != M2H_OK) return rc;
  if (!g_unet_attr_set[dev]) {
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ug::Cfg<32>::kSmemBytes));
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ug::Cfg<64>::kSmemBytes));
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ug::Cfg<128>::kSmemBytes));
    // the default carve-out left room for one CTA per SM (first ncu capture: 100 KB config, 5 warps/SM): ask for the maximum
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_gemm_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    g_unet_attr_set[dev] = true;
  }
  return M2H_OK;
}
static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

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
    pack_conv_w_kernel<<<256, 256, 0, st>>>(p->conv_w[i], (__nv_bfloat16*)(base + L.w[i]), kEncCout[i], cin_full, kEncCin[i], kEncTaps[i]);
    fold_bn_kernel<<<(kEncCout[i] + 127) / 128, 128, 0, st>>>(p->bn_w[i], p->bn_b[i], p->bn_mean[i], p->bn_var[i], bn_eps,
                                                             (float*)(base + L.scale[i]), (float*)(base + L.bias[i]), kEncCout[i], kEncCout[i]);
  }
  if (!convert_bin2mono) pack_cls_w_kernel<<<(9 * 64 + 127) / 128, 128, 0, st>>>(p->conv_w[0], (float*)(base + L.cls_w), 64, 33);
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
    if (i == 0) pack_deconv0_w_kernel<<<256, 256, 0, st>>>(p->deconv_w[0], (__nv_bfloat16*)(base + L.w[0]), 512, 512);
    else pack_deconv_w_kernel<<<256, 256, 0, st>>>(p->deconv_w[i], (__nv_bfloat16*)(base + L.w[i]), kDecCin[i], cout_real, kDecCoutBase[i]);
    fold_bn_kernel<<<(kDecCoutBase[i] + 127) / 128, 128, 0, st>>>(p->bn_w[i], p->bn_b[i], p->bn_mean[i], p->bn_var[i], bn_eps,
                                                                 (float*)(base + L.scale[i]), (float*)(base + L.bias[i]), cout_real, kDecCoutBase[i]);
  }
  pack_head_kernel<<<(32 * 64 + 127) / 128, 128, 0, st>>>(p->head_w, p->head_b, (__nv_bfloat16*)(base + L.head_w), (float*)(base + L.head_wf), (float*)(base + L.head_b), c_last);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

static void layer_defaults(UGLayer& g) {
  memset(&g, 0, sizeof(g));
  g.n_phase = 1; g.in_stride = 1; g.out_stride = 1; g.slope = 1.f; g.n_split = 1 << 30; g.n_split_ld = 0;
  g.out_mode = UG_OUT_NHWC_BF16;
}

extern "C" int m2h_unet_encoder_forward(const void* packed_enc, int convert_bin2mono, const float* obs,
                                        const float* target_class_plus1, const float* masks, const int64_t* mask_strides,
                                        int B, void* workspace, size_t workspace_bytes, float* bottleneck_f32, void* stream) {
  if (B < 0) { m2h_set_error("m2h_unet_encoder_forward: B < 0"); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!packed_enc || !obs || !workspace) { m2h_set_error("m2h_unet_encoder_forward: null pointer"); return M2H_ERR_ARG; }
  if (convert_bin2mono ? (!masks || !mask_strides) : !target_class_plus1) { m2h_set_error("m2h_unet_encoder_forward: %s", convert_bin2mono ? "bin2mono needs masks + strides" : "binaural separator needs target_class"); return M2H_ERR_ARG; }
  if (workspace_bytes < m2h_unet_workspace_bytes(B)) { m2h_set_error("m2h_unet_encoder_forward: workspace too small"); return M2H_ERR_WORKSPACE; }
  int rc = unet_prepare_device();
  if (rc != M2H_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  size_t off[7];
  m2h_unet_workspace_layout(B, off);
  char* ws = (char*)workspace;
  __nv_bfloat16* x0 = (__nv_bfloat16*)(ws + off[0]);
  __nv_bfloat16* cat[5] = {nullptr, (__nv_bfloat16*)(ws + off[1]), (__nv_bfloat16*)(ws + off[2]), (__nv_bfloat16*)(ws + off[3]), (__nv_bfloat16*)(ws + off[4])};
  __nv_bfloat16* bott = (__nv_bfloat16*)(ws + off[5]);
  const EncLayout PL = enc_layout();
  const char* pk = (const char*)packed_enc;

  PackInParams pp;
  pp.obs = obs; pp.masks = convert_bin2mono ? masks : nullptr; pp.x0 = x0; pp.B = B;
  for (int i = 0; i < 4; ++i) pp.ms[i] = convert_bin2mono ? (long long)mask_strides[i] : 0;
  pack_input_kernel<<<(B * 1024 + 255) / 256, 256, 0, st>>>(pp);
  M2H_CUDA_CHECK(cudaGetLastError());

  // inputs / outputs of the five layers; skips are written into the upper channel half of the concat buffers
  const __nv_bfloat16* in_ptr[5] = {x0, cat[1] + 64, cat[2] + 128, cat[3] + 256, cat[4] + 512};
  const int in_ld[5] = {32, 128, 256, 512, 1024};
  __nv_bfloat16* out_ptr[5] = {cat[1] + 64, cat[2] + 128, cat[3] + 256, cat[4] + 512, bott};
  const int out_ld[5] = {128, 256, 512, 1024, 512};
  const int hin[5] = {32, 16, 8, 4, 2};
  const int bn[5] = {64, 128, 128, 128, 128};
  for (int i = 0; i < 5; ++i) {
    UGLayer g;
    layer_defaults(g);
    g.in = in_ptr[i]; g.Hin = g.Win = hin[i]; g.Cin = kEncCin[i]; g.cin_shift = ilog2(kEncCin[i]); g.in_ld = in_ld[i];
    g.Hq = g.Wq = hin[i] / 2; g.M = B * g.Hq * g.Wq; g.in_stride = 2; g.n_taps = kEncTaps[i];
    for (int t = 0; t < kEncTaps[i]; ++t) {
      const int kh = kEncTaps[i] == 16 ? t >> 2 : 1 + (t >> 1), kw = kEncTaps[i] == 16 ? t & 3 : 1 + (t & 1);
      g.dy[0][t] = (int8_t)(kh - 1); g.dx[0][t] = (int8_t)(kw - 1);
    }
    g.w = (const __nv_bfloat16*)(pk + PL.w[i]); g.Cout = kEncCout[i]; g.Kvalid = g.Kpad = kEncTaps[i] * kEncCin[i];
    g.scale = (const float*)(pk + PL.scale[i]); g.bias = (const float*)(pk + PL.bias[i]); g.n_mod = kEncCout[i]; g.slope = 0.2f;
    if (i == 0 && !convert_bin2mono) { g.cls = target_class_plus1; g.cls_w = (const float*)(pk + PL.cls_w); }
    g.out = out_ptr[i]; g.Hout = g.Wout = g.Hq; g.out_ld = out_ld[i]; g.n_valid = g.Cout;
    rc = launch_layer(g, bn[i], st);
    if (rc != M2H_OK) return rc;
  }
  if (bottleneck_f32) {
    bf16_to_f32_kernel<<<(B * 512 + 255) / 256, 256, 0, st>>>(bott, bottleneck_f32, (size_t)B * 512);
    M2H_CUDA_CHECK(cudaGetLastError());
  }
  return M2H_OK;
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
  size_t off[7];
  m2h_unet_workspace_layout(B, off);
  char* ws = (char*)workspace;
  __nv_bfloat16* cat[5] = {nullptr, (__nv_bfloat16*)(ws + off[1]), (__nv_bfloat16*)(ws + off[2]), (__nv_bfloat16*)(ws + off[3]), (__nv_bfloat16*)(ws + off[4])};
  __nv_bfloat16* bott = (__nv_bfloat16*)(ws + off[5]);
  __nv_bfloat16* d4 = (__nv_bfloat16*)(ws + off[6]);
  const DecLayout PL = dec_layout();
  const char* pk = (const char*)packed_dec;

  {  // layer 0: 1x1 -> 2x2, a plain GEMM [B, 512] x [2048, 512]^T whose column blocks are the 4 output pixels
    UGLayer g;
    layer_defaults(g);
    g.in = bott; g.Hin = g.Win = 1; g.Cin = 512; g.cin_shift = 9; g.in_ld = 512;
    g.Hq = g.Wq = 1; g.M = B; g.n_taps = 1;
    g.w = (const __nv_bfloat16*)(pk + PL.w[0]); g.Cout = 2048; g.Kvalid = g.Kpad = 512;
    g.scale = (const float*)(pk + PL.scale[0]); g.bias = (const float*)(pk + PL.bias[0]); g.n_mod = 512; g.slope = 0.f;
    g.out = cat[4]; g.Hout = g.Wout = 1; g.out_ld = 4 * 1024; g.n_split = 512; g.n_split_ld = 1024; g.n_valid = 2048;
    rc = launch_layer(g, 128, st);
    if (rc != M2H_OK) return rc;
  }
  const __nv_bfloat16* in_ptr[5] = {nullptr, cat[4], cat[3], cat[2], cat[1]};
  __nv_bfloat16* out_ptr[5] = {nullptr, cat[3], cat[2], cat[1], d4};
  const int out_ld[5] = {0, 512, 256, 128, 32};
  const int hin[5] = {1, 2, 4, 8, 16};
  const int bn[5] = {0, 128, 128, 64, 32};
  for (int i = 1; i < 5; ++i) {
    UGLayer g;
    layer_defaults(g);
    g.in = in_ptr[i]; g.Hin = g.Win = hin[i]; g.Cin = kDecCin[i]; g.cin_shift = ilog2(kDecCin[i]); g.in_ld = kDecCin[i];
    g.Hq = g.Wq = hin[i]; g.M = B * hin[i] * hin[i]; g.n_taps = 4; g.n_phase = 4;
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < 4; ++t) { g.dy[ph][t] = (int8_t)deconv_d(ph >> 1, t >> 1); g.dx[ph][t] = (int8_t)deconv_d(ph & 1, t & 1); }
    g.w = (const __nv_bfloat16*)(pk + PL.w[i]); g.Cout = kDecCoutBase[i]; g.Kvalid = g.Kpad = 4 * kDecCin[i];
    g.scale = (const float*)(pk + PL.scale[i]); g.bias = (const float*)(pk + PL.bias[i]); g.n_mod = kDecCoutBase[i]; g.slope = 0.f;
    g.out = out_ptr[i]; g.Hout = g.Wout = 2 * hin[i]; g.out_stride = 2; g.out_ld = out_ld[i]; g.n_valid = g.Cout;
    if (i == 4 && !M2H_UNFUSED_HEAD) {   // last up-conv + BN + ReLU + 1x1 head + bias -> fp32 [B, C, 512, 32] in one kernel
      g.out = mask_out; g.out_mode = UG_OUT_HEAD_NCHW_F32; g.n_valid = convert_bin2mono ? 16 : 32;
      g.head_w = (const float*)(pk + PL.head_wf); g.head_b = (const float*)(pk + PL.head_b);
    }
    rc = launch_layer(g, bn[i], st);
    if (rc != M2H_OK) return rc;
  }
  if (M2H_UNFUSED_HEAD) {  // 1x1 head + bias -> fp32 NCHW = [B, C, 512, 32] (channel = ear*16 + slice, row = freq % 32)
    UGLayer g;
    layer_defaults(g);
    g.in = d4; g.Hin = g.Win = 32; g.Cin = 32; g.cin_shift = 5; g.in_ld = 32;
    g.Hq = g.Wq = 32; g.M = B * 1024; g.n_taps = 1;
    g.w = (const __nv_bfloat16*)(pk + PL.head_w); g.Cout = 32; g.Kvalid = 32; g.Kpad = 64;
    g.bias = (const float*)(pk + PL.head_b); g.n_mod = 32; g.slope = 1.f;
    g.out = mask_out; g.out_mode = UG_OUT_NCHW_F32; g.Hout = g.Wout = 32; g.n_valid = convert_bin2mono ? 16 : 32;
    rc = launch_layer(g, 32, st);
    if (rc != M2H_OK) return rc;
  }
  return M2H_OK;
}

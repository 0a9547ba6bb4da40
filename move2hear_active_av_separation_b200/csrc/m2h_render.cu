// C-ABI entry points of the mixture render (include/m2h.h): table set-up + kernel launches.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <vector>

#include "../../include/m2h.h"
#include "m2h_internal.h"
#include "render_conv.cuh"
#include "render_stft.cuh"
#include "m2h_tables.h"

// ---------------------------------------------------------------------------------------------
// error plumbing (shared with the other translation units through m2h_internal.h)
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void m2h_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* m2h_last_error(void) { return g_err; }
extern "C" int m2h_abi_version(void) { return M2H_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------
struct DeviceTables {
  M2hConvTables* conv = nullptr;
  M2hStftTables* stft = nullptr;
  int n_sms = 0;
  bool ready = false;
};
static DeviceTables g_tables[M2H_MAX_DEVICES];
static std::mutex g_init_mutex;

int m2h_current_device_checked(int* dev) {
  cudaError_t e = cudaGetDevice(dev);
  if (e != cudaSuccess) { m2h_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return M2H_ERR_CUDA; }
  if (*dev < 0 || *dev >= M2H_MAX_DEVICES) { m2h_set_error("device %d out of range", *dev); return M2H_ERR_ARG; }
  return M2H_OK;
}

int m2h_num_sms(int dev) { return g_tables[dev].n_sms; }

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// VEC = packed row accesses (even N / Lmax, aligned base pointers: conv_can_vectorise); the scalar variant serves the rest
template <bool VEC>
__global__ void __launch_bounds__(M2H_CONV_THREADS, 1)
m2h_rir_conv_kernel(const ConvParams P, const M2hConvTables* __restrict__ tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ConvSmem& s = *reinterpret_cast<ConvSmem*>(smem_raw);
  conv_cta_body<VEC>(P, s, tab, blockIdx.x, gridDim.x);
}

// 64 registers / thread (no spills): two 512-thread CTAs per SM, 2 x 86 KB of shared memory
template <bool F16>
__global__ void __launch_bounds__(M2H_STFT_THREADS, 2)
m2h_stft_kernel(const StftParams P, const M2hStftTables* __restrict__ tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
  stft_cta_body<F16>(P, s, tab, blockIdx.x, gridDim.x);
}

extern "C" int m2h_init(int device) {
  if (device < 0 || device >= M2H_MAX_DEVICES) { m2h_set_error("device %d out of range", device); return M2H_ERR_ARG; }
  std::lock_guard<std::mutex> lock(g_init_mutex);
  DeviceTables& d = g_tables[device];
  if (d.ready) return M2H_OK;
  int prev = 0;
  M2H_CUDA_CHECK(cudaGetDevice(&prev));
  M2H_CUDA_CHECK(cudaSetDevice(device));
  M2H_CUDA_CHECK(cudaDeviceGetAttribute(&d.n_sms, cudaDevAttrMultiProcessorCount, device));
  static M2hConvTables hc;
  static M2hStftTables hs;
  static bool host_built = false;
  if (!host_built) { build_conv_tables(hc); build_stft_tables(hs); host_built = true; }
  M2H_CUDA_CHECK(cudaMalloc(&d.conv, sizeof(M2hConvTables)));
  M2H_CUDA_CHECK(cudaMalloc(&d.stft, sizeof(M2hStftTables)));
  M2H_CUDA_CHECK(cudaMemcpy(d.conv, &hc, sizeof(hc), cudaMemcpyHostToDevice));
  M2H_CUDA_CHECK(cudaMemcpy(d.stft, &hs, sizeof(hs), cudaMemcpyHostToDevice));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_stft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_stft_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem)));
  // the driver's default carve-out (102 KB, ncu r1b) left room for ONE STFT CTA per SM: ask for the maximum
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_stft_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_stft_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaSetDevice(prev));
  d.ready = true;
  return M2H_OK;
}

static int ensure_init(int* dev) {
  int rc = m2h_current_device_checked(dev);
  if (rc != M2H_OK) return rc;
  if (!g_tables[*dev].ready) return m2h_init(*dev);
  return M2H_OK;
}

extern "C" size_t m2h_render_workspace_bytes(int device) {
  int n_sms = 148;
  if (device >= 0 && device < M2H_MAX_DEVICES && g_tables[device].ready) n_sms = g_tables[device].n_sms;
  else if (device >= 0) { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) n_sms = v; }
  return (size_t)n_sms * M2H_CONV_SCRATCH_F4 * sizeof(float4);
}

extern "C" int m2h_rir_convolve(const int16_t* mono, const float* rir, const int32_t* rir_len,
                                int B, int K, int N, int Lmax, int16_t* waves,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || K <= 0 || N <= 0 || N > M2H_MAX_SAMPLES || Lmax < 0) { m2h_set_error("m2h_rir_convolve: bad sizes B=%d K=%d N=%d Lmax=%d (N <= %d)", B, K, N, Lmax, M2H_MAX_SAMPLES); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!mono || !rir_len || !waves || (!rir && Lmax > 0) || !workspace) { m2h_set_error("m2h_rir_convolve: null pointer"); return M2H_ERR_ARG; }
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  const int n_items = B * K;
  int grid = g_tables[dev].n_sms;
  if (grid > n_items) grid = n_items;
  if (workspace_bytes < (size_t)grid * M2H_CONV_SCRATCH_F4 * sizeof(float4)) { m2h_set_error("m2h_rir_convolve: workspace %zu B < %zu B", workspace_bytes, (size_t)grid * M2H_CONV_SCRATCH_F4 * sizeof(float4)); return M2H_ERR_WORKSPACE; }
  ConvParams P;
  P.mono = mono; P.rir = rir; P.rir_len = rir_len; P.waves = waves; P.scratch = (float4*)workspace;
  P.n_items = n_items; P.N = N; P.Lmax = Lmax;
  if (conv_can_vectorise(P)) m2h_rir_conv_kernel<true><<<grid, M2H_CONV_THREADS, sizeof(ConvSmem), (cudaStream_t)stream>>>(P, g_tables[dev].conv);
  else m2h_rir_conv_kernel<false><<<grid, M2H_CONV_THREADS, sizeof(ConvSmem), (cudaStream_t)stream>>>(P, g_tables[dev].conv);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

static int launch_stft(const StftParams& P, int dev, void* stream) {
  const int n_items = P.B * P.n_signals;
  if (n_items == 0) return M2H_OK;
  int grid = 2 * g_tables[dev].n_sms;           // 110 KB of shared memory per CTA -> two CTAs per SM
  if (grid > n_items) grid = n_items;
  // outputs are written as (left, right) pairs
  const void* outs[5] = {P.mix_mag, P.mix_phase, P.gt_mag, P.f_mag, P.f_phase};
  for (int i = 0; i < 5; ++i)
    if (((uintptr_t)outs[i] & 7) != 0) { m2h_set_error("stft: output buffers must be 8-byte aligned"); return M2H_ERR_ARG; }
  if (P.flags & M2H_OUT_F16) m2h_stft_kernel<true><<<grid, M2H_STFT_THREADS, sizeof(StftSmem), (cudaStream_t)stream>>>(P, g_tables[dev].stft);
  else m2h_stft_kernel<false><<<grid, M2H_STFT_THREADS, sizeof(StftSmem), (cudaStream_t)stream>>>(P, g_tables[dev].stft);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

extern "C" int m2h_stft_waves(const int16_t* waves, int B, int K, int N, int n_gt, int gt_with_phase, int flags,
                              void* mix_mag, void* mix_phase, void* gt, void* stream) {
  if (B < 0 || K <= 0 || K > 256 || N < M2H_HOP || N > M2H_MAX_SAMPLES || n_gt < 0 || n_gt > K) { m2h_set_error("m2h_stft_waves: bad sizes B=%d K=%d N=%d n_gt=%d", B, K, N, n_gt); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!waves || (n_gt > 0 && !gt)) { m2h_set_error("m2h_stft_waves: null pointer"); return M2H_ERR_ARG; }
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.waves = waves; P.B = B; P.K = K; P.N = N; P.T = 1 + (N - 1) / M2H_HOP;
  P.n_gt = n_gt; P.n_signals = 1 + n_gt;
  P.flags = ((flags & M2H_RENDER_RAW_MAG) ? 0 : M2H_OUT_MIX_LOG1P) | ((flags & M2H_RENDER_F16_OBS) ? M2H_OUT_F16 : 0);
  P.mix_mag = mix_mag; P.mix_phase = mix_phase; P.gt_mag = gt; P.gt_interleaved_phase = gt_with_phase ? 1 : 0;
  return launch_stft(P, dev, stream);
}

extern "C" int m2h_stft_f32(const float* wave, int B, int C, int N, float* mag, float* phase, void* stream) {
  if (B < 0 || (C != 1 && C != 2) || N < M2H_HOP || N > M2H_MAX_SAMPLES) { m2h_set_error("m2h_stft_f32: bad sizes B=%d C=%d N=%d", B, C, N); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!wave || !mag) { m2h_set_error("m2h_stft_f32: null pointer"); return M2H_ERR_ARG; }
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.fwave = wave; P.fwave_channels = C; P.B = B; P.K = 0; P.N = N; P.T = 1 + (N - 1) / M2H_HOP;
  P.n_signals = 1; P.f_mag = mag; P.f_phase = phase;
  return launch_stft(P, dev, stream);
}

extern "C" int m2h_render_mixture(const int16_t* mono, const float* rir, const int32_t* rir_len,
                                  int B, int K, int N, int Lmax, int n_gt, int gt_with_phase, int flags,
                                  int16_t* waves, void* mix_mag, void* mix_phase, void* gt,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  int rc = m2h_rir_convolve(mono, rir, rir_len, B, K, N, Lmax, waves, workspace, workspace_bytes, stream);
  if (rc != M2H_OK) return rc;
  return m2h_stft_waves(waves, B, K, N, n_gt, gt_with_phase, flags, mix_mag, mix_phase, gt, stream);
}

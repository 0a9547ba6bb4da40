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
#include "render_conv24.cuh"
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
  M2hConv24Tables* conv24 = nullptr;
  M2hStftTables* stft = nullptr;
  M2hStftTables* istft = nullptr;       // Bluestein tables of the 1022-point inverse DFT (librosa.istft's inferred n_fft)
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
// launch with programmatic stream serialisation (M2H_PDL_WAIT in the kernels): the STFT kernel stages its tables while the
// last convolution items finish
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, void* stream, Args&&... args) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("M2H_NO_PDL"); enabled = (e && atoi(e)) ? 0 : 1; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// VEC = packed row accesses (even N / Lmax, aligned base pointers: conv_can_vectorise); the scalar variant serves the rest
template <bool VEC>
__global__ void __launch_bounds__(M2H_CONV_THREADS, 1)
m2h_rir_conv_kernel(const ConvParams P, const M2hConvTables* __restrict__ tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ConvSmem& s = *reinterpret_cast<ConvSmem*>(smem_raw);
  conv_cta_body<VEC>(P, s, tab, blockIdx.x, gridDim.x);
}

// RIRs of at most 16384 taps: circular length 24576, two CTAs per SM (render_conv24.cuh)
template <bool VEC>
__global__ void __launch_bounds__(M2H_C24_THREADS, M2H_C24_CTAS_PER_SM)
m2h_rir_conv24_kernel(const ConvParams P, const M2hConv24Tables* __restrict__ tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Conv24Smem& s = *reinterpret_cast<Conv24Smem*>(smem_raw);
  conv24_cta_body<VEC>(P, s, tab, blockIdx.x, gridDim.x);
}

// 64 registers / thread (no spills): two 512-thread CTAs per SM, 2 x 86 KB of shared memory
__global__ void __launch_bounds__(M2H_STFT_THREADS, M2H_STFT_CTAS_PER_SM)
m2h_istft_kernel(const StftParams P, const M2hStftTables* __restrict__ tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
  istft_cta_body(P, s, tab, blockIdx.x, gridDim.x);
}

// CTAS = resident CTAs per SM the register allocation targets: 3 (80 registers, no spills) or 4 (64 registers, 72 bytes of
// spills, 32 warps per SM); M2H_STFT_CTAS selects at run time (measured: see DESIGN.md 4.2)
template <bool F16, int CTAS>
__global__ void __launch_bounds__(M2H_STFT_THREADS, CTAS)
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
  static M2hConv24Tables hc24;
  static M2hStftTables hs, hi;
  static bool host_built = false;
  if (!host_built) { build_conv_tables(hc); build_conv24_tables(hc24); build_stft_tables(hs); build_istft_tables(hi); host_built = true; }
  M2H_CUDA_CHECK(cudaMalloc(&d.conv24, sizeof(M2hConv24Tables)));
  M2H_CUDA_CHECK(cudaMemcpy(d.conv24, &hc24, sizeof(hc24), cudaMemcpyHostToDevice));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv24_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Conv24Smem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv24_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Conv24Smem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv24_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv24_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaMalloc(&d.conv, sizeof(M2hConvTables)));
  M2H_CUDA_CHECK(cudaMalloc(&d.stft, sizeof(M2hStftTables)));
  M2H_CUDA_CHECK(cudaMemcpy(d.conv, &hc, sizeof(hc), cudaMemcpyHostToDevice));
  M2H_CUDA_CHECK(cudaMemcpy(d.stft, &hs, sizeof(hs), cudaMemcpyHostToDevice));
  M2H_CUDA_CHECK(cudaMalloc(&d.istft, sizeof(M2hStftTables)));
  M2H_CUDA_CHECK(cudaMemcpy(d.istft, &hi, sizeof(hi), cudaMemcpyHostToDevice));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_istft_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem)));
  M2H_CUDA_CHECK(cudaFuncSetAttribute(m2h_rir_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem)));
  // the driver's default carve-out (102 KB, ncu r1b) left room for ONE STFT CTA per SM: ask for the maximum
#define M2H_STFT_ATTR(K) \
  M2H_CUDA_CHECK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem))); \
  M2H_CUDA_CHECK(cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  M2H_STFT_ATTR((m2h_stft_kernel<false, 3>)) M2H_STFT_ATTR((m2h_stft_kernel<true, 3>))
  M2H_STFT_ATTR((m2h_stft_kernel<false, 4>)) M2H_STFT_ATTR((m2h_stft_kernel<true, 4>))
#undef M2H_STFT_ATTR
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
  // one spectrum-scratch slot per resident CTA of either convolution kernel (F = 32768: 128 KB x 1 per SM; F = 24576: 96 KB x 2 per SM)
  const size_t a = (size_t)n_sms * M2H_CONV_SCRATCH_F4 * sizeof(float4);
  const size_t b = (size_t)n_sms * M2H_C24_CTAS_PER_SM * M2H_C24_SCRATCH_F2 * sizeof(float2);
  return a > b ? a : b;
}

static int launch_conv(const int16_t* mono, const int32_t* mono_index, const float* rir, const int32_t* rir_len, const int32_t* rir_index,
                       int B, int K, int N, int Lmax, int16_t* waves, void* workspace, size_t workspace_bytes, void* stream, const char* who) {
  if (B < 0 || K <= 0 || N <= 0 || N > M2H_MAX_SAMPLES || Lmax < 0) { m2h_set_error("%s: bad sizes B=%d K=%d N=%d Lmax=%d (N <= %d)", who, B, K, N, Lmax, M2H_MAX_SAMPLES); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!mono || !rir_len || !waves || (!rir && Lmax > 0) || !workspace) { m2h_set_error("%s: null pointer", who); return M2H_ERR_ARG; }
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  const int n_items = B * K;
  static int force_f32768 = -1;
  if (force_f32768 < 0) { const char* e = getenv("M2H_CONV_F32768"); force_f32768 = (e && atoi(e)) ? 1 : 0; }
  if (Lmax <= M2H_C24_MAX_TAPS && !force_f32768) {       // every RIR fits the F = 24576 kernel (the reference's RIRs are 1 s = 16000 taps)
    int grid24 = M2H_C24_CTAS_PER_SM * g_tables[dev].n_sms;
    // fewer items than half the CTA slots: one CTA per (item, ear) - the clip's transform is repeated, the item's critical path
    // drops from 5 to 3 transforms (M2H_CONV_EAR_SPLIT=0 disables)
    static int allow_split = -1;
    if (allow_split < 0) { const char* e = getenv("M2H_CONV_EAR_SPLIT"); allow_split = (e && !atoi(e)) ? 0 : 1; }
    const int ear_split = allow_split && 2 * n_items <= grid24 ? 1 : 0;
    const int n_units = ear_split ? 2 * n_items : n_items;
    if (grid24 > n_units) grid24 = n_units;
    if (workspace_bytes < (size_t)grid24 * M2H_C24_SCRATCH_F2 * sizeof(float2)) { m2h_set_error("%s: workspace %zu B < %zu B", who, workspace_bytes, (size_t)grid24 * M2H_C24_SCRATCH_F2 * sizeof(float2)); return M2H_ERR_WORKSPACE; }
    ConvParams P;
    P.mono = mono; P.rir = rir; P.rir_len = rir_len; P.mono_index = mono_index; P.rir_index = rir_index;
    P.waves = waves; P.scratch = (float4*)workspace;
    P.n_items = n_items; P.N = N; P.Lmax = Lmax; P.ear_split = ear_split;
    const M2hConv24Tables* tab24 = g_tables[dev].conv24;
    if (conv_can_vectorise(P)) M2H_CUDA_CHECK(launch_pdl(m2h_rir_conv24_kernel<true>, grid24, M2H_C24_THREADS, sizeof(Conv24Smem), stream, P, tab24));
    else M2H_CUDA_CHECK(launch_pdl(m2h_rir_conv24_kernel<false>, grid24, M2H_C24_THREADS, sizeof(Conv24Smem), stream, P, tab24));
    return M2H_OK;
  }
  int grid = g_tables[dev].n_sms;
  if (grid > n_items) grid = n_items;
  if (workspace_bytes < (size_t)grid * M2H_CONV_SCRATCH_F4 * sizeof(float4)) { m2h_set_error("%s: workspace %zu B < %zu B", who, workspace_bytes, (size_t)grid * M2H_CONV_SCRATCH_F4 * sizeof(float4)); return M2H_ERR_WORKSPACE; }
  ConvParams P;
  P.mono = mono; P.rir = rir; P.rir_len = rir_len; P.mono_index = mono_index; P.rir_index = rir_index;
  P.waves = waves; P.scratch = (float4*)workspace;
  P.n_items = n_items; P.N = N; P.Lmax = Lmax; P.ear_split = 0;
  const M2hConvTables* tab = g_tables[dev].conv;
  if (conv_can_vectorise(P)) M2H_CUDA_CHECK(launch_pdl(m2h_rir_conv_kernel<true>, grid, M2H_CONV_THREADS, sizeof(ConvSmem), stream, P, tab));
  else M2H_CUDA_CHECK(launch_pdl(m2h_rir_conv_kernel<false>, grid, M2H_CONV_THREADS, sizeof(ConvSmem), stream, P, tab));
  return M2H_OK;
}

extern "C" int m2h_rir_convolve(const int16_t* mono, const float* rir, const int32_t* rir_len,
                                int B, int K, int N, int Lmax, int16_t* waves,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return launch_conv(mono, nullptr, rir, rir_len, nullptr, B, K, N, Lmax, waves, workspace, workspace_bytes, stream, "m2h_rir_convolve");
}

extern "C" int m2h_rir_convolve_indexed(const int16_t* clip_pool, const int32_t* clip_index, const float* rir_pool, const int32_t* rir_pool_len,
                                        const int32_t* rir_index, int B, int K, int N, int Lmax, int16_t* waves,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  if (B > 0 && (!clip_index || !rir_index)) { m2h_set_error("m2h_rir_convolve_indexed: null index"); return M2H_ERR_ARG; }
  return launch_conv(clip_pool, clip_index, rir_pool, rir_pool_len, rir_index, B, K, N, Lmax, waves, workspace, workspace_bytes, stream, "m2h_rir_convolve_indexed");
}

// ---------------------------------------------------------------------------------------------
// wav ingest: what scipy.io.wavfile.read + the [L, 2] -> per-ear split do on the host in
// habitat_audio/simulator_train.py:396-415 and audio_separation/pretrain/datasets/dataset.py:155-176
// ---------------------------------------------------------------------------------------------
static uint32_t rd_u32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t rd_u16(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

extern "C" int m2h_wav_parse(const void* bytes, size_t n_bytes, m2h_wav_info* out) {
  if (!bytes || !out) { m2h_set_error("m2h_wav_parse: null pointer"); return M2H_ERR_ARG; }
  const unsigned char* p = (const unsigned char*)bytes;
  memset(out, 0, sizeof(*out));
  if (n_bytes < 12 || memcmp(p, "RIFF", 4) != 0 || memcmp(p + 8, "WAVE", 4) != 0) { m2h_set_error("m2h_wav_parse: not a RIFF/WAVE file"); return M2H_ERR_ARG; }
  size_t pos = 12;
  bool have_fmt = false;
  uint32_t tag = 0, block_align = 0;
  while (pos + 8 <= n_bytes) {
    const uint32_t size = rd_u32(p + pos + 4);
    const unsigned char* body = p + pos + 8;
    if (memcmp(p + pos, "fmt ", 4) == 0) {
      if (size < 16 || pos + 8 + 16 > n_bytes) { m2h_set_error("m2h_wav_parse: truncated fmt chunk"); return M2H_ERR_ARG; }
      tag = rd_u16(body); out->channels = (int)rd_u16(body + 2); out->sample_rate = (int)rd_u32(body + 4);
      block_align = rd_u16(body + 12); out->bits_per_sample = (int)rd_u16(body + 14);
      if (tag == 0xFFFEu && size >= 26) tag = rd_u16(body + 24);         // WAVE_FORMAT_EXTENSIBLE: first two bytes of the sub-format GUID
      have_fmt = true;
    } else if (memcmp(p + pos, "data", 4) == 0) {
      if (!have_fmt) { m2h_set_error("m2h_wav_parse: data chunk before fmt chunk"); return M2H_ERR_ARG; }
      size_t avail = n_bytes - (pos + 8);
      size_t len = size < avail ? size : avail;                           // scipy tolerates a short final chunk
      if (tag == 1 && out->bits_per_sample == 16) out->format = M2H_WAV_PCM16;
      else if (tag == 1 && out->bits_per_sample == 32) out->format = M2H_WAV_PCM32;
      else if (tag == 3 && out->bits_per_sample == 32) out->format = M2H_WAV_FLOAT32;
      else { m2h_set_error("m2h_wav_parse: unsupported sample format (tag %u, %d bits)", tag, out->bits_per_sample); return M2H_ERR_ARG; }
      if (out->channels < 1 || block_align != (uint32_t)(out->channels * out->bits_per_sample / 8)) { m2h_set_error("m2h_wav_parse: inconsistent block alignment"); return M2H_ERR_ARG; }
      out->data_offset = pos + 8;
      out->frames = (long long)(len / block_align);
      return M2H_OK;
    }
    pos += 8 + (size_t)size + (size & 1u);
  }
  m2h_set_error("m2h_wav_parse: no data chunk");
  return M2H_ERR_ARG;
}

// interleaved wav frames [frames][channels] (PCM16 / PCM32 / FLOAT32) -> fp32 ear-planar pool row [2][Lmax]; taps beyond
// `frames` are zeroed.  One-channel RIRs (never produced by SoundSpaces) feed both ears.
__global__ void rir_ingest_kernel(const unsigned char* __restrict__ raw, int format, int channels, int frames, float* __restrict__ row, int Lmax) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * Lmax; i += gridDim.x * blockDim.x) {
    const int ear = i / Lmax, j = i - ear * Lmax;
    float v = 0.f;
    if (j < frames) {
      const size_t e = (size_t)j * channels + (ear < channels ? ear : channels - 1);
      if (format == M2H_WAV_PCM16) v = (float)reinterpret_cast<const int16_t*>(raw)[e];
      else if (format == M2H_WAV_PCM32) v = (float)reinterpret_cast<const int32_t*>(raw)[e];
      else v = reinterpret_cast<const float*>(raw)[e];
    }
    row[i] = v;
  }
}
// int16 clip: channel 0 of the wav, truncated / zero-padded to N samples
__global__ void clip_ingest_kernel(const int16_t* __restrict__ raw, int channels, int frames, int16_t* __restrict__ row, int N) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) row[i] = i < frames ? raw[(size_t)i * channels] : (int16_t)0;
}

extern "C" int m2h_rir_ingest(const void* raw_frames, int format, int channels, int frames, float* pool_row, int Lmax, void* stream) {
  if (!pool_row || Lmax <= 0 || frames < 0 || frames > Lmax || channels < 1 || (frames > 0 && !raw_frames) ||
      (format != M2H_WAV_PCM16 && format != M2H_WAV_PCM32 && format != M2H_WAV_FLOAT32)) { m2h_set_error("m2h_rir_ingest: bad arguments (frames %d, Lmax %d, channels %d, format %d)", frames, Lmax, channels, format); return M2H_ERR_ARG; }
  if (((uintptr_t)raw_frames & 3) != 0) { m2h_set_error("m2h_rir_ingest: raw frames must be 4-byte aligned"); return M2H_ERR_ARG; }
  rir_ingest_kernel<<<(2 * Lmax + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)raw_frames, format, channels, frames, pool_row, Lmax);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}
extern "C" int m2h_clip_ingest(const void* raw_frames, int format, int channels, int frames, int16_t* pool_row, int N, void* stream) {
  if (!pool_row || N <= 0 || frames < 0 || channels < 1 || (frames > 0 && !raw_frames)) { m2h_set_error("m2h_clip_ingest: bad arguments"); return M2H_ERR_ARG; }
  if (format != M2H_WAV_PCM16) { m2h_set_error("m2h_clip_ingest: the reference's clips are 16-bit PCM (format %d given)", format); return M2H_ERR_ARG; }
  if (((uintptr_t)raw_frames & 1) != 0) { m2h_set_error("m2h_clip_ingest: raw frames must be 2-byte aligned"); return M2H_ERR_ARG; }
  clip_ingest_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const int16_t*)raw_frames, channels, frames, pool_row, N);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

static int launch_stft(const StftParams& P, int dev, void* stream) {
  const int n_items = P.B * P.n_signals * ((P.T + M2H_STFT_GROUPS - 1) / M2H_STFT_GROUPS);      // (mixture, signal, M2H_STFT_GROUPS frames)
  if (n_items == 0) return M2H_OK;
  static int ctas = -1;
  if (ctas < 0) { const char* e = getenv("M2H_STFT_CTAS"); ctas = (e && atoi(e) == 4) ? 4 : (e && atoi(e) == 3 ? 3 : M2H_STFT_CTAS_PER_SM); }
  int grid = ctas * g_tables[dev].n_sms;
  if (grid > n_items) grid = n_items;
  // outputs are written as (left, right) pairs
  const void* outs[5] = {P.mix_mag, P.mix_phase, P.gt_mag, P.f_mag, P.f_phase};
  for (int i = 0; i < 5; ++i)
    if (((uintptr_t)outs[i] & 7) != 0) { m2h_set_error("stft: output buffers must be 8-byte aligned"); return M2H_ERR_ARG; }
  const M2hStftTables* tab = g_tables[dev].stft;
  if (ctas == 4) {
    if (P.flags & M2H_OUT_F16) M2H_CUDA_CHECK(launch_pdl(m2h_stft_kernel<true, 4>, grid, M2H_STFT_THREADS, sizeof(StftSmem), stream, P, tab));
    else M2H_CUDA_CHECK(launch_pdl(m2h_stft_kernel<false, 4>, grid, M2H_STFT_THREADS, sizeof(StftSmem), stream, P, tab));
  } else {
    if (P.flags & M2H_OUT_F16) M2H_CUDA_CHECK(launch_pdl(m2h_stft_kernel<true, 3>, grid, M2H_STFT_THREADS, sizeof(StftSmem), stream, P, tab));
    else M2H_CUDA_CHECK(launch_pdl(m2h_stft_kernel<false, 3>, grid, M2H_STFT_THREADS, sizeof(StftSmem), stream, P, tab));
  }
  return M2H_OK;
}

extern "C" int m2h_stft_waves(const int16_t* waves, int B, int K, int N, int n_gt, int gt_with_phase, int flags,
                              void* mix_mag, void* mix_phase, void* gt, void* stream) {
  if (B < 0 || K <= 0 || K > 256 || N < M2H_HOP || N > M2H_MAX_SAMPLES || n_gt < 0 || n_gt > K) { m2h_set_error("m2h_stft_waves: bad sizes B=%d K=%d N=%d n_gt=%d", B, K, N, n_gt); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!waves || (n_gt > 0 && !gt)) { m2h_set_error("m2h_stft_waves: null pointer"); return M2H_ERR_ARG; }
  const bool gt_only = (flags & M2H_RENDER_GT_ONLY) != 0;
  if (gt_only && n_gt == 0) return M2H_OK;
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.waves = waves; P.B = B; P.K = K; P.N = N; P.T = 1 + (N - 1) / M2H_HOP;
  P.n_gt = n_gt; P.n_signals = gt_only ? n_gt : 1 + n_gt; P.sig0 = gt_only ? 1 : 0;
  P.flags = ((flags & M2H_RENDER_RAW_MAG) ? 0 : M2H_OUT_MIX_LOG1P) | ((flags & M2H_RENDER_F16_OBS) ? M2H_OUT_F16 : 0);
  P.mix_mag = mix_mag; P.mix_phase = mix_phase; P.gt_mag = gt; P.gt_interleaved_phase = gt_with_phase ? 1 : 0;
  return launch_stft(P, dev, stream);
}

extern "C" int m2h_stft_waves_x0(const int16_t* waves, int B, int K, int N, void* x0, float* mix_mag, void* stream) {
  if (B < 0 || K <= 0 || K > 256 || N < M2H_HOP || N > M2H_MAX_SAMPLES) { m2h_set_error("m2h_stft_waves_x0: bad sizes B=%d K=%d N=%d", B, K, N); return M2H_ERR_ARG; }
  if (1 + (N - 1) / M2H_HOP != 32) { m2h_set_error("m2h_stft_waves_x0: the separator's input has 32 frames, N=%d gives %d", N, 1 + (N - 1) / M2H_HOP); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!waves || !x0 || ((uintptr_t)x0 & 15) != 0) { m2h_set_error("m2h_stft_waves_x0: null or misaligned pointer"); return M2H_ERR_ARG; }
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.waves = waves; P.B = B; P.K = K; P.N = N; P.T = 32; P.n_signals = 1; P.flags = M2H_OUT_MIX_LOG1P;
  P.x0 = x0; P.mix_mag = mix_mag;
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

extern "C" int m2h_istft(const float* mag, const float* phase, int B, int C, int T, int length, float* out, void* stream) {
  if (B < 0 || (C != 1 && C != 2) || T < 1 || T > 4096 || length < 1) { m2h_set_error("m2h_istft: bad sizes B=%d C=%d T=%d length=%d", B, C, T, length); return M2H_ERR_ARG; }
  if (B == 0) return M2H_OK;
  if (!mag || !phase || !out) { m2h_set_error("m2h_istft: null pointer"); return M2H_ERR_ARG; }
  int dev;
  int rc = ensure_init(&dev);
  if (rc != M2H_OK) return rc;
  M2H_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)B * C * length * sizeof(float), (cudaStream_t)stream));     // frames are overlap-added
  StftParams P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.T = T; P.i_mag = mag; P.i_phase = phase; P.i_out = out; P.i_channels = C; P.i_length = length;
  const int n_items = B * ((T + M2H_STFT_GROUPS - 1) / M2H_STFT_GROUPS);
  int grid = M2H_STFT_CTAS_PER_SM * g_tables[dev].n_sms;
  if (grid > n_items) grid = n_items;
  m2h_istft_kernel<<<grid, M2H_STFT_THREADS, sizeof(StftSmem), (cudaStream_t)stream>>>(P, g_tables[dev].istft);
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

extern "C" int m2h_render_mixture(const int16_t* mono, const float* rir, const int32_t* rir_len,
                                  int B, int K, int N, int Lmax, int n_gt, int gt_with_phase, int flags,
                                  int16_t* waves, void* mix_mag, void* mix_phase, void* gt,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  int rc = m2h_rir_convolve(mono, rir, rir_len, B, K, N, Lmax, waves, workspace, workspace_bytes, stream);
  if (rc != M2H_OK) return rc;
  return m2h_stft_waves(waves, B, K, N, n_gt, gt_with_phase, flags, mix_mag, mix_phase, gt, stream);
}

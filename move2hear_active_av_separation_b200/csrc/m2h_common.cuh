// Shared helpers for the render kernels.
//
// The kernel bodies in render_conv.cuh / render_stft.cuh are written as a sequence of
// barrier-separated phases.  Under nvcc each phase runs once per thread (tid = threadIdx.x)
// with __syncthreads() between phases.  Under -DM2H_HOST_EMU (plain g++, tests only) the same
// source runs each phase as a loop over all thread ids, which lets the CPU test-suite check
// every index computation of the kernels against the oracle without a GPU.  The emulation is
// never part of the product library.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef M2H_HOST_EMU
  #include <cmath>
  #include <cstring>
  struct float2 { float x, y; };
  struct float4 { float x, y, z, w; };
  static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
  static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
  #define M2H_DEV static inline
  #define M2H_HD static inline
  #define M2H_FOR_THREADS(tid, nthreads) for (int tid = 0; tid < (nthreads); ++tid)
  #define M2H_SYNC() do {} while (0)
  #define M2H_SYNC_GROUP128() do {} while (0)
  #define M2H_SYNC_WARP() do {} while (0)
  #define M2H_PDL_WAIT() do {} while (0)
  #define M2H_UNROLL
  #define M2H_RESTRICT
  M2H_DEV float m2h_rintf(float v) { return nearbyintf(v); }
  M2H_DEV float m2h_ldg(const float* p) { return *p; }
  M2H_DEV float2 m2h_ldg2(const float2* p) { return *p; }
  M2H_DEV float2 m2h_ldg2_nc(const float2* p) { return *p; }
  M2H_DEV uint32_t m2h_ldg_u32(const uint32_t* p) { return *p; }
  M2H_DEV float2 m2h_ld2_keep(const float2* p) { return *p; }
  M2H_DEV void m2h_st2_keep(float2* p, float2 v) { *p = v; }
  M2H_DEV float2 m2h_ldg2_stream(const float2* p) { return *p; }
  M2H_DEV float m2h_ldg_stream(const float* p) { return *p; }
  M2H_DEV uint32_t m2h_ldg_u32_stream(const uint32_t* p) { return *p; }
  M2H_DEV void m2h_st_u32_stream(uint32_t* p, uint32_t v) { *p = v; }
  M2H_DEV int m2h_f2i_rn(float v) {           // cvt.rni.s32.f32: round half to even, saturate
    const float r = nearbyintf(v);
    return r >= 2147483648.f ? 2147483647 : (r <= -2147483648.f ? (int)(-2147483647 - 1) : (int)r);
  }
  M2H_DEV uint32_t m2h_funnel_r(uint32_t lo, uint32_t hi, int sh) { return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
  M2H_DEV void m2h_prefetch_l2(const void*) {}
  M2H_DEV float m2h_sinpif(float x) { return (float)sin(3.14159265358979323846 * (double)x); }
  M2H_DEV void m2h_prefetch_l1(const void*) {}
#else
  #include <cuda_runtime.h>
  #include <cuda_fp16.h>
  #define M2H_DEV __device__ __forceinline__
  #define M2H_HD __host__ __device__ __forceinline__
  #define M2H_FOR_THREADS(tid, nthreads) for (int tid = threadIdx.x, m2h_once_ = 1; m2h_once_; m2h_once_ = 0)
  #define M2H_SYNC() __syncthreads()
  // named barrier of the 128-thread group a thread belongs to (ids 1..): groups of a CTA that work on independent data
  // drift apart instead of meeting at every phase boundary
  // (immediate ids: with a register operand ptxas reserves all 16 barriers of the CTA)
  __device__ __forceinline__ void m2h_sync_group128() {
    switch (threadIdx.x >> 7) {
      case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
      case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
      case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
      default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
    }
  }
  #define M2H_SYNC_GROUP128() m2h_sync_group128()
  #define M2H_SYNC_WARP() __syncwarp()
  // programmatic dependent launch (see unet_gemm.cuh::pdl_wait): wait for the predecessor, then let the successor's prologue start
  #define M2H_PDL_WAIT() asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory")
  #define M2H_UNROLL _Pragma("unroll")
  #define M2H_RESTRICT __restrict__
  M2H_DEV float m2h_rintf(float v) { return rintf(v); }
  M2H_DEV float m2h_ldg(const float* p) { return __ldg(p); }
  M2H_DEV float2 m2h_ldg2(const float2* p) { return __ldg(p); }
  // plain (coherent) load: data written earlier in the SAME kernel by the same thread must not go through the read-only path
  M2H_DEV float2 m2h_ldg2_nc(const float2* p) { return *p; }
  M2H_DEV uint32_t m2h_ldg_u32(const uint32_t* p) { return __ldg(p); }
  // L2 eviction priorities (cache-hint policies: compile-time descriptors).  "keep": a CTA's L2-resident scratch slot, rewritten
  // for every work item; "stream": inputs read once / outputs written once.  Without them the streaming traffic of a launch
  // pushes dirty scratch lines out to DRAM (conv24: 555 MB of DRAM traffic for 459 MB of compulsory bytes).  No "memory" clobbers:
  // the scratch is re-read only after a CTA barrier, the outputs never; a clobber would serialise the copy-out loop's
  // shared-memory loads behind each store.
  M2H_DEV unsigned long long m2h_policy_keep() { unsigned long long p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
  M2H_DEV unsigned long long m2h_policy_stream() { unsigned long long p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
  M2H_DEV float2 m2h_ld2_keep(const float2* p) {
    float2 v; asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(m2h_policy_keep())); return v;
  }
  M2H_DEV void m2h_st2_keep(float2* p, float2 v) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(m2h_policy_keep()));
  }
  M2H_DEV float2 m2h_ldg2_stream(const float2* p) {
    float2 v; asm("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(m2h_policy_stream())); return v;
  }
  M2H_DEV float m2h_ldg_stream(const float* p) {
    float v; asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(m2h_policy_stream())); return v;
  }
  M2H_DEV uint32_t m2h_ldg_u32_stream(const uint32_t* p) {
    uint32_t v; asm("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(m2h_policy_stream())); return v;
  }
  M2H_DEV void m2h_st_u32_stream(uint32_t* p, uint32_t v) {
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(m2h_policy_stream()));
  }
  M2H_DEV int m2h_f2i_rn(float v) { return __float2int_rn(v); }
  M2H_DEV uint32_t m2h_funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, sh); }
  M2H_DEV void m2h_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
  M2H_DEV float m2h_sinpif(float x) { return sinpif(x); }
  M2H_DEV void m2h_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

// ---- complex arithmetic on float2 ---------------------------------------------------------
// Device build: Blackwell's packed fp32 pipe (add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  A float2 is one 64-bit
// register pair; the SASS operands carry a half swap (.LO_HI), a per-half negation (.NP / .PN) and a 32-bit broadcast
// (.F32) for free, so a complex add is ONE instruction, a complex multiply TWO (FMUL2 + FFMA2) and a rotation by +-i none
// (it folds into the consumer's operand modifiers) - half the issue slots of the scalar forms, which is what bounds these
// kernels (ncu r1g: FP32 issue, not HBM).  Rounding is the same as the scalar code nvcc contracts (one product rounded,
// the other fused).  The host emulation keeps the scalar spelling.
#ifdef M2H_HOST_EMU
M2H_DEV float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
M2H_DEV float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
M2H_DEV float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
M2H_DEV float2 c_mulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// a * (c - i s) and a * (c + i s) for real scalars c, s
M2H_DEV float2 c_rot_fwd(float2 a, float c, float s) { return make_float2(a.x * c + a.y * s, a.y * c - a.x * s); }
M2H_DEV float2 c_rot_inv(float2 a, float c, float s) { return make_float2(a.x * c - a.y * s, a.y * c + a.x * s); }
M2H_DEV float2 c_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
#else
M2H_DEV float2 c_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
M2H_DEV float2 c_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
M2H_DEV float2 c_mul(float2 a, float2 b) { return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x))); }
M2H_DEV float2 c_mulc(float2 a, float2 b) { return __ffma2_rn(make_float2(a.y, a.x), make_float2(b.y, -b.y), __fmul2_rn(a, make_float2(b.x, b.x))); }
M2H_DEV float2 c_rot_fwd(float2 a, float c, float s) { return __ffma2_rn(make_float2(a.y, a.x), make_float2(s, -s), __fmul2_rn(a, make_float2(c, c))); }
M2H_DEV float2 c_rot_inv(float2 a, float c, float s) { return __ffma2_rn(make_float2(a.y, a.x), make_float2(-s, s), __fmul2_rn(a, make_float2(c, c))); }
M2H_DEV float2 c_scale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
#endif
M2H_DEV float2 c_conj(float2 a) { return make_float2(a.x, -a.y); }
template <bool CONJ> M2H_DEV float2 c_mul_sel(float2 a, float2 w) { return CONJ ? c_mulc(a, w) : c_mul(a, w); }

// ---- problem constants ----------------------------------------------------------------------
// Reference STFT: n_fft = 1023, hop = 512 (simulator_train.py:426, dataset.py:20-21).
#define M2H_NFFT 1023
#define M2H_HOP 512
#define M2H_NBINS 512
// Linear-convolution FFT length.  For N <= 16384 samples only RIR taps inside
// [c-N+1, c+N-1] (c = (L-1)/2) can reach the "same"-mode output window, i.e. at most 2N-1 taps,
// and a circular length F >= 2N-1 keeps that window alias-free for every L.
#define M2H_CONV_F 32768
#define M2H_CONV_M 16384              // complex length of the packed real FFT
#define M2H_MAX_SAMPLES 16384

// twiddle tables living in device global memory (copied to shared memory by each CTA)
// Per-thread twiddle anchors of the conv kernel (512 threads), correctly rounded from double:
//   thr[0][t] = W_16384^t        thr[1][t] = W_16384^(8t)            passes A / A'   (radix 32, Q = 8)
//   thr[2][t] = W_512^(t&15)     thr[3][t] = W_512^(8*(t&15))        pass B          (radix 32)
//   thr[4][t] = W_512^(t&31)     thr[5][t] = W_512^(4*(t&31))        pass C'         (radix 16, Q = 4)
//   thr[6][t] = W_32768^(k1 + 32*k2) of the thread's row A in the fused pass (render_conv.cuh::conv_rows)
//   unit_tw[u]: twiddle of pair unit u on the two self-paired rows: u < 8: W_32768^(1024u); u >= 8: W_32768^(512 + 1024(u-8))
struct M2hConvTables {
  float2 thr[7][512];
  float2 unit_tw[16];
};
// thr[0][l] = W_2048^l, thr[1][l] = W_2048^(4l)  (P1/P5);  thr[2][l] = W_128^(l&7), thr[3][l] = W_128^(4*(l&7))  (P2);
// thr[4][l] = W_128^(l&15), thr[5][l] = W_128^(4*(l&15))  (P3)      -- l = thread id within its 128-thread group
struct M2hStftTables {
  float2 thr[6][128];
  float2 bf[2048];      // FFT_2048 of the Bluestein chirp, stored in (k1,k2,k3) order: index (k1*16+k2)*8+k3 <-> f = k1+16*k2+256*k3
  float2 wc[1024];      // hann[n] * conj(chirp[n]), n < 1023 (entry 1023 = 0)
  float2 cc[1024];      // conj(chirp[m]) / 4096   (1/2048 inverse-FFT scale, 1/2 ear-untangle scale)
};

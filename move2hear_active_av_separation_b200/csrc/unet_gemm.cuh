// Shared definitions of the tcgen05 implicit-GEMM conv / transposed-conv kernels of the passive-separator U-Net (sm_100a):
// the layer descriptor (UGLayer), mbarrier / tcgen05 / TMEM helpers, shared-memory operand descriptors.
//
// The kernels replace nn.Conv2d(k4,s2,p1) / nn.ConvTranspose2d(k4,s2,p1) + BatchNorm2d (inference statistics) +
// LeakyReLU(0.2) / ReLU, and the 1x1 output conv, of audio_separation/rl/models/separator_cnn.py:5-24, 44-50, 128-135:
//
//   D[m, n] = sum_k A[m, k] * W[n, k]       m = (image, output pixel of this phase), n = output channel,
//                                             k = (tap, input channel)
// A is never materialised: TMA boxes of the haloed NHWC bf16 activation tensor land as the swizzled K-major tiles tcgen05.mma
// consumes (unet_gemm_tma.cuh, unet_slab.cuh).  A transposed conv runs as its 4 sub-pixel phases: each is a 2x2-tap stride-1 conv.
// (The round-1 kernel that gathered A with cp.async lived here; its last call site - the 32-channel data gradient of the last
// decoder layer - moved to the TMA kernel with 64B-swizzled boxes in round 2.)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define UG_BM 128
#define UG_BK 64
#define UG_MAX_TAPS 16

enum { UG_OUT_NHWC_BF16 = 0, UG_OUT_NCHW_F32 = 1, UG_OUT_HEAD_NCHW_F32 = 2 };

struct UGLayer {
  // input activation, NHWC bf16 with a zero border of in_pad pixels:
  // element (b, y, x, c) at in[((b*(Hin+2p) + y+p)*(Win+2p) + x+p)*in_ld + c]
  const __nv_bfloat16* in;
  int Hin, Win, Cin, cin_shift, in_ld, in_pad;
  // GEMM rows: m = (b*Hq + qy)*Wq + qx ; input pixel = (qy*in_stride + dy[tap], qx*in_stride + dx[tap])
  int M, Hq, Wq, in_stride;
  int n_taps;
  int8_t dy[4][UG_MAX_TAPS], dx[4][UG_MAX_TAPS];      // per phase
  // weights, bf16 [n_phase][Cout][Kpad], k = tap*Cin + c
  const __nv_bfloat16* w;
  int Cout, Kvalid, Kpad, n_phase;
  // epilogue
  const float* scale;      // [n_mod] or null (=1)
  const float* bias;       // [n_mod] or null (=0)
  int n_mod;               // scale/bias index = n % n_mod
  float slope;             // y = v > 0 ? v : v*slope   (1 = identity, 0 = ReLU, 0.2 = LeakyReLU)
  const float* cls;        // [B] (target_class + 1) or null
  const float* cls_w;      // [3][3][Cout] sums of the class-plane weights over the valid taps, or null
  // output: pixel (qy*out_stride + py, qx*out_stride + px) of a grid with a border of out_pad pixels;
  // column n -> pix_off[n / n_split] + n % n_split (the 1x1-input decoder layer spreads column blocks over 4 pixels)
  void* out;
  int out_mode, Hout, Wout, out_stride, out_ld, out_pad, n_split, pix_off[4];
  int n_valid;             // NCHW modes: channels >= n_valid are padding and are not stored; Cout_real = n_valid
  // UG_OUT_HEAD_NCHW_F32 (BN = 32 only): the activated 32-channel pixel goes through the 1x1 head conv in the epilogue
  //   z[n] = head_b[n] + sum_c head_w[n][c] * y[c]   (separator_cnn.py:135) and z is stored as fp32 NCHW
  const float* head_w;     // fp32 [32][32] (zero padded)
  const float* head_b;     // fp32 [32]
};

namespace ug {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream drains; pdl_wait() blocks until the predecessor has completed and flushed (no-op for an
// ordinary launch).  Every kernel here runs its prologue (barrier init, TMEM allocation, parameter staging), waits, and
// only then lets ITS dependents launch - so a dependent's prologue can overlap this kernel's tail, never an older kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B (64 bf16), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UG_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major tile of 64-byte rows (32 bf16 of K per row) in the 64B swizzle: 8-row atoms of 512 B
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                          // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ug


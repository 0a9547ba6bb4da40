// Implicit-GEMM conv / transposed-conv layer of the passive-separator U-Net on tcgen05 (sm_100a).
//
// Replaces nn.Conv2d(k4,s2,p1) / nn.ConvTranspose2d(k4,s2,p1) + BatchNorm2d (inference statistics) +
// LeakyReLU(0.2) / ReLU, and the 1x1 output conv, of audio_separation/rl/models/separator_cnn.py:5-24, 44-50, 128-135.
//
//   D[m, n] = sum_k A[m, k] * W[n, k]       m = (image, output pixel of this phase), n = output channel,
//                                             k = (tap, input channel)
// A is never materialised: four producer warps gather it straight from the NHWC bf16 activation tensor with
// 16-byte cp.async (zero-filled outside the image = the conv padding) into the 128B-swizzled K-major shared-memory
// layout tcgen05.mma consumes; the same warps stream the [n, k] weight tile.  One elected thread issues
// tcgen05.mma (M = 128, N = BN, K = 16 per instruction, bf16 x bf16 -> fp32 in TMEM); completion of the MMAs
// that read a stage is signalled back to the producers with tcgen05.commit on an mbarrier.  After the last
// k-block the producer warps become the epilogue: tcgen05.ld the accumulator rows, apply
// (acc + class-plane term) * scale[n] + bias[n], the activation, and store bf16 NHWC (possibly into the channel
// slice of a concat buffer, which is how the skip connections avoid torch.cat) or fp32 NCHW (the mask layout).
//
// A transposed conv runs as its 4 sub-pixel phases (blockIdx.z): each phase is a 2x2-tap stride-1 conv.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define UG_BM 128
#define UG_BK 64
#define UG_PRODUCER_THREADS 128
#define UG_THREADS 160
#define UG_MAX_TAPS 16
#define UG_LOOKAHEAD 2

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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
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

template <int BN> struct Cfg {
  static constexpr int kStages = 3;
  static constexpr int kMinBlocks = BN >= 128 ? 2 : 3;
  static constexpr int kABytes = UG_BM * UG_BK * 2;          // 16 KB
  static constexpr int kBBytes = BN * UG_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + (BN == 32 ? 4224 : 0) /*head weights*/;
};

}  // namespace ug

template <int BN>
__global__ void __launch_bounds__(UG_THREADS, ug::Cfg<BN>::kMinBlocks) unet_gemm_kernel(const UGLayer L) {
  ug::pdl_wait();                 // launched with programmatic stream serialisation like the rest of the chain
  ug::pdl_launch_dependents();
  using C = ug::Cfg<BN>;
  extern __shared__ unsigned char ug_smem_raw[];
  const uint32_t raw = ug::smem_u32(ug_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B atoms need 1024 B alignment
  unsigned char* tiles_ptr = ug_smem_raw + (tiles - raw);
  const uint32_t bars = tiles + C::kStages * C::kStageBytes;           // full[S], empty[S], accum, tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tiles_ptr + C::kStages * C::kStageBytes + 8 * (2 * C::kStages + 1));
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::kStages + s); };
  const uint32_t accum_bar = bars + 8u * (2 * C::kStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * UG_BM, n0 = blockIdx.y * BN, phase = blockIdx.z;
  const int nkb = L.Kpad / UG_BK;

  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) { ug::mbar_init(full_bar(s), UG_PRODUCER_THREADS); ug::mbar_init(empty_bar(s), 1); }
    ug::mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float* head_smem = reinterpret_cast<float*>(tiles_ptr + C::kStages * C::kStageBytes + 256);
  if (BN == 32 && L.out_mode == UG_OUT_HEAD_NCHW_F32) {
    for (int i = tid; i < 32 * 32 + 32; i += UG_THREADS) head_smem[i] = i < 1024 ? L.head_w[i] : L.head_b[i - 1024];
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ug::smem_u32(tmem_slot)), "r"((uint32_t)C::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ug::tc_fence_before();
  __syncthreads();
  ug::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------ producers: gather A row `tid`, stream B rows ------------------------------
    const int m = m0 + tid;
    const bool row_ok = m < L.M;
    int b = 0, qy = 0, qx = 0;
    if (row_ok) { qx = m % L.Wq; const int t = m / L.Wq; qy = t % L.Hq; b = t / L.Hq; }
    const int by = qy * L.in_stride, bx = qx * L.in_stride;
    const int ip = L.in_pad, Wpi = L.Win + 2 * ip;
    const __nv_bfloat16* in_b = L.in + ((size_t)b * (L.Hin + 2 * ip) * Wpi + (size_t)ip * Wpi + ip) * L.in_ld;
    const __nv_bfloat16* w_tile = L.w + ((size_t)phase * L.Cout + n0) * L.Kpad;
    const uint32_t a_row = (uint32_t)tid * 128u;
    const uint32_t sw = (uint32_t)(tid & 7);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % C::kStages;
      if (kb >= C::kStages) ug::mbar_wait(empty_bar(s), (uint32_t)((kb / C::kStages - 1) & 1));
      const uint32_t a_st = tiles + s * C::kStageBytes, b_st = a_st + C::kABytes;
      // A: 8 chunks of 8 channels.  With Cin >= 64 a k-block lies inside one tap: one address / bounds computation.
      if (L.Cin >= UG_BK) {
        const int kk = kb * UG_BK;
        const int tap = kk >> L.cin_shift, c = kk & (L.Cin - 1);
        const int iy = by + L.dy[phase][tap & (UG_MAX_TAPS - 1)], ix = bx + L.dx[phase][tap & (UG_MAX_TAPS - 1)];
        const bool ok = row_ok && kk < L.Kvalid && (ip || (iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win));
        const __nv_bfloat16* src = ok ? in_b + ((ptrdiff_t)iy * Wpi + ix) * L.in_ld + c : L.in;
        const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) ug::cp_async16(a_st + a_row + (((uint32_t)j ^ sw) << 4), ok ? src + j * 8 : src, nb);
      } else {   // Cin == 32: two taps per k-block
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int kk = kb * UG_BK + h * 32;
          const int tap = kk >> L.cin_shift;
          const int iy = by + L.dy[phase][tap & (UG_MAX_TAPS - 1)], ix = bx + L.dx[phase][tap & (UG_MAX_TAPS - 1)];
          const bool ok = row_ok && kk < L.Kvalid && (ip || (iy >= 0 && iy < L.Hin && ix >= 0 && ix < L.Win));
          const __nv_bfloat16* src = ok ? in_b + ((ptrdiff_t)iy * Wpi + ix) * L.in_ld : L.in;
          const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) ug::cp_async16(a_st + a_row + (((uint32_t)(h * 4 + j) ^ sw) << 4), ok ? src + j * 8 : src, nb);
        }
      }
      // B: BN rows x 8 chunks, 128 threads
#pragma unroll
      for (int i = 0; i < BN * 8 / UG_PRODUCER_THREADS; ++i) {
        const int q = i * UG_PRODUCER_THREADS + tid;
        const int r = q >> 3, j = q & 7;
        ug::cp_async16(b_st + (uint32_t)r * 128u + (((uint32_t)j ^ (uint32_t)(r & 7)) << 4),
                       w_tile + (size_t)r * L.Kpad + kb * UG_BK + j * 8, 16u);
      }
      ug::cp_async_commit();
      if (kb >= UG_LOOKAHEAD) {
        ug::cp_async_wait<UG_LOOKAHEAD>();
        ug::fence_proxy_async();
        ug::mbar_arrive(full_bar((kb - UG_LOOKAHEAD) % C::kStages));
      }
    }
    // drain
    ug::cp_async_wait<0>();
    ug::fence_proxy_async();
    for (int kb = (nkb > UG_LOOKAHEAD ? nkb - UG_LOOKAHEAD : 0); kb < nkb; ++kb) ug::mbar_arrive(full_bar(kb % C::kStages));

    // ------------------------------ epilogue: TMEM -> registers -> global ------------------------------
    ug::mbar_wait(accum_bar, 0);
    ug::tc_fence_after();
    float clsv = 0.f;
    const float* clsw = nullptr;
    if (L.cls != nullptr && row_ok) {
      clsv = L.cls[b];
      const int cy = qy == 0 ? 0 : (qy == L.Hq - 1 ? 2 : 1), cx = qx == 0 ? 0 : (qx == L.Wq - 1 ? 2 : 1);
      clsw = L.cls_w + (size_t)(cy * 3 + cx) * L.Cout;
    }
    const int oy = qy * L.out_stride + (phase >> 1), ox = qx * L.out_stride + (phase & 1);
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      uint32_t v[32];
      ug::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
      if (!row_ok) continue;
      float y[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = n0 + cb + i, nm = n % L.n_mod;
        float a = __uint_as_float(v[i]);
        if (clsw) a = fmaf(clsv, clsw[n], a);
        a = fmaf(a, L.scale ? L.scale[nm] : 1.f, L.bias ? L.bias[nm] : 0.f);
        y[i] = a > 0.f ? a : a * L.slope;
      }
      const int nn = n0 + cb;
      const size_t col = (size_t)L.pix_off[nn / L.n_split] + (nn % L.n_split);
      if (L.out_mode == UG_OUT_NHWC_BF16) {
        const int op = L.out_pad;
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(L.out) + (((size_t)b * (L.Hout + 2 * op) + oy + op) * (L.Wout + 2 * op) + ox + op) * L.out_ld + col;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(y[i], y[i + 1]), t1 = __floats2bfloat162_rn(y[i + 2], y[i + 3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(y[i + 4], y[i + 5]), t3 = __floats2bfloat162_rn(y[i + 6], y[i + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(o + i) = pk;
        }
      } else if (BN == 32 && L.out_mode == UG_OUT_HEAD_NCHW_F32) {
        // 1x1 head on the activated pixel (operands rounded to bf16 like the unfused layer pair), fp32 NCHW store
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __bfloat162float(__float2bfloat16_rn(y[i]));
        float* o = reinterpret_cast<float*>(L.out) + ((size_t)b * L.n_valid * L.Hout + oy) * L.Wout + ox;
#pragma unroll 1
        for (int n8 = 0; n8 < L.n_valid; n8 += 8) {
          float z[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = head_smem[1024 + n8 + u];
#pragma unroll
          for (int c4 = 0; c4 < 32; c4 += 4) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float4 w4 = *reinterpret_cast<const float4*>(head_smem + (n8 + u) * 32 + c4);
              z[u] = fmaf(w4.x, y[c4], fmaf(w4.y, y[c4 + 1], fmaf(w4.z, y[c4 + 2], fmaf(w4.w, y[c4 + 3], z[u]))));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) o[(size_t)(n8 + u) * L.Hout * L.Wout] = z[u];
        }
      } else {   // fp32 NCHW: out[((b*Cout + n)*Hout + oy)*Wout + ox]
        float* o = reinterpret_cast<float*>(L.out) + (((size_t)b * L.n_valid + nn) * L.Hout + oy) * L.Wout + ox;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (nn + i < L.n_valid) o[(size_t)i * L.Hout * L.Wout] = y[i];
      }
    }
    ug::tc_fence_before();
  } else if (lane == 0) {
    // ------------------------------ MMA issuer (one thread) ------------------------------
    const uint32_t idesc = ug::instr_desc_bf16(BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % C::kStages;
      ug::mbar_wait(full_bar(s), (uint32_t)((kb / C::kStages) & 1));
      ug::tc_fence_after();
      const uint32_t a_st = tiles + s * C::kStageBytes, b_st = a_st + C::kABytes;
      const uint64_t adesc = ug::smem_desc_sw128(a_st), bdesc = ug::smem_desc_sw128(b_st);
#pragma unroll
      for (int k = 0; k < UG_BK / 16; ++k)      // advance 16 bf16 = 32 B inside the swizzle atom: +2 in the >>4 address field
        ug::umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
      ug::umma_commit(empty_bar(s));
    }
    ug::umma_commit(accum_bar);
  }
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    ug::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
  }
}

// "Slab" implicit-GEMM kernel for the high-resolution conv / transposed-conv layers of the passive-separator U-Net.
//
// Replaces nn.Conv2d(k4,s2,p1) (first encoder layer) and nn.ConvTranspose2d(k4,s2,p1) (last decoder layers) + BatchNorm2d
// (inference statistics) + LeakyReLU / ReLU (+ the 1x1 head) of audio_separation/rl/models/separator_cnn.py:5-24, 44-50,
// 128-135 for the layers whose im2col gather is bound by L2 -> shared-memory traffic (small Cout, many pixels).
//
// Activations live in HBM as *haloed* NHWC bf16: [B][H+2][W+2][C] with a zero border, i.e. one flat list of
// "positions" q = (b*(H+2) + y)*(W+2) + x.  With the zero border, a tap (dy, dx) of a stride-1 window is a plain
// shift of the position index by dy*(W+2) + dx, so ONE shared-memory copy of the positions [q0-(W+3), q0+128+(W+3))
// (the slab, fetched by a single TMA box per 64-channel chunk) serves every tap and every sub-pixel phase:
//   D[phase][m, n] = sum_{tap, c} slab[m + shift(phase, tap), c] * W[phase][n, tap, c]
// Each tap is a tcgen05.mma whose A descriptor is the slab's base address plus shift*128 B (K-major, 128B-swizzled rows;
// the swizzle is a function of the absolute shared-memory address - measured: parity holds with the descriptor's
// base-offset field left 0 - so row-shifted start addresses stay consistent with what TMA wrote).  Phases that share a
// shift are multiplied in one MMA (N = phases x Cout).  The stride-2 encoder layer uses the space-to-depth form of its
// input ([B][H/2+2][W/2+2][4*C], channel = (y parity, x parity, c)), where every kernel tap is again a unit shift plus a
// channel sub-block.  Rows of the 128-position tile that fall on the border are computed and discarded (79 % useful at 16x16).
//
// Persistent: one CTA per SM walks the tiles.  Weight tiles are either resident in shared memory (loaded once per CTA)
// or stream through a TMA ring; slab chunks stream through their own ring, prefetched across tiles; the accumulators
// are double-buffered in TMEM so that the epilogue of tile i (two groups of 4 warps alternate) overlaps the MMAs of tile i+1.
// Warp roles: 0-7 epilogue (BatchNorm scale/bias, activation, class-plane term, 1x1 head, stores), 8 MMA issuer,
// 9 weight loader, 10 slab loader.
#pragma once
#include <cuda.h>
#include "unet_gemm.cuh"

#define US_THREADS 384
#define US_MAX_ITEMS 40
#define US_B_STAGE_BYTES 16384        // up to 128 weight rows x 64 k (bf16)
#define US_MAX_STAGES 8
#define US_AUX_BYTES 4096             // scale, bias (2 x 128 fp32) | class-plane sums (9 x 64 fp32) | head bias (32 fp32)
#define US_HEAD_BYTES (4096 + 2 * 16384)   // fused 1x1 head: weight tile [32 n][64 k] + one activation tile [128 pos][64 k] per epilogue group

struct USPart { int16_t b_row, n, tmem_col, init; };   // init: first MMA writing these accumulator columns
struct USItem {                       // one weight tile and the MMAs that consume it
  int16_t a_shift;                    // slab row of output row 0 for this tap: (Wp+1) + dy*Wp + dx
  int8_t chunk;                       // 64-channel chunk of the slab (items are ordered by chunk)
  int8_t a_k16, n_k16;                // first 16-channel step inside the chunk, number of steps
  int8_t n_parts;
  int16_t b_rows;                     // weight rows of this tile (multiple of box rows)
  int16_t w_row;                      // first row of the tile in the packed weight tensor
  USPart part[2];
};

// Issue record of an item: everything the MMA-issuing warp needs, pre-combined on the host into twelve 32-bit words so that
// one record is three 16-byte constant-bank loads into uniform registers (fetched one item ahead of its use) and the issue
// path holds no shifts, masks or ring arithmetic (ncu r2: the issuing warps, not the tensor pipe, bound the slab layers -
// ~100 uniform-datapath instructions and ~830 cycles per item).
struct alignas(16) USIssue {
  uint32_t a_off16;                   // ((a_shift*128 + a_k16*32) >> 4) + low descriptor word constant; + slab stage address >> 4 = A descriptor low word
  uint32_t b_off16;                   // resident weights: (w_row*128) >> 4; streamed: ring stage (item % n_b_stages) * stage bytes >> 4
  uint32_t ctl;                       // n_k16 | n_parts << 8 | first_of_chunk << 16 | last_of_chunk << 17 | ring parity of the item << 18
  uint32_t b_bar;                     // streamed: 8 * ring stage (byte offset of the stage's full / empty barriers)
  uint32_t idesc0, col0, brow0, acc0; // part 0: instruction descriptor, accumulator column, (b_row*128) >> 4 + low descriptor word constant, accumulate-from-start
  uint32_t idesc1, col1, brow1, acc1; // part 1
};
struct USLayer {
  int n_items, n_chunks, n_tiles;
  int slab_rows;                      // 128 + 2*(Wp+1)
  int slab_bytes;                     // slab_rows*128 rounded up to 1024
  int n_slab_stages, n_b_stages;      // ring depths; n_b_stages == 0: weights resident (w_rows*128 B)
  int w_rows;
  int Hp, Wp, PI, B;                  // padded geometry of the input grid, positions per image, batch
  int in_coff;                        // first input channel inside the position's channel vector
  int w_box_rows;                     // rows per weight TMA box
  int tile_cols, tmem_cols;           // accumulator columns per tile (n_phase*Cout), allocation (power of two >= n_buf*tile_cols)
  int n_buf;                          // accumulator buffers in TMEM: 4 when they fit (two per issuing warp), else 2
  int n_phase, Cout;
  // epilogue
  const float* scale; const float* bias; float slope;
  const float* cls; const float* cls_w;           // class-plane term (encoder layer 0): cls[b] = target class + 1 as fp32 ...
  const long long* cls_i64;                         // ... or the raw int64 class indices (the + 1 and the conversion happen here)
  void* out; int out_mode;                          // UG_OUT_NHWC_BF16 | UG_OUT_HEAD_NCHW_F32
  int out_stride, Hop, Wop, out_ld, out_coff;       // output pixel (y-1)*out_stride + py + 1 in a haloed [Hop][Wop] grid
  int n_valid; const __nv_bfloat16* head_w; const float* head_b;   // head weights: bf16 [32 n][64 k], zero padded (pack_head_kernel)
  int head_bytes;                     // US_HEAD_BYTES when the 1x1 head is fused (out_mode == UG_OUT_HEAD_NCHW_F32), else 0
  USItem items[US_MAX_ITEMS];
  USIssue rec[US_MAX_ITEMS];
};

namespace us {
using namespace ug;

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t instr_desc_bf16_n(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// Whole-warp (convergent) variants: every lane holds the same operands, one elected lane issues.  Keeping the issuing
// warp convergent lets the compiler keep descriptors in uniform registers instead of per-thread registers + R2UR loops.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar) : "memory");
}
__host__ __device__ constexpr int smem_bytes(int slab_bytes, int n_slab_stages, int n_b_stages, int w_rows, int head_bytes) {
  return n_slab_stages * slab_bytes + (n_b_stages ? n_b_stages * US_B_STAGE_BYTES : w_rows * 128) + head_bytes + 1024 /*align*/ + 512 /*barriers*/ + US_AUX_BYTES;
}
// byte offset of 16-byte chunk c of row r inside a K-major 128B-swizzled tile (8-row atoms of 1024 B)
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace us

__global__ void __launch_bounds__(US_THREADS, 1)
unet_slab_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ USLayer L) {
  extern __shared__ unsigned char us_smem_raw[];
  const uint32_t raw = ug::smem_u32(us_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = us_smem_raw + (base - raw);
  const bool resident = L.n_b_stages == 0;
  const uint32_t w_bytes = resident ? (uint32_t)L.w_rows * 128u : (uint32_t)L.n_b_stages * US_B_STAGE_BYTES;
  const uint32_t b0 = base;                                   // weights (resident block or ring)
  const uint32_t slab0 = base + w_bytes;
  const uint32_t head0 = slab0 + (uint32_t)L.n_slab_stages * L.slab_bytes;      // head weight tile, then the two activation tiles
  unsigned char* head_ptr = base_ptr + w_bytes + (size_t)L.n_slab_stages * L.slab_bytes;
  const uint32_t bars = head0 + (uint32_t)L.head_bytes;
  unsigned char* bars_ptr = head_ptr + L.head_bytes;
  auto s_full = [&](int s) { return bars + 8u * s; };
  auto s_empty = [&](int s) { return bars + 8u * (US_MAX_STAGES + s); };
  auto b_full = [&](int s) { return bars + 8u * (2 * US_MAX_STAGES + s); };      // slot 0 doubles as "resident weights landed"
  auto b_empty = [&](int s) { return bars + 8u * (3 * US_MAX_STAGES + s); };
  auto t_full = [&](int s) { return bars + 8u * (4 * US_MAX_STAGES + s); };       // 4 accumulator buffers
  auto t_empty = [&](int s) { return bars + 8u * (4 * US_MAX_STAGES + 4 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * (4 * US_MAX_STAGES + 8));
  auto h_done = [&](int s) { return bars + 8u * (4 * US_MAX_STAGES + 9 + s); };   // head MMAs of epilogue group s have completed
  float* sb_smem = reinterpret_cast<float*>(bars_ptr + 512);             // (scale, bias) pairs [Cout]: one 16-byte load = two channels
  float* cls_smem = sb_smem + 2 * 128;                                    // [9][Cout]
  float* hb_smem = cls_smem + 9 * 64;                                     // head bias [32]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role code stays on the uniform datapath

  if (tid == 0) {
    for (int s = 0; s < US_MAX_STAGES; ++s) {
      ug::mbar_init(s_full(s), 1); ug::mbar_init(s_empty(s), 1); ug::mbar_init(b_full(s), 1); ug::mbar_init(b_empty(s), 2);   // both issuers release a weight stage
    }
    for (int s = 0; s < 4; ++s) { ug::mbar_init(t_full(s), 1); ug::mbar_init(t_empty(s), 4); }
    for (int s = 0; s < 2; ++s) ug::mbar_init(h_done(s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (L.out_mode == UG_OUT_HEAD_NCHW_F32) {
    // head weights [32 n][64 k] bf16 -> K-major 128B-swizzled operand tile (written by threads: made visible to the tensor
    // core's async proxy before the barrier below)
    for (int i = tid; i < 32 * 8; i += US_THREADS)
      *reinterpret_cast<uint4*>(head_ptr + us::sw128_off(i >> 3, i & 7)) = reinterpret_cast<const uint4*>(L.head_w)[i];
    if (tid < 32) hb_smem[tid] = L.head_b[tid];
    us::fence_async_smem();
  }
  for (int i = tid; i < L.Cout; i += US_THREADS) { sb_smem[2 * i] = L.scale ? L.scale[i] : 1.f; sb_smem[2 * i + 1] = L.bias ? L.bias[i] : 0.f; }
  if (L.cls != nullptr || L.cls_i64 != nullptr)
    for (int i = tid; i < 9 * L.Cout; i += US_THREADS) cls_smem[i] = L.cls_w[i];
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ug::smem_u32(tmem_slot)), "r"((uint32_t)L.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ug::tc_fence_before();
  __syncthreads();
  ug::tc_fence_after();
  ug::pdl_wait();                      // everything above touched shared memory, TMEM and packed parameters only
  ug::pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int NS = L.n_slab_stages, NB = L.n_b_stages;

  // Tiles of this CTA: t = blockIdx.x + iter * gridDim.x, iter < n_my.  They are processed in PAIRS (iter = 2p, 2p+1): two
  // MMA-issuing warps (q = 0, 1) work on the two tiles of a pair at the same time - one warp cannot issue faster than
  // ~600 cycles per item on the uniform datapath - and consume the SAME weight stage for both tiles (streamed weights are
  // read once per pair from L2); slab chunks are fetched in the order (c, q).  Issuer q / epilogue group q own the
  // accumulator buffers q, q+2, ...
  const int n_my = ((int)L.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_pairs = (n_my + 1) >> 1;
  const int bufs_per_q = L.n_buf >> 1;

  if (warp == 10) {
    // ------------------------------ slab loader: one TMA box per (tile, 64-channel chunk) ------------------------------
    if (lane == 0) {
      int cnt = 0;
      for (int p = 0; p < n_pairs; ++p) {
        for (int c = 0; c < L.n_chunks; ++c) {
          for (int q = 0; q < 2; ++q) {
            const int iter = 2 * p + q;
            if (iter >= n_my) continue;
            const int t = blockIdx.x + iter * gridDim.x;
            const int st = cnt % NS;
            if (cnt >= NS) ug::mbar_wait(s_empty(st), (uint32_t)((cnt / NS - 1) & 1));
            us::mbar_expect_tx(s_full(st), (uint32_t)L.slab_rows * 128u);
            us::tma_load_2d(slab0 + st * L.slab_bytes, &tm_in, L.in_coff + c * 64, t * 128 - (L.Wp + 1), s_full(st));
            ++cnt;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------ weight loader ------------------------------
    if (lane == 0) {
      if (resident) {
        us::mbar_expect_tx(b_full(0), w_bytes);
        for (int r = 0; r < L.w_rows; r += L.w_box_rows) us::tma_load_2d(b0 + r * 128, &tm_w, 0, r, b_full(0));
      } else {
        int cnt = 0;
        for (int p = 0; p < n_pairs; ++p) {
          for (int i = 0; i < L.n_items; ++i, ++cnt) {
            const int st = cnt % NB;
            if (cnt >= NB) ug::mbar_wait(b_empty(st), (uint32_t)((cnt / NB - 1) & 1));
            const int rows = L.items[i].b_rows, row = L.items[i].w_row;
            us::mbar_expect_tx(b_full(st), (uint32_t)rows * 128u);
            for (int r = 0; r < rows; r += L.w_box_rows)
              us::tma_load_2d(b0 + st * US_B_STAGE_BYTES + r * 128, &tm_w, 0, row + r, b_full(st));
          }
        }
      }
    }
  } else if (warp == 8 || warp == 11) {
    // ------------------------------ MMA issuers (convergent warps, one elected lane issues) ------------------------------
    // Everything the issue path needs is a function of kernel parameters and warp-uniform counters: per-item issue records
    // are read from the constant bank with uniform indices (one item ahead) and the K16 steps of a part are unrolled, so
    // that an MMA costs two uniform adds + UTCHMMA (the first version walked a shared-memory list through per-thread
    // registers: ~45 instructions and several R2UR round trips per MMA, ~300 cycles - the tensor pipe idled).
    const int q = warp == 8 ? 0 : 1;
    if (resident) { ug::mbar_wait(b_full(0), 0); ug::tc_fence_after(); }
    const uint64_t desc_hi32 = ug::smem_desc_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint32_t slab0_16 = slab0 >> 4, slab_bytes16 = (uint32_t)L.slab_bytes >> 4, b0_16 = b0 >> 4;
    const uint32_t b_full0 = b_full(0), b_empty0 = b_empty(0);
    // weight ring: n_items is a multiple of the ring depth (host), so an item's stage is fixed and its parity is the pair's
    // base parity xor a per-item bit; slab ring: this issuer's boxes are NS-periodic with step 2 (1 in a ragged last pair)
    const uint32_t turns_odd = NB ? (uint32_t)(L.n_items / NB) & 1u : 0u;
    uint32_t pair_par = 0;
    int cs_next = q;                               // ring position of this issuer's next slab box (NS >= 2 > q)
    uint32_t s_par = 0;
    const uint4* recs = reinterpret_cast<const uint4*>(&L.rec[0]);
    for (int p = 0; p < n_pairs; ++p) {
      const bool active = 2 * p + q < n_my;        // the last pair may hold a single tile: issuer 1 then only recycles weight stages
      const bool full = 2 * p + 1 < n_my;
      const int buf = q + 2 * (p % bufs_per_q), use = p / bufs_per_q;
      if (active && use >= 1) { ug::mbar_wait(t_empty(buf), (uint32_t)((use - 1) & 1)); ug::tc_fence_after(); }
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * L.tile_cols);
      const int s_step = full ? 2 : 1;
      int cs = 0;
      uint32_t a_base16 = 0;
      auto do_item = [&](const uint4 r0, const uint4 r1, const uint4 r2) {
        const uint32_t ctl = r0.z;
        if (active && (ctl & (1u << 16))) {       // first item of a chunk: its slab must have landed
          cs = cs_next;
          ug::mbar_wait(s_full(cs), s_par);
          a_base16 = slab0_16 + (uint32_t)cs * slab_bytes16;
          cs_next += s_step;
          if (cs_next >= NS) { cs_next -= NS; s_par ^= 1u; }       // the wait above used the parity of the box it waited for
          ug::tc_fence_after();
        }
        if (!resident) {
          ug::mbar_wait(b_full0 + r0.w, pair_par ^ ((ctl >> 18) & 1u));
          ug::tc_fence_after();
        }
        if (active) {
          // descriptor = (hi word, constant) : (lo word = LBO bit | start address >> 4); shared-memory addresses are < 2^18, so the
          // constant bits never meet a carry and a K16 step is one 32-bit add of 2 to the lo word
          const uint32_t a_lo = a_base16 + r0.x;
          const uint32_t b_base16 = b0_16 + r0.y;
          const bool nk4 = (ctl & 0xffu) == 4u;
          {
            const uint32_t d = tmem_d + r1.y, b_lo = b_base16 + r1.z;
            us::umma_bf16_elect(d, desc_hi32 | (uint64_t)a_lo, desc_hi32 | (uint64_t)b_lo, r1.x, r1.w);
            us::umma_bf16_elect(d, desc_hi32 | (uint64_t)(a_lo + 2u), desc_hi32 | (uint64_t)(b_lo + 2u), r1.x, 1u);
            if (nk4) {
              us::umma_bf16_elect(d, desc_hi32 | (uint64_t)(a_lo + 4u), desc_hi32 | (uint64_t)(b_lo + 4u), r1.x, 1u);
              us::umma_bf16_elect(d, desc_hi32 | (uint64_t)(a_lo + 6u), desc_hi32 | (uint64_t)(b_lo + 6u), r1.x, 1u);
            }
          }
          if (ctl & (2u << 8)) {
            const uint32_t d = tmem_d + r2.y, b_lo = b_base16 + r2.z;
            us::umma_bf16_elect(d, desc_hi32 | (uint64_t)a_lo, desc_hi32 | (uint64_t)b_lo, r2.x, r2.w);
            us::umma_bf16_elect(d, desc_hi32 | (uint64_t)(a_lo + 2u), desc_hi32 | (uint64_t)(b_lo + 2u), r2.x, 1u);
            if (nk4) {
              us::umma_bf16_elect(d, desc_hi32 | (uint64_t)(a_lo + 4u), desc_hi32 | (uint64_t)(b_lo + 4u), r2.x, 1u);
              us::umma_bf16_elect(d, desc_hi32 | (uint64_t)(a_lo + 6u), desc_hi32 | (uint64_t)(b_lo + 6u), r2.x, 1u);
            }
          }
        }
        if (!resident) us::umma_commit_elect(b_empty0 + r0.w);      // arrives at once when this warp issued nothing
        if (active && (ctl & (1u << 17))) us::umma_commit_elect(s_empty(cs));
      };
      // two items per trip, records fetched one item ahead: their constant-bank latency hides behind the previous item's MMAs and
      // no register rotation is needed
      uint4 a0 = recs[0], a1 = recs[1], a2 = recs[2];
      const int last = L.n_items - 1;
      for (int item = 0; item < L.n_items; item += 2) {
        const int i1 = item + 1 < last ? item + 1 : last;       // clamped: an odd item count re-reads the last record and skips it
        const uint4 c0 = recs[3 * i1], c1 = recs[3 * i1 + 1], c2 = recs[3 * i1 + 2];
        do_item(a0, a1, a2);
        const int i2 = item + 2 < last ? item + 2 : last;
        a0 = recs[3 * i2]; a1 = recs[3 * i2 + 1]; a2 = recs[3 * i2 + 2];
        if (item + 1 < L.n_items) do_item(c0, c1, c2);
      }
      pair_par ^= turns_odd;
      if (active) us::umma_commit_elect(t_full(buf));
    }
  } else {
    // ------------------------------ epilogue: two groups of 4 warps alternate tiles / TMEM buffers ------------------------------
    const int grp = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
    const bool relu = L.slope == 0.f;   // warp-uniform: the decoder's ReLU needs no slope multiply
    for (int j = 0;; ++j) {             // the group's j-th tile: iter = 2j + grp, accumulator buffer grp + 2*(j % bufs_per_q)
      const int iter = 2 * j + grp;
      if (iter >= n_my) break;
      const int t = blockIdx.x + iter * gridDim.x;
      const int buf = grp + 2 * (j % bufs_per_q);
      const int q = t * 128 + row;
      const int b = q / L.PI, r = q - b * L.PI;
      const int y = r / L.Wp, x = r - y * L.Wp;
      const bool ok = b < L.B && y >= 1 && y <= L.Hp - 2 && x >= 1 && x <= L.Wp - 2;
      ug::mbar_wait(t_full(buf), (uint32_t)((j / bufs_per_q) & 1));
      ug::tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * L.tile_cols);
      if (L.out_mode == UG_OUT_NHWC_BF16) {
        float clsv = 0.f;
        const float* clsw = nullptr;
        if ((L.cls != nullptr || L.cls_i64 != nullptr) && ok) {
          clsv = L.cls != nullptr ? L.cls[b] : (float)(L.cls_i64[b] + 1);
          const int cy = y == 1 ? 0 : (y == L.Hp - 2 ? 2 : 1), cx = x == 1 ? 0 : (x == L.Wp - 2 ? 2 : 1);
          clsw = cls_smem + (cy * 3 + cx) * L.Cout;
        }
        for (int ph = 0; ph < L.n_phase; ++ph) {
          const int oy = (y - 1) * L.out_stride + (ph >> 1) + 1, ox = (x - 1) * L.out_stride + (ph & 1) + 1;
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(L.out) + (((size_t)b * L.Hop + oy) * L.Wop + ox) * L.out_ld + L.out_coff;
#pragma unroll 1
          for (int cb = 0; cb < L.Cout; cb += 32) {
            uint32_t v[32];
            ug::tmem_ld32(trow + (uint32_t)(ph * L.Cout + cb), v);
            if (!ok) continue;
            float yv[32];
            // (Leaky)ReLU with slope in [0, 1]: max(a, slope * a)
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 s01 = *reinterpret_cast<const float4*>(sb_smem + 2 * (cb + i));
              const float4 s23 = *reinterpret_cast<const float4*>(sb_smem + 2 * (cb + i) + 4);
              float a0 = __uint_as_float(v[i]), a1 = __uint_as_float(v[i + 1]), a2 = __uint_as_float(v[i + 2]), a3 = __uint_as_float(v[i + 3]);
              if (clsw) {
                const float4 c4 = *reinterpret_cast<const float4*>(clsw + cb + i);
                a0 = fmaf(clsv, c4.x, a0); a1 = fmaf(clsv, c4.y, a1); a2 = fmaf(clsv, c4.z, a2); a3 = fmaf(clsv, c4.w, a3);
              }
              a0 = fmaf(a0, s01.x, s01.y); a1 = fmaf(a1, s01.z, s01.w); a2 = fmaf(a2, s23.x, s23.y); a3 = fmaf(a3, s23.z, s23.w);
              if (relu) { yv[i] = fmaxf(a0, 0.f); yv[i + 1] = fmaxf(a1, 0.f); yv[i + 2] = fmaxf(a2, 0.f); yv[i + 3] = fmaxf(a3, 0.f); }
              else {
                yv[i] = fmaxf(a0, a0 * L.slope); yv[i + 1] = fmaxf(a1, a1 * L.slope);
                yv[i + 2] = fmaxf(a2, a2 * L.slope); yv[i + 3] = fmaxf(a3, a3 * L.slope);
              }
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(yv[i], yv[i + 1]), t1 = __floats2bfloat162_rn(yv[i + 2], yv[i + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(yv[i + 4], yv[i + 5]), t3 = __floats2bfloat162_rn(yv[i + 6], yv[i + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(o + cb + i) = pk;
            }
          }
        }
        ug::tc_fence_before();
        __syncwarp();
        if (lane == 0) ug::mbar_arrive(t_empty(buf));
      } else {
        // last up-conv (Cout = 32) + BN + ReLU, then the 1x1 head z = head_b + head_w y (separator_cnn.py:135) as a second
        // tcgen05.mma: the activated pixels (bf16, like the unfused layer pair) of two sub-pixel phases form one K-major
        // [128 positions][2 x 32 channels] operand tile in shared memory; four MMAs (2 phases x 2 K16 steps, N = 32) write
        // z over the accumulator columns the group has just read.  Two rounds (phases 0-1, 2-3) share the tile.
        // (On CUDA cores the head cost 4096 FMAs per position: the epilogue, not the tensor pipe, bounded this layer.)
        const int Hm = (L.Hp - 2) * 2, Wm = (L.Wp - 2) * 2;
        const uint32_t ytile = head0 + 4096u + (uint32_t)grp * 16384u;
        unsigned char* yrow = head_ptr + 4096 + grp * 16384;
        const uint32_t tcol = tmem_base + (uint32_t)(buf * L.tile_cols);
        const uint64_t desc_hi = ug::smem_desc_sw128(0);
        const uint32_t idesc = us::instr_desc_bf16_n(32);
        uint32_t hpar = 0;                                      // every tile of a group consumes exactly two phases of its barrier
        uint32_t pk[2][16];
#pragma unroll
        for (int rd = 0; rd < 2; ++rd) {
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            uint32_t v[32];
            ug::tmem_ld32(trow + (uint32_t)((rd * 2 + pl) * 32), v);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float4 s01 = *reinterpret_cast<const float4*>(sb_smem + 2 * i);
              float a0 = fmaf(__uint_as_float(v[i]), s01.x, s01.y);
              float a1 = fmaf(__uint_as_float(v[i + 1]), s01.z, s01.w);
              if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }       // ReLU: no slope multiply (every decoder layer)
              else { a0 = fmaxf(a0, a0 * L.slope); a1 = fmaxf(a1, a1 * L.slope); }
              __nv_bfloat162 t2 = __floats2bfloat162_rn(a0, a1);
              pk[pl][i >> 1] = *reinterpret_cast<uint32_t*>(&t2);
            }
          }
          if (rd == 1) {      // the tile still feeds the first round's MMAs
            ug::mbar_wait(h_done(grp), hpar);
            hpar ^= 1u;
          }
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(yrow + us::sw128_off(row, pl * 4 + c)) = make_uint4(pk[pl][4 * c], pk[pl][4 * c + 1], pk[pl][4 * c + 2], pk[pl][4 * c + 3]);
          us::fence_async_smem();
          ug::tc_fence_before();
          us::group_sync(1 + grp);
          if (wq == 0) {
            ug::tc_fence_after();
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
              for (int k = 0; k < 2; ++k)
                us::umma_bf16_elect(tcol + (uint32_t)((rd * 2 + pl) * 32), desc_hi | (uint64_t)(((ytile + pl * 64 + k * 32) >> 4) & 0x3FFFu),
                                    desc_hi | (uint64_t)(((head0 + k * 32) >> 4) & 0x3FFFu), idesc, (uint32_t)k);
            us::umma_commit_elect(h_done(grp));
          }
        }
        ug::mbar_wait(h_done(grp), hpar);
        ug::tc_fence_after();
        float* o = reinterpret_cast<float*>(L.out) + ((size_t)b * L.n_valid * Hm + (size_t)((y - 1) * 2)) * Wm + (x - 1) * 2;
#pragma unroll
        for (int rd = 0; rd < 2; ++rd) {
          uint32_t z0[32], z1[32];
          ug::tmem_ld32(trow + (uint32_t)(rd * 64), z0);
          ug::tmem_ld32(trow + (uint32_t)(rd * 64 + 32), z1);
          if (ok) {
            float* on = o + rd * Wm;
            const size_t plane = (size_t)Hm * Wm;
#pragma unroll
            for (int n4 = 0; n4 < 32; n4 += 4) {
              if (n4 < L.n_valid) {              // n_valid is 16 or 32
                const float4 hb = *reinterpret_cast<const float4*>(hb_smem + n4);
                *reinterpret_cast<float2*>(on) = make_float2(__uint_as_float(z0[n4]) + hb.x, __uint_as_float(z1[n4]) + hb.x);
                *reinterpret_cast<float2*>(on + plane) = make_float2(__uint_as_float(z0[n4 + 1]) + hb.y, __uint_as_float(z1[n4 + 1]) + hb.y);
                *reinterpret_cast<float2*>(on + 2 * plane) = make_float2(__uint_as_float(z0[n4 + 2]) + hb.z, __uint_as_float(z1[n4 + 2]) + hb.z);
                *reinterpret_cast<float2*>(on + 3 * plane) = make_float2(__uint_as_float(z0[n4 + 3]) + hb.w, __uint_as_float(z1[n4 + 3]) + hb.w);
                on += 4 * plane;
              }
            }
          }
        }
        ug::tc_fence_before();
        __syncwarp();
        if (lane == 0) ug::mbar_arrive(t_empty(buf));
      }
    }
  }
  ug::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    ug::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L.tmem_cols) : "memory");
  }
}

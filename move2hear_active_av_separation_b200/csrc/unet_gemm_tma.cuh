// TMA-fed implicit-GEMM conv / transposed-conv layer (Cin >= 64) of the passive-separator U-Net on tcgen05.
//
// Same arithmetic and epilogue as unet_gemm.cuh (nn.Conv2d / nn.ConvTranspose2d k4 s2 p1 + BatchNorm + activation,
// audio_separation/rl/models/separator_cnn.py:5-24), but the A operand is no longer gathered with cp.async by four warps:
// with haloed NHWC feature maps every (tap, 64-channel chunk) of a 128-row tile is ONE dense TMA box
//   transposed conv (stride-1 taps): box [nb][rows][W][64 ch] of the 4-D map at pixel offset (1 + dy, 1 + dx)
//   stride-2 conv: box of the 5-D parity view [B][Hp/2][2][Wp/2][2*ld] (pixel (2y + kh, 2x + kw) = cell (y + kh/2, x + kw/2),
//   parity (kh & 1, kw & 1))
// which lands as the 128B-swizzled K-major tile tcgen05.mma reads.  One thread feeds a deep ring of (A box, weight box)
// stages; the old kernel's 1024 cp.async per k-block and its two-deep look-ahead were the bottleneck of the deep layers
// (1.2 us per 64-wide k-block at any batch size).
#pragma once
#include "unet_slab.cuh"

#define UT_THREADS 192
struct UTGeom {
  int rows_y, nb;          // M tile = [nb images][rows_y rows][Wq columns] = 128 positions
  int tiles_per_grp;       // Hq / rows_y
  int conv;                // 1: stride-2 conv (5-D parity view), 0: stride-1 taps (4-D map)
  int B;
  int k32;                 // 1: Cin = 32 - a k-block is one tap x 32 channels, boxes of 64-byte rows in the 64B swizzle (the data
                           // gradient of the last decoder layer; was the one call site of the cp.async gather kernel)
};

// DEEP = false: 3-4 stages, two CTAs per SM (grids of several waves).  DEEP = true: a grid that leaves at most one CTA per
// SM (the 1x1 / 2x2 / 4x4 layers: 32-128 CTAs at B = 1024) keeps 6-7 stages in flight instead - the ring depth, not the
// tensor pipe, sets the pace of a lone CTA (TMA round trip ~1 us per stage).
template <int BN, bool DEEP = false> struct UTCfg {
  static constexpr int kStages = BN >= 256 ? 4 : (DEEP ? (BN >= 128 ? 6 : 7) : (BN >= 128 ? 3 : 4));    // 256-wide tiles: one CTA per SM, 48 KB per stage
  static constexpr int kABytes = 16384, kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN, bool DEEP = false>
__global__ void __launch_bounds__(UT_THREADS, DEEP ? 1 : 2)
unet_gemm_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ UGLayer L,
                     const __grid_constant__ UTGeom G) {
  using C = UTCfg<BN, DEEP>;
  extern __shared__ unsigned char ut_smem_raw[];
  const uint32_t raw = ug::smem_u32(ut_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* tiles_ptr = ut_smem_raw + (tiles - raw);
  const uint32_t bars = tiles + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::kStages + s); };
  const uint32_t accum_bar = bars + 8u * (2 * C::kStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tiles_ptr + C::kStages * C::kStageBytes + 8 * (2 * C::kStages + 1));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int n0 = blockIdx.y * BN, phase = blockIdx.z;
  const int grp = blockIdx.x / G.tiles_per_grp, b0 = grp * G.nb, y0 = (blockIdx.x % G.tiles_per_grp) * G.rows_y;
  const int kpt = G.k32 ? 1 : L.Cin >> 6;         // 64-channel chunks per tap (k32: one 32-channel chunk)
  const int nkb = L.n_taps * kpt;

  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) { ug::mbar_init(full_bar(s), 1); ug::mbar_init(empty_bar(s), 1); }
    ug::mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ug::smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ug::tc_fence_before();
  __syncthreads();
  ug::tc_fence_after();
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::kStages;
        if (kb >= C::kStages) ug::mbar_wait(empty_bar(s), (uint32_t)((kb / C::kStages - 1) & 1));
        const int tap = kb / kpt, c0 = (kb - tap * kpt) * 64;
        const int dy = L.dy[phase][tap], dx = L.dx[phase][tap];
        const uint32_t st = tiles + s * C::kStageBytes;
        us::mbar_expect_tx(full_bar(s), G.k32 ? (uint32_t)(C::kStageBytes / 2) : (uint32_t)C::kStageBytes);
        if (G.conv) {     // dy, dx hold kh - 1, kw - 1: haloed pixel (2y + kh, 2x + kw)
          const int kh = dy + 1, kw = dx + 1;
          us::tma_load_5d(st, &tm_a, (kw & 1) * L.in_ld + c0, (kw >> 1), (kh & 1), y0 + (kh >> 1), b0, full_bar(s));
        } else {
          us::tma_load_4d(st, &tm_a, c0, L.in_pad + dx, L.in_pad + y0 + dy, b0, full_bar(s));
        }
        us::tma_load_2d(st + C::kABytes, &tm_w, kb * (G.k32 ? 32 : 64), phase * L.Cout + n0, full_bar(s));
      }
    }
  } else if (warp == 4) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = ug::instr_desc_bf16(BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % C::kStages;
      ug::mbar_wait(full_bar(s), (uint32_t)((kb / C::kStages) & 1));
      ug::tc_fence_after();
      const uint32_t st = tiles + s * C::kStageBytes;
      if (G.k32) {
        const uint64_t adesc = ug::smem_desc_sw64(st), bdesc = ug::smem_desc_sw64(st + C::kABytes);
#pragma unroll
        for (int k = 0; k < 2; ++k) us::umma_bf16_elect(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
      } else {
        const uint64_t adesc = ug::smem_desc_sw128(st), bdesc = ug::smem_desc_sw128(st + C::kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) us::umma_bf16_elect(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
      }
      us::umma_commit_elect(empty_bar(s));
    }
    us::umma_commit_elect(accum_bar);
  } else {
    // ------------------------------ epilogue: TMEM -> registers -> global (bf16 NHWC) ------------------------------
    const int r = warp * 32 + lane;
    const int per_img = G.rows_y * L.Wq;
    const int ib = r / per_img, rr = r - ib * per_img;
    const int b = b0 + ib, qy = y0 + rr / L.Wq, qx = rr % L.Wq;
    const bool row_ok = b < G.B;
    ug::mbar_wait(accum_bar, 0);
    ug::tc_fence_after();
    const int oy = qy * L.out_stride + (phase >> 1), ox = qx * L.out_stride + (phase & 1);
    const int op = L.out_pad;
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      uint32_t v[32];
      ug::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
      if (!row_ok) continue;
      float y[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int nm = (n0 + cb + i) % L.n_mod;
        const float a = fmaf(__uint_as_float(v[i]), L.scale ? L.scale[nm] : 1.f, L.bias ? L.bias[nm] : 0.f);
        y[i] = a > 0.f ? a : a * L.slope;
      }
      const int nn = n0 + cb;
      const size_t col = (size_t)L.pix_off[nn / L.n_split] + (nn % L.n_split);
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
    }
    ug::tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    ug::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// Persistent form of the same layer kernel with 256 x BN super-tiles: one CTA per SM walks the (m-tile pair, n-tile,
// phase) list; a stage holds the A boxes of TWO neighbouring m-tiles and one weight box, so the weight tile is fetched
// once per 2 x 128 output rows (the one-tile kernel moves 32 KB from L2 per 2.1 MFLOP and is bound by that, ~46 B/clk/SM:
// 24 KB here).  The TMA ring (4 stages of 48 KB) never drains between tiles, the two accumulators are double-buffered in
// TMEM (4 x BN columns) so that the epilogue of one pair overlaps the MMAs of the next, and the per-CTA set-up is paid once.
// Tile order: the n-tiles of one (phase, m-pair) are consecutive, i.e. neighbouring CTAs read the same A boxes from L2.
// ------------------------------------------------------------------------------------------------
template <int BN> struct UTPCfg {
  static constexpr int kStages = BN >= 128 ? 4 : 5;
  static constexpr int kABytes = 16384, kBBytes = BN * 128;
  static constexpr int kStageBytes = 2 * kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(UT_THREADS, 1)
unet_gemm_tma_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ UGLayer L,
                                const __grid_constant__ UTGeom G, const int n_mpairs, const int n_ntiles, const int n_tiles) {
  using C = UTPCfg<BN>;
  extern __shared__ unsigned char ut_smem_raw[];
  const uint32_t raw = ug::smem_u32(ut_smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* tiles_ptr = ut_smem_raw + (tiles - raw);
  const uint32_t bars = tiles + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::kStages + s); };
  auto t_full = [&](int s) { return bars + 8u * (2 * C::kStages + s); };
  auto t_empty = [&](int s) { return bars + 8u * (2 * C::kStages + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tiles_ptr + C::kStages * C::kStageBytes + 8 * (2 * C::kStages + 4));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int kpt = L.Cin >> 6;                     // 64-channel chunks per tap
  const int nkb = L.n_taps * kpt;

  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) { ug::mbar_init(full_bar(s), 1); ug::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { ug::mbar_init(t_full(s), 1); ug::mbar_init(t_empty(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ug::smem_u32(tmem_slot)), "r"((uint32_t)(4 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ug::tc_fence_before();
  __syncthreads();
  ug::tc_fence_after();
  ug::pdl_wait();
  ug::pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;

  // super-tile -> (m-tile pair bp, n-tile by, phase bz), n-tiles fastest
  auto decode = [&](int t, int& bp, int& by, int& bz) { by = t % n_ntiles; const int q = t / n_ntiles; bp = q % n_mpairs; bz = q / n_mpairs; };

  if (warp == 5) {
    // ------------------------------ TMA producer: the ring runs across tile boundaries ------------------------------
    if (lane == 0) {
      int cnt = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int bp, by, phase;
        decode(t, bp, by, phase);
        const int n0 = by * BN;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int s = cnt % C::kStages;
          if (cnt >= C::kStages) ug::mbar_wait(empty_bar(s), (uint32_t)((cnt / C::kStages - 1) & 1));
          const int tap = kb / kpt, c0 = (kb - tap * kpt) * 64;
          const int dy = L.dy[phase][tap], dx = L.dx[phase][tap];
          const uint32_t st = tiles + s * C::kStageBytes;
          us::mbar_expect_tx(full_bar(s), (uint32_t)C::kStageBytes);
#pragma unroll
          for (int im = 0; im < 2; ++im) {
            const int bx = 2 * bp + im;
            const int grp = bx / G.tiles_per_grp, b0 = grp * G.nb, y0 = (bx % G.tiles_per_grp) * G.rows_y;
            if (G.conv) {     // dy, dx hold kh - 1, kw - 1: haloed pixel (2y + kh, 2x + kw)
              const int kh = dy + 1, kw = dx + 1;
              us::tma_load_5d(st + im * C::kABytes, &tm_a, (kw & 1) * L.in_ld + c0, (kw >> 1), (kh & 1), y0 + (kh >> 1), b0, full_bar(s));
            } else {
              us::tma_load_4d(st + im * C::kABytes, &tm_a, c0, L.in_pad + dx, L.in_pad + y0 + dy, b0, full_bar(s));
            }
          }
          us::tma_load_2d(st + 2 * C::kABytes, &tm_w, kb * 64, phase * L.Cout + n0, full_bar(s));
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = ug::instr_desc_bf16(BN);
    int cnt = 0, iter = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++iter) {
      const int buf = iter & 1;
      if (iter >= 2) { ug::mbar_wait(t_empty(buf), (uint32_t)(((iter >> 1) - 1) & 1)); ug::tc_fence_after(); }
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 2 * BN);
      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        const int s = cnt % C::kStages;
        ug::mbar_wait(full_bar(s), (uint32_t)((cnt / C::kStages) & 1));
        ug::tc_fence_after();
        const uint32_t st = tiles + s * C::kStageBytes;
        const uint64_t bdesc = ug::smem_desc_sw128(st + 2 * C::kABytes);
#pragma unroll
        for (int im = 0; im < 2; ++im) {
          const uint64_t adesc = ug::smem_desc_sw128(st + im * C::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            us::umma_bf16_elect(tmem_d + (uint32_t)(im * BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        us::umma_commit_elect(empty_bar(s));
      }
      us::umma_commit_elect(t_full(buf));
    }
  } else {
    // ------------------------------ epilogue: TMEM -> registers -> global (bf16 NHWC) ------------------------------
    const int r = warp * 32 + lane;
    const int per_img = G.rows_y * L.Wq;
    const int ib = r / per_img, rr = r - ib * per_img;
    const int ry = rr / L.Wq, qx = rr % L.Wq;
    const int op = L.out_pad;
    int iter = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++iter) {
      int bp, by, phase;
      decode(t, bp, by, phase);
      const int n0 = by * BN, buf = iter & 1;
      ug::mbar_wait(t_full(buf), (uint32_t)((iter >> 1) & 1));
      ug::tc_fence_after();
#pragma unroll 1
      for (int im = 0; im < 2; ++im) {
        const int bx = 2 * bp + im;
        const int grp = bx / G.tiles_per_grp, b0 = grp * G.nb, y0 = (bx % G.tiles_per_grp) * G.rows_y;
        const int b = b0 + ib, qy = y0 + ry;
        const bool row_ok = b < G.B;
        const int oy = qy * L.out_stride + (phase >> 1), ox = qx * L.out_stride + (phase & 1);
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
          uint32_t v[32];
          ug::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((buf * 2 + im) * BN + cb), v);
          if (!row_ok) continue;
          float y[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int nm = (n0 + cb + i) % L.n_mod;
            const float a = fmaf(__uint_as_float(v[i]), L.scale ? L.scale[nm] : 1.f, L.bias ? L.bias[nm] : 0.f);
            y[i] = fmaxf(a, a * L.slope);
          }
          const int nn = n0 + cb;
          const size_t col = (size_t)L.pix_off[nn / L.n_split] + (nn % L.n_split);
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
        }
      }
      ug::tc_fence_before();
      __syncwarp();
      if (lane == 0) ug::mbar_arrive(t_empty(buf));
    }
  }
  ug::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    ug::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(4 * BN)) : "memory");
  }
}

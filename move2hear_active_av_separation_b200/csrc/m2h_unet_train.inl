// Training entry points of the passive-separator U-Net (included at the end of m2h_unet.cu): train-mode forward
// (batch-statistics BatchNorm), backward (BatchNorm, data gradients through the forward GEMM kernels with re-packed
// weights, weight gradients on tcgen05), and the L1 mask loss.
#include "unet_train.cuh"

// ------------------------------------------------------------------------------------------------
// training workspace: slots 0-5 as the inference workspace, then raw conv outputs z, gradients, scratch
// ------------------------------------------------------------------------------------------------
enum {
  TW_X0S = 0, TW_CAT1, TW_CAT2, TW_CAT3, TW_CAT4, TW_BOTT, TW_UNUSED, TW_A4, TW_X0P,
  TW_ZCAT1, TW_ZCAT2, TW_ZCAT3, TW_ZCAT4, TW_ZBOTT, TW_Z4,
  TW_DCAT1, TW_DCAT2, TW_DCAT3, TW_DCAT4, TW_DBOTT, TW_DA4,
  TW_DENC1, TW_DENC2, TW_DENC3, TW_DENC4,
  TW_STATS, TW_AUX, TW_PACK_ENC, TW_PACK_ENC_DG, TW_PACK_DEC, TW_PACK_DEC_DG, TW_N
};
static const size_t kCatElems[5] = {0, 18 * 18 * 128, 10 * 10 * 256, 6 * 6 * 512, 4 * 4 * 1024};
static const size_t kOutElems = 34 * 34 * 32;
// dgrad weight blobs: encoder layers 1-3 deconv-packed [4][Cin_fwd][4*Cout_fwd], layer 4 [4*Cin_fwd][Cout_fwd]; decoder layers conv-packed
static size_t enc_dg_off(int i) { size_t o = 0; for (int j = 1; j < i; ++j) o = align256(o + (size_t)16 * kEncCin[j] * kEncCout[j] * 2); return o; }
static size_t dec_dg_off(int i) { size_t o = 0; for (int j = 0; j < i; ++j) o = align256(o + (size_t)(j == 0 ? 4 : 16) * kDecCin[j] * kDecCoutBase[j] * 2); return o; }
static size_t tw_slot_bytes(int slot, int B) {
  const size_t b = (size_t)B;
  switch (slot) {
    case TW_X0S: return b * 18 * 18 * 128 * 2;
    case TW_CAT1: case TW_ZCAT1: case TW_DCAT1: return b * kCatElems[1] * 2;
    case TW_CAT2: case TW_ZCAT2: case TW_DCAT2: return b * kCatElems[2] * 2;
    case TW_CAT3: case TW_ZCAT3: case TW_DCAT3: return b * kCatElems[3] * 2;
    case TW_CAT4: case TW_ZCAT4: case TW_DCAT4: return b * kCatElems[4] * 2;
    case TW_BOTT: case TW_ZBOTT: case TW_DBOTT: return b * 512 * 2;
    case TW_UNUSED: return 0;
    case TW_A4: case TW_X0P: case TW_Z4: case TW_DA4: return b * kOutElems * 2;
    case TW_DENC1: return b * 18 * 18 * 64 * 2;
    case TW_DENC2: return b * 10 * 10 * 128 * 2;
    case TW_DENC3: return b * 6 * 6 * 256 * 2;
    case TW_DENC4: return b * 4 * 4 * 512 * 2;
    case TW_STATS: return (size_t)10 * 4 * 512 * 8;      // per BN layer: [sum | sumsq | dgamma | dbeta] doubles
    case TW_AUX: return (size_t)10 * 4 * 512 * 4;        // per BN layer: [scale | shift | mean | invstd]
    case TW_PACK_ENC: return enc_layout().total;
    case TW_PACK_ENC_DG: return enc_dg_off(5);
    case TW_PACK_DEC: return dec_layout().total;
    case TW_PACK_DEC_DG: return dec_dg_off(5);
  }
  return 0;
}
extern "C" int m2h_unet_train_workspace_layout(int B, size_t offsets[32]) {
  if (B < 0 || !offsets) { m2h_set_error("m2h_unet_train_workspace_layout: bad arguments"); return M2H_ERR_ARG; }
  size_t o = 0;
  for (int i = 0; i < 32; ++i) {
    offsets[i] = o;
    if (i < TW_N) o = align256(o + tw_slot_bytes(i, B));
  }
  return M2H_OK;
}
extern "C" size_t m2h_unet_train_workspace_bytes(int B) {
  size_t off[32];
  m2h_unet_train_workspace_layout(B > 0 ? B : 0, off);
  return off[31];
}
extern "C" int m2h_unet_train_workspace_init(int B, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || !workspace) { m2h_set_error("m2h_unet_train_workspace_init: bad arguments"); return M2H_ERR_ARG; }
  if (workspace_bytes < m2h_unet_train_workspace_bytes(B)) { m2h_set_error("m2h_unet_train_workspace_init: workspace too small"); return M2H_ERR_WORKSPACE; }
  M2H_CUDA_CHECK(cudaMemsetAsync(workspace, 0, m2h_unet_train_workspace_bytes(B), (cudaStream_t)stream));   // zero borders once; kernels only write interiors
  return M2H_OK;
}

struct TW {
  char* base; size_t off[32]; int B;
  __nv_bfloat16* bf(int slot) const { return (__nv_bfloat16*)(base + off[slot]); }
  double* stats(int layer) const { return (double*)(base + off[TW_STATS]) + (size_t)layer * 4 * 512; }
  float* aux(int layer) const { return (float*)(base + off[TW_AUX]) + (size_t)layer * 4 * 512; }
};
static TW tw_make(void* ws, int B) { TW t; t.base = (char*)ws; t.B = B; m2h_unet_train_workspace_layout(B, t.off); return t; }
static FMap fmap(__nv_bfloat16* p, int ld, int coff, int Hp, int B) { FMap m; m.p = p; m.ld = ld; m.coff = coff; m.Hp = Hp; m.P = (long long)B * (Hp ? Hp * Hp : 1); return m; }

static int grid_for(long long work, int per_block) {
  long long g = (work + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}
// BatchNorm forward of one layer: stats over z, running-stat update, a = act(BN(z)).  layer = BN slot (0-4 encoder, 5-9 decoder)
// The fp64 reduction slots (forward sums | backward sums) of layers [first, first + count) are zeroed by ONE memset at the
// start of a forward function (they were one memset per layer and pass: 40 graph nodes per step)
static int bn_zero_slots(const TW& W, int first, int count, cudaStream_t st) {
  M2H_CUDA_CHECK(cudaMemsetAsync(W.stats(first), 0, (size_t)count * 4 * 512 * 8, st));
  return M2H_OK;
}
static int bn_forward(const TW& W, int layer, const FMap& z, const FMap& a, int C, int Creal, double n, const float* gamma, const float* beta,
                      float* run_mean, float* run_var, long long* nbt, float eps, float momentum, float slope, cudaStream_t st) {
  double* stats = W.stats(layer);
  if (z.P * (C / 8) >= (1LL << 31) || (C & (C - 1)) != 0) { m2h_set_error("BatchNorm: map of %lld positions x %d channels is outside the kernels' 32-bit index range", z.P, C); return M2H_ERR_ARG; }
  M2H_CUDA_CHECK(launch_pdl(bn_stats_kernel, dim3(grid_for(z.P, 4 * (256 / (C / 8)))), dim3(256), (size_t)(2 * 256 * 8 * sizeof(float)), st, z, C, stats));
  // (a fused finalize + apply kernel was measured and dropped: every thread of the wide apply grid re-derives eight channels'
  // scale / shift in fp64, which costs more than the 20 tiny finalize launches it saves: 3.33 vs 3.18 ms per step)
  M2H_CUDA_CHECK(launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), (size_t)(0), st, stats, n, gamma, beta, eps, momentum, run_mean, run_var, nbt, W.aux(layer), Creal, C));
  M2H_CUDA_CHECK(launch_pdl(bn_apply_kernel, dim3(grid_for(z.P * (C / 8), 256)), dim3(256), (size_t)(0), st, z, a, C, W.aux(layer), slope));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}
// BatchNorm backward of one layer: dz written over z; fp32 dgamma / dbeta out
static int bn_backward(const TW& W, int layer, const FMap& da1, const FMap& da2, const FMap& a, const FMap& z, int C, int Creal, double n,
                       float slope, float* dgamma, float* dbeta, cudaStream_t st) {
  double* red = W.stats(layer) + 2 * 512;          // zeroed by the forward pass of the same step (bn_zero_slots)
  static int gate_from_a = -1;                     // M2H_BN_GATE_FROM_A=1: read the stored activation for the gate (first version)
  if (gate_from_a < 0) { const char* e = getenv("M2H_BN_GATE_FROM_A"); gate_from_a = e && atoi(e) ? 1 : 0; }
  // reductions use [C] strides inside the 512-wide slots: kernels index red[c] and red[C + c]
  M2H_CUDA_CHECK(launch_pdl(bn_bwd_reduce_kernel, dim3(grid_for(z.P, 4 * (256 / (C / 8)))), dim3(256), (size_t)(2 * 256 * 8 * sizeof(float)), st, da1, da2, a, z, C, W.aux(layer), slope, red, gate_from_a));
  M2H_CUDA_CHECK(launch_pdl(bn_bwd_apply_kernel, dim3(grid_for(z.P * (C / 8), 256)), dim3(256), (size_t)(0), st, da1, da2, a, z, C, Creal, W.aux(layer), slope, red, n, dgamma, dbeta, gate_from_a));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

// ------------------------------------------------------------------------------------------------
// weight-gradient launch
// ------------------------------------------------------------------------------------------------
static bool g_wgrad_attr_set[M2H_MAX_DEVICES];
// dense: [B][Hq+2][Hq+2][d_ld] haloed (or plain [B][d_ld] when Hq == 1 && !d_haloed); sampled: [B][2Hq+2][2Hq+2][s_ld] haloed
static int launch_wgrad(const __nv_bfloat16* dense, int d_ld, int d_coff, int Cd, bool d_haloed, const __nv_bfloat16* sampled, int s_ld, int s_coff,
                        int Cs, int Cs_full, int Hq, int B, float* dw, cudaStream_t st) {
  int dev;
  int rc = m2h_current_device_checked(&dev);
  if (rc != M2H_OK) return rc;
  if (!g_wgrad_attr_set[dev]) {
    M2H_CUDA_CHECK(cudaFuncSetAttribute(unet_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes));
    g_wgrad_attr_set[dev] = true;
  }
  WGLayer L;
  memset(&L, 0, sizeof(L));
  L.Cd = Cd; L.Cs = Cs; L.Cs_full = Cs_full; L.d_coff = d_coff; L.s_coff = s_coff; L.d_pad = d_haloed ? 1 : 0; L.s_ld = s_ld; L.Hq = Hq;
  L.rows_y = Hq >= 8 ? 8 : Hq;
  L.nb = 128 / (Hq * L.rows_y);
  L.n_kchunks = ((B + L.nb - 1) / L.nb) * (Hq / L.rows_y);
  L.s_sw64 = Cs <= 32 ? 1 : 0;
  L.BN = Cs <= 32 ? 32 : 64;
  L.TC = 256 / L.BN;
  L.tap_mask = 0;
  for (int kh = 0; kh < 4; ++kh)
    for (int kw = 0; kw < 4; ++kw) {
      // sampled pixel 2y + kh - 1 must fall inside [0, 2Hq) for some y in [0, Hq): always for Hq > 1, only kh in {1, 2} for a 1x1 map
      const bool oky = Hq > 1 || kh == 1 || kh == 2, okx = Hq > 1 || kw == 1 || kw == 2;
      if (oky && okx) L.tap_mask |= 1 << (kh * 4 + kw);
    }
  L.dw = dw;
  const int m_tiles = (Cd + 127) / 128, n_tiles = (Cs + L.BN - 1) / L.BN;
  int live_taps = 0;                  // tap groups (CTAs along grid.x) with at least one live tap
  for (int t = 0; t < 16; t += L.TC) live_taps += ((L.tap_mask >> t) & ((1 << L.TC) - 1)) != 0;
  // split the positions so that ~2 waves of CTAs are in flight
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Position split: every slice adds a full copy of the gradient tile to L2 with atomics, and the launch runs next to the
  // data-gradient chain on another stream - measured at B = 256 (one B200): 2 waves of CTAs 3.09 ms per step, 1 wave 2.94,
  // half a wave 2.88, a third 2.85 (tap groups: 2.73), a ninth 3.14.  M2H_WGRAD_TARGET_CTAS overrides.
  static int target = -1, target32 = -1;
  if (target < 0) { const char* e = getenv("M2H_WGRAD_TARGET_CTAS"); target = e && atoi(e) > 0 ? atoi(e) : sms / 3; }
  if (target32 < 0) { const char* e = getenv("M2H_WGRAD_TARGET_CTAS_32"); target32 = e && atoi(e) > 0 ? atoi(e) : sms; }
  // the 32-channel layers (first encoder / last decoder layer: 65536 positions at B = 256, a 128 x 32 x 16 gradient) are all
  // main loop and almost no epilogue: they take a full wave
  int ksplit = ((Cs <= 32 ? target32 : target) + live_taps * m_tiles * n_tiles - 1) / (live_taps * m_tiles * n_tiles);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > L.n_kchunks) ksplit = L.n_kchunks;
  L.kchunks_per_cta = (L.n_kchunks + ksplit - 1) / ksplit;
  ksplit = (L.n_kchunks + L.kchunks_per_cta - 1) / L.kchunks_per_cta;

  alignas(64) CUtensorMap tm_d, tm_s;
  {
    const int hp = d_haloed ? Hq + 2 : 1;
    cuuint64_t dims[4] = {(cuuint64_t)d_ld, (cuuint64_t)hp, (cuuint64_t)hp, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)d_ld * 2, (cuuint64_t)hp * d_ld * 2, (cuuint64_t)hp * hp * d_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Hq, (cuuint32_t)L.rows_y, (cuuint32_t)L.nb};
    rc = make_tmap_nd(&tm_d, dense, 4, dims, strides, box, false);
    if (rc != M2H_OK) return rc;
  }
  {
    const int hp = 2 * Hq + 2;
    cuuint64_t dims[5] = {(cuuint64_t)2 * s_ld, (cuuint64_t)hp / 2, 2, (cuuint64_t)hp / 2, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)2 * s_ld * 2, (cuuint64_t)hp * s_ld * 2, (cuuint64_t)2 * hp * s_ld * 2, (cuuint64_t)hp * hp * s_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)(L.s_sw64 ? 32 : 64), (cuuint32_t)Hq, 1, (cuuint32_t)L.rows_y, (cuuint32_t)L.nb};
    rc = make_tmap_nd(&tm_s, sampled, 5, dims, strides, box, L.s_sw64 != 0);
    if (rc != M2H_OK) return rc;
  }
  dim3 grid(16 / L.TC, m_tiles * n_tiles, ksplit);
  M2H_CUDA_CHECK(launch_pdl(unet_wgrad_kernel, dim3(grid), dim3(WG_THREADS), (size_t)(wg::kSmemBytes), st, tm_d, tm_s, L));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

// ------------------------------------------------------------------------------------------------
// forward (train mode)
// ------------------------------------------------------------------------------------------------
static int check_train_args(const char* who, int B, const void* ws, size_t ws_bytes) {
  if (B <= 0) { m2h_set_error("%s: B must be positive", who); return M2H_ERR_ARG; }
  if (!ws) { m2h_set_error("%s: null workspace", who); return M2H_ERR_ARG; }
  if (ws_bytes < m2h_unet_train_workspace_bytes(B)) { m2h_set_error("%s: workspace too small", who); return M2H_ERR_WORKSPACE; }
  return unet_prepare_device();
}

// bf16 GEMM operands of the current weights (BatchNorm is NOT folded in train mode)
static int enc_pack_train(const m2h_unet_enc_params* p, int convert_bin2mono, char* pk, cudaStream_t st) {
  const EncLayout PL = enc_layout();
  for (int i = 0; i < 5; ++i) {
    const int cin_full = i == 0 ? (convert_bin2mono ? 32 : 33) : kEncCin[i];
    if (i == 0) {
      USLayer sl; SlabPackDesc pd;
      build_conv0_items(18, cin_full, sl, pd);
      M2H_CUDA_CHECK(launch_pdl(pack_slab_conv_w_kernel, dim3(96), dim3(256), (size_t)(0), st, p->conv_w[0], (__nv_bfloat16*)(pk + PL.w[0]), pd));
    } else {
      M2H_CUDA_CHECK(launch_pdl(pack_conv_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->conv_w[i], (__nv_bfloat16*)(pk + PL.w[i]), kEncCout[i], cin_full, kEncCin[i], kEncTaps[i]));
    }
  }
  if (!convert_bin2mono) M2H_CUDA_CHECK(launch_pdl(pack_cls_w_kernel, dim3((9 * 64 + 127) / 128), dim3(128), (size_t)(0), st, p->conv_w[0], (float*)(pk + PL.cls_w), 64, 33));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

extern "C" int m2h_unet_encoder_forward_train(const m2h_unet_enc_params* p, int64_t* const* num_batches_tracked, int convert_bin2mono,
                                              float bn_eps, float bn_momentum, const float* obs, const float* target_class_plus1,
                                              const float* masks, const int64_t* mask_strides, int B, void* workspace, size_t workspace_bytes,
                                              float* bottleneck_f32, void* stream) {
  const int weights_packed = convert_bin2mono & M2H_UNET_WEIGHTS_PACKED;
  convert_bin2mono &= 1;
  int rc = check_train_args("m2h_unet_encoder_forward_train", B, workspace, workspace_bytes);
  if (rc != M2H_OK) return rc;
  if (!p || !obs || (convert_bin2mono ? (!masks || !mask_strides) : !target_class_plus1)) { m2h_set_error("m2h_unet_encoder_forward_train: null pointer"); return M2H_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const TW W = tw_make(workspace, B);
  char* pk = W.base + W.off[TW_PACK_ENC];
  const EncLayout PL = enc_layout();
  rc = bn_zero_slots(W, 0, 5, st);
  if (rc != M2H_OK) return rc;
  if (!weights_packed) { rc = enc_pack_train(p, convert_bin2mono, pk, st); if (rc != M2H_OK) return rc; }
  PackInParams pp;
  pp.obs = obs; pp.masks = convert_bin2mono ? masks : nullptr; pp.x0 = W.bf(TW_X0S); pp.B = B; pp.x0_plain = W.bf(TW_X0P);
  for (int i = 0; i < 4; ++i) pp.ms[i] = convert_bin2mono ? (long long)mask_strides[i] : 0;
  M2H_CUDA_CHECK(launch_pdl(pack_input_kernel, dim3((B * 1024 + 255) / 256), dim3(256), (size_t)(0), st, pp));
  M2H_CUDA_CHECK(cudaGetLastError());

  const int cat_slot[5] = {0, TW_CAT1, TW_CAT2, TW_CAT3, TW_CAT4}, zcat_slot[5] = {0, TW_ZCAT1, TW_ZCAT2, TW_ZCAT3, TW_ZCAT4};
  const int cat_ld[5] = {0, 128, 256, 512, 1024}, hp[5] = {0, 18, 10, 6, 4};
  {  // layer 0 (slab kernel, identity epilogue): x0 -> z in zcat1[..., 64:128]
    USLayer sl; SlabPackDesc pd;
    memset(&sl, 0, sizeof(sl));
    build_conv0_items(18, 0, sl, pd);
    sl.slope = 1.f;
    if (!convert_bin2mono) { sl.cls = target_class_plus1; sl.cls_w = (const float*)(pk + PL.cls_w); }
    sl.out = W.bf(TW_ZCAT1); sl.out_mode = UG_OUT_NHWC_BF16; sl.out_stride = 1; sl.Hop = sl.Wop = 18; sl.out_ld = 128; sl.out_coff = 64;
    rc = launch_slab(sl, W.bf(TW_X0S), 128, 18, B, (const __nv_bfloat16*)(pk + PL.w[0]), st);
    if (rc != M2H_OK) return rc;
  }
  for (int i = 0; i < 5; ++i) {
    if (i > 0) {   // layer i: a in cat_i[..., C:] -> z in zcat_{i+1}[..., C':] (layer 4 -> zbott)
      UGLayer g;
      layer_defaults(g);
      const int hin = hp[i] - 2;
      g.in = W.bf(cat_slot[i]) + cat_ld[i] / 2; g.Hin = g.Win = hin; g.Cin = kEncCin[i]; g.cin_shift = ilog2(kEncCin[i]); g.in_ld = cat_ld[i]; g.in_pad = 1;
      g.Hq = g.Wq = hin / 2; g.M = B * g.Hq * g.Wq; g.in_stride = 2; g.n_taps = kEncTaps[i];
      for (int t = 0; t < kEncTaps[i]; ++t) {
        const int kh = kEncTaps[i] == 16 ? t >> 2 : 1 + (t >> 1), kw = kEncTaps[i] == 16 ? t & 3 : 1 + (t & 1);
        g.dy[0][t] = (int8_t)(kh - 1); g.dx[0][t] = (int8_t)(kw - 1);
      }
      g.w = (const __nv_bfloat16*)(pk + PL.w[i]); g.Cout = kEncCout[i]; g.Kvalid = g.Kpad = kEncTaps[i] * kEncCin[i];
      g.n_mod = kEncCout[i]; g.slope = 1.f;
      if (i < 4) { g.out = W.bf(zcat_slot[i + 1]) + cat_ld[i + 1] / 2; g.out_ld = cat_ld[i + 1]; g.out_pad = 1; }
      else { g.out = W.bf(TW_ZBOTT); g.out_ld = 512; g.out_pad = 0; }
      g.Hout = g.Wout = g.Hq; g.n_valid = g.Cout;
      rc = launch_layer_auto(g, 128, B, st);
      if (rc != M2H_OK) return rc;
    }
    const int C = kEncCout[i];
    const int hq = i < 4 ? hp[i + 1] - 2 : 1;
    const FMap z = i < 4 ? fmap(W.bf(zcat_slot[i + 1]), cat_ld[i + 1], cat_ld[i + 1] / 2, hp[i + 1], B) : fmap(W.bf(TW_ZBOTT), 512, 0, 0, B);
    const FMap a = i < 4 ? fmap(W.bf(cat_slot[i + 1]), cat_ld[i + 1], cat_ld[i + 1] / 2, hp[i + 1], B) : fmap(W.bf(TW_BOTT), 512, 0, 0, B);
    rc = bn_forward(W, i, z, a, C, C, (double)B * hq * hq, p->bn_w[i], p->bn_b[i], const_cast<float*>(p->bn_mean[i]), const_cast<float*>(p->bn_var[i]),
                    num_batches_tracked ? (long long*)num_batches_tracked[i] : nullptr, bn_eps, bn_momentum, 0.2f, st);
    if (rc != M2H_OK) return rc;
  }
  if (bottleneck_f32) {
    M2H_CUDA_CHECK(launch_pdl(bf16_to_f32_kernel, dim3((B * 512 + 255) / 256), dim3(256), (size_t)(0), st, W.bf(TW_BOTT), bottleneck_f32, (size_t)B * 512));
    M2H_CUDA_CHECK(cudaGetLastError());
  }
  return M2H_OK;
}

// geometry of the gather-kernel transposed-conv launch (4 sub-pixel phases), shared by forward layers and data gradients
static void deconv_phases(UGLayer& g, int hin, int B) {
  g.Hin = g.Win = hin; g.in_pad = 1; g.Hq = g.Wq = hin; g.M = B * hin * hin; g.n_taps = 4; g.n_phase = 4;
  for (int ph = 0; ph < 4; ++ph)
    for (int t = 0; t < 4; ++t) { g.dy[ph][t] = (int8_t)deconv_d(ph >> 1, t >> 1); g.dx[ph][t] = (int8_t)deconv_d(ph & 1, t & 1); }
  g.Hout = g.Wout = 2 * hin; g.out_stride = 2; g.out_pad = 1;
}

static int dec_pack_train(const m2h_unet_dec_params* p, int convert_bin2mono, char* pk, cudaStream_t st) {
  const DecLayout PL = dec_layout();
  const int c_last = convert_bin2mono ? 16 : 32;
  for (int i = 0; i < 5; ++i) {
    const int cout_real = i == 4 ? c_last : kDecCoutBase[i];
    if (i == 0) M2H_CUDA_CHECK(launch_pdl(pack_deconv0_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->deconv_w[0], (__nv_bfloat16*)(pk + PL.w[0]), 512, 512));
    else if (i >= 3) {
      USLayer sl; SlabPackDesc pd;
      build_deconv_items(kDecCin[i], kDecCoutBase[i], cout_real, i == 3 ? 10 : 18, sl, pd);
      M2H_CUDA_CHECK(launch_pdl(pack_slab_deconv_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->deconv_w[i], (__nv_bfloat16*)(pk + PL.w[i]), pd));
    } else M2H_CUDA_CHECK(launch_pdl(pack_deconv_w_kernel, dim3(256), dim3(256), (size_t)(0), st, p->deconv_w[i], (__nv_bfloat16*)(pk + PL.w[i]), kDecCin[i], cout_real, kDecCoutBase[i]));
  }
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

extern "C" int m2h_unet_decoder_forward_train(const m2h_unet_dec_params* p, int64_t* const* num_batches_tracked, int convert_bin2mono,
                                              float bn_eps, float bn_momentum, int B, void* workspace, size_t workspace_bytes,
                                              float* mask_out, void* stream) {
  const int weights_packed = convert_bin2mono & M2H_UNET_WEIGHTS_PACKED;
  convert_bin2mono &= 1;
  int rc = check_train_args("m2h_unet_decoder_forward_train", B, workspace, workspace_bytes);
  if (rc != M2H_OK) return rc;
  if (!p || !mask_out) { m2h_set_error("m2h_unet_decoder_forward_train: null pointer"); return M2H_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const TW W = tw_make(workspace, B);
  char* pk = W.base + W.off[TW_PACK_DEC];
  rc = bn_zero_slots(W, 5, 5, st);
  if (rc != M2H_OK) return rc;
  const DecLayout PL = dec_layout();
  const int c_last = convert_bin2mono ? 16 : 32;
  if (!weights_packed) { rc = dec_pack_train(p, convert_bin2mono, pk, st); if (rc != M2H_OK) return rc; }
  const int cat_slot[5] = {0, TW_CAT1, TW_CAT2, TW_CAT3, TW_CAT4}, zcat_slot[5] = {0, TW_ZCAT1, TW_ZCAT2, TW_ZCAT3, TW_ZCAT4};
  const int cat_ld[5] = {0, 128, 256, 512, 1024}, hp[5] = {0, 18, 10, 6, 4};
  // decoder layer i reads cat_{4-i+1}.. : in level = 5 - i (i >= 1), out level = 4 - i
  for (int i = 0; i < 5; ++i) {
    const int lvl_out = 4 - i;                      // z lives in zcat_{lvl_out}[..., :C]; layer 4 -> z4
    if (i == 0) {
      UGLayer g;
      layer_defaults(g);
      g.in = W.bf(TW_BOTT); g.Hin = g.Win = 1; g.Cin = 512; g.cin_shift = 9; g.in_ld = 512;
      g.Hq = g.Wq = 1; g.M = B; g.n_taps = 1;
      g.w = (const __nv_bfloat16*)(pk + PL.w[0]); g.Cout = 2048; g.Kvalid = g.Kpad = 512; g.n_mod = 512; g.slope = 1.f;
      g.out = W.bf(TW_ZCAT4); g.Hout = g.Wout = 1; g.out_ld = 16 * 1024; g.n_split = 512; g.n_valid = 2048;
      for (int pix = 0; pix < 4; ++pix) g.pix_off[pix] = (((pix >> 1) + 1) * 4 + (pix & 1) + 1) * 1024;
      rc = launch_layer_auto(g, 128, B, st);
    } else if (i < 3) {
      UGLayer g;
      layer_defaults(g);
      const int lvl_in = 5 - i;
      g.in = W.bf(cat_slot[lvl_in]); g.Cin = kDecCin[i]; g.cin_shift = ilog2(kDecCin[i]); g.in_ld = kDecCin[i];
      deconv_phases(g, hp[lvl_in] - 2, B);
      g.w = (const __nv_bfloat16*)(pk + PL.w[i]); g.Cout = kDecCoutBase[i]; g.Kvalid = g.Kpad = 4 * kDecCin[i]; g.n_mod = g.Cout; g.slope = 1.f;
      g.out = W.bf(zcat_slot[lvl_out]); g.out_ld = cat_ld[lvl_out]; g.n_valid = g.Cout;
      rc = launch_layer_auto(g, 128, B, st);
    } else {
      USLayer sl; SlabPackDesc pd;
      memset(&sl, 0, sizeof(sl));
      const int lvl_in = 5 - i, hpi = hp[lvl_in];
      build_deconv_items(kDecCin[i], kDecCoutBase[i], kDecCoutBase[i], hpi, sl, pd);
      sl.slope = 1.f; sl.out_mode = UG_OUT_NHWC_BF16; sl.out_stride = 2; sl.out_coff = 0;
      if (i == 3) { sl.out = W.bf(TW_ZCAT1); sl.Hop = sl.Wop = 18; sl.out_ld = 128; }
      else { sl.out = W.bf(TW_Z4); sl.Hop = sl.Wop = 34; sl.out_ld = 32; }
      rc = launch_slab(sl, W.bf(cat_slot[lvl_in]), kDecCin[i], hpi, B, (const __nv_bfloat16*)(pk + PL.w[i]), st);
    }
    if (rc != M2H_OK) return rc;
    const int C = kDecCoutBase[i], Creal = i == 4 ? c_last : C;
    const int hq = i < 4 ? hp[lvl_out] - 2 : 32;
    const FMap z = i < 4 ? fmap(W.bf(zcat_slot[lvl_out]), cat_ld[lvl_out], 0, hp[lvl_out], B) : fmap(W.bf(TW_Z4), 32, 0, 34, B);
    const FMap a = i < 4 ? fmap(W.bf(cat_slot[lvl_out]), cat_ld[lvl_out], 0, hp[lvl_out], B) : fmap(W.bf(TW_A4), 32, 0, 34, B);
    rc = bn_forward(W, 5 + i, z, a, C, Creal, (double)B * hq * hq, p->bn_w[i], p->bn_b[i], const_cast<float*>(p->bn_mean[i]), const_cast<float*>(p->bn_var[i]),
                    num_batches_tracked ? (long long*)num_batches_tracked[i] : nullptr, bn_eps, bn_momentum, 0.f, st);
    if (rc != M2H_OK) return rc;
  }
  M2H_CUDA_CHECK(launch_pdl(head_fwd_kernel, dim3((B * 1024 + 127) / 128), dim3(128), (size_t)(0), st, W.bf(TW_A4), p->head_w, p->head_b, mask_out, B, c_last));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradients on a forked stream
// ------------------------------------------------------------------------------------------------
// A layer's weight gradient is a leaf of the backward pass: it needs dz of its own layer and the layer's (read-only) input
// activation, and nothing downstream needs it before the optimizer.  The backward functions therefore fork: after the
// BatchNorm backward of layer i the weight-gradient kernel goes to a side stream (event fork), the data gradient and the
// next layers continue on the caller's stream, and the function joins the side stream before it returns.  The step is a
// chain of short, latency-bound kernels (B = 256: ~10 us each), so the two chains fill each other's gaps; the same
// fork / join is captured as parallel branches by pretrain.GraphedTrainStep.  M2H_NO_WGRAD_FORK=1 keeps one stream.
struct WgradFork {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, ready = nullptr;
  bool used = false;
};
static WgradFork g_forks[M2H_MAX_DEVICES][8];
static std::mutex g_fork_mutex;

static WgradFork* wgrad_fork_get(cudaStream_t st) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("M2H_NO_WGRAD_FORK"); disabled = (e && atoi(e)) ? 1 : 0; }
  if (disabled) return nullptr;
  int dev;
  if (m2h_current_device_checked(&dev) != M2H_OK) return nullptr;
  std::lock_guard<std::mutex> lock(g_fork_mutex);
  WgradFork* slot = nullptr;
  for (int i = 0; i < 8; ++i) {
    WgradFork& f = g_forks[dev][i];
    if (f.used && f.main == st) return &f;
    if (!f.used && !slot) slot = &f;
  }
  if (!slot) return nullptr;                       // more than 8 distinct caller streams: stay on the caller's stream
  if (cudaStreamCreateWithFlags(&slot->side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&slot->fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&slot->join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&slot->ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  slot->main = st; slot->used = true;
  return slot;
}
// stream for the weight gradient of the layer whose dz has just been produced on `st`
static cudaStream_t wgrad_stream(WgradFork* f, cudaStream_t st) {
  if (!f) return st;
  if (cudaEventRecord(f->fork, st) != cudaSuccess || cudaStreamWaitEvent(f->side, f->fork, 0) != cudaSuccess) return st;
  return f->side;
}
// The data-gradient weight operands only depend on the weights: they are packed on the side stream while the caller's
// stream zeroes the gradients and runs the first BatchNorm backward; the first data-gradient GEMM waits for `ready`.
static int fork_mark_ready(WgradFork* f) {                   // right after the packing launches on the side stream
  if (!f) return M2H_OK;
  M2H_CUDA_CHECK(cudaEventRecord(f->ready, f->side));
  return M2H_OK;
}
static int fork_wait_ready(WgradFork* f, cudaStream_t st) {   // before the first consumer on the caller's stream
  if (!f) return M2H_OK;
  M2H_CUDA_CHECK(cudaStreamWaitEvent(st, f->ready, 0));
  return M2H_OK;
}
static int wgrad_join(WgradFork* f, cudaStream_t st) {
  if (!f) return M2H_OK;
  M2H_CUDA_CHECK(cudaEventRecord(f->join, f->side));
  M2H_CUDA_CHECK(cudaStreamWaitEvent(st, f->join, 0));
  return M2H_OK;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
extern "C" int m2h_unet_pack_train_weights(const m2h_unet_enc_params* enc, const m2h_unet_dec_params* dec, int convert_bin2mono, int B,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_train_args("m2h_unet_pack_train_weights", B, workspace, workspace_bytes);
  if (rc != M2H_OK) return rc;
  if (!enc && !dec) { m2h_set_error("m2h_unet_pack_train_weights: null pointer"); return M2H_ERR_ARG; }
  const TW W = tw_make(workspace, B);
  if (enc) { rc = enc_pack_train(enc, convert_bin2mono & 1, W.base + W.off[TW_PACK_ENC], (cudaStream_t)stream); if (rc != M2H_OK) return rc; }
  if (dec) { rc = dec_pack_train(dec, convert_bin2mono & 1, W.base + W.off[TW_PACK_DEC], (cudaStream_t)stream); if (rc != M2H_OK) return rc; }
  return M2H_OK;
}

extern "C" int m2h_unet_decoder_backward(const m2h_unet_dec_params* p, int convert_bin2mono, const float* dmask, int B, void* workspace,
                                         size_t workspace_bytes, const m2h_unet_dec_grads* grads, void* stream) {
  int rc = check_train_args("m2h_unet_decoder_backward", B, workspace, workspace_bytes);
  if (rc != M2H_OK) return rc;
  if (!p || !dmask || !grads) { m2h_set_error("m2h_unet_decoder_backward: null pointer"); return M2H_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const TW W = tw_make(workspace, B);
  char* pk = W.base + W.off[TW_PACK_DEC_DG];
  const int c_last = convert_bin2mono ? 16 : 32;
  const int cat_slot[5] = {0, TW_CAT1, TW_CAT2, TW_CAT3, TW_CAT4}, zcat_slot[5] = {0, TW_ZCAT1, TW_ZCAT2, TW_ZCAT3, TW_ZCAT4};
  const int dcat_slot[5] = {0, TW_DCAT1, TW_DCAT2, TW_DCAT3, TW_DCAT4};
  const int cat_ld[5] = {0, 128, 256, 512, 1024}, hp[5] = {0, 18, 10, 6, 4};
  // data-gradient operands: ConvTranspose2d weight [Cin][Cout][4][4] read as a Conv2d weight [out = Cin][in = Cout]
  WgradFork* fk = wgrad_fork_get(st);
  const cudaStream_t pst = wgrad_stream(fk, st);            // packed next to the memsets / head backward / first BatchNorm backward
  for (int i = 4; i >= 0; --i) {                            // in the order the data gradients need them
    const int cout_real = i == 4 ? c_last : kDecCoutBase[i];
    M2H_CUDA_CHECK(launch_pdl(pack_conv_w_kernel, dim3(256), dim3(256), (size_t)(0), pst, p->deconv_w[i], (__nv_bfloat16*)(pk + dec_dg_off(i)), kDecCin[i], cout_real, kDecCoutBase[i], i == 0 ? 4 : 16));
  }
  rc = fork_mark_ready(fk);
  if (rc != M2H_OK) return rc;
  bool packs_joined = false;
  for (int i = 0; i < 5; ++i) {
    const int cout_real = i == 4 ? c_last : kDecCoutBase[i];
    M2H_CUDA_CHECK(cudaMemsetAsync(grads->deconv_w[i], 0, (size_t)kDecCin[i] * cout_real * 16 * 4, st));
  }
  M2H_CUDA_CHECK(cudaMemsetAsync(grads->head_w, 0, (size_t)c_last * c_last * 4, st));
  M2H_CUDA_CHECK(cudaMemsetAsync(grads->head_b, 0, (size_t)c_last * 4, st));
  const int hb_chunks = (B * 1024 + 127) / 128;       // persistent: ~4 blocks per SM, each accumulates dW over its chunks
  M2H_CUDA_CHECK(launch_pdl(head_bwd_kernel, dim3(hb_chunks < 592 ? hb_chunks : 592), dim3(128), (size_t)(0), st, W.bf(TW_A4), p->head_w, dmask, W.bf(TW_DA4), grads->head_w, grads->head_b, B, c_last));
  M2H_CUDA_CHECK(cudaGetLastError());
  const FMap none = fmap(nullptr, 0, 0, 0, 0);
  for (int i = 4; i >= 0; --i) {
    const int lvl_out = 4 - i, lvl_in = 5 - i;          // lvl_in == 5: the bottleneck
    const int C = kDecCoutBase[i], Creal = i == 4 ? c_last : C;
    const int hq_out = i < 4 ? hp[lvl_out] - 2 : 32;
    const FMap z = i < 4 ? fmap(W.bf(zcat_slot[lvl_out]), cat_ld[lvl_out], 0, hp[lvl_out], B) : fmap(W.bf(TW_Z4), 32, 0, 34, B);
    const FMap a = i < 4 ? fmap(W.bf(cat_slot[lvl_out]), cat_ld[lvl_out], 0, hp[lvl_out], B) : fmap(W.bf(TW_A4), 32, 0, 34, B);
    const FMap da = i < 4 ? fmap(W.bf(dcat_slot[lvl_out]), cat_ld[lvl_out], 0, hp[lvl_out], B) : fmap(W.bf(TW_DA4), 32, 0, 34, B);
    rc = bn_backward(W, 5 + i, da, none, a, z, C, Creal, (double)B * hq_out * hq_out, 0.f, grads->bn_w[i], grads->bn_b[i], st);
    if (rc != M2H_OK) return rc;
    // weight gradient: dense = layer input (cat_{lvl_in} or the bottleneck), sampled = dz (now in z)
    const int hq_in = i == 0 ? 1 : hp[lvl_in] - 2;
    const cudaStream_t wst = wgrad_stream(fk, st);
    if (i == 0) rc = launch_wgrad(W.bf(TW_BOTT), 512, 0, 512, false, z.p, z.ld, 0, Creal, Creal, 1, B, grads->deconv_w[0], wst);
    else rc = launch_wgrad(W.bf(cat_slot[lvl_in]), cat_ld[lvl_in], 0, kDecCin[i], true, z.p, z.ld, 0, Creal, Creal, hq_in, B, grads->deconv_w[i], wst);
    if (rc != M2H_OK) return rc;
    // data gradient: Conv2d(k4, s2, p1) over dz -> d(cat_{lvl_in}) (all its channels: previous decoder output | skip) or d(bottleneck)
    if (!packs_joined) { rc = fork_wait_ready(fk, st); if (rc != M2H_OK) return rc; packs_joined = true; }
    UGLayer g;
    layer_defaults(g);
    g.in = z.p; g.Hin = g.Win = hq_out; g.Cin = C; g.cin_shift = ilog2(C); g.in_ld = z.ld; g.in_pad = 1;
    g.Hq = g.Wq = hq_out / 2; g.M = B * g.Hq * g.Wq; g.in_stride = 2; g.n_taps = i == 0 ? 4 : 16;
    for (int t = 0; t < g.n_taps; ++t) {
      const int kh = g.n_taps == 16 ? t >> 2 : 1 + (t >> 1), kw = g.n_taps == 16 ? t & 3 : 1 + (t & 1);
      g.dy[0][t] = (int8_t)(kh - 1); g.dx[0][t] = (int8_t)(kw - 1);
    }
    g.w = (const __nv_bfloat16*)(pk + dec_dg_off(i)); g.Cout = kDecCin[i]; g.Kvalid = g.Kpad = g.n_taps * C; g.n_mod = g.Cout; g.slope = 1.f;
    if (i == 0) { g.out = W.bf(TW_DBOTT); g.out_ld = 512; g.out_pad = 0; }
    else { g.out = W.bf(dcat_slot[lvl_in]); g.out_ld = cat_ld[lvl_in]; g.out_pad = 1; }
    g.Hout = g.Wout = g.Hq; g.n_valid = g.Cout;
    rc = launch_layer_auto(g, 128, B, st);
    if (rc != M2H_OK) return rc;
  }
  return wgrad_join(fk, st);
}

extern "C" int m2h_unet_encoder_backward(const m2h_unet_enc_params* p, int convert_bin2mono, const float* target_class_plus1, int B,
                                         void* workspace, size_t workspace_bytes, const m2h_unet_enc_grads* grads, void* stream) {
  int rc = check_train_args("m2h_unet_encoder_backward", B, workspace, workspace_bytes);
  if (rc != M2H_OK) return rc;
  if (!p || !grads || (!convert_bin2mono && !target_class_plus1)) { m2h_set_error("m2h_unet_encoder_backward: null pointer"); return M2H_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const TW W = tw_make(workspace, B);
  char* pk = W.base + W.off[TW_PACK_ENC_DG];
  const int cat_slot[5] = {0, TW_CAT1, TW_CAT2, TW_CAT3, TW_CAT4}, zcat_slot[5] = {0, TW_ZCAT1, TW_ZCAT2, TW_ZCAT3, TW_ZCAT4};
  const int dcat_slot[5] = {0, TW_DCAT1, TW_DCAT2, TW_DCAT3, TW_DCAT4}, denc_slot[5] = {0, TW_DENC1, TW_DENC2, TW_DENC3, TW_DENC4};
  const int cat_ld[5] = {0, 128, 256, 512, 1024}, hp[5] = {0, 18, 10, 6, 4};
  // data-gradient operands: Conv2d weight [Cout][Cin][4][4] read as a ConvTranspose2d weight [in = Cout][out = Cin]
  WgradFork* fk = wgrad_fork_get(st);
  const cudaStream_t pst = wgrad_stream(fk, st);
  for (int i = 4; i >= 1; --i) {                            // in the order the data gradients need them
    if (i == 4) M2H_CUDA_CHECK(launch_pdl(pack_deconv0_w_kernel, dim3(256), dim3(256), (size_t)(0), pst, p->conv_w[4], (__nv_bfloat16*)(pk + enc_dg_off(4)), 512, 512));
    else M2H_CUDA_CHECK(launch_pdl(pack_deconv_w_kernel, dim3(256), dim3(256), (size_t)(0), pst, p->conv_w[i], (__nv_bfloat16*)(pk + enc_dg_off(i)), kEncCout[i], kEncCin[i], kEncCin[i]));
  }
  rc = fork_mark_ready(fk);
  if (rc != M2H_OK) return rc;
  bool packs_joined = false;
  for (int i = 0; i < 5; ++i) {
    const int cin_full = i == 0 ? (convert_bin2mono ? 32 : 33) : kEncCin[i];
    M2H_CUDA_CHECK(cudaMemsetAsync(grads->conv_w[i], 0, (size_t)kEncCout[i] * cin_full * 16 * 4, st));
  }
  const FMap none = fmap(nullptr, 0, 0, 0, 0);
  for (int i = 4; i >= 0; --i) {
    const int C = kEncCout[i];
    const int lvl = i + 1;                          // output level of layer i (5 = bottleneck)
    const int hq = i < 4 ? hp[lvl] - 2 : 1;
    const FMap z = i < 4 ? fmap(W.bf(zcat_slot[lvl]), cat_ld[lvl], cat_ld[lvl] / 2, hp[lvl], B) : fmap(W.bf(TW_ZBOTT), 512, 0, 0, B);
    const FMap a = i < 4 ? fmap(W.bf(cat_slot[lvl]), cat_ld[lvl], cat_ld[lvl] / 2, hp[lvl], B) : fmap(W.bf(TW_BOTT), 512, 0, 0, B);
    // gradient w.r.t. the activation: decoder's concat consumer (+ the next encoder layer), or d(bottleneck)
    const FMap da1 = i < 4 ? fmap(W.bf(dcat_slot[lvl]), cat_ld[lvl], cat_ld[lvl] / 2, hp[lvl], B) : fmap(W.bf(TW_DBOTT), 512, 0, 0, B);
    const FMap da2 = i < 4 ? fmap(W.bf(denc_slot[lvl]), C, 0, hp[lvl], B) : none;
    rc = bn_backward(W, i, da1, da2, a, z, C, C, (double)B * hq * hq, 0.2f, grads->bn_w[i], grads->bn_b[i], st);
    if (rc != M2H_OK) return rc;
    // weight gradient: dense = dz (output resolution), sampled = layer input
    const cudaStream_t wst = wgrad_stream(fk, st);
    if (i == 0) {
      rc = launch_wgrad(z.p, z.ld, z.coff, 64, true, W.bf(TW_X0P), 32, 0, 32, convert_bin2mono ? 32 : 33, 16, B, grads->conv_w[0], wst);
      if (rc != M2H_OK) return rc;
      if (!convert_bin2mono) {
        M2H_CUDA_CHECK(launch_pdl(wgrad_cls_kernel, dim3(dim3(64, 16)), dim3(256), (size_t)(0), wst, z.p, z.ld, z.coff, target_class_plus1, B, grads->conv_w[0], 33));
        M2H_CUDA_CHECK(cudaGetLastError());
      }
    } else {
      rc = launch_wgrad(z.p, z.ld, z.coff, C, i < 4, W.bf(cat_slot[i]), cat_ld[i], cat_ld[i] / 2, kEncCin[i], kEncCin[i], hq, B, grads->conv_w[i], wst);
      if (rc != M2H_OK) return rc;
      // data gradient: ConvTranspose2d(k4, s2, p1) over dz -> denc_i (gradient of the skip feature at level i)
      if (!packs_joined) { rc = fork_wait_ready(fk, st); if (rc != M2H_OK) return rc; packs_joined = true; }
      UGLayer g;
      layer_defaults(g);
      g.w = (const __nv_bfloat16*)(pk + enc_dg_off(i)); g.slope = 1.f;
      if (i == 4) {
        g.in = z.p; g.Hin = g.Win = 1; g.Cin = 512; g.cin_shift = 9; g.in_ld = 512;
        g.Hq = g.Wq = 1; g.M = B; g.n_taps = 1; g.Cout = 2048; g.Kvalid = g.Kpad = 512; g.n_mod = 512;
        g.out = W.bf(TW_DENC4); g.Hout = g.Wout = 1; g.out_ld = 16 * 512; g.n_split = 512; g.n_valid = 2048;
        for (int pix = 0; pix < 4; ++pix) g.pix_off[pix] = (((pix >> 1) + 1) * 4 + (pix & 1) + 1) * 512;
      } else {
        g.in = z.p + z.coff; g.Cin = C; g.cin_shift = ilog2(C); g.in_ld = z.ld;
        deconv_phases(g, hq, B);
        g.Cout = kEncCin[i]; g.Kvalid = g.Kpad = 4 * C; g.n_mod = g.Cout;
        g.out = W.bf(denc_slot[i]); g.out_ld = kEncCin[i]; g.n_valid = g.Cout;
      }
      rc = launch_layer_auto(g, kEncCin[i] >= 128 ? 128 : 64, B, st);
      if (rc != M2H_OK) return rc;
    }
  }
  return wgrad_join(fk, st);
}

// ------------------------------------------------------------------------------------------------
// loss
// ------------------------------------------------------------------------------------------------
extern "C" int m2h_l1_mask_loss(const float* mask, const float* obs, const float* gt, int B, int C, int use_mix, float grad_scale,
                                double* loss_sum, float* dmask, void* stream) {
  if (B <= 0 || (C != 1 && C != 2) || !mask || !gt || !loss_sum || (use_mix && !obs)) { m2h_set_error("m2h_l1_mask_loss: bad arguments"); return M2H_ERR_ARG; }
  int dev;
  int rc = m2h_current_device_checked(&dev);
  if (rc != M2H_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  M2H_CUDA_CHECK(cudaMemsetAsync(loss_sum, 0, sizeof(double), st));
  M2H_CUDA_CHECK(launch_pdl(l1_mask_loss_kernel, dim3(grid_for((long long)B * C * 16384, 1024)), dim3(256), (size_t)(0), st, mask, obs, gt, B, C, use_mix, grad_scale, loss_sum, dmask));
  M2H_CUDA_CHECK(cudaGetLastError());
  return M2H_OK;
}

/*
 * m2h.h - C ABI of the B200-native Move2Hear audio hot path (libm2h_b200.so).
 *
 * The reference (SAGNIKMJR/move2hear-active-AV-separation) is pure Python and has no FFI for this
 * path; the entry points below are what a ctypes binding placed at the reference's own call sites
 * binds (INTEGRATION.md shows the stubs).  Every function takes raw DEVICE pointers, sizes and a
 * CUDA stream (cudaStream_t passed as void*; NULL = legacy default stream), launches asynchronously
 * on that stream, allocates nothing, and returns 0 on success or a negative m2h_status code
 * (message via m2h_last_error()).  There is no CPU fallback: without a CUDA device every compute
 * call fails with M2H_ERR_CUDA.
 *
 * Array layouts use the reference's own axis orders (C-contiguous).
 */
#ifndef M2H_H_
#define M2H_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M2H_ABI_VERSION 4

enum m2h_status {
  M2H_OK = 0,
  M2H_ERR_ARG = -1,        /* bad size / null pointer / unsupported shape */
  M2H_ERR_CUDA = -2,       /* CUDA runtime error (no device, launch failure, ...) */
  M2H_ERR_WORKSPACE = -3   /* caller-supplied workspace too small */
};

/* flags for m2h_render_mixture / m2h_stft_waves */
enum m2h_render_flags {
  M2H_RENDER_F16_OBS = 1,      /* simulator dtypes: magnitude -> fp16 -> log1p -> fp16, phases / GT in fp16
                                  (habitat_audio/simulator_train.py:437-441, 483-486).  Otherwise the dataset's
                                  fp32 outputs (audio_separation/pretrain/datasets/dataset.py:198, 226-228). */
  M2H_RENDER_RAW_MAG = 2,      /* write |STFT| of the mixture instead of log1p(|STFT|) */
  M2H_RENDER_GT_ONLY = 4       /* m2h_stft_waves: only the n_gt per-source spectrograms (mix_mag / mix_phase ignored): the
                                  mixture's transform comes from another launch, e.g. on another stream */
};

int m2h_abi_version(void);
const char* m2h_last_error(void);

/* Upload the twiddle / chirp tables to `device` (idempotent; called lazily by every entry point). */
int m2h_init(int device);

/* Bytes of device workspace m2h_render_mixture / m2h_rir_convolve need (spectrum scratch, one slot per
 * resident CTA; independent of B). */
size_t m2h_render_workspace_bytes(int device);

/*
 * RIR convolution + quantisation for B*K (mixture, source) pairs.
 * Replaces: scipy.signal.fftconvolve(mono, rir[:, ch], mode="same") -> np.round -> astype(int16)
 *   habitat_audio/simulator_train.py:417-421, habitat_audio/simulator_eval.py:429-433,
 *   audio_separation/pretrain/datasets/dataset.py:178-183.
 *   mono     int16  [B, K, N]          source clips (N <= 16384; the reference uses N = 16000)
 *   rir      float  [B, K, 2, Lmax]    binaural RIRs, ear-planar (transpose of the wav's [L, 2])
 *   rir_len  int32  [B, K]             taps per RIR (<= Lmax); 0 = unreadable/empty file -> zeros
 *                                      (simulator_train.py:407-415)
 *   waves    int16  [B, K, 2, N]       OUT: rounded binaural waveform of each source, int16 full-scale units
 *                                      (the reference's float value is waves * 2^-15)
 */
int m2h_rir_convolve(const int16_t* mono, const float* rir, const int32_t* rir_len,
                     int B, int K, int N, int Lmax, int16_t* waves,
                     void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same convolution with the inputs referenced by row in device-resident pools (device-side
 * LRU caches of the clips, keyed by sound name like `_cached_mono_waveforms`, simulator_train.py:396-400, and of the RIRs,
 * keyed by (scene, azimuth, receiver, source), simulator_train.py:402-415): per step only the index arrays cross PCIe.
 *   clip_pool     int16 [n_clips, N]        clip_index  int32 [B, K]
 *   rir_pool      float [n_rirs, 2, Lmax]   rir_pool_len int32 [n_rirs]   rir_index int32 [B, K]
 */
int m2h_rir_convolve_indexed(const int16_t* clip_pool, const int32_t* clip_index, const float* rir_pool, const int32_t* rir_pool_len,
                             const int32_t* rir_index, int B, int K, int N, int Lmax, int16_t* waves,
                             void* workspace, size_t workspace_bytes, void* stream);

/*
 * wav ingest.  Replaces scipy.io.wavfile.read (simulator_train.py:397, 409; dataset.py:157, 170) for the sample formats
 * it returns without conversion (PCM16 -> int16, PCM32 -> int32, IEEE float32).  m2h_wav_parse is host-only: it locates
 * the fmt / data chunks of a RIFF/WAVE image.  m2h_rir_ingest / m2h_clip_ingest run on the device: the raw interleaved
 * frames (device pointer, e.g. an H2D copy of the file body from pinned memory) become one pool row - fp32 ear-planar
 * [2, Lmax] (the transpose of the wav's [L, 2]; zero tail) or the int16 clip [N] (channel 0, zero tail).
 */
enum { M2H_WAV_PCM16 = 1, M2H_WAV_PCM32 = 2, M2H_WAV_FLOAT32 = 3 };
typedef struct m2h_wav_info {
  int format;            /* M2H_WAV_* */
  int channels;
  int sample_rate;
  int bits_per_sample;
  long long frames;      /* samples per channel */
  size_t data_offset;    /* byte offset of the first frame inside the file image */
} m2h_wav_info;
int m2h_wav_parse(const void* file_bytes, size_t n_bytes, m2h_wav_info* out);
int m2h_rir_ingest(const void* raw_frames, int format, int channels, int frames, float* pool_row, int Lmax, void* stream);
int m2h_clip_ingest(const void* raw_frames, int format, int channels, int frames, int16_t* pool_row, int N, void* stream);

/*
 * STFT (n_fft 1023, hop 512, periodic Hann, centre/reflect) of the mixture of K quantised sources and,
 * optionally, of the first n_gt sources alone.
 * Replaces: librosa.stft / librosa.magphase / np.angle / np.log1p at
 *   habitat_audio/simulator_train.py:425-441, 469-486, habitat_audio/simulator_eval.py:437-500,
 *   audio_separation/pretrain/datasets/dataset.py:190-198, 216-228.
 *   waves      int16 [B, K, 2, N]
 *   mix_mag    OUT [B, 512, T, 2]  log1p(|STFT(mix)|)  (fp32, or fp16 with M2H_RENDER_F16_OBS), T = 1 + (N-1)/512
 *   mix_phase  OUT [B, 512, T, 2]  angle(STFT(mix)) or NULL
 *   gt         OUT or NULL.  gt_with_phase = 0: [B, n_gt, 512, T, 2] magnitudes (dataset.py:198);
 *              gt_with_phase = 1: [B, n_gt, 512, T, 4] = (mag_l, phase_l, mag_r, phase_r)
 *              (simulator_train.py:437-441).  dtype follows M2H_RENDER_F16_OBS.
 */
int m2h_stft_waves(const int16_t* waves, int B, int K, int N, int n_gt, int gt_with_phase, int flags,
                   void* mix_mag, void* mix_phase, void* gt, void* stream);

/*
 * STFT of the mixture written as the binaural separator's first-layer operand: x0 = bf16 [B][18][18][128], the haloed
 * space-to-depth NHWC form of log1p(|STFT(mix)|) [B,512,32,2] (cell (h/2+1, t/2+1), channel ((h&1)*2 + (t&1))*32 + ear*16 +
 * slice, h = bin % 32, slice = bin / 32: audio_separation/rl/models/separator_cnn.py:84-90).  Only interior cells are
 * written: the caller zeroes the buffer once.  N must give 32 frames.  mix_mag (fp32 [B,512,32,2]) is optional.
 * Feeds m2h_unet_encoder_forward_x0: the fp32 observation never travels through HBM and no repacking kernel runs.
 */
int m2h_stft_waves_x0(const int16_t* waves, int B, int K, int N, void* x0, float* mix_mag, void* stream);

/*
 * The whole per-step render: m2h_rir_convolve followed by m2h_stft_waves on the same stream.
 * Replaces HabitatSimAudioEnabledTrain.get_current_mixed_bin_audio_mag_spec
 * (habitat_audio/simulator_train.py:386-486; eval twin simulator_eval.py:396-500) and
 * PassiveDataset.compute_audiospects (audio_separation/pretrain/datasets/dataset.py:162-228)
 * minus the file reads and the cached monaural GT (see m2h_stft_f32).
 */
int m2h_render_mixture(const int16_t* mono, const float* rir, const int32_t* rir_len,
                       int B, int K, int N, int Lmax, int n_gt, int gt_with_phase, int flags,
                       int16_t* waves, void* mix_mag, void* mix_phase, void* gt,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * STFT magnitude (and optional phase) of arbitrary fp32 waveforms [B, C, N], C in {1, 2} -> [B, 512, T, C] fp32.
 * Used for the per-sound cached monaural GT spectrograms (simulator_train.py:443-449, dataset.py:201-207).
 */
int m2h_stft_f32(const float* wave, int B, int C, int N, float* mag, float* phase, void* stream);


/*
 * Inverse STFT as the reference's evaluation calls it:
 * Replaces librosa.istft(mag * exp(1j * phase), hop_length=512, length=16000) in eval_metrics.istft
 *   (audio_separation/common/eval_metrics.py:232-252).  librosa infers n_fft = 2*(512-1) = 1022 from the 512 bins (not the
 *   forward transform's 1023): 1022-point inverse real DFT per frame, periodic Hann(1022) synthesis window, overlap-add at
 *   hop 512, division by the window's sum of squares, first 511 samples dropped, cut / zero-padded to `length`.
 *   mag, phase  float [B, 512, T, C]  (C = 1: monaural, C = 2: both ears in one pass)
 *   out         float [B, C, length]  OUT
 */
int m2h_istft(const float* mag, const float* phase, int B, int C, int T, int length, float* out, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Passive-separator U-Net, inference statistics (eval-mode BatchNorm).
 * Replaces PassiveSepEncCNN.forward / PassiveSepDecCNN.forward
 * (audio_separation/rl/models/separator_cnn.py:70-108, 153-170) as composed by Policy.get_binSepMasks /
 * convert_bin2mono (audio_separation/pretrain/passive/policy.py:61-71, audio_separation/rl/ppo/policy.py:183-193).
 * Conv / transposed-conv layers run as bf16 implicit GEMMs on tcgen05 with fp32 accumulation; BatchNorm,
 * (Leaky)ReLU, the target-class plane, the skip concatenations and the 1x1 head are fused into them.
 * The spectrogram must be [512, 32] (N = 16000 samples).
 * ------------------------------------------------------------------------------------------------ */

/* fp32 DEVICE pointers to the reference module's own parameters/buffers (state_dict layouts). */
typedef struct m2h_unet_enc_params {
  const float* conv_w[5];    /* cnn.{i}.0.weight  [Cout, Cin, 4, 4]  (Cin = 33 for layer 0 of the binaural separator) */
  const float* bn_w[5];      /* cnn.{i}.1.weight */
  const float* bn_b[5];      /* cnn.{i}.1.bias */
  const float* bn_mean[5];   /* cnn.{i}.1.running_mean */
  const float* bn_var[5];    /* cnn.{i}.1.running_var */
} m2h_unet_enc_params;

typedef struct m2h_unet_dec_params {
  const float* deconv_w[5];  /* cnn.{i}.0.weight  [Cin, Cout, 4, 4] (ConvTranspose2d) */
  const float* bn_w[5];
  const float* bn_b[5];
  const float* bn_mean[5];
  const float* bn_var[5];
  const float* head_w;       /* cnn.5.0.weight [C, C, 1, 1], C = 32 (binaural) or 16 (bin2mono) */
  const float* head_b;       /* cnn.5.0.bias   [C] */
} m2h_unet_dec_params;

size_t m2h_unet_enc_packed_bytes(int convert_bin2mono);
size_t m2h_unet_dec_packed_bytes(int convert_bin2mono);
/* Fold BatchNorm into per-channel scale/bias and re-lay the weights as bf16 GEMM operands into `packed`. */
int m2h_unet_pack_encoder(const m2h_unet_enc_params* p, int convert_bin2mono, float bn_eps, void* packed, void* stream);
int m2h_unet_pack_decoder(const m2h_unet_dec_params* p, int convert_bin2mono, float bn_eps, void* packed, void* stream);

/* Activation workspace shared by one encoder/decoder pair: bf16 NHWC feature maps, each stored with a zero border of one
 * pixel ("haloed": [B, H+2, W+2, C]) so that conv padding and tap shifts need no bounds logic; the skip connections
 * live in the upper channel half of the decoder's concat buffers.  offsets[7] = byte offsets of
 * {x0 [B,18,18,128] (network input, space-to-depth: cell (h/2, w/2), channel ((h&1)*2 + (w&1))*32 + c),
 *  cat1 [B,18,18,128], cat2 [B,10,10,256], cat3 [B,6,6,512], cat4 [B,4,4,1024], bottleneck [B,512], (unused)}. */
size_t m2h_unet_workspace_bytes(int B);
int m2h_unet_workspace_layout(int B, size_t offsets[7]);
/* Zero the borders of the feature maps.  Contract (ABI version 2): call it ONCE after allocating a workspace (and again
 * if anything but this library wrote to it) - m2h_unet_encoder_forward / m2h_unet_decoder_forward write interiors only
 * and rely on the borders being zero; they no longer clear them per call (43 us per 1024 samples). */
int m2h_unet_workspace_clear_halos(int B, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Encoder.  obs: fp32 [B, 512, 32, 2] log1p magnitudes (observations["mixed_bin_audio_mag"], or `mixed_audio`
 * when convert_bin2mono).  target_class_plus1: fp32 [B] = target_class + 1 (binaural separator only).
 * masks (+ element strides of its [B, 512, 32, 2] view): predicted ratio masks, convert_bin2mono only; the encoder
 * input is then log1p(clamp(masks * (exp(obs) - 1), 0)) (separator_cnn.py:73-79).
 * bottleneck_f32: optional fp32 [B, 512] copy of the bottleneck features.
 */
int m2h_unet_encoder_forward(const void* packed_enc, int convert_bin2mono, const float* obs,
                             const float* target_class_plus1, const float* masks, const int64_t* mask_strides,
                             int B, void* workspace, size_t workspace_bytes, float* bottleneck_f32, void* stream);
/* Same, for the binaural separator, with the first-layer operand x0 (see m2h_stft_waves_x0) instead of the fp32 observation.
 * target_class: the observation's raw int64 class indices [B] (separator_cnn.py:91-100 adds 1 and broadcasts the plane: done
 * in the first layer's epilogue here).  bottleneck_f32 may be NULL: the decoder reads the bf16 copy in the workspace. */
int m2h_unet_encoder_forward_x0(const void* packed_enc, const void* x0, const int64_t* target_class, int B,
                                void* workspace, size_t workspace_bytes, float* bottleneck_f32, void* stream);
/* Decoder.  mask_out: fp32 [B, C, 512, 32] (C = 2 or 1) - the memory layout behind the reference's permuted
 * [B, 512, 32, C] return value (separator_cnn.py:163-170). */
int m2h_unet_decoder_forward(const void* packed_dec, int convert_bin2mono, int B, void* workspace,
                             size_t workspace_bytes, float* mask_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Passive-separator U-Net, training (batch-statistics BatchNorm) forward + backward, and the pretraining loss.
 * Replaces the .train()-mode forward/backward of PassiveSepEncCNN / PassiveSepDecCNN
 * (audio_separation/rl/models/separator_cnn.py:27-170) under PassiveTrainer.train
 * (audio_separation/pretrain/passive/passive_trainer.py:226-236) and the L1 losses of optimize_supervised_loss (:269-275).
 * Convolutions, their data gradients (the forward kernels with re-packed weights) and weight gradients run as bf16
 * tcgen05 GEMMs with fp32 accumulation; BatchNorm statistics are reduced in fp64.  One workspace holds the activations,
 * the raw conv outputs, the gradients and the packed weights of one encoder/decoder pair between forward and backward.
 * ------------------------------------------------------------------------------------------------ */
typedef struct m2h_unet_enc_grads {
  float* conv_w[5];   /* OUT, same layouts as the parameters */
  float* bn_w[5];
  float* bn_b[5];
} m2h_unet_enc_grads;
typedef struct m2h_unet_dec_grads {
  float* deconv_w[5];
  float* bn_w[5];
  float* bn_b[5];
  float* head_w;
  float* head_b;
} m2h_unet_dec_grads;

size_t m2h_unet_train_workspace_bytes(int B);
/* offsets[32]: byte offsets of the slots; 0-5 as m2h_unet_workspace_layout (activations), 7 a4 [B,34,34,32], 8 plain input copy,
 * 9-12 raw conv outputs of the concat levels (overwritten by their gradients in backward), 13 z bottleneck, 14 z4,
 * 15-18 gradients of the concat buffers [B,H+2,W+2,2C] (lower half: decoder activation, upper half: skip), 19 d bottleneck [B,512],
 * 20 d a4, 21-24 encoder-side skip gradients, 25+ scratch.  offsets[31] = total bytes. */
int m2h_unet_train_workspace_layout(int B, size_t offsets[32]);
/* Zero a fresh workspace once (feature-map borders must be zero; the kernels only write interiors). */
int m2h_unet_train_workspace_init(int B, void* workspace, size_t workspace_bytes, void* stream);

/* Train-mode weight packing ahead of time: the bf16 GEMM operands of `enc` and / or `dec` (either may be NULL) are written
 * into the train workspace - the work m2h_unet_*_forward_train does first.  A training loop issues it on a side stream while
 * the step's audio is still being rendered, then passes `convert_bin2mono | M2H_UNET_WEIGHTS_PACKED` to the forward
 * calls (the weights must not change in between; the caller orders the streams). */
#define M2H_UNET_WEIGHTS_PACKED 2
int m2h_unet_pack_train_weights(const m2h_unet_enc_params* enc, const m2h_unet_dec_params* dec, int convert_bin2mono, int B,
                                void* workspace, size_t workspace_bytes, void* stream);
/* Train-mode forward.  The running_mean / running_var tensors behind p->bn_mean / p->bn_var are UPDATED in place
 * (momentum, unbiased variance) and num_batches_tracked[i] (int64 device scalars, may be NULL) incremented,
 * as nn.BatchNorm2d does.  Other arguments as m2h_unet_encoder_forward / m2h_unet_decoder_forward. */
int m2h_unet_encoder_forward_train(const m2h_unet_enc_params* p, int64_t* const* num_batches_tracked, int convert_bin2mono,
                                   float bn_eps, float bn_momentum, const float* obs, const float* target_class_plus1,
                                   const float* masks, const int64_t* mask_strides, int B, void* workspace, size_t workspace_bytes,
                                   float* bottleneck_f32, void* stream);
int m2h_unet_decoder_forward_train(const m2h_unet_dec_params* p, int64_t* const* num_batches_tracked, int convert_bin2mono,
                                   float bn_eps, float bn_momentum, int B, void* workspace, size_t workspace_bytes,
                                   float* mask_out, void* stream);
/* Backward of the decoder given dmask = dL/d(mask_out) (fp32 [B, C, 512, 32]): parameter gradients into `grads`,
 * gradients of the bottleneck / skip features left in the workspace (slots 15-19) for m2h_unet_encoder_backward. */
int m2h_unet_decoder_backward(const m2h_unet_dec_params* p, int convert_bin2mono, const float* dmask, int B, void* workspace,
                              size_t workspace_bytes, const m2h_unet_dec_grads* grads, void* stream);
int m2h_unet_encoder_backward(const m2h_unet_enc_params* p, int convert_bin2mono, const float* target_class_plus1, int B,
                              void* workspace, size_t workspace_bytes, const m2h_unet_enc_grads* grads, void* stream);

/* L1 loss of optimize_supervised_loss: use_mix = 1: pred = mask * (exp(obs) - 1) vs gt (binaural, C = 2);
 * use_mix = 0: pred = mask vs gt (monaural, C = 1).  mask: fp32 [B, C, 512, 32]; obs: fp32 [B, 512, 32, 2]; gt: fp32 [B, 512, 32, C].
 * loss_sum (device double) = sum |pred - gt|; dmask (may be NULL) = sign(pred - gt) * mix * grad_scale. */
int m2h_l1_mask_loss(const float* mask, const float* obs, const float* gt, int B, int C, int use_mix, float grad_scale,
                     double* loss_sum, float* dmask, void* stream);

/*
 * ---- consumers of the masks on the rollout / evaluation step ----
 *
 * STFT_L2_distance(mixed_audio, pred_binSepMasks, gt_bin_comps, pred_mono, gt_mono_comps)
 *   audio_separation/common/eval_metrics.py:306-366; call sites rl/ppo/ppo_trainer.py:407-420.
 *   mixed_audio   float [B, 512, T, 2]  log1p magnitudes (contiguous, 8-byte aligned)
 *   pred_masks    float [B, 512, T, 2]  any strides (element strides in mask_strides[4])
 *   gt_bin_comps  float [B, 512, T, c_bin]   components 0..3 = mag_l, phase_l, mag_r, phase_r (c_bin >= 4)
 *   pred_mono     float [B, 512, T, 1]  any strides (mono_strides[4])
 *   gt_mono_comps float [B, 512, T, c_mono]  components 0..1 = mag, phase (c_mono >= 2)
 *   bin_dist, mono_dist  float [B]  OUT: the reference's [B, 1] tensors
 */
int m2h_stft_l2_distance(const float* mixed_audio, const float* pred_masks, const int64_t* mask_strides,
                         const float* gt_bin_comps, int c_bin, const float* pred_mono, const int64_t* mono_strides,
                         const float* gt_mono_comps, int c_mono, int B, int T, float* bin_dist, float* mono_dist,
                         void* stream);

/*
 * Raw moments of a (reference, estimate, mixture) waveform triple, from which every scale-invariant BSS metric of
 * scale_bss_eval_helper / scale_bss_eval / preprocess (audio_separation/common/eval_metrics.py:12-57, 60-122, 167-196)
 * follows in closed form for the reference's single-source evaluation (compute_waveform_quality, :256-303).
 *   reference, estimate float [B, L];  mixture float [B, mix_channels, L] (its channel mean is the "mixture" estimate)
 *   moments double [B, 9] OUT = { L, S r, S e, S m, S rr, S ee, S mm, S re, S rm }
 */
int m2h_wave_moments(const float* reference, const float* estimate, const float* mixture, int mix_channels,
                     int B, int L, double* moments, void* stream);

/*
 * AcousticMem.forward(pred_mono, prev_pred_monoFromMem_masked)  (audio_separation/rl/models/memory_nets.py:5-69;
 * call sites rl/ppo/ppo_trainer.py:310-319, rl/ppo/policy.py:195-196), inference statistics:
 * frequency slicing into 16 chunks, channel concat, Conv2d(32,32,3,p1) [+ BatchNorm2d(32), absent when use_ddppo] + ReLU,
 * Conv2d(32,16,3,p1), de-slicing - one launch, fp32.
 *   pred_mono, prev_pred  float [B, 512, 32, 1]  any strides;  conv1_w [32,32,3,3], conv2_w [16,32,3,3]
 *   bn_*  [32] or all NULL;  out float [B, 512, 32, 1] contiguous
 */
int m2h_acoustic_mem_forward(const float* pred_mono, const int64_t* pred_mono_strides, const float* prev_pred,
                             const int64_t* prev_pred_strides, const float* conv1_w, const float* conv2_w,
                             const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var, float bn_eps,
                             int B, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* M2H_H_ */

/*
 * vidtok_b200 -- C ABI of the B200-native VidTok causal tokenizer hot path
 * (encode -> KL/FSQ regularize -> decode).
 *
 * The reference (microsoft/VidTok @ d6ad92d) has no FFI / plugin registry: the only indirection on this
 * path is `instantiate_from_config` (vidtok/modules/util.py:69-86), which builds
 * vidtok.models.autoencoder[_v1_1].AutoencodingEngine from a YAML `target:` string, and the engine's
 * encode()/decode()/forward() methods (vidtok/models/autoencoder.py:197-229,
 * vidtok/models/autoencoder_v1_1.py:230-342).  The entry points below are what a binding for THAT
 * interface needs: one opaque model handle per (config, device), a parameter manifest that uses the
 * reference's checkpoint key names, and encode/decode calls that take raw device pointers in the
 * reference's tensor layout ([B,C,T,H,W], fp32).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative vt_status and
 * records a message retrievable with vt_last_error() (thread-local).  The library never allocates
 * caller-visible outputs: the caller owns inputs, outputs and the scratch workspace (size queried
 * first).  All device work is enqueued on the caller's cudaStream_t (passed as void*); calls on one
 * handle must be serialised by the caller (the reference's modules are not re-entrant either:
 * chunk caches live on the modules, vidtok/modules/model_3dcausal_v1_1.py:155-157).
 */
#ifndef VIDTOK_B200_H
#define VIDTOK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_MAX_LEVELS 8

typedef enum {
  VT_OK = 0,
  VT_ERR_INVALID = -1,    /* bad argument / unsupported configuration */
  VT_ERR_CUDA = -2,       /* CUDA runtime or driver error */
  VT_ERR_NOT_READY = -3,  /* parameters missing or vt_model_finalize not called */
  VT_ERR_WORKSPACE = -4,  /* workspace too small */
  VT_ERR_NO_DEVICE = -5   /* no sm_100 device: there is deliberately no CPU fallback */
} vt_status;

/* Precision modes.
 * EXACT_TC (the parity mode): fp32-class results on the tcgen05 tensor cores.  Activations and weights are kept as
 *   two fp16 planes (hi = fp16(v), lo = fp16(v - hi): 11 + 11 mantissa bits); every K step issues hi*hi + lo*hi + hi*lo
 *   into the fp32 TMEM accumulator ("fp16x3": products good to ~2^-21), LayerNorm / SiLU / regularizers in fp32.
 *   Gate: 1e-3 max-abs, FSQ codes equal.
 * BF16 (the throughput mode): bf16 activations / weights, fp32 accumulation.  Gate: PSNR within 0.01 dB.
 * MIXED: encoder in EXACT_TC (bit-exact FSQ codes / 1e-3 latents), decoder in BF16.
 * FMA32 (VT_PREC_EXACT, kept for cross-checks): fp32 activations on fp32 FMA kernels, no tensor cores. */
#define VT_PREC_EXACT 0
#define VT_PREC_FMA32 0
#define VT_PREC_BF16 1
#define VT_PREC_EXACT_TC 2
#define VT_PREC_MIXED 3

#define VT_NORM_LAYERNORM 0
#define VT_NORM_GROUPNORM 1
#define VT_REG_KL 0
#define VT_REG_FSQ 1
#define VT_INTERP_NEAREST 0
#define VT_INTERP_TRILINEAR 1

/* Mirrors the `encoder_config.params` block of configs/STAR.yaml plus the regularizer choice
 * (e.g. configs/vidtok_kl_causal_488_4chn.yaml:11-36).  n_* == -1 selects the reference default
 * (vidtok/modules/model_3dcausal.py:539-540,756-757). */
typedef struct vt_model_desc {
  int32_t version;                /* 0 = v1.0 (model_3dcausal.py), 1 = v1.1 (model_3dcausal_v1_1.py) */
  int32_t ch;
  int32_t num_levels;             /* len(ch_mult) */
  int32_t ch_mult[VT_MAX_LEVELS];
  int32_t num_res_blocks;
  int32_t in_channels;
  int32_t out_ch;
  int32_t z_channels;
  int32_t double_z;
  int32_t norm_type;              /* VT_NORM_* */
  int32_t time_downsample_factor;
  int32_t n_spatial_ds, spatial_ds[VT_MAX_LEVELS];
  int32_t n_tempo_ds, tempo_ds[VT_MAX_LEVELS];
  int32_t n_spatial_us, spatial_us[VT_MAX_LEVELS];
  int32_t n_tempo_us, tempo_us[VT_MAX_LEVELS];
  int32_t interpolation_mode;     /* VT_INTERP_* (v1.1 only) */
  int32_t regularizer;            /* VT_REG_* */
  int32_t fsq_num_levels;
  int32_t fsq_levels[VT_MAX_LEVELS];
  int32_t kl_sample;              /* DiagonalGaussianRegularizer(sample=...) regularizers.py:75 */
  int32_t noncausal;              /* 1 = the non-causal family (vidtok/modules/model_3dnoncausal.py: Encoder3D / Decoder3D; v1.0
                                     only): symmetric zero padding in time, plain nn.Conv3d/1d checkpoint keys, no front
                                     padding of the clip, no dropped frames */
} vt_model_desc;

typedef struct vt_model vt_model;

const char* vt_last_error(void);
int32_t vt_abi_version(void);
/* number of kernels launched by this library on this thread since the last call with reset != 0 */
int64_t vt_launch_count(int32_t reset);

/* Optional per-kernel profile of everything this thread launches between start and stop: CUDA events on the launch
 * stream around every kernel (adds launch gaps: use it to attribute time, not to measure throughput).
 * vt_profile_stop synchronises the device and writes a JSON object
 * {"kernel": {"launches": n, "ms": total, "flops": algorithmic, "bytes": algorithmic}, ...}; returns its length. */
void vt_profile_start(void);
void vt_profile_start_detailed(void); /* keys additionally carry the layer geometry */
int32_t vt_profile_stop(char* json, int32_t cap);

/* diagnostics: co-resident 2-CTA clusters of the conv kernel for a given dynamic shared-memory size */
int32_t vt_debug_cluster_query(int32_t smem_bytes, char* msg, int32_t cap);

/* ---- model lifetime (replaces AutoencodingEngine.__init__, autoencoder.py:103-144) ---- */
int32_t vt_model_create(const vt_model_desc* desc, int32_t device, vt_model** out);
void vt_model_destroy(vt_model* m);

/* Parameter manifest: names are the reference checkpoint keys ("encoder.conv_in.conv.weight", ...;
 * SURVEY.md section 8b), shapes are the reference's (OIDHW / OIHW / OIW / [C] / [1]). */
int32_t vt_model_num_params(const vt_model* m);
int32_t vt_model_param_info(const vt_model* m, int32_t index, char* name, int32_t name_cap, int64_t* shape5,
                            int32_t* ndim);
/* Copies one fp32 parameter (host or device pointer, `numel` floats) into the model
 * (replaces load_state_dict, autoencoder.py:164). */
int32_t vt_model_load_param(vt_model* m, const char* name, const float* data, int64_t numel, int32_t is_device,
                            void* stream);
/* Repacks all parameters for the kernels (K-major bf16 tiles for tcgen05, [K][Cout] fp32 for the FMA
 * path).  Must be called after the last vt_model_load_param and before encode/decode. */
int32_t vt_model_finalize(vt_model* m, void* stream);

/* ---- whole-clip path (AutoencodingEngine.encode/decode, autoencoder.py:197-219;
 *      untiled v1.1: autoencoder_v1_1.py:230-241,286-300) ---- */
/* Latent geometry for an input of T x H x W: frames (after the encoder's front padding), height, width. */
int32_t vt_latent_shape(const vt_model* m, int32_t T, int32_t H, int32_t W, int32_t* Tz, int32_t* Hz, int32_t* Wz);
/* Number of frames decode() returns for Tz latent frames (v1.0 drops tdf-1, model_3dcausal.py:885). */
int32_t vt_decoded_frames(const vt_model* m, int32_t Tz);
int64_t vt_workspace_bytes(const vt_model* m, int32_t precision, int32_t B, int32_t T, int32_t H, int32_t W);

/* x: device fp32 [B,C,T,H,W]; C must equal in_channels (a mismatch is rejected: the reference raises a shape error in
 * conv_in, model_3dcausal.py:634).  noise: device fp32 [B,z,Tz,Hz,Wz] = the reference's
 * torch.randn(mean.shape) (distributions.py:17), required for KL with kl_sample, else NULL.
 * Outputs (device, caller-allocated): z fp32 [B,z,Tz,Hz,Wz]; indices int32 [B,Tz,Hz,Wz] (FSQ, may be
 * NULL); kl_loss 1 float (KL, may be NULL); h_pre fp32 [B,2z|z,Tz,Hz,Wz] encoder output before the
 * regularizer (may be NULL). */
int32_t vt_encode(vt_model* m, int32_t precision, const float* x, int32_t B, int32_t C, int32_t T, int32_t H, int32_t W,
                  const float* noise, float* z, int32_t* indices, float* kl_loss, float* h_pre, void* workspace,
                  int64_t workspace_bytes, void* stream);
/* z: device fp32 [B,Cz,Tz,Hz,Wz] with Cz == z_channels, or (from_indices; Cz ignored) int32 [B,Tz,Hz,Wz]
 * (autoencoder.py:205-217).
 * x_out: device fp32 [B,out_ch,vt_decoded_frames(Tz),Hz*s,Wz*s]. */
int32_t vt_decode(vt_model* m, int32_t precision, const void* z, int32_t from_indices, int32_t B, int32_t Cz, int32_t Tz,
                  int32_t Hz, int32_t Wz, float* x_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- temporal tiling with causal caches (v1.1: tile_encode / tile_decode,
 *      autoencoder_v1_1.py:244-264,302-331; per-layer caches model_3dcausal_v1_1.py:159-178,216-236,
 *      289-302,325-343).  One state object per in-flight video. ---- */
typedef struct vt_chunk_state vt_chunk_state;
int32_t vt_chunk_state_create(vt_model* m, int32_t precision, int32_t B, int32_t H, int32_t W, int32_t is_decoder,
                              int32_t use_overlap, vt_chunk_state** out);
void vt_chunk_state_destroy(vt_chunk_state* s);
int64_t vt_chunk_workspace_bytes(const vt_chunk_state* s, int32_t T_chunk);
/* x_chunk: device fp32 [B,C,Tc,H,W] (dense), C == in_channels.  Outputs as vt_encode, for this chunk only. */
int32_t vt_encode_chunk(vt_chunk_state* s, int32_t is_first, const float* x_chunk, int32_t C, int32_t Tc, const float* noise,
                        float* z, int32_t* indices, float* kl_loss, void* workspace, int64_t workspace_bytes,
                        void* stream);
/* z_chunk: device fp32 [B,Cz,Tzc,Hz,Wz] (Cz == z_channels) including the look-ahead frame when overlap applies; x_out receives
 * all decoded frames of this chunk (the caller trims the look-ahead tail as autoencoder_v1_1.py:327-328). */
int32_t vt_decode_chunk(vt_chunk_state* s, int32_t is_first, const float* z_chunk, int32_t Cz, int32_t Tzc, float* x_out,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* ---- whole-video tiling below the ABI (tile_encode / tile_decode, autoencoder_v1_1.py:218-228,244-264,302-331): the chunk
 *      schedule, the causal caches and the chunk staging run inside the library -- one call per video, no host
 *      synchronisation, no per-chunk allocation.  Chunk i+1 is staged on the library's own copy stream while chunk i
 *      computes on the caller's stream (double buffering); with x_on_host / out_on_host the staging copies are the
 *      host <-> device transfers themselves (pinned memory recommended).  The call returns with all work enqueued; the
 *      caller's stream is made to wait for the copy stream. ---- */
int64_t vt_encode_video_workspace_bytes(const vt_model* m, int32_t precision, int32_t B, int32_t T, int32_t H, int32_t W,
                                        int32_t t_chunk_enc);
/* x fp32 [B,C,T,H,W] (device, or host when x_on_host); noise device fp32 [B,z,Tz,Hz,Wz] = the per-chunk torch.randn draws
 * concatenated along T (KL with sampling; else NULL); z / indices as vt_encode for the whole video (Tz = sum of the chunks'
 * latent frames); kl_loss = mean of the per-chunk values (autoencoder_v1_1.py:261-264). */
int32_t vt_encode_video(vt_model* m, int32_t precision, const float* x, int32_t x_on_host, int32_t B, int32_t C, int32_t T,
                        int32_t H, int32_t W, int32_t t_chunk_enc, const float* noise, float* z, int32_t* indices,
                        float* kl_loss, void* workspace, int64_t workspace_bytes, void* stream);
int64_t vt_decode_video_workspace_bytes(const vt_model* m, int32_t precision, int32_t B, int32_t Tz, int32_t Hz, int32_t Wz,
                                        int32_t t_chunk_dec, int32_t use_overlap);
/* frames vt_decode_video writes (look-ahead tails dropped; the v1.1 forward then keeps the last T_in frames) */
int32_t vt_decode_video_frames(const vt_model* m, int32_t Tz, int32_t t_chunk_dec, int32_t use_overlap);
/* z device fp32 [B,Cz,Tz,Hz,Wz]; x_out fp32 [B,out_ch,vt_decode_video_frames(...),H,W] (device, or host when out_on_host) */
int32_t vt_decode_video(vt_model* m, int32_t precision, const float* z, int32_t B, int32_t Cz, int32_t Tz, int32_t Hz, int32_t Wz,
                        int32_t t_chunk_dec, int32_t use_overlap, float* x_out, int32_t out_on_host, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* ---- video I/O adjacent steps (scripts/inference_reconstruct.py:41-47,71-82,231-239): the tokenizer runs at > 1000
 *      frames/s, so the uint8 <-> float conversions around it belong on the device too ---- */
/* frames: device uint8 [T,Hs,Ws,C] (decord's HWC frames); clip: device fp32 [C,T,H,W] = Normalize(.5,.5)(frames/255)
 * of the crop window at (h0, w0) (CenterCrop; the optional Resize is not part of this entry point). */
int32_t vt_video_u8_to_clip(const uint8_t* frames, float* clip, int32_t T, int32_t Hs, int32_t Ws, int32_t C, int32_t h0,
                            int32_t w0, int32_t H, int32_t W, void* stream);
/* clip: device fp32 [C,T,H,W]; frames: device uint8 [T,H,W,C] = uint8(255 * (clamp(clip,-1,1) + 1) / 2) (tensor_to_uint8). */
int32_t vt_clip_to_video_u8(const float* clip, uint8_t* frames, int32_t C, int32_t T, int32_t H, int32_t W, void* stream);

/* ---- single operators, exposed for the parity tests (same kernels the model path launches) ---- */
typedef struct vt_conv_desc {
  int32_t B, Ti, Hi, Wi, Ci;        /* input, channels-last [B,Ti,Hi,Wi,Ci] */
  int32_t Co, kt, kh, kw;
  int32_t st, sh, sw;               /* strides */
  int32_t pt;                       /* causal front pad in time (zeros) */
  int32_t ph0, ph1, pw0, pw1;       /* spatial zero pads (top,bottom,left,right) */
  int32_t ut, uh, uw;               /* nearest-neighbour upsampling of the input folded into the gather */
  int32_t res_mode;                 /* 0 none, 1 out = conv + res, 2 out = a*res[t/2] + (1-a)*conv, 3 out = a*avgpool3(res) + (1-a)*conv */
  float alpha;
} vt_conv_desc;
/* x/res/out are channels-last activations in the precision's activation type: fp32 (FMA32), bf16 (BF16), or hi|lo split
 * bf16 rows [..., hi(C) | lo(C)] (EXACT_TC; value = hi + lo); w fp32 [Co,Ci,kt,kh,kw]; bias fp32 [Co].
 * BF16 / EXACT_TC run the tcgen05 kernel (an unsupported geometry is an error, never a silent fallback) unless
 * force_simt != 0, which runs the fp32-FMA kernel on the same operands. */
int32_t vt_op_conv(int32_t precision, int32_t force_simt, const vt_conv_desc* d, const void* x, const float* w,
                   const float* bias, const void* res, void* out, void* stream);
/* The variants of the convolution the model path uses beyond vt_conv_desc: v1.1 time padding (replicated first frame /
 * per-layer cache, model_3dcausal_v1_1.py:216-236), LayerNorm(+SiLU) fused into the epilogue (model_3dcausal.py:62-80),
 * dropped leading output frames (:883-885), fp32 [B,C,T,H,W] output (the heads), residual as a mix. */
typedef struct vt_conv_ex {
  vt_conv_desc d;
  int32_t force_simt;
  int32_t t_mode;          /* 0 zeros, 1 replicate frame 0, 2 `cache` holds cacheT frames [B,cacheT,H,W,C] in front of x */
  int32_t cacheT;
  int32_t ln_mode;         /* 0 none, 1 out := act(LN(v)), 2 out := v and out2 := act(LN(v)) */
  int32_t ln_silu;
  int32_t to_off;          /* leading output frames dropped */
  int32_t out_f32_ncdhw;   /* out is fp32 [B,Co,To,Ho,Wo] */
  int32_t res_mix;         /* res_mode 1 computes alpha*res + (1-alpha)*conv instead of res + conv */
  int32_t res_t_mode;      /* res_mode 3 front pad: 0 zero, 1 frame 0, 2 `cache` = 1 frame [B,1,H,W,C] */
} vt_conv_ex;
int32_t vt_op_conv_ex(int32_t precision, const vt_conv_ex* e, const void* x, const void* cache, const float* w,
                      const float* bias, const void* res, const float* gamma, const float* beta, void* out, void* out2,
                      void* stream);
/* Encoder conv_out with the KL / FSQ regularizer applied in the convolution's epilogue (regularizers.py:82-92,153-178,
 * distributions.py:8-18): reg_mode 1 = KL (Co = 2*zc, zc in {4,8,16}; noise NULL = the mode), 2 = FSQ (Co = zc = len(levels)).
 * h_out (fp32 [B,Co,T,H,W]) may be NULL; z fp32 [B,zc,T,H,W]; indices int32 [B,T,H,W] (FSQ); kl_loss 1 float (KL). */
int32_t vt_op_conv_regularize(int32_t precision, const vt_conv_desc* d, const void* x, const float* w, const float* bias,
                              int32_t reg_mode, int32_t zc, const int32_t* fsq_levels, const float* noise, float* h_out,
                              float* z, int32_t* indices, float* kl_loss, void* stream);
/* Encoder stem from the caller's fp32 [B,Ci,T,H,W] (t_rep replicated leading frames); out channels-last
 * [B,T+t_rep,H,W,Co] in the precision's activation type (BF16 / EXACT_TC). */
int32_t vt_op_conv_stem(int32_t precision, const float* x, const float* w, const float* bias, void* out, int32_t B,
                        int32_t Ci, int32_t T, int32_t H, int32_t W, int32_t Co, int32_t t_rep, void* stream);
/* Decoder head of the BF16 mode (tap-planes GEMM + gather): x bf16 [B,T,H,W,Ci] -> out fp32 [B,Co,T-to_off,H,W]. */
int32_t vt_op_head_planes(const void* x, const float* w, const float* bias, float* out, int32_t B, int32_t T, int32_t H,
                          int32_t W, int32_t Ci, int32_t Co, int32_t to_off, void* stream);
/* "nearest 2x upsample then conv" through the phase-collapsed weights (kind 0: Upsample, w [Co,Ci,3,3];
 * kind 1: v1.0 TimeUpsampleResCausal2x, w [C,C,3,3,3], alpha = sigmoid(mix_factor)); gamma/beta/out2 optional: the
 * following LayerNorm(+SiLU) fused into the phase convolutions. */
int32_t vt_op_upsample_conv(int32_t precision, int32_t kind, const void* x, const float* w, const float* bias, float alpha,
                            const float* gamma, const float* beta, int32_t ln_silu, void* out, void* out2, int32_t B,
                            int32_t T, int32_t H, int32_t W, int32_t Ci, int32_t Co, void* stream);
/* ResnetCausalBlock1D (model_3dcausal.py:427-499) as the BF16 mode runs it for 128 channels (one fused launch):
 * n1 = silu(LN1(x)) and x bf16 channels-last [B,T,H,W,C]; w1, w2 fp32 [C,C,3]; out = x + conv2(silu(LN2(conv1(n1))));
 * optional out2 = act(LN3(out)) (g3/be3/out2 may be NULL). */
int32_t vt_op_tblock(const void* n1, const void* x, const float* w1, const float* b1, const float* g2, const float* be2,
                     const float* w2, const float* b2, const float* g3, const float* be3, int32_t out_silu, void* out,
                     void* out2, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C, void* stream);
/* y = silu?(norm(x)) over channels-last x [rows, C]; groupnorm variants take frame geometry. */
int32_t vt_op_layernorm(int32_t precision, const void* x, const float* gamma, const float* beta, void* y,
                        int64_t rows, int32_t C, int32_t apply_silu, void* stream);
int32_t vt_op_groupnorm(int32_t precision, const void* x, const float* gamma, const float* beta, void* y,
                        int64_t frames, int64_t positions_per_frame, int32_t C, int32_t per_position,
                        int32_t apply_silu, void* workspace, int64_t workspace_bytes, void* stream);
/* per-frame single-head attention core: q,k,v,o channels-last [frames, tokens, C]; scale = C^-0.5.  Runs what the
 * model path runs: tcgen05 GEMMs in BF16 / EXACT_TC when tokens % 64 == 0 and C % 64 == 0, fp32 FMAs otherwise.
 * workspace: frames*tokens*(8*tokens + 24*C) + 65536 bytes is always enough. */
int32_t vt_op_attention(int32_t precision, const void* q, const void* k, const void* v, void* o, int32_t frames,
                        int32_t tokens, int32_t C, void* workspace, int64_t workspace_bytes, void* stream);
int32_t vt_op_fsq(const float* h, int32_t d, const int32_t* levels, int64_t positions_per_batch, int32_t B,
                  float* codes, int32_t* indices, void* stream);
int32_t vt_op_fsq_indices_to_codes(const int32_t* indices, int32_t d, const int32_t* levels,
                                   int64_t positions_per_batch, int32_t B, float* codes, void* stream);
int32_t vt_op_kl(const float* h, const float* noise, int32_t zc, int64_t positions_per_batch, int32_t B, int32_t sample,
                 float* z, float* kl_loss, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDTOK_B200_H */

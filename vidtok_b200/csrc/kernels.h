// Host-callable launchers of the vidtok_b200 kernels (internal; the public surface is include/vidtok_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace vt {

// DT_SPLIT: fp32-class activations stored as two fp16 planes side by side in the channel dimension,
// [..., hi(C) | lo(C)] with hi = fp16(v), lo = fp16(v - hi) (common.cuh: split_store; the operand format of the EXACT_TC mode).
// One logical element = 4 bytes; pointers to such tensors are bf16*, strides in ConvP count bf16 elements.
enum DType { DT_F32 = 0, DT_BF16 = 1, DT_SPLIT = 2 };
inline size_t dtype_size(DType t) { return t == DT_BF16 ? 2 : 4; }

void prof_start();
void prof_set_detail(bool on);
int prof_stop(char* buf, int cap);

// conv_simt.cu
cudaError_t launch_conv_simt(const ConvP& p, DType tin, DType tout, DType tres, const void* x, const float* w_kn,
                             void* out, cudaStream_t s);
cudaError_t launch_gemm_simt(DType ta, DType tb, DType tc, const void* A, const void* B, void* C, int M, int N, int K,
                             long long lda, long long sbn, long long sbk, long long ldc, int batch, long long bsA,
                             long long bsB, long long bsC, float scale, cudaStream_t s);

// elementwise.cu
cudaError_t launch_layernorm(DType t, const void* x, const float* gamma, const float* beta, void* y, long long rows,
                             int C, bool silu, bool exact, cudaStream_t s);
// Element counts / strides of the data-movement launchers (upsample_nearest, cache_update, copy_frames) are LOGICAL
// elements; a DT_SPLIT element is moved as one 4-byte unit (whole rows keep their hi|lo layout).
// stats: float2 [frames*32] scratch (per-frame mode)
cudaError_t launch_groupnorm(DType t, const void* x, const float* gamma, const float* beta, void* y, long long frames,
                             long long pos_per_frame, int C, bool per_position, bool silu, bool exact, float* stats,
                             cudaStream_t s);
cudaError_t launch_softmax_rows(DType tout, const float* S, void* P, long long rows, int N, cudaStream_t s);
cudaError_t launch_kl(const float* h, const float* noise, int zc, long long P, int B, bool sample, float* z,
                      float* kl_loss, double* scratch, cudaStream_t s);
cudaError_t launch_fsq(const float* h, int d, const int* levels_host, long long P, int B, float* codes, int* indices,
                       cudaStream_t s);
cudaError_t launch_fsq_indices_to_codes(const int* indices, int d, const int* levels_host, long long P, int B,
                                        float* codes, cudaStream_t s);
// weight repacking: w [Co][Ci][taps] (reference OIDHW flattened) -> [K = tap*Ci + ci][Co] fp32
cudaError_t launch_pack_w_kn(const float* w, float* out, int Co, int Ci, int taps, cudaStream_t s);
// -> [Co][K = tap*Ci + ci] bf16 (K-major rows for the tcgen05 B operand)
// wscale != 0: split rows [hi(Kpad) | lo(Kpad)] (fp16 planes, DT_SPLIT operand format) of w * wscale (see split_weight_scale)
cudaError_t launch_pack_w_nk_bf16(const float* w, bf16* out, int Co, int Co_pad, int Ci, int taps, int Kpad, cudaStream_t s,
                                  float wscale = 0.f);
cudaError_t launch_absmax(const float* w, long long n, float* out, cudaStream_t s);
// Power-of-two scale for the split (fp16 hi|lo) copy of a weight tensor with the given max |w|: largest 2^s with
// 2^s * maxabs * headroom < 4096 (headroom: phase-collapsed weights sum up to 4 taps), so that values at 1e-4 of the maximum
// still have a normal fp16 lo plane.  The conv epilogue multiplies the accumulator by 1 / scale.
inline float split_weight_scale(float maxabs, float headroom = 1.0f) {
  if (!(maxabs > 0.f)) return 1.0f;
  float sc = 1.0f;
  while (sc * maxabs * headroom < 2048.0f && sc < 1.0e30f) sc *= 2.0f;
  while (sc * maxabs * headroom >= 4096.0f && sc > 1.0e-30f) sc *= 0.5f;
  return sc;
}
// phase-collapsed weights of "nearest-2x upsample then conv": original taps (a,b,c) of a kt x kh x kw kernel are
// summed into tap (mt[a], mh[b], mw[c]) of a kt2 x kh2 x kw2 kernel; output [Co_pad][kt2*kh2*kw2*Ci] bf16
cudaError_t launch_pack_w_collapsed(const float* w, bf16* out, int Co, int Co_pad, int Ci, int kt, int kh, int kw,
                                    const int* mt, const int* mh, const int* mw, int kt2, int kh2, int kw2, cudaStream_t s,
                                    float wscale = 0.f);
// trilinear (align_corners=False) 2x upsampling along T of channels-last x [B,T,HWC] -> [B,2T,HWC]
cudaError_t launch_time_interp2x(DType t, const void* x, void* y, int B, int T, long long hw, int C, cudaStream_t s);
cudaError_t launch_upsample_nearest(DType t, const void* x, void* y, int B, int T, int H, int W, int C, int ut, int uh,
                                    int uw, cudaStream_t s);
cudaError_t launch_cache_update(DType t, const void* x, const void* old_cache, void* new_cache, int B, int Tc, int P,
                                int off, bool first, long long frame_elems, long long x_bs, cudaStream_t s);
cudaError_t launch_ncdhw_to_cl(DType t, const float* x, void* y, int B, int C, int T, int H, int W, int t_rep,
                               cudaStream_t s);
cudaError_t launch_copy_frames(DType t, const void* src, void* dst, int B, long long src_bs, long long dst_bs,
                               long long n_per_batch, cudaStream_t s);

// video I/O adjacent steps: decoded uint8 frames [T,Hs,Ws,C] -> cropped normalised clip fp32 [C,T,H,W]; clip -> uint8 frames
cudaError_t launch_u8_frames_to_clip(const uint8_t* src, float* dst, int T, int Hs, int Ws, int C, int h0, int w0, int H, int W,
                                    cudaStream_t s);
cudaError_t launch_clip_to_u8_frames(const float* src, uint8_t* dst, int C, int T, int H, int W, cudaStream_t s);
// hi|lo split rows [rows][hi(C) | lo(C)] <-> fp32 rows [rows][C]
cudaError_t launch_split_to_f32(const bf16* x, float* y, long long rows, int C, cudaStream_t s);
cudaError_t launch_f32_to_split(const float* x, bf16* y, long long rows, int C, cudaStream_t s);

// conv_tc.cu (tcgen05 / TMA implicit GEMM)
// LayerNorm(+SiLU) of the output row fused into the conv epilogue (the row is complete in TMEM when Cout <= 256):
// mode 1: out := act(LN(v));  mode 2: out := v, out2 := act(LN(v))   (out2 uses the strides of out)
struct TcLnFusion {
  int mode = 0;
  bool silu = true;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  void* out2 = nullptr;
};
// Regularizer fused into the epilogue of the encoder's conv_out (fp32 heads, Cout <= 32: the thread that owns an output
// position holds all of its channels): KL reparameterisation (distributions.py:8-18) or FSQ bound/round/index
// (regularizers.py:153-178), written straight to the caller's z / indices tensors ([B,zc,T,H,W] / [B,T,H,W]).
struct TcRegFusion {
  int mode = 0;                  // 0 none, 1 KL, 2 FSQ
  int zc = 0;
  int sample = 1;                // KL: z = mean + std * noise (else the mode)
  const float* noise = nullptr;  // KL, [B,zc,T,H,W]
  float* z = nullptr;
  int* indices = nullptr;        // FSQ (may be null)
  double* kl_acc = nullptr;      // KL: sum over all elements of mean^2 + var - 1 - logvar (cleared by the caller)
  int fsq_levels[VT_MAX_FSQ] = {0};
};
bool conv_tc_can_fuse_ln(const ConvP& p);
// planning = true: geometry-only answer (workspace dry runs: no device pointers, possibly no driver)
bool conv_tc_supported(const ConvP& p, DType tout, bool planning = false);
// out may be null when a regularizer epilogue consumes the result (reg != nullptr)
cudaError_t launch_conv_tc(const ConvP& p, const bf16* x, const bf16* w_nk, int Kpad, void* out, DType tout, cudaStream_t s,
                           int w_batches = 1, long long w_batch_stride = 0, const TcLnFusion* ln = nullptr,
                           const TcRegFusion* reg = nullptr);
cudaError_t launch_kl_clear(double* scratch, cudaStream_t s);
cudaError_t launch_kl_finish(const double* scratch, int B, float* kl_loss, cudaStream_t s);
// decoder head through per-tap partial outputs (see elementwise.cu)
cudaError_t launch_tap_planes_gather(const bf16* P, const float* bias, float* out, int B, int Ti, int H, int W, int NP, int Co,
                                     int to_off, cudaStream_t s, int pt = 2);
cudaError_t launch_pack_w_tap_planes(const float* w, bf16* out, int Co, int Ci, int NP, cudaStream_t s);
// x [batch][rows][cols] -> y [batch][cols][rows] (bf16), rows and cols multiples of 32
cudaError_t launch_transpose_bf16(const bf16* x, bf16* y, int batch, int rows, int cols, cudaStream_t s, bool split = false);
const char* conv_tc_last_error();
void conv_tc_set_pair(bool on);
int conv_tc_cluster_query(int smem, char* msg, int cap);

// tblock_tc.cu: fused ResnetCausalBlock1D (k311 conv -> LayerNorm -> SiLU -> k311 conv + residual) for C = 128, v1.0 padding
bool tblock_tc_supported(int B, int T, int H, int W, int C, bool planning = false);
cudaError_t launch_tblock_tc(const bf16* n1, const bf16* x, const bf16* w1, const float* bias1, const float* gamma2,
                             const float* beta2, const bf16* w2, const float* bias2, bf16* out, bf16* out2,
                             const float* gamma_out, const float* beta_out, bool out_silu, int B, int T, int H, int W,
                             cudaStream_t s);
const char* tblock_tc_last_error();

// conv_stem.cu (thread-built im2col A tile + tcgen05 for the Cin=3 stem)
bool conv_stem_supported(const ConvP& p);
cudaError_t launch_conv_stem(const ConvP& p, const float* x, const bf16* wpk, bf16* out, cudaStream_t s);
cudaError_t launch_stem_cache_update(const float* x, float* cache, int B, int Ci, int T, int t_rep, int H, int W, cudaStream_t s);

}  // namespace vt

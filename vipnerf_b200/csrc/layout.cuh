// Packed-weight layout and network constants shared by every kernel of the render path.
//
// The network is the reference's MLP (src/models/VipNeRF01.py:451-492) at the one shape every shipped
// config uses: D=8, W=256, skip after layer 4, L_pts=10 (63-d encoding), L_view=4 (27-d encoding),
// view-dependent rgb + visibility head.  It is evaluated as ten "matrix layers" (M0..M9):
//
//   M0      pts_linears.0        K= 64 (63 encoding cols + 1 zero)          N=256  ReLU
//   M1..M4  pts_linears.1..4     K=256                                       N=256  ReLU
//   M5      pts_linears.5        K=320 (64 encoding cols first, then 256 h)  N=256  ReLU   (:543-544)
//   M6,M7   pts_linears.6,7      K=256                                       N=256  ReLU
//   M8      feature_linear       K=256                                       N=256  (no activation, :564)
//   M9      views_linears.0      K=256 (feature columns only)                N=128  ReLU after adding the
//                                                       view-direction columns' contribution in fp32 (:576-580)
// plus three small fp32 heads done on CUDA cores: sigma (pts_output_linear, :546-553), the 27
// view-encoding columns of views_linears.0, and views_output_linear (128->4, :582-594).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace vipnerf {

constexpr int kWidth = 256;
constexpr int kHalfWidth = 128;
constexpr int kLPts = 10;
constexpr int kLView = 4;
constexpr int kEncPts = 3 + 6 * kLPts;    // 63
constexpr int kEncPtsPad = 64;
constexpr int kEncView = 3 + 6 * kLView;  // 27
constexpr int kNumMatLayers = 10;
// Tensor-core builds carry one more chunk image behind M9, "layer" 10: the 27 view-direction columns of
// views_linears.0 (+ its bias in column 31, facing a constant-one input column), K = 32, N = 128.  The fused kernel
// multiplies it with the per-SAMPLE direction encodings of the secondary views (visibility2, VipNeRF01.py:527-530);
// the primary view's direction is per ray and stays an fp32 CUDA-core term.
constexpr int kViewChunkLayer = 10;
constexpr int kNumTcLayers = 11;

__host__ __device__ constexpr int layer_k(int l) { return l == 0 ? 64 : (l == 5 ? 320 : (l == 10 ? 32 : 256)); }
__host__ __device__ constexpr int layer_n(int l) { return l >= 9 ? 128 : 256; }

// ---- small fp32 parameters (first region of every packed buffer), offsets in floats
constexpr int kOffBias = 0;                          // [9][256]: M0..M8 biases (M8 = feature_linear.bias)
constexpr int kOffBiasViews = kOffBias + 9 * 256;    // [128]  views_linears.0.bias
constexpr int kOffWSigma = kOffBiasViews + 128;      // [256]  pts_output_linear.weight
constexpr int kOffBSigma = kOffWSigma + 256;         // [4]    pts_output_linear.bias, max |pts_output_linear.weight|, pad
constexpr int kOffWViewDir = kOffBSigma + 4;         // [27][128] views_linears.0.weight[:, 256+j] transposed
constexpr int kOffWOut = kOffWViewDir + 27 * 128;    // [128][4]  views_output_linear.weight transposed
constexpr int kOffBOut = kOffWOut + 128 * 4;         // [4]    views_output_linear.bias
// tensor-core kernels (feature_linear folded into views_linears.0, see below): views_linears.0.bias +
// views_linears.0.weight[:, :256] . feature_linear.bias
constexpr int kOffBiasViewsFused = kOffBOut + 4;     // [128]
constexpr int kSmallFloats = kOffBiasViewsFused + 128;
constexpr int kSmallBytes = ((kSmallFloats * 4 + 1023) / 1024) * 1024;  // keep the big region 1 KiB aligned

// ---- fp32 big region: per matrix layer the transposed weight Wt[k][n] (k = A column), floats
__host__ __device__ constexpr int fp32_layer_offset(int l) {
  int off = 0;
  for (int i = 0; i < l; ++i) off += layer_k(i) * layer_n(i);
  return off;
}
constexpr int kFp32BigFloats = fp32_layer_offset(kNumMatLayers);  // 589,824
// ---- fp32 backward region (training, behind the forward region): the matrices of the backward-data chain
// dX = dY . W, i.e. the nn.Linear weights in their original [out][in] orientation restricted to the 256 hidden /
// feature input columns, in the order the chain consumes them: views_linears.0[:, :256] (128 rows),
// feature_linear, pts_linears.7 .. pts_linears.1 (pts_linears.5: columns 63..318).
constexpr int kBwdOffViews = 0;
constexpr int kBwdOffFeature = kBwdOffViews + 128 * 256;
constexpr int kBwdOffTrunk = kBwdOffFeature + 256 * 256;
constexpr int kBwdOffEnc0 = kBwdOffTrunk + 7 * 256 * 256;         // pts_linears.0 padded to [256][64] (column 63 = 0)
constexpr int kBwdOffEnc5 = kBwdOffEnc0 + 256 * 64;               // pts_linears.5[:, :63] padded to [256][64]
constexpr int kFp32BwdFloats = kBwdOffEnc5 + 256 * 64;            // 589,824; the two encoding blocks serve the
                                                                  // tensor-core forward, whose B operands are [out][in]

// ---- fp16 mirror (training, behind the backward region): the forward and backward regions once more as fp16, element
// for element - the B operands of the fp16 training mode (VIPNERF_FLAG_TRAIN_F16, tcgen05 kind::f16)
constexpr int kF16MirrorHalves = kFp32BigFloats + kFp32BwdFloats;
// ... followed by the 27 view-direction columns of views_linears.0 as an fp16 [128][64] matrix (columns 27..63 zero): the
// second operand pair of the fp16 forward's views layer, whose A rows are the 64-column direction encodings
constexpr int kF16ViewDirHalves = 128 * 64;
constexpr size_t kFp32PackBytes = (size_t)kSmallBytes + (size_t)(kFp32BigFloats + kFp32BwdFloats) * 4 +
                                  (size_t)(kF16MirrorHalves + kF16ViewDirHalves) * 2;

// ---- tensor-core big region: "chunk images".  One chunk = ALL output rows (n) of a layer x 32 k-columns of
// bf16 in the canonical K-major SWIZZLE_64B shared-memory layout tcgen05.mma reads (64-byte rows, 8-row / 512-byte
// swizzle atoms):
//   byte(n, k) = n*64 + ((((k & 31) >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7)*2
// so a chunk is 16 KiB for the 256-row layers (two N=256 MMAs of K=16) and 8 KiB for M9 (N=128).  Chunks are
// stored in the order the kernel consumes them: layer, then k; in BF16X3 mode every chunk is followed by its
// "lo" image (bf16(w - float(bf16(w)))).
constexpr int kChunkK = 32;
constexpr int kChunkBytes = 16384;   // ring stage size (largest chunk)
__host__ __device__ constexpr int layer_chunks(int l) { return layer_k(l) / kChunkK; }
// The bias is added by the tensor core as well: the encoding buffer's pad column (A column 63 of M0 and of the
// first k-block of M5) holds the constant 1.0, so for M0 / M5 the bias is simply weight column 63; every other
// 256-wide layer gets one extra "bias chunk" whose only non-zero column (31, facing encoding column 63) is the
// bias, consumed by ONE K=16 MMA against the last 16 encoding columns.  M9's bias travels with the
// view-direction term (fp32, per ray).  In bf16 mode the bias is therefore rounded to bf16; in BF16X3 mode the
// lo image carries its residual.
//
// feature_linear (M8) has NO activation (VipNeRF01.py:564) and its output feeds only views_linears.0 (:576-579; the
// `feature` tensor itself is dropped, :533-534), so on the tensor path the two are ONE layer:
//   views_linears.0[:, :256] . (W8 h7 + b8) = (Wv_f W8) h7 + Wv_f b8
// The packer forms Wv_f W8 [128 x 256] in double precision and stores it as "M9"; Wv_f b8 joins the views bias
// (kOffBiasViewsFused).  The tensor-core stream therefore has no M8 chunks: 2,112 of 18,432 tensor cycles per tile and
// one epilogue pass less.  (The fp32 kernels and the training path keep the two layers apart.)
__host__ __device__ constexpr bool layer_has_bias_chunk(int l) { return l != 0 && l != 5 && l < 8; }
__host__ __device__ constexpr int layer_stream_chunks(int l) {
  return l == 8 ? 0 : layer_chunks(l) + (layer_has_bias_chunk(l) ? 1 : 0);
}
__host__ __device__ constexpr int layer_chunk_bytes(int l) { return layer_n(l) * kChunkK * 2; }
__host__ __device__ constexpr int tc_layer_byte_offset(int l) {    // of the hi image set
  int off = 0;
  for (int i = 0; i < l; ++i) off += layer_stream_chunks(i) * layer_chunk_bytes(i);
  return off;
}
constexpr int kTcBigBytes = tc_layer_byte_offset(kNumTcLayers);    // 1,155,072 (68 weight + 6 bias chunks + the view-direction chunk)

// Maps A column k of matrix layer l to the source weight column (or -1 for a zero pad column).
__host__ __device__ inline int source_col(int l, int k) {
  if (l == 0) return k < kEncPts ? k : -1;
  if (l == 5) return k < kEncPts ? k : (k == 63 ? -1 : k - 1);  // 64 + j -> 63 + j
  return k;
}
// Index into params[24] of the bias of matrix layer l (M8 = feature_linear, M9 = views_linears.0).
__host__ __device__ inline int bias_param(int l) { return l < 8 ? 2 * l + 1 : (l == 8 ? 21 : 17); }
// Source tensor (index into params[24]) and its in-features of matrix layer l.
__host__ __device__ inline int source_param(int l) { return l < 8 ? 2 * l : (l == 8 ? 20 : 16); }
__host__ __device__ inline int source_in_features(int l) {
  return l == 0 ? kEncPts : (l == 5 ? kWidth + kEncPts : (l == 9 ? kWidth + kEncView : kWidth));
}

}  // namespace vipnerf

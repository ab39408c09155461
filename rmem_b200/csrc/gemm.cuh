// Generic t16 GEMM / implicit-GEMM convolution with a fused epilogue.
//   C[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )
// A is either a row-major matrix (linear mode) or an NHWC activation map gathered on the fly
// (conv mode, K ordered (ky,kx,ci)); B is always the [N,K] row-major weight (K-major both sides).
// This is the general-purpose workhorse for the ResNet-50 encoder, FPN decoder and the GPM
// projections; the attention contractions have their own kernels.
#pragma once
#include "common.cuh"

namespace rmem {

struct GemmParams {
  const t16* A = nullptr;
  long long lda = 0;
  const t16* B = nullptr;
  long long ldb = 0;
  int M = 0, N = 0, K = 0;
  // conv mode (A = NHWC [Hin,Win,Cin]); output pixel m -> (m / Wout, m % Wout)
  // conv = 2, "stem" mode (the 7x7 stride-2 first convolution on 3 -> 8 channels, tcgen05 kernel only): A is the
  // zero-PADDED image [Hin][Win][8] (3 pad rows / pixels before the first real one, >= 5 pad pixels after the last), one
  // k-block per window row: the 64 contiguous elements (8 pixels x 8 channels) starting at padded pixel (oy*2 + ky,
  // ox*2); B = [N][kw rows][8 pixels][8 channels] with zero weights for the 8th pixel, K = kw * 64.
  int conv = 0, Hin = 0, Win = 0, Cin = 0, Wout = 0, kw = 1, stride = 1, pad = 0;
  // conv modes, tcgen05 kernel only: nimg images of [Hin,Win,Cin] stacked in A and [Hout*Wout, N] blocks stacked in C / res
  // (M = nimg * Hout * Wout); padding and tile edges are handled per image.
  int nimg = 1;
  // epilogue
  float alpha = 1.f;
  const float* bias = nullptr;  // [N] (or [M] if bias_m)
  int bias_m = 0;
  int act = ACT_NONE;           // applied to columns >= act_from
  int act_from = 0;
  const t16* res = nullptr;    // added before the activation
  long long ldr = 0;
  const t16* gate = nullptr;   // multiplied after the activation
  long long ldg = 0;
  int accumulate = 0;           // C += (fp32 outputs only)
  void* C = nullptr;
  long long ldc = 0;
  int c_fp32 = 0;
  void* C2 = nullptr;           // columns >= n_split go to C2[:, col - n_split]
  long long ldc2 = 0;
  int c2_fp32 = 0;
  int n_split = 1 << 30;
  // N % 32 != 0 only: columns N .. round_up(N,32)-1 of C may be written with padding values (no residual / gate /
  // per-column bias in that case).  Lets the tcgen05 kernel store whole 32-column chunks.
  int pad_n_ok = 0;
  // tcgen05 kernel only: GroupNorm statistics of the (t16, single-destination, un-activated) output over gn_groups
  // channel groups, written to gn_stats[2 * group] = (sum, sum of squares) in double -- the scratch block of
  // groupnorm_t16 (ops.cuh, kGnScratchDoubles doubles), so that groupnorm_apply_t16 can follow without a statistics pass.
  double* gn_stats = nullptr;
  int gn_groups = 0;
  // batched mode (legacy kernel only): blockIdx.z selects a problem; element strides between problems
  int batch = 1;
  long long sA = 0, sB = 0, sC = 0;
};

// Launches the best kernel for the shape: the tcgen05/TMA kernel (gemm_tc.cu) whenever its alignment rules hold,
// else the legacy mma.sync kernel (gemm.cu).  Returns RMEM_OK or an error code.
int gemm_launch(const GemmParams& p, cudaStream_t stream);

// tcgen05 + TMA implementation (gemm_tc.cu): K % 64 == 0, 16-byte aligned operands / outputs, conv Cin % 64 == 0.
bool gemm_tc_supported(const GemmParams& p);
int gemm_tc_launch(const GemmParams& p, cudaStream_t stream);
int gemm_legacy_launch(const GemmParams& p, cudaStream_t stream);
// Debug: clock64 event trace of the first 64 CTAs of every gemm_tc launch ([cta][8] long long); nullptr disables.
int gemm_tc_set_trace(long long* dev_buf);
// Tuning aids, thread-local (tools/tune_gemm.py): force the tile width (0 | 64 | 128 | 256), split-K factor (0 = auto;
// BN = 64 only) and ring depth (0 = auto) of every following tcgen05 launch; log the shapes launched into a HOST buffer
// of 16-int records (see gemm_tc.cu).
int gemm_tc_set_force(int bn, int splitk, int stages);
int gemm_tc_set_log(int* host_buf, int cap_records);
int gemm_tc_log_count();
// 0 = auto (tcgen05 when supported), 1 = force the legacy mma.sync kernel (parity tests).  Thread-local.
int& gemm_impl_switch();

}  // namespace rmem

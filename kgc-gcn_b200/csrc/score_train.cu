// Training-time 1-N scoring, backward through the sigmoid (autograd of reference model.py:177-179).
//
// The forward (kgc_score_1n_fwd, gemm_tf32x3.cu) produced pred[B, N] = sigmoid(X E^T + bias).  The two gradient GEMMs
// that follow contract over the batch (d_E = d_logit^T X) and over the entities (d_X = d_logit E); both tensor-core
// kernels stream their big operand row-major with the ENTITY index as the row, so this kernel writes the logit gradient
// transposed, d_logitT[N, ldt], in the same pass that applies p (1 - p) and reduces the bias gradient:
//   d_logitT[n, b] = d_pred[b, n] * pred[b, n] * (1 - pred[b, n]),     d_bias[n] = sum_b d_logitT[n, b]
// One CTA owns 32 entities and walks the batch in tiles of 32 queries: coalesced 128-byte reads along n, a padded
// shared-memory transpose, coalesced 128-byte writes along b.  The bias sum is a fixed shuffle tree per tile, tiles added
// in order: deterministic, no atomics.
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kThreadsST = 256;

__global__ void __launch_bounds__(kThreadsST)
score_1n_bwd_logit_kernel(const float* __restrict__ d_pred, int64_t ld_dp, const float* __restrict__ pred, int64_t ld_p,
                          int64_t n_ent, int B, int ldt, float* __restrict__ d_logitT, float* __restrict__ d_bias) {
  __shared__ float tile[32][33];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;      // 8 warps x 4 rows
  const int64_t n0 = (int64_t)blockIdx.x * 32;
  float bias_acc[4] = {0.f, 0.f, 0.f, 0.f};                         // lane 0 of each warp: rows warp * 4 + i
  for (int b0 = 0; b0 < ldt; b0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + warp * 4 + i;
      const int64_t n = n0 + lane;
      float g = 0.f;
      if (b < B && n < n_ent) {
        const float p = pred[(int64_t)b * ld_p + n];
        g = d_pred[(int64_t)b * ld_dp + n] * p * (1.f - p);
      }
      tile[warp * 4 + i][lane] = g;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = warp * 4 + i;                                   // entity n0 + r, query b0 + lane
      const float g = tile[lane][r];
      if (n0 + r < n_ent && b0 + lane < ldt) d_logitT[(n0 + r) * ldt + b0 + lane] = g;   // pad columns [B, ldt) get zeros
      float s = g;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      bias_acc[i] += s;
    }
    __syncthreads();
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + warp * 4 + i < n_ent) d_bias[n0 + warp * 4 + i] = bias_acc[i];
  }
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int kgc_score_1n_bwd_logit(const float* d_pred, int64_t ld_dp, const float* pred, int64_t ld_p, int64_t n_ent,
                                      int32_t B, int32_t ldt, float* d_logitT, float* d_bias, void* stream) {
  KGC_REQUIRE(n_ent > 0 && B > 0 && ldt >= B && ld_dp >= n_ent && ld_p >= n_ent, "bad sizes");
  KGC_REQUIRE(d_pred && pred && d_logitT && d_bias, "null buffer");
  score_1n_bwd_logit_kernel<<<(unsigned)ceil_div(n_ent, 32), kThreadsST, 0, as_stream(stream)>>>(
      d_pred, ld_dp, pred, ld_p, n_ent, B, ldt, d_logitT, d_bias);
  KGC_LAUNCH_CHECK();
  return 0;
}

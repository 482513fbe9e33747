// Weight-gradient reductions of the layer:  C[Ka, Nb] = A[M, Ka]^T @ B[M, Nb]   (M = node rows, Ka x Nb = 100 x 200)
//
// Autograd of model.py:116 w.r.t. the weights (dW_h = agg_h^T @ d_res_h).  The output is tiny and the contraction
// runs over the node rows, so a library GEMM falls back to split-K SIMT kernels at ~13 TFLOP/s (3 x 120 us of a
// 1 ms step, profiles/r01_launches_conv_step.md).  Here every CTA owns a slab of rows and keeps the WHOLE Ka x Nb
// partial product in registers (thread tile TI x TO), streaming the two row-major operands through a double-buffered
// cp.async shared-memory stage; the per-CTA partials are then added in CTA order - plain fp32 FMAs, deterministic,
// no atomics.  (A tcgen05 version needs MN-major operand descriptors and an in-kernel 3xTF32 split of BOTH operands;
// it is HBM-bound at ~8 us per product and is listed as the next step in DESIGN.md.)
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kThreadsTN = 256;
constexpr int kRowsPerStage = 16;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// thread (wi = warp id, lane): output rows i = wi * TI + a, output columns o = lane * TO + b
template <int TI, int TO>
__global__ void __launch_bounds__(kThreadsTN)
gemm_tn_partial_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int64_t M,
                       int Ka, int Nb, int64_t rows_per_cta, float* __restrict__ partial) {
  constexpr int KaP = 8 * TI, NbP = 32 * TO;
  extern __shared__ __align__(16) float sm[];
  float* sa = sm;                                         // [2][kRowsPerStage][KaP]
  float* sb = sm + 2 * kRowsPerStage * KaP;               // [2][kRowsPerStage][NbP]
  const int tid = threadIdx.x, wi = tid / 32, lane = tid % 32;
  for (int i = tid; i < 2 * kRowsPerStage * (KaP + NbP); i += kThreadsTN) sm[i] = 0.f;   // padding columns stay zero
  __syncthreads();
  const int64_t m0 = blockIdx.x * rows_per_cta;
  const int64_t m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  const int ka4 = Ka / 4, nb4 = Nb / 4;

  auto load_stage = [&](int buf, int64_t row0) {
    // kRowsPerStage rows of A and B as 16-byte vectors; rows past m1 are zero-filled by plain stores
    for (int v = tid; v < kRowsPerStage * (ka4 + nb4); v += kThreadsTN) {
      const int r = v / (ka4 + nb4), c = v % (ka4 + nb4);
      const int64_t m = row0 + r;
      float* dst = c < ka4 ? sa + (buf * kRowsPerStage + r) * KaP + c * 4 : sb + (buf * kRowsPerStage + r) * NbP + (c - ka4) * 4;
      if (m < m1) {
        const float* src = c < ka4 ? A + m * lda + c * 4 : B + m * ldb + (c - ka4) * 4;
        cp_async16(dst, src);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };

  float acc[TI][TO];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TO; ++b) acc[a][b] = 0.f;

  const int64_t n_stages = m1 > m0 ? (m1 - m0 + kRowsPerStage - 1) / kRowsPerStage : 0;
  if (n_stages > 0) load_stage(0, m0);
  for (int64_t s = 0; s < n_stages; ++s) {
    const int buf = (int)(s & 1);
    if (s + 1 < n_stages) {
      load_stage(buf ^ 1, m0 + (s + 1) * kRowsPerStage);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* pa = sa + buf * kRowsPerStage * KaP + wi * TI;
    const float* pb = sb + buf * kRowsPerStage * NbP + lane * TO;
#pragma unroll 4
    for (int r = 0; r < kRowsPerStage; ++r) {
      float av[TI], bv[TO];
#pragma unroll
      for (int a = 0; a < TI; ++a) av[a] = pa[r * KaP + a];          // warp-wide broadcast
#pragma unroll
      for (int b = 0; b < TO; ++b) bv[b] = pb[r * NbP + b];          // stride TO across lanes (odd TO: conflict-free)
#pragma unroll
      for (int a = 0; a < TI; ++a)
#pragma unroll
        for (int b = 0; b < TO; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* out = partial + (int64_t)blockIdx.x * Ka * Nb;
#pragma unroll
  for (int a = 0; a < TI; ++a) {
    const int i = wi * TI + a;
#pragma unroll
    for (int b = 0; b < TO; ++b) {
      const int o = lane * TO + b;
      if (i < Ka && o < Nb) out[i * Nb + o] = acc[a][b];
    }
  }
}

// C[e] = sum over the CTA partials in a FIXED order (deterministic): 8 part-lanes per element stride over the
// partials, then their 8 sums are added in lane order.  blockDim = (32 elements, 8 part-lanes).
__global__ void __launch_bounds__(256)
gemm_tn_reduce_kernel(const float* __restrict__ partial, int n_parts, int n_elem, float* __restrict__ C) {
  __shared__ float sm[8][33];
  const int e = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (e < n_elem)
    for (int g = threadIdx.y; g < n_parts; g += 8) s += partial[(int64_t)g * n_elem + e];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && e < n_elem) {
    float t = sm[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += sm[k][threadIdx.x];
    C[e] = t;
  }
}

inline int n_ctas(int64_t M) {
  int64_t g = ceil_div(M, 64);                 // at least 64 rows per CTA
  if (g > 2 * kNumSMs) g = 2 * kNumSMs;
  return (int)(g < 1 ? 1 : g);
}

template <int TI, int TO>
int launch_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int Ka, int Nb, float* C, float* ws,
              cudaStream_t st) {
  const int g = n_ctas(M);
  const int64_t rows_per_cta = ceil_div(M, g);
  const size_t smem = (size_t)2 * kRowsPerStage * (8 * TI + 32 * TO) * sizeof(float);
  auto kern = gemm_tn_partial_kernel<TI, TO>;
  static SmemAttrCache attr;                           // one per <TI, TO> instantiation
  KGC_CUDA_TRY(attr.ensure(kern, smem));
  kern<<<g, kThreadsTN, smem, st>>>(A, lda, B, ldb, M, Ka, Nb, rows_per_cta, ws);
  KGC_LAUNCH_CHECK();
  gemm_tn_reduce_kernel<<<(Ka * Nb + 31) / 32, dim3(32, 8), 0, st>>>(ws, g, Ka * Nb, C);
  KGC_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" size_t kgc_gemm_tn_workspace_bytes(int64_t M, int32_t Ka, int32_t Nb) {
  return (size_t)n_ctas(M) * Ka * Nb * sizeof(float);
}

extern "C" int kgc_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int32_t Ka, int32_t Nb,
                           float* C, void* workspace, size_t workspace_bytes, void* stream) {
  KGC_REQUIRE(M > 0 && Ka > 0 && Nb > 0 && Ka <= 128 && Nb <= 256, "supported: Ka <= 128, Nb <= 256");
  KGC_REQUIRE(Ka % 4 == 0 && Nb % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, "dimensions and leading dimensions must be multiples of 4");
  KGC_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0, "operands must be 16-byte aligned");
  KGC_REQUIRE(workspace && workspace_bytes >= kgc_gemm_tn_workspace_bytes(M, Ka, Nb), "workspace too small");
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  if (Ka <= 8 * 13 && Nb <= 32 * 7) return launch_tn<13, 7>(A, lda, B, ldb, M, Ka, Nb, C, ws, st);
  return launch_tn<16, 8>(A, lda, B, ldb, M, Ka, Nb, C, ws, st);
}

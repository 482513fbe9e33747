// K8: ConvE's one-input-channel 2-D convolution, forward and backward (SURVEY.md "next" row N2).
// Replaces reference model.py:166 (x = self.conv_e(x): nn.Conv2d(1, num_filter, (k, k), stride 1, padding 0) over the
// stacked [B, 1, 2 k_w, k_h] image of the (source, relation) embeddings) and its autograd.  At the reference defaults
// (B = 128, 200 filters, 7 x 7, 20 x 20 image -> 14 x 14 maps) this is 0.49 GFLOP per pass over a 20 MB feature map; cuDNN
// spends 32 us (implicit sgemm) + 13 us (layout conversions) on the forward, 36 us on the data gradient and 69 us on the
// weight gradient.  Direct kernels, fp32 FMA in a fixed order (deterministic, no atomics):
//   * forward : one CTA per image; filters (F k k floats) and the image in shared memory; a thread owns one output ROW of
//               one filter (OW accumulators in registers), rows of consecutive threads are consecutive in memory;
//   * d_x     : one CTA per image; d_y staged through shared memory in chunks of filters; thread = (input row, filter group),
//               W accumulators in registers, the filter groups are added in a fixed order through shared memory;
//   * d_w     : CTA = (32 filters) x (a group of images); thread = (filter, kernel row), K accumulators in registers across
//               the CTA's images; per-group partials are added in group order by a second small kernel.
// Layouts are torch's: x [B, 1, H, W], w [F, 1, K, K], y [B, F, OH, OW], all contiguous fp32.
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kThreadsCv = 256;     // d_w
constexpr int kThreadsImg = 1024;   // forward / d_x: one CTA per image, 32 warps to hide the shared-memory latency
constexpr int kDyChunk = 50;        // d_x: filters per shared-memory chunk of d_y
constexpr int kDwFilters = 32;      // d_w: filters per CTA
constexpr int kDwGroupsMax = 64;

template <int K, int W>
__global__ void __launch_bounds__(kThreadsImg)
conv1ch_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int F, int H,
                   float* __restrict__ y) {
  constexpr int OW = W - K + 1, WP = W + 1;
  extern __shared__ float sm[];
  float* w_s = sm;                      // [F][K*K]   (K*K odd: consecutive filters fall in different banks)
  float* x_s = sm + F * K * K;          // [H][W + 1]
  const int b = blockIdx.x, OH = H - K + 1;
  for (int i = threadIdx.x; i < F * K * K; i += kThreadsImg) w_s[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < H * W; i += kThreadsImg) x_s[(i / W) * WP + i % W] = __ldg(x + (int64_t)b * H * W + i);
  __syncthreads();
  float* yb = y + (int64_t)b * F * OH * OW;
  for (int item = threadIdx.x; item < F * OH; item += kThreadsImg) {
    const int f = item / OH, oy = item % OH;
    float acc[OW];
    const float b0 = bias != nullptr ? __ldg(bias + f) : 0.f;
#pragma unroll
    for (int ox = 0; ox < OW; ++ox) acc[ox] = b0;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      float xr[W];
#pragma unroll
      for (int j = 0; j < W; ++j) xr[j] = x_s[(oy + ky) * WP + j];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float wv = w_s[f * K * K + ky * K + kx];
#pragma unroll
        for (int ox = 0; ox < OW; ++ox) acc[ox] = fmaf(xr[ox + kx], wv, acc[ox]);
      }
    }
    float* out = yb + (int64_t)item * OW;
    if (OW % 2 == 0) {
#pragma unroll
      for (int ox = 0; ox < OW; ox += 2) *reinterpret_cast<float2*>(out + ox) = make_float2(acc[ox], acc[ox + 1]);
    } else {
#pragma unroll
      for (int ox = 0; ox < OW; ++ox) out[ox] = acc[ox];
    }
  }
}

// d_x[b][iy][ix] = sum_f sum_ky,kx d_y[b][f][iy - ky][ix - kx] * w[f][ky][kx]
// Thread = (filter group g, input row iy) with g FASTEST: the lanes of a warp share iy, so the row-validity test is warp-
// uniform (invalid kernel rows are skipped, not predicated: 30% of them at 20 / 14 / 7), consecutive lanes read consecutive
// filters - conflict-free with the odd filter strides K K of w_s and OH OW + 1 of dy_s.
template <int K, int W>
__global__ void __launch_bounds__(kThreadsImg)
conv1ch_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, int F, int H, float* __restrict__ dx) {
  constexpr int OW = W - K + 1;
  extern __shared__ float sm[];
  const int OH = H - K + 1, G = kThreadsImg / H, FS = OH * OW + 1;
  float* w_s = sm;                                  // [F][K*K]
  float* dy_s = w_s + F * K * K;                    // [kDyChunk][OH*OW + 1]
  float* red = dy_s + kDyChunk * FS;                // [G][H][W]
  const int b = blockIdx.x;
  const int g = threadIdx.x % G, iy = threadIdx.x / G;
  const bool work = iy < H;
  for (int i = threadIdx.x; i < F * K * K; i += kThreadsImg) w_s[i] = __ldg(w + i);
  float acc[W];
#pragma unroll
  for (int j = 0; j < W; ++j) acc[j] = 0.f;
  const float* dyb = dy + (int64_t)b * F * OH * OW;
  for (int c0 = 0; c0 < F; c0 += kDyChunk) {
    const int fc = F - c0 < kDyChunk ? F - c0 : kDyChunk;
    __syncthreads();                                // previous chunk consumed (and w_s written, first time)
    for (int i = threadIdx.x; i < fc * OH * OW; i += kThreadsImg)
      dy_s[(i / (OH * OW)) * FS + i % (OH * OW)] = __ldg(dyb + (int64_t)c0 * OH * OW + i);
    __syncthreads();
    if (work) {
      for (int fl = g; fl < fc; fl += G) {
        const int f = c0 + fl;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int oy = iy - ky;
          if (oy >= 0 && oy < OH) {
            float dr[OW];
#pragma unroll
            for (int ox = 0; ox < OW; ++ox) dr[ox] = dy_s[fl * FS + oy * OW + ox];
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
              const float wv = w_s[f * K * K + ky * K + kx];
#pragma unroll
              for (int ox = 0; ox < OW; ++ox) acc[ox + kx] = fmaf(dr[ox], wv, acc[ox + kx]);
            }
          }
        }
      }
    }
  }
  if (work) {
#pragma unroll
    for (int j = 0; j < W; ++j) red[(g * H + iy) * W + j] = acc[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * W; i += kThreadsImg) {
    float s = red[i];
    for (int q = 1; q < G; ++q) s += red[q * H * W + i];          // fixed order
    dx[(int64_t)b * H * W + i] = s;
  }
}

// partial[grp][f][ky][kx] = sum over the group's images of sum_oy,ox d_y[b][f][oy][ox] * x[b][oy + ky][ox + kx]
template <int K, int W>
__global__ void __launch_bounds__(kThreadsCv)
conv1ch_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int F, int H,
                          float* __restrict__ partial) {
  constexpr int OW = W - K + 1, WP = W + 1;
  extern __shared__ float sm[];
  const int OH = H - K + 1;
  float* dy_s = sm;                                 // [kDwFilters][OH][OW]
  float* x_s = dy_s + kDwFilters * OH * OW;         // [H][W + 1]
  const int f0 = blockIdx.x * kDwFilters, grp = blockIdx.y, n_grp = gridDim.y;
  const int fc = F - f0 < kDwFilters ? F - f0 : kDwFilters;
  const int fl = threadIdx.x / K, ky = threadIdx.x % K;
  const bool work = fl < fc;
  float acc[K];
#pragma unroll
  for (int kx = 0; kx < K; ++kx) acc[kx] = 0.f;
  for (int b = grp; b < B; b += n_grp) {
    __syncthreads();
    const float* src = dy + ((int64_t)b * F + f0) * OH * OW;
    for (int i = threadIdx.x; i < fc * OH * OW; i += kThreadsCv) dy_s[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < H * W; i += kThreadsCv) x_s[(i / W) * WP + i % W] = __ldg(x + (int64_t)b * H * W + i);
    __syncthreads();
    if (work) {
      for (int oy = 0; oy < OH; ++oy) {
        float dr[OW], xr[W];
#pragma unroll
        for (int ox = 0; ox < OW; ++ox) dr[ox] = dy_s[(fl * OH + oy) * OW + ox];
#pragma unroll
        for (int j = 0; j < W; ++j) xr[j] = x_s[(oy + ky) * WP + j];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
#pragma unroll
          for (int ox = 0; ox < OW; ++ox) acc[kx] = fmaf(dr[ox], xr[ox + kx], acc[kx]);
        }
      }
    }
  }
  if (work) {
    float* out = partial + ((int64_t)grp * F + f0 + fl) * K * K + ky * K;
#pragma unroll
    for (int kx = 0; kx < K; ++kx) out[kx] = acc[kx];
  }
}

__global__ void conv1ch_dw_reduce_kernel(const float* __restrict__ partial, int n_grp, int n, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = partial[i];
  for (int g = 1; g < n_grp; ++g) s += partial[(int64_t)g * n + i];   // fixed order
  dw[i] = s;
}

inline int dw_groups(int64_t B, int F) {
  const int chunks = (int)ceil_div(F, kDwFilters);
  int64_t g = 3 * kNumSMs / chunks;                 // three resident CTAs per SM
  if (g < 1) g = 1;
  if (g > B) g = B;
  if (g > kDwGroupsMax) g = kDwGroupsMax;
  return (int)g;
}

struct Shape { int F, K, H, W; };

inline size_t smem_fwd(const Shape& s) { return ((size_t)s.F * s.K * s.K + (size_t)s.H * (s.W + 1)) * 4; }
inline size_t smem_bwd_data(const Shape& s) {
  const int OH = s.H - s.K + 1, OW = s.W - s.K + 1;
  return ((size_t)s.F * s.K * s.K + (size_t)kDyChunk * (OH * OW + 1) + (size_t)(kThreadsImg / s.H) * s.H * s.W) * 4;
}
inline size_t smem_bwd_weight(const Shape& s) {
  const int OH = s.H - s.K + 1, OW = s.W - s.K + 1;
  return ((size_t)kDwFilters * OH * OW + (size_t)s.H * (s.W + 1)) * 4;
}

inline bool supported(const Shape& s) {
  if (!(s.W == 20 && (s.K == 3 || s.K == 5 || s.K == 7))) return false;
  if (s.H < s.K || s.H > 64 || s.F < 1) return false;
  if (s.K * kDwFilters > kThreadsCv) return false;
  return smem_fwd(s) <= 200 * 1024 && smem_bwd_data(s) <= 200 * 1024 && smem_bwd_weight(s) <= 200 * 1024;
}

#define KGC_CONV_DISPATCH(KV, ...)                               \
  switch (KV) {                                                  \
    case 3: { constexpr int K = 3, W = 20; __VA_ARGS__; } break; \
    case 5: { constexpr int K = 5, W = 20; __VA_ARGS__; } break; \
    default: { constexpr int K = 7, W = 20; __VA_ARGS__; } break; \
  }

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int kgc_conv1ch_supported(int32_t F, int32_t K, int32_t H, int32_t W) {
  return supported(Shape{F, K, H, W}) ? 1 : 0;
}

extern "C" size_t kgc_conv1ch_bwd_workspace_bytes(int64_t B, int32_t F, int32_t K) {
  return (size_t)dw_groups(B, F) * F * K * K * sizeof(float);
}

extern "C" int kgc_conv1ch_fwd(const float* x, const float* w, const float* bias, int64_t B, int32_t F, int32_t K,
                               int32_t H, int32_t W, float* y, void* stream) {
  const Shape s{F, K, H, W};
  KGC_REQUIRE(supported(s), "unsupported convolution shape (W = 20, K in {3,5,7}, filters must fit shared memory)");
  if (B == 0) return 0;
  const size_t sm = smem_fwd(s);
  KGC_CONV_DISPATCH(K, {
    KGC_CUDA_TRY(cudaFuncSetAttribute(conv1ch_fwd_kernel<K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    conv1ch_fwd_kernel<K, W><<<(unsigned)B, kThreadsImg, sm, as_stream(stream)>>>(x, w, bias, F, H, y);
  });
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_conv1ch_bwd(const float* dy, const float* x, const float* w, int64_t B, int32_t F, int32_t K, int32_t H,
                               int32_t W, float* dx, float* dw, void* workspace, void* stream) {
  const Shape s{F, K, H, W};
  KGC_REQUIRE(supported(s), "unsupported convolution shape (W = 20, K in {3,5,7}, filters must fit shared memory)");
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dx != nullptr) {
    const size_t sm = smem_bwd_data(s);
    KGC_CONV_DISPATCH(K, {
      KGC_CUDA_TRY(cudaFuncSetAttribute(conv1ch_bwd_data_kernel<K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      conv1ch_bwd_data_kernel<K, W><<<(unsigned)B, kThreadsImg, sm, st>>>(dy, w, F, H, dx);
    });
    KGC_LAUNCH_CHECK();
  }
  if (dw != nullptr) {
    KGC_REQUIRE(workspace != nullptr, "workspace of kgc_conv1ch_bwd_workspace_bytes() is required for dw");
    const int groups = dw_groups(B, F);
    const dim3 grid((unsigned)ceil_div(F, kDwFilters), (unsigned)groups);
    const size_t sm = smem_bwd_weight(s);
    KGC_CONV_DISPATCH(K, {
      KGC_CUDA_TRY(cudaFuncSetAttribute(conv1ch_bwd_weight_kernel<K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      conv1ch_bwd_weight_kernel<K, W><<<grid, kThreadsCv, sm, st>>>(dy, x, (int)B, F, H, (float*)workspace);
    });
    KGC_LAUNCH_CHECK();
    const int n = F * K * K;
    conv1ch_dw_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((const float*)workspace, groups, n, dw);
    KGC_LAUNCH_CHECK();
  }
  return 0;
}

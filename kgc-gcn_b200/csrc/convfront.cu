// K7: ConvE feature-map normalisation - BatchNorm2d + ReLU + feature dropout, forward and backward.
// Replaces reference model.py:168-170 (x = bn1(x); x = relu(x); x = feature_drop(x)) over the [B, C, H, W] output of the
// 7x7 convolution (B = 128, C = 200, H x W = 14 x 14: 5 M values, 20 MB) and its autograd.  cuDNN spends 27 us on the
// forward and ~60 us on the backward of this BatchNorm alone, plus separate ReLU / dropout passes; here the statistics
// are one pass (fp64 accumulation per thread, fixed-order block and cross-block reduction: deterministic), the
// normalisation + ReLU + dropout one elementwise pass, and the backward two passes.  The dropout keep mask is a
// counter-based Philox4x32-10 stream keyed by a device-resident seed (regenerated in the backward: no mask tensor).
//
// Layout: x[b][c][hw], hw contiguous, HW % 4 == 0 (float4 accesses); channel of float4 index i = (i / HW4) % C.
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kThreadsBN = 256;
constexpr int kSplitsBN = 4;             // CTAs per channel in the reduction passes

struct BnArgs {
  const float4* x;          // conv output
  const float4* dy;         // backward: upstream gradient (of the dropout output)
  const float* stats;       // [2][C]: mean, rstd
  const float* gamma;
  const float* beta;
  const int64_t* seed;      // NULL: no dropout
  uint32_t drop_thr;
  float keep_scale;
  int B, C, HW4;
  int relu;                 // 1: ReLU after the normalisation (bn1), 0: plain BatchNorm (bn0)
};

// y = dropout(relu((x - mean) * rstd * gamma + beta)); returns also the keep-scale (for the backward)
__device__ __forceinline__ float4 bn_relu_drop(const BnArgs& a, int64_t i, const float4 v, float4* keep) {
  const int c = (int)((i / a.HW4) % a.C);
  const float m = __ldg(a.stats + c), rs = __ldg(a.stats + a.C + c), g = __ldg(a.gamma + c), b = __ldg(a.beta + c);
  const float sc = rs * g, sh = b - m * sc;
  float4 y = make_float4(fmaf(v.x, sc, sh), fmaf(v.y, sc, sh), fmaf(v.z, sc, sh), fmaf(v.w, sc, sh));
  if (a.relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
  const float4 k = a.seed != nullptr ? philox_mask4(i, 7u, (uint64_t)__ldg(a.seed), a.drop_thr, a.keep_scale)
                                     : make_float4(1.f, 1.f, 1.f, 1.f);
  *keep = k;
  return make_float4(y.x * k.x, y.y * k.y, y.z * k.z, y.w * k.w);
}

// block-wide sum of two doubles in a fixed order: warp shuffle tree, then the warp sums in warp order
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0) { sm[2 * warp] = a; sm[2 * warp + 1] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0, sb = 0;
    for (int w = 0; w < kThreadsBN / 32; ++w) { sa += sm[2 * w]; sb += sm[2 * w + 1]; }
    sm[0] = sa; sm[1] = sb;
  }
  __syncthreads();
  a = sm[0]; b = sm[1];
}

// MODE 0: sum x, sum x^2 (forward statistics).  MODE 1: sum g, sum g * xhat with g = dy * keep * [bn > 0] (backward).
// grid (C, kSplitsBN): block (c, s) covers the samples b = s, s + kSplitsBN, ... of channel c.
template <int MODE>
__global__ void __launch_bounds__(kThreadsBN) bn2d_reduce_kernel(const BnArgs a, double* __restrict__ partials) {
  __shared__ double sm[2 * kThreadsBN / 32];
  const int c = blockIdx.x, s = blockIdx.y;
  double s0 = 0, s1 = 0;
  float m = 0.f, rs = 0.f, sc = 0.f, sh = 0.f;
  if (MODE == 1) {
    m = a.stats[c]; rs = a.stats[a.C + c];
    sc = rs * a.gamma[c]; sh = a.beta[c] - m * sc;
  }
  const int per = (a.B - s + kSplitsBN - 1) / kSplitsBN;                  // samples of this split
  for (int j = threadIdx.x; j < per * a.HW4; j += kThreadsBN) {
    const int b = s + (j / a.HW4) * kSplitsBN, q = j % a.HW4;
    const int64_t i = ((int64_t)b * a.C + c) * a.HW4 + q;
    const float4 v = __ldg(a.x + i);
    if (MODE == 0) {
      s0 += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      s1 += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      const float4 d = __ldg(a.dy + i);
      const float4 k = a.seed != nullptr ? philox_mask4(i, 7u, (uint64_t)__ldg(a.seed), a.drop_thr, a.keep_scale)
                                         : make_float4(1.f, 1.f, 1.f, 1.f);
      const bool nr = !a.relu;
      const float gx = nr || fmaf(v.x, sc, sh) > 0.f ? d.x * k.x : 0.f, gy = nr || fmaf(v.y, sc, sh) > 0.f ? d.y * k.y : 0.f;
      const float gz = nr || fmaf(v.z, sc, sh) > 0.f ? d.z * k.z : 0.f, gw = nr || fmaf(v.w, sc, sh) > 0.f ? d.w * k.w : 0.f;
      s0 += (double)gx + (double)gy + (double)gz + (double)gw;
      s1 += (double)gx * ((v.x - m) * rs) + (double)gy * ((v.y - m) * rs) + (double)gz * ((v.z - m) * rs) +
            (double)gw * ((v.w - m) * rs);
    }
  }
  block_sum2(s0, s1, sm);
  if (threadIdx.x == 0) {
    partials[((int64_t)c * kSplitsBN + s) * 2 + 0] = s0;
    partials[((int64_t)c * kSplitsBN + s) * 2 + 1] = s1;
  }
}

// forward: partial sums -> stats = {mean, rstd} (batch statistics in training, running statistics otherwise) and the
// running-statistics update of nn.BatchNorm2d (momentum, unbiased variance)
__global__ void bn2d_finalize_fwd_kernel(const double* __restrict__ partials, int C, double count, float eps, float momentum,
                                         int training, float* running_mean, float* running_var, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (!training) {
    stats[c] = running_mean[c];
    stats[C + c] = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
    return;
  }
  double s0 = 0, s1 = 0;
  for (int s = 0; s < kSplitsBN; ++s) { s0 += partials[((int64_t)c * kSplitsBN + s) * 2]; s1 += partials[((int64_t)c * kSplitsBN + s) * 2 + 1]; }
  const double mean = s0 / count;
  double var = s1 / count - mean * mean;
  if (var < 0) var = 0;
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (count / (count > 1 ? count - 1 : 1)));
  }
}

// backward: partial sums -> sums[2][C] (fp32: d_beta = sum g, d_gamma = sum g * xhat)
__global__ void bn2d_finalize_bwd_kernel(const double* __restrict__ partials, int C, float* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0, s1 = 0;
  for (int s = 0; s < kSplitsBN; ++s) { s0 += partials[((int64_t)c * kSplitsBN + s) * 2]; s1 += partials[((int64_t)c * kSplitsBN + s) * 2 + 1]; }
  sums[c] = (float)s0;
  sums[C + c] = (float)s1;
}

__global__ void __launch_bounds__(kThreadsBN) bn2d_fwd_kernel(const BnArgs a, int64_t total4, float4* __restrict__ y) {
  const int64_t i = blockIdx.x * (int64_t)kThreadsBN + threadIdx.x;
  if (i >= total4) return;
  float4 keep;
  y[i] = bn_relu_drop(a, i, __ldg(a.x + i), &keep);
}

// dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)) with batch statistics, gamma * rstd * g otherwise
__global__ void __launch_bounds__(kThreadsBN) bn2d_bwd_kernel(const BnArgs a, const float* __restrict__ sums, float inv_count,
                                                                int training, int64_t total4, float4* __restrict__ dx) {
  const int64_t i = blockIdx.x * (int64_t)kThreadsBN + threadIdx.x;
  if (i >= total4) return;
  const int c = (int)((i / a.HW4) % a.C);
  const float m = __ldg(a.stats + c), rs = __ldg(a.stats + a.C + c), g = __ldg(a.gamma + c), b = __ldg(a.beta + c);
  const float sc = rs * g, sh = b - m * sc;
  const float mg = training ? __ldg(sums + c) * inv_count : 0.f, mgx = training ? __ldg(sums + a.C + c) * inv_count : 0.f;
  const float4 v = __ldg(a.x + i), d = __ldg(a.dy + i);
  const float4 k = a.seed != nullptr ? philox_mask4(i, 7u, (uint64_t)__ldg(a.seed), a.drop_thr, a.keep_scale)
                                     : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 r;
  const bool nr = !a.relu;
  {
    const float gg = nr || fmaf(v.x, sc, sh) > 0.f ? d.x * k.x : 0.f;
    r.x = sc * (gg - mg - (v.x - m) * rs * mgx);
  }
  {
    const float gg = nr || fmaf(v.y, sc, sh) > 0.f ? d.y * k.y : 0.f;
    r.y = sc * (gg - mg - (v.y - m) * rs * mgx);
  }
  {
    const float gg = nr || fmaf(v.z, sc, sh) > 0.f ? d.z * k.z : 0.f;
    r.z = sc * (gg - mg - (v.z - m) * rs * mgx);
  }
  {
    const float gg = nr || fmaf(v.w, sc, sh) > 0.f ? d.w * k.w : 0.f;
    r.w = sc * (gg - mg - (v.w - m) * rs * mgx);
  }
  dx[i] = r;
}

int fill_args(BnArgs* a, const float* x, const float* dy, const float* stats, const float* gamma, const float* beta,
              const int64_t* seed, float drop_p, int64_t B, int32_t C, int32_t HW, int32_t relu) {
  KGC_REQUIRE(B > 0 && C > 0 && HW > 0 && HW % 4 == 0, "feature maps need H * W to be a multiple of 4");
  KGC_REQUIRE(B <= (1 << 24), "batch too large");
  KGC_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability must lie in [0, 1)");
  KGC_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "tensors must be 16-byte aligned");
  a->x = reinterpret_cast<const float4*>(x); a->dy = reinterpret_cast<const float4*>(dy);
  a->stats = stats; a->gamma = gamma; a->beta = beta;
  a->seed = drop_p > 0.f ? seed : nullptr;
  a->drop_thr = (uint32_t)((double)drop_p * 4294967296.0);
  a->keep_scale = 1.f / (1.f - drop_p);
  a->B = (int)B; a->C = C; a->HW4 = HW / 4; a->relu = relu;
  return 0;
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" size_t kgc_bn2d_partials_bytes(int32_t C) { return (size_t)C * kSplitsBN * 2 * sizeof(double); }

extern "C" int kgc_bn2d_relu_drop_fwd(const float* x, int64_t B, int32_t C, int32_t HW, const float* gamma, const float* beta,
                                      float* running_mean, float* running_var, float eps, float momentum, int32_t training,
                                      int32_t relu, const int64_t* seed, float drop_p, double* partials, float* stats, float* y,
                                      void* stream) {
  BnArgs a;
  if (fill_args(&a, x, nullptr, stats, gamma, beta, training ? seed : nullptr, training ? drop_p : 0.f, B, C, HW, relu)) return 1;
  KGC_REQUIRE(gamma && beta && stats && y && partials, "null buffer");
  KGC_REQUIRE(training || (running_mean && running_var), "evaluation mode needs running statistics");
  cudaStream_t st = as_stream(stream);
  if (training) {
    bn2d_reduce_kernel<0><<<dim3(C, kSplitsBN), kThreadsBN, 0, st>>>(a, partials);
    KGC_LAUNCH_CHECK();
  }
  bn2d_finalize_fwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(partials, C, (double)B * HW, eps, momentum, training, running_mean,
                                                            running_var, stats);
  KGC_LAUNCH_CHECK();
  const int64_t total4 = B * C * (int64_t)a.HW4;
  bn2d_fwd_kernel<<<(unsigned)ceil_div(total4, kThreadsBN), kThreadsBN, 0, st>>>(a, total4, reinterpret_cast<float4*>(y));
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_bn2d_relu_drop_bwd(const float* dy, const float* x, int64_t B, int32_t C, int32_t HW, const float* gamma,
                                      const float* beta, const float* stats, int32_t training, int32_t relu, const int64_t* seed,
                                      float drop_p, double* partials, float* sums, float* dx, void* stream) {
  BnArgs a;
  if (fill_args(&a, x, dy, stats, gamma, beta, training ? seed : nullptr, training ? drop_p : 0.f, B, C, HW, relu)) return 1;
  KGC_REQUIRE(dy && gamma && beta && stats && sums && dx && partials, "null buffer");
  cudaStream_t st = as_stream(stream);
  bn2d_reduce_kernel<1><<<dim3(C, kSplitsBN), kThreadsBN, 0, st>>>(a, partials);
  KGC_LAUNCH_CHECK();
  bn2d_finalize_bwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(partials, C, sums);
  KGC_LAUNCH_CHECK();
  const int64_t total4 = B * C * (int64_t)a.HW4;
  bn2d_bwd_kernel<<<(unsigned)ceil_div(total4, kThreadsBN), kThreadsBN, 0, st>>>(a, sums, (float)(1.0 / ((double)B * HW)), training,
                                                                               total4, reinterpret_cast<float4*>(dx));
  KGC_LAUNCH_CHECK();
  return 0;
}

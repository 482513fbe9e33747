// K2 / K3: relation-aware aggregation, forward and backward (HBM-bound segment reductions).
//
// Reference arithmetic replaced (model.py:111-118 + PyG aggr='add', model.py:50,99-100):
//   m_e = (x[src_e] * rel[type_e] * edge_emb[e]) @ W * norm_e ;  out[dst] += m_e
// restated as aggregate-then-transform (the matmul is linear):
//   agg[dst] = sum_e norm_e * x[src_e] (.) rel[type_e] (.) ee[e]      (this file)
//   res      = agg @ W                                                   (dense GEMM, host side)
//
// Layout: every row is D floats = D4 float4.  A GROUP of 8 lanes owns one work item (a run of
// <= 32 sorted edge records of one output row); lane g of the group owns float4 columns
// g, g+8, g+16, ... (NF of them), so the 8 lanes read 128 contiguous bytes per step.  A warp
// therefore works on 4 items at once, which keeps enough independent 128-bit loads in flight
// on the low-degree rows that dominate knowledge graphs.  Rows longer than one item are reduced
// through partial rows by kgc_rows_reduce in a fixed order: no float atomics anywhere, results
// are bit-reproducible run to run.
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kGroup = 8;
constexpr int kThreads = 256;
constexpr int kMaxNF = 8;   // D <= 8 * 8 * 4 = 256

__device__ __forceinline__ float4 mul3s(float s, const float4& a, const float4& b, const float4& c) {
  // s * ((a*b)*c), the reference's product order (model.py:115) followed by the norm (model.py:118)
  return make_float4(s * ((a.x * b.x) * c.x), s * ((a.y * b.y) * c.y), s * ((a.z * b.z) * c.z),
                     s * ((a.w * b.w) * c.w));
}
__device__ __forceinline__ void add4(float4& acc, const float4& v) {
  acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}
__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 scale4(float s, const float4& a) {
  return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
}

// ------------------------------------------------------------------------------------------------ forward
template <int NF>
__global__ void __launch_bounds__(kThreads)
agg_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ rel, const float4* __restrict__ ee,
               const kgc_edge_rec_t* __restrict__ rec, const kgc_item_t* __restrict__ items, int64_t n_items,
               float4* __restrict__ out_final, float4* __restrict__ out_part, int D4) {
  const int64_t item = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / kGroup;
  const int g = threadIdx.x % kGroup;
  if (item >= n_items) return;
  const int4 it = __ldg(reinterpret_cast<const int4*>(items + item));
  float4 acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);

  int e = it.x;
  for (; e + 1 < it.y; e += 2) {   // two edges per trip: 4*NF independent 128-bit loads in flight per lane
    const int4 r0 = ld_rec(rec + e), r1 = ld_rec(rec + e + 1);
    float4 a0[NF], b0[NF], a1[NF], b1[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        a0[f] = ld_stream(ee + (int64_t)r0.x * D4 + c);
        b0[f] = __ldg(x + (int64_t)r0.y * D4 + c);
        a1[f] = ld_stream(ee + (int64_t)r1.x * D4 + c);
        b1[f] = __ldg(x + (int64_t)r1.y * D4 + c);
      }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        add4(acc[f], mul3s(__int_as_float(r0.w), b0[f], __ldg(rel + (int64_t)r0.z * D4 + c), a0[f]));
        add4(acc[f], mul3s(__int_as_float(r1.w), b1[f], __ldg(rel + (int64_t)r1.z * D4 + c), a1[f]));
      }
    }
  }
  if (e < it.y) {
    const int4 r0 = ld_rec(rec + e);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        const float4 a = ld_stream(ee + (int64_t)r0.x * D4 + c);
        const float4 b = __ldg(x + (int64_t)r0.y * D4 + c);
        add4(acc[f], mul3s(__int_as_float(r0.w), b, __ldg(rel + (int64_t)r0.z * D4 + c), a));
      }
    }
  }
  float4* out = ((it.w & 1) ? out_final : out_part) + (int64_t)it.z * D4;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const int c = g + f * kGroup;
    if (c < D4) out[c] = acc[f];
  }
}

// ------------------------------------------------------------------------------------------------ backward (src rows)
template <int NF>
__global__ void __launch_bounds__(kThreads)
agg_bwd_src_kernel(const float4* __restrict__ x, const float4* __restrict__ rel, const float4* __restrict__ ee,
                   const float4* __restrict__ g3, const kgc_edge_rec_t* __restrict__ rec,
                   const kgc_item_t* __restrict__ items, int64_t n_items, int64_t n_nodes, int32_t half_edges,
                   const float4* __restrict__ loop_addend, float4* __restrict__ d_ee, float4* __restrict__ dx_final,
                   float4* __restrict__ dx_part, int D4) {
  const int64_t item = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / kGroup;
  const int g = threadIdx.x % kGroup;
  if (item >= n_items) return;
  const int4 it = __ldg(reinterpret_cast<const int4*>(items + item));
  const bool final_row = (it.w & 1) != 0;
  // flags >> 1 carries the source row j when the item writes a partial (out is then a slot id)
  const int64_t j = final_row ? it.z : (it.w >> 1);
  const int64_t plane = n_nodes * (int64_t)D4;
  float4 acc[NF], xj[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const int c = g + f * kGroup;
    acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D4) xj[f] = __ldg(x + j * D4 + c);
  }
  int e = it.x;
  for (; e + 1 < it.y; e += 2) {
    const int4 r0 = ld_rec(rec + e), r1 = ld_rec(rec + e + 1);
    const float4* gp0 = g3 + (r0.x >= half_edges ? plane : 0) + (int64_t)r0.y * D4;
    const float4* gp1 = g3 + (r1.x >= half_edges ? plane : 0) + (int64_t)r1.y * D4;
    float4 a0[NF], b0[NF], a1[NF], b1[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        a0[f] = ld_stream(ee + (int64_t)r0.x * D4 + c);
        b0[f] = __ldg(gp0 + c);
        a1[f] = ld_stream(ee + (int64_t)r1.x * D4 + c);
        b1[f] = __ldg(gp1 + c);
      }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        const float4 p0 = scale4(__int_as_float(r0.w), mul4(b0[f], __ldg(rel + (int64_t)r0.z * D4 + c)));
        const float4 p1 = scale4(__int_as_float(r1.w), mul4(b1[f], __ldg(rel + (int64_t)r1.z * D4 + c)));
        st_stream(d_ee + (int64_t)r0.x * D4 + c, mul4(p0, xj[f]));
        st_stream(d_ee + (int64_t)r1.x * D4 + c, mul4(p1, xj[f]));
        add4(acc[f], mul4(p0, a0[f]));
        add4(acc[f], mul4(p1, a1[f]));
      }
    }
  }
  if (e < it.y) {
    const int4 r0 = ld_rec(rec + e);
    const float4* gp0 = g3 + (r0.x >= half_edges ? plane : 0) + (int64_t)r0.y * D4;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        const float4 a = ld_stream(ee + (int64_t)r0.x * D4 + c);
        const float4 p0 = scale4(__int_as_float(r0.w), mul4(__ldg(gp0 + c), __ldg(rel + (int64_t)r0.z * D4 + c)));
        st_stream(d_ee + (int64_t)r0.x * D4 + c, mul4(p0, xj[f]));
        add4(acc[f], mul4(p0, a));
      }
    }
  }
  if (final_row) {
    float4* out = dx_final + j * D4;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        float4 v = acc[f];
        if (loop_addend != nullptr) add4(v, __ldg(loop_addend + j * D4 + c));     // self-loop term of d_x
        out[c] = v;
      }
    }
  } else {
    float4* out = dx_part + (int64_t)it.z * D4;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) out[c] = acc[f];
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward (type rows)
template <int NF>
__global__ void __launch_bounds__(kThreads)
agg_bwd_rel_kernel(const float4* __restrict__ x, const float4* __restrict__ ee, const float4* __restrict__ g3,
                   const kgc_edge_rec_t* __restrict__ rec, const kgc_item_t* __restrict__ items, int64_t n_items,
                   int64_t n_nodes, int32_t half_edges, float4* __restrict__ out_final,
                   float4* __restrict__ out_part, int D4) {
  const int64_t item = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / kGroup;
  const int g = threadIdx.x % kGroup;
  if (item >= n_items) return;
  const int4 it = __ldg(reinterpret_cast<const int4*>(items + item));
  const int64_t plane = n_nodes * (int64_t)D4;
  float4 acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
  int e = it.x;
  for (; e + 1 < it.y; e += 2) {
    const int4 r0 = ld_rec(rec + e), r1 = ld_rec(rec + e + 1);
    const float4* gp0 = g3 + (r0.x >= half_edges ? plane : 0) + (int64_t)r0.z * D4;
    const float4* gp1 = g3 + (r1.x >= half_edges ? plane : 0) + (int64_t)r1.z * D4;
    float4 a0[NF], a1[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        a0[f] = ld_stream(ee + (int64_t)r0.x * D4 + c);
        a1[f] = ld_stream(ee + (int64_t)r1.x * D4 + c);
      }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        add4(acc[f], mul3s(__int_as_float(r0.w), __ldg(gp0 + c), __ldg(x + (int64_t)r0.y * D4 + c), a0[f]));
        add4(acc[f], mul3s(__int_as_float(r1.w), __ldg(gp1 + c), __ldg(x + (int64_t)r1.y * D4 + c), a1[f]));
      }
    }
  }
  if (e < it.y) {
    const int4 r0 = ld_rec(rec + e);
    const float4* gp0 = g3 + (r0.x >= half_edges ? plane : 0) + (int64_t)r0.z * D4;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        const float4 a = ld_stream(ee + (int64_t)r0.x * D4 + c);
        add4(acc[f], mul3s(__int_as_float(r0.w), __ldg(gp0 + c), __ldg(x + (int64_t)r0.y * D4 + c), a));
      }
    }
  }
  float4* out = ((it.w & 1) ? out_final : out_part) + (int64_t)it.z * D4;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const int c = g + f * kGroup;
    if (c < D4) out[c] = acc[f];
  }
}

// ------------------------------------------------------------------------------------------------ higher levels
// One block per item; 32 groups stride over the item's partial rows, then group 0 adds the 32
// group sums in index order.  Fan-in up to 1024 rows per item keeps the level count at <= 3 even
// for a 9M-edge hub row.
template <int NF>
__global__ void __launch_bounds__(kThreads)
rows_reduce_kernel(const float4* __restrict__ part_in, const kgc_item_t* __restrict__ items,
                   float4* __restrict__ out_final, float4* __restrict__ out_part,
                   const float4* __restrict__ addend, int D4) {
  __shared__ float4 sm[kThreads / kGroup][kMaxNF * kGroup];
  const int4 it = __ldg(reinterpret_cast<const int4*>(items + blockIdx.x));
  const int grp = threadIdx.x / kGroup, g = threadIdx.x % kGroup;
  float4 acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kStride = kThreads / kGroup;
  int r = it.x + grp;
  for (; r + 3 * kStride < it.y; r += 4 * kStride) {      // 4 independent rows in flight, added in row order
    float4 v[4][NF];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const int c = g + f * kGroup;
        if (c < D4) v[u][f] = __ldg(part_in + (int64_t)(r + u * kStride) * D4 + c);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const int c = g + f * kGroup;
        if (c < D4) add4(acc[f], v[u][f]);
      }
  }
  for (; r < it.y; r += kStride) {
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) add4(acc[f], __ldg(part_in + (int64_t)r * D4 + c));
    }
  }
#pragma unroll
  for (int f = 0; f < NF; ++f) sm[grp][g + f * kGroup] = acc[f];
  __syncthreads();
  if (grp == 0) {
    const bool final_row = (it.w & 1) != 0;
    float4* out = (final_row ? out_final : out_part) + (int64_t)it.z * D4;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int c = g + f * kGroup;
      if (c < D4) {
        float4 v = sm[0][c];
        for (int k = 1; k < kThreads / kGroup; ++k) add4(v, sm[k][c]);
        if (final_row && addend != nullptr) add4(v, __ldg(addend + (int64_t)it.z * D4 + c));
        out[c] = v;
      }
    }
  }
}

inline int check_dim(int32_t D, int* D4, int* NF) {
  if (D <= 0 || D % 4 != 0 || D > kMaxNF * kGroup * 4) return 1;
  *D4 = D / 4;
  *NF = (*D4 + kGroup - 1) / kGroup;
  return 0;
}

#define KGC_DISPATCH_NF(NFV, ...)                         \
  switch (NFV) {                                          \
    case 1: { constexpr int NF = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NF = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NF = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int NF = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int NF = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int NF = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int NF = 7; __VA_ARGS__; } break; \
    default: { constexpr int NF = 8; __VA_ARGS__; } break; \
  }

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int kgc_agg_fwd(const float* x, const float* rel, const float* ee, const kgc_edge_rec_t* rec_dst,
                           const kgc_item_t* items, int64_t n_items, float* out_final, float* out_part, int32_t D,
                           void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  if (n_items == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(n_items * kGroup, kThreads);
  KGC_DISPATCH_NF(nf, (agg_fwd_kernel<NF><<<grid, kThreads, 0, as_stream(stream)>>>(
                          (const float4*)x, (const float4*)rel, (const float4*)ee, rec_dst, items, n_items,
                          (float4*)out_final, (float4*)out_part, D4)));
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_rows_reduce(const float* part_in, const kgc_item_t* items, int64_t n_items, float* out_final,
                               float* out_part, const float* addend, int32_t D, void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  if (n_items == 0) return 0;
  KGC_DISPATCH_NF(nf, (rows_reduce_kernel<NF><<<(unsigned)n_items, kThreads, 0, as_stream(stream)>>>(
                          (const float4*)part_in, items, (float4*)out_final, (float4*)out_part,
                          (const float4*)addend, D4)));
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_agg_bwd_src(const float* x, const float* rel, const float* ee, const float* g3,
                               const kgc_edge_rec_t* rec_src, const kgc_item_t* items, int64_t n_items,
                               int64_t n_dst_rows, int64_t n_edges_in, const float* loop_addend, float* d_ee,
                               float* dx_final, float* dx_part, int32_t D, void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  if (n_items == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(n_items * kGroup, kThreads);
  KGC_DISPATCH_NF(nf, (agg_bwd_src_kernel<NF><<<grid, kThreads, 0, as_stream(stream)>>>(
                          (const float4*)x, (const float4*)rel, (const float4*)ee, (const float4*)g3, rec_src, items,
                          n_items, n_dst_rows, (int32_t)n_edges_in, (const float4*)loop_addend, (float4*)d_ee,
                          (float4*)dx_final, (float4*)dx_part, D4)));
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_agg_bwd_rel(const float* x, const float* ee, const float* g3, const kgc_edge_rec_t* rec_type,
                               const kgc_item_t* items, int64_t n_items, int64_t n_dst_rows, int64_t n_edges_in,
                               float* drel_final, float* drel_part, int32_t D, void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  if (n_items == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(n_items * kGroup, kThreads);
  KGC_DISPATCH_NF(nf, (agg_bwd_rel_kernel<NF><<<grid, kThreads, 0, as_stream(stream)>>>(
                          (const float4*)x, (const float4*)ee, (const float4*)g3, rec_type, items, n_items, n_dst_rows,
                          (int32_t)n_edges_in, (float4*)drel_final, (float4*)drel_part, D4)));
  KGC_LAUNCH_CHECK();
  return 0;
}

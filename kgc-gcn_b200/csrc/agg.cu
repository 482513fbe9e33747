// K2 / K3: relation-aware aggregation, forward and backward (HBM-bound segment reductions).
//
// Reference arithmetic replaced (model.py:111-118 + PyG aggr='add', model.py:50,99-100):
//   m_e = (x[src_e] * rel[type_e] * edge_emb[e]) @ W * norm_e ;  out[dst] += m_e
// restated as aggregate-then-transform (the matmul is linear):
//   agg[dst] = sum_e norm_e * x[src_e] (.) rel[type_e] (.) ee[e]      (this file)
//   res      = agg @ W                                                   (dense GEMM, host side)
//
// Streaming design (round 1, second version - the first one gave one 8-lane group a whole output row and
// was latency-bound at 25-40% of HBM peak, profiles/r01_ncu_agg_kernels.md):
//   * the sorted edge records are cut into CHUNKS of 32 consecutive records, one warp per chunk, whatever
//     the row boundaries: every warp has the same amount of work and there is no per-row launch overhead;
//   * lane l owns float4 columns l and l + 32 of a row (D <= 256), so a 400-byte row is one fully
//     coalesced 25-lane request; 4 edges are in flight per warp (records prefetched one trip ahead);
//   * a row that lies inside one chunk is written straight to the output; the (at most two) rows a chunk
//     shares with its neighbours go to CARRY rows, which kgc_rows_reduce adds in a fixed order.
// No float atomics anywhere: results are bit-reproducible run to run.
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kGroup = 8;          // kgc_rows_reduce: lanes per partial-row group
constexpr int kThreads = 256;
constexpr int kMaxNF = 8;          // kgc_rows_reduce: D <= 8 * 8 * 4 = 256
constexpr int kChunk = KGC_CHUNK_EDGES;
constexpr uint32_t kRowMask = 0x3FFFFFFFu, kFirst = 0x40000000u, kLast = 0x80000000u;

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

enum Mode { kFwd = 0, kBwdSrc = 1, kBwdRel = 2 };

struct StreamArgs {
  const float4* x;            // node rows (global ids)
  const float4* rel;          // relation rows (fwd, bwd_src)
  const float4* ee;           // edge-embedding rows (global edge ids of this rank)
  const float4* g3;           // bwd: [2, n_dst_rows, D] upstream planes
  const float4* addend;       // bwd_src: self-loop term added on rows finished inside a chunk (or null)
  const kgc_edge_rec_t* rec;
  const uint32_t* rowflags;   // row | first << 30 | last << 31 per sorted record
  const kgc_chunk_t* chunks;
  float4* out_final;
  float4* carry;
  float4* d_ee;               // bwd_src: per-edge output
  int64_t n_rec;
  int64_t plane;              // n_dst_rows * D4 (bwd)
  int32_t n_edges_in;         // records with eid >= n_edges_in belong to the out half (bwd)
  int32_t D4;
};

// ---- PTX: mbarrier + bulk async copy (global -> shared, completes on an mbarrier; SASS: UBLKCP) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One warp walks one chunk of kChunk (= 32) sorted records (third version; the register-staged second version
// was co-limited by instruction issue and by 16 warps x 8 edges of loads in flight, ~0.5 of the HBM peak).
//   1. lane l loads record cb + l (one coalesced request for the whole chunk);
//   2. lane l issues one bulk async copy per operand ROW of its edge (edge-embedding row, gathered x / g row and,
//      for large relation tables, the relation row) from global memory straight into the warp's shared-memory
//      stage: 2-3 x 32 rows (25-38 KB for D = 100) in flight per warp without a single staging register; the copies
//      complete on the warp's mbarrier (expect_tx = the chunk's bytes);
//   3. the warp consumes the 32 edges from shared memory in record order (lane c owns float4 column c; a row is
//      one conflict-free 400-byte LDS wave), writes rows that lie inside the chunk straight to the output and the
//      <= 2 boundary rows to carry rows.
// One CTA per SM owns as many such warps as shared memory allows (8 for D = 100); the grid is persistent.
constexpr int kSmemBudget = 216 * 1024;

template <int MODE, bool kRelBulk, int NF>
__global__ void __launch_bounds__(512, 1)
agg_bulk_kernel(const StreamArgs A, const int n_arr, const int64_t n_chunks) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  static_assert(kChunk == 32, "one record per lane");
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, n_warps = blockDim.x / 32;
  const int D4 = A.D4;
  const uint32_t row_bytes = (uint32_t)D4 * 16u;
  const uint32_t arr_bytes = row_bytes * kChunk;
  uint8_t* stage = smem_raw + (size_t)warp * n_arr * arr_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)n_warps * n_arr * arr_bytes) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const float4* s_ee = reinterpret_cast<const float4*>(stage);
  const float4* s_b = reinterpret_cast<const float4*>(stage + arr_bytes);
  const float4* s_c = reinterpret_cast<const float4*>(stage + 2 * arr_bytes);     // rel (fwd/bwd_src bulk) or x (bwd_rel)
  bool active[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) active[f] = lane + f * 32 < D4;
  uint32_t phase = 0;

  for (int64_t chunk = blockIdx.x * (int64_t)n_warps + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * n_warps) {
    const int64_t cb = chunk * kChunk;
    const int cnt = (int)(cb + kChunk < A.n_rec ? kChunk : A.n_rec - cb);
    const int2 slots = __ldg(reinterpret_cast<const int2*>(A.chunks + chunk));
    const int4 myrec = ld_rec(A.rec + cb + (lane < cnt ? lane : cnt - 1));
    const uint32_t myflag = __ldg(A.rowflags + cb + (lane < cnt ? lane : cnt - 1));
    // ---- issue the chunk's row copies
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * (uint32_t)n_arr * row_bytes);
    __syncwarp();
    if (lane < cnt) {
      const uint32_t eid = (uint32_t)myrec.x, ra = (uint32_t)myrec.y, rb = (uint32_t)myrec.z;
      bulk_g2s(stage + lane * row_bytes, A.ee + (uint64_t)eid * D4, row_bytes, bar);
      if (MODE == kFwd) {
        bulk_g2s(stage + arr_bytes + lane * row_bytes, A.x + (uint64_t)ra * D4, row_bytes, bar);
        if (kRelBulk) bulk_g2s(stage + 2 * arr_bytes + lane * row_bytes, A.rel + (uint64_t)rb * D4, row_bytes, bar);
      } else if (MODE == kBwdSrc) {
        bulk_g2s(stage + arr_bytes + lane * row_bytes, A.g3 + ((int)eid >= A.n_edges_in ? A.plane : 0) + (uint64_t)ra * D4,
                 row_bytes, bar);
        if (kRelBulk) bulk_g2s(stage + 2 * arr_bytes + lane * row_bytes, A.rel + (uint64_t)rb * D4, row_bytes, bar);
      } else {
        bulk_g2s(stage + arr_bytes + lane * row_bytes, A.g3 + ((int)eid >= A.n_edges_in ? A.plane : 0) + (uint64_t)rb * D4,
                 row_bytes, bar);
        bulk_g2s(stage + 2 * arr_bytes + lane * row_bytes, A.x + (uint64_t)ra * D4, row_bytes, bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    // ---- consume in record order
    float4 acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool started_here = (__shfl_sync(0xffffffffu, myflag, 0) & kFirst) != 0;
    bool open_row = false;
#pragma unroll 4
    for (int e = 0; e < cnt; ++e) {
      // shuffles stay outside every lane predicate: all 32 lanes take part
      const uint32_t flags = __shfl_sync(0xffffffffu, myflag, e);
      const float nrm = __int_as_float(__shfl_sync(0xffffffffu, myrec.w, e));
      const uint32_t rb = (uint32_t)__shfl_sync(0xffffffffu, myrec.z, e);
      const uint32_t eid = (uint32_t)__shfl_sync(0xffffffffu, myrec.x, e);
      (void)rb; (void)eid;
      const int64_t row = flags & kRowMask;
      if (flags & kFirst) {
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        started_here = true;
      }
      open_row = true;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const int c = lane + f * 32;
        if (active[f]) {
          const float4 va = s_ee[e * D4 + c];
          const float4 vb = s_b[e * D4 + c];
          if (MODE == kFwd) {
            const float4 vr = kRelBulk ? s_c[e * D4 + c] : __ldg(A.rel + (uint64_t)rb * D4 + c);
            add4(acc[f], mul3s(nrm, vb, vr, va));                                    // norm * ((x * rel) * ee)
          } else if (MODE == kBwdSrc) {
            const float4 vr = kRelBulk ? s_c[e * D4 + c] : __ldg(A.rel + (uint64_t)rb * D4 + c);
            const float4 pe = scale4(nrm, mul4(vb, vr));                              // norm * g[dst] * rel
            st_stream(A.d_ee + (uint64_t)eid * D4 + c, mul4(pe, __ldg(A.x + row * D4 + c)));   // * x[src]
            add4(acc[f], mul4(pe, va));
          } else {
            add4(acc[f], mul3s(nrm, vb, s_c[e * D4 + c], va));                       // norm * ((g * x) * ee)
          }
        }
      }
      if (flags & kLast) {                                       // the row ends here
        float4* out = started_here ? A.out_final + row * D4 : A.carry + (int64_t)slots.x * D4;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const int c = lane + f * 32;
          if (active[f]) {
            float4 v = acc[f];
            if (MODE == kBwdSrc && started_here && A.addend != nullptr) add4(v, __ldg(A.addend + row * D4 + c));
            out[c] = v;
          }
        }
        open_row = false;
      }
    }
    if (open_row) {                                              // the row continues in the next chunk
      float4* out = A.carry + (int64_t)slots.y * D4;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const int c = lane + f * 32;
        if (active[f]) out[c] = acc[f];
      }
    }
    __syncwarp();                                                // every lane is done with the stage before it is refilled
  }
}

// out[rows[i]] = addend ? addend[rows[i]] : 0   (rows without any edge record)
__global__ void rows_fill_kernel(const int32_t* __restrict__ rows, int64_t n_rows, const float4* __restrict__ addend,
                                 float4* __restrict__ out, int D4) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (i >= n_rows) return;
  const int64_t r = rows[i];
  for (int c = lane; c < D4; c += 32)
    out[r * D4 + c] = addend ? __ldg(addend + r * D4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
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

template <int MODE, bool kRelBulk, int NF>
int launch_bulk(const StreamArgs& A, int n_arr, cudaStream_t st) {
  const int64_t n_chunks = ceil_div(A.n_rec, kChunk);
  const size_t stage = (size_t)n_arr * kChunk * A.D4 * 16;
  int warps = (int)(kSmemBudget / stage);
  if (warps > 16) warps = 16;
  if (warps < 1) return fail("launch_bulk", "row too wide for one shared-memory stage");
  const size_t smem = warps * stage + 16 * sizeof(uint64_t);
  auto kern = agg_bulk_kernel<MODE, kRelBulk, NF>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    KGC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int64_t grid = ceil_div(n_chunks, warps);
  if (grid > kNumSMs) grid = kNumSMs;
  kern<<<(unsigned)grid, warps * 32, smem, st>>>(A, n_arr, n_chunks);
  KGC_LAUNCH_CHECK();
  return 0;
}

template <int MODE>
int launch_stream(const StreamArgs& A, int64_t n_types, cudaStream_t st) {
  if (A.n_rec == 0) return 0;
  // relation rows ride the bulk copies when the table would not stay in the (small) L1 left beside the stages
  const bool rel_bulk = MODE != kBwdRel && n_types * (int64_t)A.D4 * 16 > 12 * 1024;
  const int n_arr = (MODE == kBwdRel || rel_bulk) ? 3 : 2;
  if (A.D4 <= 32) {
    return rel_bulk ? launch_bulk<MODE, true, 1>(A, n_arr, st) : launch_bulk<MODE, false, 1>(A, n_arr, st);
  }
  return rel_bulk ? launch_bulk<MODE, true, 2>(A, n_arr, st) : launch_bulk<MODE, false, 2>(A, n_arr, st);
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int kgc_agg_fwd(const float* x, const float* rel, int64_t n_types, const float* ee,
                           const kgc_edge_rec_t* rec_dst, const uint32_t* rowflags, const kgc_chunk_t* chunks,
                           int64_t n_rec, float* out_final, float* carry, int32_t D, void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  StreamArgs A = {};
  A.x = (const float4*)x; A.rel = (const float4*)rel; A.ee = (const float4*)ee;
  A.rec = rec_dst; A.rowflags = rowflags; A.chunks = chunks;
  A.out_final = (float4*)out_final; A.carry = (float4*)carry;
  A.n_rec = n_rec; A.D4 = D4;
  return launch_stream<kFwd>(A, n_types, as_stream(stream));
}

extern "C" int kgc_agg_bwd_src(const float* x, const float* rel, int64_t n_types, const float* ee, const float* g3,
                               const kgc_edge_rec_t* rec_src, const uint32_t* rowflags, const kgc_chunk_t* chunks,
                               int64_t n_rec, int64_t n_dst_rows, int64_t n_edges_in, const float* loop_addend,
                               float* d_ee, float* dx_final, float* carry, int32_t D, void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  StreamArgs A = {};
  A.x = (const float4*)x; A.rel = (const float4*)rel; A.ee = (const float4*)ee; A.g3 = (const float4*)g3;
  A.addend = (const float4*)loop_addend;
  A.rec = rec_src; A.rowflags = rowflags; A.chunks = chunks;
  A.out_final = (float4*)dx_final; A.carry = (float4*)carry; A.d_ee = (float4*)d_ee;
  A.n_rec = n_rec; A.plane = n_dst_rows * (int64_t)D4; A.n_edges_in = (int32_t)n_edges_in; A.D4 = D4;
  return launch_stream<kBwdSrc>(A, n_types, as_stream(stream));
}

extern "C" int kgc_agg_bwd_rel(const float* x, const float* ee, const float* g3, const kgc_edge_rec_t* rec_type,
                               const uint32_t* rowflags, const kgc_chunk_t* chunks, int64_t n_rec, int64_t n_dst_rows,
                               int64_t n_edges_in, float* drel_final, float* carry, int32_t D, void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  StreamArgs A = {};
  A.x = (const float4*)x; A.ee = (const float4*)ee; A.g3 = (const float4*)g3;
  A.rec = rec_type; A.rowflags = rowflags; A.chunks = chunks;
  A.out_final = (float4*)drel_final; A.carry = (float4*)carry;
  A.n_rec = n_rec; A.plane = n_dst_rows * (int64_t)D4; A.n_edges_in = (int32_t)n_edges_in; A.D4 = D4;
  return launch_stream<kBwdRel>(A, 0, as_stream(stream));
}

extern "C" int kgc_rows_fill(const int32_t* rows, int64_t n_rows, const float* addend, float* out, int32_t D,
                             void* stream) {
  int D4, nf;
  KGC_REQUIRE(check_dim(D, &D4, &nf) == 0, "D must be a multiple of 4 and <= 256");
  if (n_rows == 0) return 0;
  rows_fill_kernel<<<(unsigned)ceil_div(n_rows * 32, 256), 256, 0, as_stream(stream)>>>(rows, n_rows, (const float4*)addend,
                                                                                   (float4*)out, D4);
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

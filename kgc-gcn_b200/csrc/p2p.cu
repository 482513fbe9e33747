// K10: halo exchange of the dst-partitioned layer over NVLink peer memory - no NCCL on the data path.
//
// The partitioned layer (conv.py forward_partitioned) needs, per step, the source rows x[src] of the edges a rank owns
// and returns the source-row gradients to their owners.  The library formulation is an all-gather of ALL node rows and a
// reduce-scatter of a dense [N, D] partial - N x 400 bytes in and out per rank whatever the graph.  A rank only ever
// touches the rows its edges reference (73% / 46% / 26% of the nodes at 2 / 4 / 8 ranks of the WN18RR-shape workload), and
// every GPU of the NVSwitch domain can load any peer's memory directly, so:
//   * kgc_p2p_halo_gather  - every rank publishes its row block at the head of a symmetric node table; after a barrier each
//                            rank PULLS exactly the remote rows its records reference (one warp per 400-byte row, 8 rows
//                            in flight per warp) into the compact tail of its own table (own rows, then the halo: every
//                            per-rank structure is O(own rows + halo) whatever the number of ranks);
//                            The list is ascending by owner, so ranks that walk it front to back all read the SAME owner at
//                            the same time (one GPU's NVLink egress shared by seven readers); `order` is a permutation
//                            of the list positions that deals 8-row trips to the owners in turn, starting behind the
//                            reader's own rank - every peer link of every GPU carries traffic from the first trip on.
//   * kgc_p2p_halo_reduce  - every rank leaves its partial d_x in a symmetric buffer; after a barrier the owner of a row
//                            pulls the partials of the ranks that touched it (a per-row index table built with the partition)
//                            and adds them IN RANK ORDER (deterministic) together with the self-loop term - the
//                            reduce-scatter and the add that followed it, in one kernel;
//   * kgc_p2p_barrier      - flag barrier in symmetric memory (release/acquire at system scope, monotonically increasing
//                            epoch, one flag slot per peer).  Ranks may skew by any amount a training loop produces
//                            (evaluation, checkpoint writes, logging): the wait has NO data-path time-out.  A watchdog
//                            of ~2 minutes (a peer process is gone) sets *error and TRAPS - the context dies and every
//                            later CUDA call of the process raises; the kernels after the barrier never run on
//                            unpublished peer data.
// Hazards: a buffer a peer reads in step t is overwritten by its owner in step t + 1 only after another barrier of the same
// sequence has been passed by every rank (forward barrier between two backward uses and vice versa).
#include "common.cuh"

namespace kgc {
namespace {

__device__ __forceinline__ float4 add_vec(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ double2 add_vec(const double2& a, const double2& b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// flags[r] = rank r's flag array [world] (peer pointers); slot [s] of rank r's array is written by rank s.
// Called by every thread of ONE CTA (>= world threads); the CTA's earlier writes are ordered before the arrival.
__device__ __forceinline__ void cta_barrier_across_ranks(uint32_t* const* __restrict__ flags, int rank, int world,
                                                         uint32_t* __restrict__ epoch, int* __restrict__ error) {
  __shared__ uint32_t e_s;
  __threadfence_system();                                   // this thread's earlier writes are visible before the arrival
  __syncthreads();
  if (threadIdx.x == 0) e_s = *epoch + 1;
  __syncthreads();
  const uint32_t e = e_s;
  const int r = threadIdx.x;
  if (r < world) {
    st_release_sys(flags[r] + rank, e);
    const uint32_t* mine = flags[rank] + r;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 240000000000ll) {                // ~2 min at 1.9 GHz: a peer process is gone
        *error = 1;
        __threadfence_system();
        __trap();                                           // never continue on peer buffers that were not published
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
}

__global__ void p2p_barrier_kernel(uint32_t* const* __restrict__ flags, int rank, int world, uint32_t* __restrict__ epoch,
                                   int* __restrict__ error) {
  cta_barrier_across_ranks(flags, rank, world, epoch, error);
}

// One-shot all-reduce of a small vector (BatchNorm column sums, hub rows, replicated-parameter gradients): every rank
// copies its contribution into its symmetric staging slot, the ranks meet at a flag barrier, then every rank reads all
// slots over NVLink and adds them IN RANK ORDER - the same bits on every rank.  One CTA; n4 = 16-byte elements.
constexpr int kArThreads = 1024;
constexpr int kArUnroll = 4;

template <typename T4, int WMAX>
__global__ void __launch_bounds__(kArThreads)
p2p_allreduce_kernel(char* const* __restrict__ stage, int64_t offset, uint32_t* const* __restrict__ flags, int rank, int world,
                     uint32_t* __restrict__ epoch, int* __restrict__ error, const T4* __restrict__ in, T4* __restrict__ out,
                     int64_t n4) {
  T4* mine = reinterpret_cast<T4*>(stage[rank] + offset);
  for (int64_t i = threadIdx.x; i < n4; i += kArThreads) mine[i] = in[i];
  cta_barrier_across_ranks(flags, rank, world, epoch, error);
  for (int64_t i0 = threadIdx.x; i0 < n4; i0 += (int64_t)kArThreads * kArUnroll) {
    T4 v[kArUnroll][WMAX];
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      const int64_t i = i0 + (int64_t)u * kArThreads;
#pragma unroll
      for (int r = 0; r < WMAX; ++r)
        if (r < world && i < n4) v[u][r] = reinterpret_cast<const T4*>(stage[r] + offset)[i];
    }
#pragma unroll
    for (int u = 0; u < kArUnroll; ++u) {
      const int64_t i = i0 + (int64_t)u * kArThreads;
      if (i < n4) {
        T4 acc = v[u][0];
#pragma unroll
        for (int r = 1; r < WMAX; ++r)
          if (r < world) acc = add_vec(acc, v[u][r]);
        out[i] = acc;
      }
    }
  }
}

// The same sum for payloads one CTA is too slow for (the replicated-parameter gradients: ~1 MB from each of 8 ranks):
// stage with many CTAs, ONE CTA meets the other ranks, many CTAs add - three launches on the same stream.
template <typename T4>
__global__ void __launch_bounds__(256) p2p_stage_kernel(char* const* __restrict__ stage, int64_t offset, int rank,
                                                        const T4* __restrict__ in, int64_t n4) {
  T4* mine = reinterpret_cast<T4*>(stage[rank] + offset);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) mine[i] = in[i];
}
template <typename T4, int WMAX>
__global__ void __launch_bounds__(256) p2p_sum_kernel(char* const* __restrict__ stage, int64_t offset, int world,
                                                      T4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    T4 v[WMAX];
#pragma unroll
    for (int r = 0; r < WMAX; ++r)
      if (r < world) v[r] = reinterpret_cast<const T4*>(stage[r] + offset)[i];
    T4 acc = v[0];
#pragma unroll
    for (int r = 1; r < WMAX; ++r)
      if (r < world) acc = add_vec(acc, v[r]);
    out[i] = acc;
  }
}

constexpr int kHaloThreads = 256;
constexpr int kHaloUnroll = 8;

// table[r] = rank r's COMPACT node table: its own block_rows rows, then its halo; rows[i] = renumbered id g of the i-th
// halo row: owner g / block_rows, local row g % block_rows; it lands in row block_rows + i of this rank's table
__global__ void __launch_bounds__(kHaloThreads)
halo_gather_kernel(float4* const* __restrict__ table, int rank, const int32_t* __restrict__ rows,
                   const int32_t* __restrict__ order, int64_t n_rows, int block_rows, int D4) {
  const int lane = threadIdx.x % 32;
  const int64_t warp = (blockIdx.x * (int64_t)kHaloThreads + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * (kHaloThreads / 32);
  float4* __restrict__ mine = table[rank];
  for (int64_t i0 = warp * kHaloUnroll; i0 < n_rows; i0 += n_warps * kHaloUnroll) {
    float4 v[kHaloUnroll][2];
    int64_t g[kHaloUnroll];
#pragma unroll
    for (int u = 0; u < kHaloUnroll; ++u) {
      int64_t i = i0 + u < n_rows ? i0 + u : n_rows - 1;
      if (order != nullptr) i = __ldg(order + i);                                   // pull schedule (see kgc_p2p_halo_gather)
      g[u] = __ldg(rows + i);
      const float4* src = table[g[u] / block_rows] + (g[u] % block_rows) * D4;     // the owner's block opens its table
      g[u] = block_rows + i;                                                        // compact row of this rank's table
      if (lane < D4) v[u][0] = src[lane];
      if (lane + 32 < D4) v[u][1] = src[lane + 32];
    }
#pragma unroll
    for (int u = 0; u < kHaloUnroll; ++u) {
      if (i0 + u < n_rows) {
        float4* dst = mine + g[u] * D4;
        if (lane < D4) dst[lane] = v[u][0];
        if (lane + 32 < D4) dst[lane + 32] = v[u][1];
      }
    }
  }
}

// out[v] = addend[v] + sum over the ranks r with idx[r][j] >= 0 (ascending r) of part[r][idx[r][j]];  v = row_ids[j], or j
// when row_ids is NULL (all rows).  addend == out is allowed (in-place add of the remote partials).
template <int WMAX>
__global__ void __launch_bounds__(kHaloThreads)
halo_reduce_kernel(const float4* const* __restrict__ part, int world, const int32_t* __restrict__ idx,
                   const int32_t* __restrict__ row_ids, int64_t n_rows, const float4* addend, float4* out, int D4) {
  const int lane = threadIdx.x % 32;
  const int64_t warp = (blockIdx.x * (int64_t)kHaloThreads + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * (kHaloThreads / 32);
  constexpr int kRows = WMAX <= 4 ? 4 : 2;                  // rows per warp trip: kRows * WMAX row loads in flight
  for (int64_t v0 = warp * kRows; v0 < n_rows; v0 += n_warps * kRows) {
    int32_t at[kRows][WMAX];
#pragma unroll
    for (int u = 0; u < kRows; ++u)
#pragma unroll
      for (int r = 0; r < WMAX; ++r) at[u][r] = (r < world && v0 + u < n_rows) ? __ldg(idx + (int64_t)r * n_rows + v0 + u) : -1;
    for (int c = lane; c < D4; c += 32) {
      float4 p[kRows][WMAX];
#pragma unroll
      for (int u = 0; u < kRows; ++u)
#pragma unroll
        for (int r = 0; r < WMAX; ++r)
          if (at[u][r] >= 0) p[u][r] = part[r][(int64_t)at[u][r] * D4 + c];
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        if (v0 + u < n_rows) {
          const int64_t v = row_ids != nullptr ? (int64_t)__ldg(row_ids + v0 + u) : v0 + u;    // addend may alias out: plain load
          const float4 acc = addend != nullptr ? addend[v * D4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < WMAX; ++r)
            if (at[u][r] >= 0) { s.x += p[u][r].x; s.y += p[u][r].y; s.z += p[u][r].z; s.w += p[u][r].w; }
          out[v * D4 + c] = make_float4(s.x + acc.x, s.y + acc.y, s.z + acc.z, s.w + acc.w);
        }
      }
    }
  }
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int kgc_p2p_barrier(void* const* flag_ptrs_dev, int32_t rank, int32_t world, uint32_t* epoch, int32_t* error,
                               void* stream) {
  KGC_REQUIRE(flag_ptrs_dev && epoch && error && world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad arguments");
  p2p_barrier_kernel<<<1, 64, 0, as_stream(stream)>>>((uint32_t* const*)flag_ptrs_dev, rank, world, epoch, error);
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_p2p_halo_gather(void* const* table_ptrs_dev, int32_t rank, const int32_t* rows, const int32_t* order,
                                   int64_t n_rows, int64_t block_rows, int32_t D, void* stream) {
  KGC_REQUIRE(table_ptrs_dev && block_rows > 0 && D > 0 && D % 4 == 0 && D <= 256, "bad arguments");
  if (n_rows == 0) return 0;
  KGC_REQUIRE(rows != nullptr, "null row list");
  int64_t blocks = ceil_div(n_rows, (kHaloThreads / 32) * kHaloUnroll);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  halo_gather_kernel<<<(unsigned)blocks, kHaloThreads, 0, as_stream(stream)>>>((float4* const*)table_ptrs_dev, rank, rows, order,
                                                                               n_rows, (int)block_rows, D / 4);
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_p2p_halo_reduce(void* const* part_ptrs_dev, int32_t world, const int32_t* idx, const int32_t* row_ids,
                                   int64_t n_rows, const float* addend, float* out, int32_t D, void* stream) {
  KGC_REQUIRE(part_ptrs_dev && idx && out && world >= 1 && world <= 64 && D > 0 && D % 4 == 0, "bad arguments");
  if (n_rows == 0) return 0;
  int64_t blocks = ceil_div(n_rows, (kHaloThreads / 32) * 2);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  const unsigned g = (unsigned)blocks;
  cudaStream_t st = as_stream(stream);
  if (world <= 2) {
    halo_reduce_kernel<2><<<g, kHaloThreads, 0, st>>>((const float4* const*)part_ptrs_dev, world, idx, row_ids, n_rows,
                                                      (const float4*)addend, (float4*)out, D / 4);
  } else if (world <= 4) {
    halo_reduce_kernel<4><<<g, kHaloThreads, 0, st>>>((const float4* const*)part_ptrs_dev, world, idx, row_ids, n_rows,
                                                      (const float4*)addend, (float4*)out, D / 4);
  } else if (world <= 8) {
    halo_reduce_kernel<8><<<g, kHaloThreads, 0, st>>>((const float4* const*)part_ptrs_dev, world, idx, row_ids, n_rows,
                                                      (const float4*)addend, (float4*)out, D / 4);
  } else {
    return fail(__func__, "more than 8 ranks are not supported by this build");
  }
  KGC_LAUNCH_CHECK();
  return 0;
}

extern "C" int kgc_p2p_allreduce(void* const* stage_ptrs_dev, int64_t offset_bytes, void* const* flag_ptrs_dev, int32_t rank,
                                 int32_t world, uint32_t* epoch, int32_t* error, const void* in, void* out, int64_t n_bytes,
                                 int32_t is_double, void* stream) {
  KGC_REQUIRE(stage_ptrs_dev && flag_ptrs_dev && epoch && error && in && out, "null argument");
  KGC_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "1..8 ranks");
  KGC_REQUIRE(n_bytes > 0 && n_bytes % 16 == 0 && offset_bytes % 16 == 0, "sizes and offsets must be multiples of 16 bytes");
  KGC_REQUIRE((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0, "buffers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int64_t n4 = n_bytes / 16;
  if (n_bytes > (64 << 10)) {                                  // wide payload: stage / barrier / add with many CTAs
    const unsigned g = (unsigned)(ceil_div(n4, 256) < 2 * kNumSMs ? ceil_div(n4, 256) : 2 * kNumSMs);
#define KGC_AR_WIDE(T4, WM)                                                                                                \
  do {                                                                                                                     \
    p2p_stage_kernel<T4><<<g, 256, 0, st>>>((char* const*)stage_ptrs_dev, offset_bytes, rank, (const T4*)in, n4);          \
    p2p_barrier_kernel<<<1, 64, 0, st>>>((uint32_t* const*)flag_ptrs_dev, rank, world, epoch, error);                      \
    p2p_sum_kernel<T4, WM><<<g, 256, 0, st>>>((char* const*)stage_ptrs_dev, offset_bytes, world, (T4*)out, n4);            \
  } while (0)
    if (is_double) {
      if (world <= 2) KGC_AR_WIDE(double2, 2); else if (world <= 4) KGC_AR_WIDE(double2, 4); else KGC_AR_WIDE(double2, 8);
    } else {
      if (world <= 2) KGC_AR_WIDE(float4, 2); else if (world <= 4) KGC_AR_WIDE(float4, 4); else KGC_AR_WIDE(float4, 8);
    }
#undef KGC_AR_WIDE
    KGC_LAUNCH_CHECK();
    return 0;
  }
#define KGC_AR_LAUNCH(T4, WM)                                                                                              \
  p2p_allreduce_kernel<T4, WM><<<1, kArThreads, 0, st>>>((char* const*)stage_ptrs_dev, offset_bytes,                       \
                                                          (uint32_t* const*)flag_ptrs_dev, rank, world, epoch, error,      \
                                                          (const T4*)in, (T4*)out, n4)
  if (is_double) {
    if (world <= 2) KGC_AR_LAUNCH(double2, 2); else if (world <= 4) KGC_AR_LAUNCH(double2, 4); else KGC_AR_LAUNCH(double2, 8);
  } else {
    if (world <= 2) KGC_AR_LAUNCH(float4, 2); else if (world <= 4) KGC_AR_LAUNCH(float4, 4); else KGC_AR_LAUNCH(float4, 8);
  }
#undef KGC_AR_LAUNCH
  KGC_LAUNCH_CHECK();
  return 0;
}

// K1: CSR / permutation builder + per-half symmetric degree normalisation.
// Replaces MGCNConv.compute_norm (reference model.py:72-80) and the index handling inside PyG
// propagate (model.py:99-101).  Integer work is exact; the stable radix sort is CUB's (a library
// primitive for a plain sort); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace kgc {
namespace {

constexpr int kThreads = 256;

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline int key_bits(int64_t n) {
  int b = 1;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

struct Layout {
  size_t src32, dst32, type32, iota, key_sorted, cnt, cnt2, flag, cub, total;
  size_t cub_bytes;
};

Layout make_layout(int64_t n2, int64_t n_nodes, int64_t n_types, size_t cub_bytes) {
  Layout L;
  size_t off = 0;
  const size_t e = align_up(size_t(n2) * 4);
  const size_t c = align_up(size_t((n_nodes > n_types ? n_nodes : n_types) + 2) * 4);
  L.src32 = off; off += e;
  L.dst32 = off; off += e;
  L.type32 = off; off += e;
  L.iota = off; off += e;
  L.key_sorted = off; off += e;
  L.cnt = off; off += c;
  L.cnt2 = off; off += c;
  L.flag = off; off += 256;
  L.cub = off; off += align_up(cub_bytes);
  L.cub_bytes = cub_bytes;
  L.total = off;
  return L;
}

cudaError_t cub_temp_bytes(int64_t n2, int64_t n_rows_max, size_t* out) {
  size_t a = 0, b = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n2, 0, 32);
  if (e != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n_rows_max + 1));
  if (e != cudaSuccess) return e;
  *out = a > b ? a : b;
  return cudaSuccess;
}

__global__ void narrow_and_check(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 const int64_t* __restrict__ type, int64_t n2, int64_t n_src_rows, int64_t n_dst_rows,
                                 int64_t n_types,
                                 int32_t* __restrict__ src32, int32_t* __restrict__ dst32,
                                 int32_t* __restrict__ type32, int32_t* __restrict__ iota, int32_t* flag) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n2) return;
  int64_t s = src[i], d = dst[i], t = type[i];
  if (s < 0 || s >= n_src_rows || d < 0 || d >= n_dst_rows || t < 0 || t >= n_types) atomicOr(flag, 1);
  src32[i] = (int32_t)s;
  dst32[i] = (int32_t)d;
  type32[i] = (int32_t)t;
  iota[i] = (int32_t)i;
}

// deg[h*N + v] = #{e in half h : src_e == v}   (model.py:73-75: row = edge_index[0])
__global__ void half_degree(const int32_t* __restrict__ src32, int64_t n2, int64_t n_in, int64_t n_nodes, int32_t* deg) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n2) return;
  int64_t h = (i >= n_in) ? 1 : 0;
  atomicAdd(&deg[h * n_nodes + src32[i]], 1);   // integer atomics: order-independent, exact
}

__device__ __forceinline__ float deg_inv_sqrt(int32_t d) {
  // deg.pow(-0.5) with inf -> 0 (model.py:76-77); evaluated in fp64 and rounded once to fp32.
  return d > 0 ? (float)(1.0 / sqrt((double)d)) : 0.0f;
}

__global__ void edge_norm(const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32,
                          const int32_t* __restrict__ deg, int64_t n2, int64_t n_in, int64_t n_nodes, int64_t dst_offset,
                          float* __restrict__ norm) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const int32_t* dh = deg + ((i >= n_in) ? n_nodes : 0);
  norm[i] = deg_inv_sqrt(dh[src32[i]]) * deg_inv_sqrt(dh[dst32[i] + dst_offset]);   // model.py:78
}

__global__ void histogram(const int32_t* __restrict__ key, int64_t n, int32_t* cnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[key[i]], 1);
}

__global__ void add_rows(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n, int32_t* out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// key of the type-sorted pass when it is blocked by subject: (block of the edge's subject row) * n_types + type.  The
// subject (src of an in-half edge, dst of its reverse) is the endpoint whose row the d_rel pass gathers at random - the
// object side of a knowledge graph is hub-heavy and stays in L2 by itself.  Inside one block the gathered x / g rows
// (2 * block_rows * 4 D bytes) are L2-resident; the per-(block, type) partial rows are summed afterwards (kgc_block_sum).
__global__ void type_block_key(const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32, int32_t* __restrict__ type32,
                               int64_t n2, int64_t n_in, int32_t n_types, int32_t block_rows) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const int32_t subj = i < n_in ? src32[i] : dst32[i];
  type32[i] = (subj / block_rows) * n_types + type32[i];
}

__global__ void gather_records(const int32_t* __restrict__ perm, const int32_t* __restrict__ a32,
                               const int32_t* __restrict__ b32, const float* __restrict__ norm, int64_t n2,
                               kgc_edge_rec_t* __restrict__ rec) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n2) return;
  int32_t e = perm[p];
  kgc_edge_rec_t r;
  r.eid = e;
  r.a = a32[e];
  r.b = b32[e];
  r.norm = norm[e];
  rec[p] = r;
}

int sorted_csr(const int32_t* key32, int64_t n2, int64_t n_rows, const int32_t* iota, int32_t* key_sorted,
               int32_t* perm, int32_t* rowptr, int32_t* cnt, void* cub_ws, size_t cub_bytes, cudaStream_t st) {
  const int grid = (int)ceil_div(n2, kThreads);
  KGC_CUDA_TRY(cudaMemsetAsync(cnt, 0, size_t(n_rows + 1) * 4, st));
  if (n2 > 0) {
    histogram<<<grid, kThreads, 0, st>>>(key32, n2, cnt);
    KGC_LAUNCH_CHECK();
  }
  size_t tb = cub_bytes;
  KGC_CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_ws, tb, cnt, rowptr, (int)(n_rows + 1), st));
  if (n2 > 0) {
    tb = cub_bytes;
    KGC_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_ws, tb, key32, key_sorted, iota, perm, (int)n2, 0,
                                                 key_bits(n_rows), st));
  }
  return 0;
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int64_t kgc_csr_type_rows(int64_t n_nodes, int64_t n_types, int64_t type_block_rows) {
  return type_block_rows > 0 ? ceil_div(n_nodes, type_block_rows) * n_types : n_types;
}

extern "C" size_t kgc_csr_workspace_bytes(int64_t n_edges2, int64_t n_nodes, int64_t n_types) {
  size_t cub_bytes = 0;
  if (cub_temp_bytes(n_edges2, n_nodes > n_types ? n_nodes : n_types, &cub_bytes) != cudaSuccess) {
    set_error("kgc_csr_workspace_bytes: CUB size query failed (no CUDA device?)");
    return 0;
  }
  return make_layout(n_edges2, n_nodes, n_types, cub_bytes).total;
}

extern "C" int kgc_csr_build(const int64_t* src, const int64_t* dst, const int64_t* type, int64_t n2, int64_t n_in,
                             int64_t n_nodes, int64_t n_dst_rows, int64_t dst_offset, int64_t n_types,
                             int32_t deg_given, int32_t* deg, float* norm, int32_t* perm_dst, int32_t* rowptr_dst,
                             int32_t* rowmid_dst, kgc_edge_rec_t* rec_dst, int32_t* perm_src, int32_t* rowptr_src,
                             kgc_edge_rec_t* rec_src, int32_t* perm_type, int32_t* rowptr_type, kgc_edge_rec_t* rec_type,
                             int64_t type_block_rows, void* workspace, size_t workspace_bytes, void* stream) {
  KGC_REQUIRE(n2 >= 0 && n_in >= 0 && n_in <= n2, "n_edges_in must lie in [0, n_edges2] (in half first, model.py:84-90)");
  KGC_REQUIRE(n_nodes > 0 && n_types > 0 && n_dst_rows > 0, "empty node or type set");
  KGC_REQUIRE(dst_offset >= 0 && dst_offset + n_dst_rows <= n_nodes, "destination range must lie inside the node set");
  KGC_REQUIRE(n2 < (int64_t(1) << 31) && n_nodes < (int64_t(1) << 31) && n_types < (int64_t(1) << 31),
              "ids must fit int32");
  // type rows of the (optionally subject-blocked) type sort: kgc_csr_type_rows(n_nodes, n_types, type_block_rows)
  KGC_REQUIRE(type_block_rows >= 0, "type_block_rows must be >= 0");
  const int64_t n_type_rows = type_block_rows > 0 ? ceil_div(n_nodes, type_block_rows) * n_types : n_types;
  KGC_REQUIRE(n_type_rows < (int64_t(1) << 30), "too many (block, type) rows");
  size_t cub_bytes = 0;
  KGC_CUDA_TRY(cub_temp_bytes(n2, n_nodes > n_type_rows ? n_nodes : n_type_rows, &cub_bytes));
  const Layout L = make_layout(n2, n_nodes, n_type_rows, cub_bytes);
  KGC_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, "workspace too small");
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  int32_t* src32 = (int32_t*)(ws + L.src32);
  int32_t* dst32 = (int32_t*)(ws + L.dst32);
  int32_t* type32 = (int32_t*)(ws + L.type32);
  int32_t* iota = (int32_t*)(ws + L.iota);
  int32_t* key_sorted = (int32_t*)(ws + L.key_sorted);
  int32_t* cnt = (int32_t*)(ws + L.cnt);
  int32_t* cnt2 = (int32_t*)(ws + L.cnt2);
  int32_t* flag = (int32_t*)(ws + L.flag);
  void* cub_ws = ws + L.cub;
  const int grid = (int)ceil_div(n2 > 0 ? n2 : 1, kThreads);
  const int grid_n = (int)ceil_div(n_dst_rows, kThreads);

  KGC_CUDA_TRY(cudaMemsetAsync(flag, 0, 4, st));
  if (!deg_given) KGC_CUDA_TRY(cudaMemsetAsync(deg, 0, size_t(2 * n_nodes) * 4, st));
  if (n2 > 0) {
    narrow_and_check<<<grid, kThreads, 0, st>>>(src, dst, type, n2, n_nodes, n_dst_rows, n_types, src32, dst32, type32,
                                                iota, flag);
    KGC_LAUNCH_CHECK();
  }
  int32_t bad = 0;
  KGC_CUDA_TRY(cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, st));
  KGC_CUDA_TRY(cudaStreamSynchronize(st));
  KGC_REQUIRE(bad == 0, "edge list holds a node or type id out of range");

  if (n2 > 0) {
    if (!deg_given) {
      half_degree<<<grid, kThreads, 0, st>>>(src32, n2, n_in, n_nodes, deg);
      KGC_LAUNCH_CHECK();
    }
    edge_norm<<<grid, kThreads, 0, st>>>(src32, dst32, deg, n2, n_in, n_nodes, dst_offset, norm);
    KGC_LAUNCH_CHECK();
  }
  // dst-sorted (forward)
  if (sorted_csr(dst32, n2, n_dst_rows, iota, key_sorted, perm_dst, rowptr_dst, cnt, cub_ws, cub_bytes, st)) return 1;
  KGC_CUDA_TRY(cudaMemsetAsync(cnt2, 0, size_t(n_dst_rows) * 4, st));
  if (n_in > 0) {
    histogram<<<(int)ceil_div(n_in, kThreads), kThreads, 0, st>>>(dst32, n_in, cnt2);   // in-half edges per dst
    KGC_LAUNCH_CHECK();
  }
  add_rows<<<grid_n, kThreads, 0, st>>>(rowptr_dst, cnt2, n_dst_rows, rowmid_dst);
  KGC_LAUNCH_CHECK();
  if (n2 > 0) {
    gather_records<<<grid, kThreads, 0, st>>>(perm_dst, src32, type32, norm, n2, rec_dst);
    KGC_LAUNCH_CHECK();
  }
  // src-sorted (backward d_x / d_ee)
  if (sorted_csr(src32, n2, n_nodes, iota, key_sorted, perm_src, rowptr_src, cnt, cub_ws, cub_bytes, st)) return 1;
  if (n2 > 0) {
    gather_records<<<grid, kThreads, 0, st>>>(perm_src, dst32, type32, norm, n2, rec_src);
    KGC_LAUNCH_CHECK();
  }
  // type-sorted (backward d_rel), optionally blocked by subject row
  if (type_block_rows > 0 && n2 > 0) {
    type_block_key<<<grid, kThreads, 0, st>>>(src32, dst32, type32, n2, n_in, (int32_t)n_types, (int32_t)type_block_rows);
    KGC_LAUNCH_CHECK();
  }
  if (sorted_csr(type32, n2, n_type_rows, iota, key_sorted, perm_type, rowptr_type, cnt, cub_ws, cub_bytes, st)) return 1;
  if (n2 > 0) {
    gather_records<<<grid, kThreads, 0, st>>>(perm_type, src32, dst32, norm, n2, rec_type);
    KGC_LAUNCH_CHECK();
  }
  return 0;
}

// Shared helpers for libkgc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "kgc_b200.h"

namespace kgc {

void set_error(const std::string& msg);

inline int fail(const char* where, const std::string& what) {
  set_error(std::string(where) + ": " + what);
  return 1;
}

#define KGC_CUDA_TRY(expr)                                                          \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) return ::kgc::fail(__func__, std::string(#expr) + " -> " + cudaGetErrorString(_e)); \
  } while (0)

#define KGC_REQUIRE(cond, msg)                          \
  do {                                                  \
    if (!(cond)) return ::kgc::fail(__func__, (msg));   \
  } while (0)

#define KGC_LAUNCH_CHECK() KGC_CUDA_TRY(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE setting of a kernel: remember, per device, the largest value
// this process has set, so that the attribute call stays off the launch path without breaking a process that drives several
// GPUs.  (Concurrent first launches from two host threads may both set it: harmless.)
struct SmemAttrCache {
  size_t set[64] = {0};
  template <typename Kernel>
  cudaError_t ensure(Kernel kern, size_t smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && smem <= set[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && tracked) set[dev] = smem;
    return e;
  }
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// 128-bit streaming load that does not allocate in L1 (the edge-embedding stream is read once).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}
__device__ __forceinline__ int4 ld_rec(const kgc_edge_rec_t* p) {
  return __ldg(reinterpret_cast<const int4*>(p));
}

// Counter-based dropout: Philox4x32-10 keyed by the step's seed, counter = (float4 index, plane).  The forward and
// the backward regenerate the same keep mask from (seed, index), so no mask tensor is written or read.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// keep-scale per lane of one float4: s where the 32-bit draw >= drop_thr (= p * 2^32), else 0
__device__ __forceinline__ float4 philox_mask4(int64_t idx4, uint32_t plane, uint64_t seed, uint32_t drop_thr, float s) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), plane, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return make_float4(r.x >= drop_thr ? s : 0.f, r.y >= drop_thr ? s : 0.f, r.z >= drop_thr ? s : 0.f,
                     r.w >= drop_thr ? s : 0.f);
}

// cuTensorMapEncodeTiled is a DRIVER entry point: it needs a context current on the calling thread.  A thread that has
// made no runtime call yet (autograd's backward thread when the first thing it runs is one of the GEMM launchers) has
// none and the encode fails with CUDA_ERROR_INVALID_CONTEXT (201); one runtime call per thread binds the primary context.
inline void ensure_thread_context() {
  static thread_local bool done = false;
  if (!done) {
    cudaFree(nullptr);
    done = true;
  }
}

}  // namespace kgc

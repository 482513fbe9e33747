// Streaming-aggregation schedules on the device (the O(2E) part of plan.py's build_stream_plan): per sorted record its
// row + first / last flags, per 32-record chunk whether it needs a head / tail carry row.  The host keeps only the
// O(#rows that span chunks) part (slot numbering is a prefix sum, the fix-up levels a few vectorised numpy lines).
// Integer work, bit-exact against the numpy restatement (tests/test_gpu_conv.py::test_stream_plan_on_device).
#include "common.cuh"

namespace kgc {
namespace {

constexpr int kThreads = 256;

// segments [beg, end) tile [0, n_rec) in order (empty segments allowed): the segment of record p is the LAST one with
// beg <= p (an empty segment that starts at p precedes the non-empty one that starts there)
__global__ void stream_flags_kernel(const int32_t* __restrict__ seg_beg, const int32_t* __restrict__ seg_end,
                                    const int32_t* __restrict__ seg_row, int64_t n_seg, int64_t n_rec,
                                    uint32_t* __restrict__ rowflags, int32_t* __restrict__ rec_seg) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_rec) return;
  int64_t lo = 0, hi = n_seg;                      // first segment with beg > p
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(seg_beg + mid) <= p) lo = mid + 1; else hi = mid;
  }
  const int64_t s = lo - 1;
  uint32_t f = (uint32_t)__ldg(seg_row + s);
  if ((int64_t)__ldg(seg_beg + s) == p) f |= 1u << 30;
  if ((int64_t)__ldg(seg_end + s) - 1 == p) f |= 1u << 31;
  rowflags[p] = f;
  rec_seg[p] = (int32_t)s;
}

// inter[2c] = chunk c needs a HEAD carry row (its first segment began earlier and ends inside it), inter[2c + 1] = a TAIL
// carry row (its last segment is still open at the end of the chunk)
__global__ void chunk_flags_kernel(const int32_t* __restrict__ seg_beg, const int32_t* __restrict__ seg_end,
                                   const int32_t* __restrict__ rec_seg, int64_t n_rec, int32_t chunk, int64_t n_chunks,
                                   int32_t* __restrict__ inter) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const int64_t cb = c * chunk, ce = cb + chunk < n_rec ? cb + chunk : n_rec;
  const int32_t lead = rec_seg[cb], trail = rec_seg[ce - 1];
  inter[2 * c] = ((int64_t)seg_beg[lead] < cb && (int64_t)seg_end[lead] <= ce) ? 1 : 0;
  inter[2 * c + 1] = ((int64_t)seg_end[trail] > ce) ? 1 : 0;
}

}  // namespace
}  // namespace kgc

using namespace kgc;

extern "C" int kgc_stream_plan_flags(const int32_t* seg_beg, const int32_t* seg_end, const int32_t* seg_row, int64_t n_seg,
                                     int64_t n_rec, int32_t chunk, uint32_t* rowflags, int32_t* rec_seg, int32_t* inter,
                                     void* stream) {
  KGC_REQUIRE(n_seg > 0 && n_rec >= 0 && chunk > 0, "bad sizes");
  if (n_rec == 0) return 0;
  KGC_REQUIRE(seg_beg && seg_end && seg_row && rowflags && rec_seg && inter, "null buffer");
  cudaStream_t st = as_stream(stream);
  stream_flags_kernel<<<(unsigned)ceil_div(n_rec, kThreads), kThreads, 0, st>>>(seg_beg, seg_end, seg_row, n_seg, n_rec, rowflags,
                                                                            rec_seg);
  KGC_LAUNCH_CHECK();
  const int64_t n_chunks = ceil_div(n_rec, chunk);
  chunk_flags_kernel<<<(unsigned)ceil_div(n_chunks, kThreads), kThreads, 0, st>>>(seg_beg, seg_end, rec_seg, n_rec, chunk, n_chunks,
                                                                              inter);
  KGC_LAUNCH_CHECK();
  return 0;
}

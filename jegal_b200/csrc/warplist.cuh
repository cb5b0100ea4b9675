// Shared by K2 (topk.cu) and the fused K2 + NVLink exchange (exchange.cu): the ordering rule, the warp-wide
// sorted list and the streaming scan of one score row.  One definition, so a sharded gallery merges to exactly
// the single-GPU list by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jegal {
namespace k2 {

constexpr int kTopkWarps = 4;

__device__ __forceinline__ bool better(float av, int32_t ai, float bv, int32_t bi) {
  return av > bv || (av == bv && ai < bi);
}

// A warp holds a descending list of 32 (value, index) items, one per lane.
struct WarpList {
  float v;
  int32_t i;
  __device__ __forceinline__ void init() {
    v = -INFINITY;
    i = 0x7fffffff;
  }
  // insert (cv, ci) — warp-uniform arguments — which must beat lane 31's item
  __device__ __forceinline__ void insert(float cv, int32_t ci, int lane) {
    const bool worse = better(cv, ci, v, i);
    const uint32_t wm = __ballot_sync(0xffffffffu, worse);
    const int pos = __ffs(wm) - 1;
    const float upv = __shfl_up_sync(0xffffffffu, v, 1);
    const int32_t upi = __shfl_up_sync(0xffffffffu, i, 1);
    if (pos >= 0) {
      if (lane > pos) {
        v = upv;
        i = upi;
      } else if (lane == pos) {
        v = cv;
        i = ci;
      }
    }
  }
  // like offer(), with the current k-th item (tv, ti) cached by the caller: an empty vote costs one ballot
  __device__ __forceinline__ void offer_cached(float cv, int32_t ci, bool valid, int k, int lane, float& tv, int32_t& ti) {
    uint32_t m = __ballot_sync(0xffffffffu, valid && better(cv, ci, tv, ti));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float bv = __shfl_sync(0xffffffffu, cv, src);
      const int32_t bi = __shfl_sync(0xffffffffu, ci, src);
      if (better(bv, bi, tv, ti)) {
        insert(bv, bi, lane);
        tv = __shfl_sync(0xffffffffu, v, k - 1);
        ti = __shfl_sync(0xffffffffu, i, k - 1);
      }
    }
  }
  // offer one candidate per lane; candidates that beat the current k-th item get inserted
  __device__ __forceinline__ void offer(float cv, int32_t ci, bool valid, int k, int lane) {
    float tv = __shfl_sync(0xffffffffu, v, k - 1);
    int32_t ti = __shfl_sync(0xffffffffu, i, k - 1);
    uint32_t m = __ballot_sync(0xffffffffu, valid && better(cv, ci, tv, ti));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float bv = __shfl_sync(0xffffffffu, cv, src);
      const int32_t bi = __shfl_sync(0xffffffffu, ci, src);
      if (better(bv, bi, tv, ti)) {
        insert(bv, bi, lane);
        tv = __shfl_sync(0xffffffffu, v, k - 1);
        ti = __shfl_sync(0xffffffffu, i, k - 1);
      }
    }
  }
};


// Stream float4 elements [lo4, n4) of `row4` (this block's share of a 16-byte aligned score row) into the
// warp's list.  kU independent 16-byte loads per lane per batch, and the NEXT batch is already in flight while
// this one is examined (software pipelining); a whole batch is skipped with one vote when nothing in it can
// enter the current top-k.
__device__ __forceinline__ void scan_row4(WarpList& L, const float4* __restrict__ row4, int32_t lo4, int32_t n4, int32_t k,
                                          int warp, int lane) {
  constexpr int kU = 4;
  constexpr int kStride = kTopkWarps * 32 * kU;
  const float4 kNegInf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  float4 x[kU], nx[kU];
  int32_t base = lo4 + warp * 32 * kU;
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int32_t j4 = base + u * 32 + lane;
    nx[u] = j4 < n4 ? __ldg(row4 + j4) : kNegInf4;
  }
  for (; base < n4; base += kStride) {
#pragma unroll
    for (int u = 0; u < kU; ++u) x[u] = nx[u];
    const int32_t nbase = base + kStride;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int32_t j4 = nbase + u * 32 + lane;
      nx[u] = j4 < n4 ? __ldg(row4 + j4) : kNegInf4;
    }
    float tv = __shfl_sync(0xffffffffu, L.v, k - 1);
    int32_t ti = __shfl_sync(0xffffffffu, L.i, k - 1);
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < kU; ++u) mx = fmaxf(mx, fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)));
    if (!__any_sync(0xffffffffu, mx >= tv)) continue;  // >= : an equal value with a lower index still wins
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int32_t j4 = base + u * 32 + lane;
      const bool valid = j4 < n4;
      const int32_t j = j4 * 4;
      const float um = fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w));
      if (!__any_sync(0xffffffffu, valid && um >= tv)) continue;
      L.offer_cached(x[u].x, j + 0, valid, k, lane, tv, ti);
      L.offer_cached(x[u].y, j + 1, valid, k, lane, tv, ti);
      L.offer_cached(x[u].z, j + 2, valid, k, lane, tv, ti);
      L.offer_cached(x[u].w, j + 3, valid, k, lane, tv, ti);
    }
  }
}

// Scalar tail [from, n_g) of a row (and whole rows that are not 16-byte aligned).
__device__ __forceinline__ void scan_row_tail(WarpList& L, const float* __restrict__ row, int32_t from, int32_t n_g, int32_t k,
                                              int warp, int lane) {
  for (int32_t b2 = from + warp * 32; b2 < n_g; b2 += kTopkWarps * 32) {
    const int32_t j = b2 + lane;
    const bool valid = j < n_g;
    const float xv = valid ? __ldg(row + j) : 0.f;
    L.offer(xv, j, valid, k, lane);
  }
}

}  // namespace k2
}  // namespace jegal

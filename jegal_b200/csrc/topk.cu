// K2 — per-query top-k, k-way merge of per-shard lists, and rank-of-positive.
// These replace the full np.sort(-x, axis=1) + diagonal lookup of
// compute_metrics (evaluation/evaluate_retrieval.py:51-65).  All three stream a
// dense fp32 score matrix once with 16-byte loads: HBM-bound.
//
// Ordering rule everywhere: descending score, ties towards the LOWER index, so a
// gallery sharded over several GPUs merges to exactly the single-GPU list.
#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "warplist.cuh"

namespace jegal {
namespace {

using namespace k2;

// A query row is cut into gridDim.y column slices (16-byte aligned); block (q, s) streams slice s of
// row q.  With more than one slice the block's list goes to a workspace and the LAST block of the
// row to finish (atomic ticket) merges the slices — short serial chains per block and enough blocks
// to keep HBM busy even for ~1000 rows on 148 SMs.
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_kernel(const float* __restrict__ scores, int32_t n_g, int64_t ld, int32_t k, int32_t idx_offset,
            float* __restrict__ out_val, int32_t* __restrict__ out_idx, float* __restrict__ ws_val,
            int32_t* __restrict__ ws_idx, uint32_t* __restrict__ ws_ticket) {
  __shared__ float sv[kTopkWarps][32];
  __shared__ int32_t si[kTopkWarps][32];
  __shared__ uint32_t s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = blockIdx.x;
  const int32_t n_slices = gridDim.y, slice = blockIdx.y;
  const float* row = scores + q * ld;
  WarpList L;
  L.init();
  // vector body: needs a 16-byte aligned row start
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
  const int32_t n4_all = vec_ok ? (n_g >> 2) : 0;
  const int32_t per = (n4_all + n_slices - 1) / n_slices;
  const int32_t lo4 = min(slice * per, n4_all);
  const int32_t n4 = min(lo4 + per, n4_all);  // this block streams float4 indices [lo4, n4)
  const float4* row4 = reinterpret_cast<const float4*>(row);
  scan_row4(L, row4, lo4, n4, k, warp, lane);
  if (slice == n_slices - 1) scan_row_tail(L, row, n4_all * 4, n_g, k, warp, lane);  // and unaligned rows
  sv[warp][lane] = L.v;
  si[warp][lane] = L.i;
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < kTopkWarps; ++w) L.offer(sv[w][lane], si[w][lane], lane < k, k, lane);
  if (n_slices > 1) {
    // publish this slice's list, take a ticket; the last slice of the row merges all of them
    const int64_t wbase = (q * n_slices + slice) * 32;
    ws_val[wbase + lane] = L.v;
    ws_idx[wbase + lane] = L.i;
    __threadfence();
    __syncwarp();
    if (lane == 0) s_last = atomicAdd(ws_ticket + q, 1u);
    __syncwarp();
    if (s_last != static_cast<uint32_t>(n_slices - 1)) return;
    __threadfence();
    for (int32_t s2 = 0; s2 < n_slices; ++s2) {
      if (s2 == slice) continue;
      const int64_t ob = (q * n_slices + s2) * 32;
      const float cv = __ldcg(ws_val + ob + lane);
      const int32_t ci = __ldcg(ws_idx + ob + lane);
      L.offer(cv, ci, lane < k && ci != 0x7fffffff, k, lane);
    }
    if (lane == 0) ws_ticket[q] = 0;  // ready for the next launch
  }
  if (lane < k) {
    const bool real = L.i != 0x7fffffff;
    out_val[q * k + lane] = L.v;
    out_idx[q * k + lane] = real ? L.i + idx_offset : -1;
  }
}

// one warp per query: merge n_lists lists of k
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idxs, int32_t n_lists,
                  int32_t n_q, int32_t k, float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_q) return;
  WarpList L;
  L.init();
  for (int32_t l = 0; l < n_lists; ++l) {
    const int64_t off = (static_cast<int64_t>(l) * n_q + q) * k;
    const bool valid = lane < k;
    float cv = valid ? __ldg(vals + off + lane) : 0.f;
    int32_t ci = valid ? __ldg(idxs + off + lane) : 0;
    const bool real = valid && ci >= 0;
    L.offer(cv, ci, real, k, lane);
  }
  if (lane < k) {
    const bool real = L.i != 0x7fffffff;
    out_val[static_cast<int64_t>(q) * k + lane] = L.v;
    out_idx[static_cast<int64_t>(q) * k + lane] = real ? L.i : -1;
  }
}

// one block per row: counts of entries strictly greater than / equal to the positive
__global__ void __launch_bounds__(256)
rank_kernel(const float* __restrict__ scores, int32_t n_g, int64_t ld_row, int64_t ld_col,
            const int32_t* __restrict__ gt, int32_t* __restrict__ n_greater, int32_t* __restrict__ n_equal) {
  __shared__ int32_t sg[8], se[8];
  const int64_t q = blockIdx.x;
  const int32_t g = gt ? __ldg(gt + q) : static_cast<int32_t>(q);
  const float* row = scores + q * ld_row;
  const float pos = __ldg(row + static_cast<int64_t>(g) * ld_col);
  int32_t cg = 0, ce = 0;
  for (int32_t j = threadIdx.x; j < n_g; j += blockDim.x) {
    const float x = __ldg(row + static_cast<int64_t>(j) * ld_col);
    cg += x > pos;
    ce += x == pos;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cg += __shfl_xor_sync(0xffffffffu, cg, o);
    ce += __shfl_xor_sync(0xffffffffu, ce, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sg[warp] = cg;
    se[warp] = ce;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t tg = 0, te = 0;
    for (int w = 0; w < 8; ++w) {
      tg += sg[w];
      te += se[w];
    }
    n_greater[q] = tg;
    if (n_equal) n_equal[q] = te;
  }
}

// The same counts for a matrix whose QUERY index is the contiguous one (x.t() of a row-major matrix: ld_row == 1):
// a block takes 32 consecutive queries, lane = query, and its 8 warps walk the candidates j = warp, warp + 8, ...
// so every load is one coalesced 128-byte line (rank_kernel would read with stride ld_col per lane).
__global__ void __launch_bounds__(256)
rank_kernel_qmajor(const float* __restrict__ scores, int32_t n_q, int32_t n_g, int64_t ld_col,
                   const int32_t* __restrict__ gt, int32_t* __restrict__ n_greater, int32_t* __restrict__ n_equal) {
  __shared__ int32_t sg[8][32], se[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 32 + lane;
  const bool ok = q < n_q;
  const int32_t g = ok ? (gt ? __ldg(gt + q) : static_cast<int32_t>(q)) : 0;
  const float pos = ok ? __ldg(scores + q + static_cast<int64_t>(g) * ld_col) : 0.f;
  int32_t cg = 0, ce = 0;
  int32_t j = warp;
  for (; j + 24 < n_g; j += 32) {  // four independent loads in flight per lane
    float x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = ok ? __ldg(scores + q + static_cast<int64_t>(j + 8 * u) * ld_col) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      cg += x[u] > pos;
      ce += x[u] == pos;
    }
  }
  for (; j < n_g; j += 8) {
    const float x = ok ? __ldg(scores + q + static_cast<int64_t>(j) * ld_col) : 0.f;
    cg += x > pos;
    ce += x == pos;
  }
  sg[warp][lane] = cg;
  se[warp][lane] = ce;
  __syncthreads();
  if (warp == 0 && ok) {
    int32_t tg = 0, te = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      tg += sg[w][lane];
      te += se[w][lane];
    }
    n_greater[q] = tg;
    if (n_equal) n_equal[q] = te;
  }
}

}  // namespace

int launch_topk(jegal_ctx* ctx, const float* scores, int32_t n_q, int32_t n_g, int64_t ld, int32_t k,
                int32_t idx_offset, float* out_val, int32_t* out_idx, cudaStream_t stream) {
  if (n_q <= 0) return JEGAL_OK;
  // rows are only sliced when there are too few of them to fill the GPU (every slice pays the start-up
  // phase of an empty list again); at least 16 K floats per block
  const char* cap_env = std::getenv("JEGAL_TOPK_BLOCKS_PER_SM");  // tuning knob; default 8
  const int64_t per_sm = cap_env && *cap_env ? std::max(1, std::atoi(cap_env)) : 8;
  const int64_t capacity = static_cast<int64_t>(ctx->sm_count) * per_sm;
  int64_t slices = capacity / n_q;
  slices = std::min<int64_t>(slices, std::max<int64_t>(1, n_g / 16384));
  slices = std::max<int64_t>(1, std::min<int64_t>(slices, 64));
  if (slices > 1) {
    const size_t need = static_cast<size_t>(n_q) * slices * 32;
    if (ctx->topk_ws_elems < need || ctx->topk_ws_rows < static_cast<size_t>(n_q)) {
      // one-time (re)allocation; stream-ordered use afterwards
      JEGAL_CUDA_OK(ctx, cudaStreamSynchronize(stream));
      if (ctx->topk_ws_val) cudaFree(ctx->topk_ws_val);
      if (ctx->topk_ws_idx) cudaFree(ctx->topk_ws_idx);
      if (ctx->topk_ws_ticket) cudaFree(ctx->topk_ws_ticket);
      ctx->topk_ws_val = nullptr; ctx->topk_ws_idx = nullptr; ctx->topk_ws_ticket = nullptr;
      ctx->topk_ws_elems = 0; ctx->topk_ws_rows = 0;
      JEGAL_CUDA_OK(ctx, cudaMalloc(&ctx->topk_ws_val, need * sizeof(float)));
      JEGAL_CUDA_OK(ctx, cudaMalloc(&ctx->topk_ws_idx, need * sizeof(int32_t)));
      JEGAL_CUDA_OK(ctx, cudaMalloc(&ctx->topk_ws_ticket, static_cast<size_t>(n_q) * sizeof(uint32_t)));
      JEGAL_CUDA_OK(ctx, cudaMemset(ctx->topk_ws_ticket, 0, static_cast<size_t>(n_q) * sizeof(uint32_t)));
      ctx->topk_ws_elems = need;
      ctx->topk_ws_rows = static_cast<size_t>(n_q);
    }
  }
  const dim3 grid(static_cast<unsigned>(n_q), static_cast<unsigned>(slices));
  topk_kernel<<<grid, kTopkWarps * 32, 0, stream>>>(scores, n_g, ld, k, idx_offset, out_val, out_idx, ctx->topk_ws_val,
                                                   ctx->topk_ws_idx, ctx->topk_ws_ticket);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

int launch_topk_merge(jegal_ctx* ctx, const float* vals, const int32_t* idxs, int32_t n_lists,
                      int32_t n_q, int32_t k, float* out_val, int32_t* out_idx, cudaStream_t stream) {
  if (n_q <= 0) return JEGAL_OK;
  const int warps = 4;
  const unsigned blocks = static_cast<unsigned>((n_q + warps - 1) / warps);
  topk_merge_kernel<<<blocks, warps * 32, 0, stream>>>(vals, idxs, n_lists, n_q, k, out_val, out_idx);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

int launch_rank_of_positive(jegal_ctx* ctx, const float* scores, int32_t n_q, int32_t n_g,
                            int64_t ld_row, int64_t ld_col, const int32_t* gt, int32_t* n_greater,
                            int32_t* n_equal, cudaStream_t stream) {
  if (n_q <= 0) return JEGAL_OK;
  if (ld_row == 1 && ld_col != 1) {
    rank_kernel_qmajor<<<static_cast<unsigned>((n_q + 31) / 32), 256, 0, stream>>>(scores, n_q, n_g, ld_col, gt, n_greater,
                                                                                 n_equal);
  } else {
    rank_kernel<<<static_cast<unsigned>(n_q), 256, 0, stream>>>(scores, n_g, ld_row, ld_col, gt, n_greater, n_equal);
  }
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // namespace jegal

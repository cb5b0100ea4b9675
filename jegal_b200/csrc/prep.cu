// K0 — one pass over the embedding rows: L2-normalise in fp32, cast to the 16-bit
// operand type the tensor cores consume, and (optionally) produce the per-clip
// 1/||mean row|| scale that turns mean/mean pooling into the reference's
// "cosine of mean-pooled embeddings" (evaluation/evaluate_retrieval.py:30-31,38-48;
// evaluation/evaluate_asd.py:31-36,43-47).  HBM-bound: every input byte is read
// once with 16-byte loads, every output byte written once with 16-byte stores.
#include "internal.h"
#include "rowio.cuh"

namespace jegal {
namespace {

constexpr int kPrepWarps = 4;

using namespace rowio;

// one block per clip; each warp walks the clip's rows with stride kPrepWarps
template <int kInDtype, int kOutDtype>
__global__ void __launch_bounds__(kPrepWarps * 32)
prep_kernel(const void* __restrict__ emb, const int32_t* __restrict__ cu, int32_t n_clips,
            int normalize_rows, float row_eps, float mean_eps, void* __restrict__ out,
            float* __restrict__ inv_meannorm, void* __restrict__ mean_rows) {
  __shared__ float colsum[kPrepWarps][kD];
  __shared__ float red[kPrepWarps];
  __shared__ float mean_s[kD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool want_mean = inv_meannorm != nullptr || mean_rows != nullptr;
  for (int32_t clip = blockIdx.x; clip < n_clips; clip += gridDim.x) {
    const int32_t r0 = __ldg(cu + clip), r1 = __ldg(cu + clip + 1);
    float cs[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cs[i] = 0.f;
    for (int32_t r = r0 + warp; r < r1; r += kPrepWarps) {
      float a[8], b[8];
      load8<kInDtype>(emb, r, lane * 8, a);
      load8<kInDtype>(emb, r, 256 + lane * 8, b);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ss += a[i] * a[i] + b[i] * b[i];
        cs[i] += a[i];
        cs[8 + i] += b[i];
      }
      if (normalize_rows) {
        ss = warp_sum(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), row_eps);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] *= inv;
          b[i] *= inv;
        }
      }
      if (out != nullptr) {  // null: statistics only (jegal_clip_means), the rows are read and nothing is written back
        store8<kOutDtype>(out, r, lane * 8, a);
        store8<kOutDtype>(out, r, 256 + lane * 8, b);
      }
    }
    if (want_mean) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        colsum[warp][lane * 8 + i] = cs[i];
        colsum[warp][256 + lane * 8 + i] = cs[8 + i];
      }
      __syncthreads();
      const float inv_len = 1.0f / static_cast<float>(max(r1 - r0, 1));
      float part = 0.f;
      for (int c = threadIdx.x; c < kD; c += kPrepWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kPrepWarps; ++w) s += colsum[w][c];
        float m = s * inv_len;
        // numpy's .mean(axis=0) of an fp16 array returns fp16 (fp32 accumulate): mirror that rounding
        if constexpr (kInDtype == JEGAL_F16) m = __half2float(__float2half_rn(m));
        mean_s[c] = m;
        part += m * m;
      }
      part = warp_sum(part);
      if (lane == 0) red[warp] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kPrepWarps; ++w) tot += red[w];
        const float inv = 1.0f / fmaxf(sqrtf(tot), mean_eps);
        if (inv_meannorm) inv_meannorm[clip] = inv;
        red[0] = inv;
      }
      __syncthreads();
      if (mean_rows) {
        const float inv = red[0];
        if (threadIdx.x < kD / 8) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = mean_s[threadIdx.x * 8 + i] * inv;
          store8<kOutDtype>(mean_rows, clip, threadIdx.x * 8, x);
        }
      }
      __syncthreads();
    }
  }
}

__global__ void rowinfo_kernel(const int32_t* __restrict__ cu, int32_t n_clips, int64_t rows,
                               int4* __restrict__ rowinfo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows; r += stride) {
    int32_t lo = 0, hi = n_clips;  // invariant: cu[lo] <= r < cu[hi]
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (__ldg(cu + mid) <= r) lo = mid; else hi = mid;
    }
    rowinfo[r] = make_int4(lo, __ldg(cu + lo), __ldg(cu + lo + 1), 0);
  }
}

// One warp per listed pair of 512-wide rows: dot product and both squared norms from one read of the two rows.
// normalize: score = a.b / (max(||a||, eps) max(||b||, eps)) (nn.CosineSimilarity, evaluate_asd.py:45-47), else a.b
template <int kDtype>
__global__ void __launch_bounds__(256)
pair_cosine_kernel(const void* __restrict__ a_rows, const void* __restrict__ b_rows, const int32_t* __restrict__ pair_a,
                   const int32_t* __restrict__ pair_b, int32_t n_pairs, int normalize, float eps,
                   float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int32_t pr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pr >= n_pairs) return;
  const int64_t ra = pair_a ? __ldg(pair_a + pr) : pr, rb = pair_b ? __ldg(pair_b + pr) : pr;
  float a0[8], a1[8], b0[8], b1[8];
  load8<kDtype>(a_rows, ra, lane * 8, a0);
  load8<kDtype>(a_rows, ra, 256 + lane * 8, a1);
  load8<kDtype>(b_rows, rb, lane * 8, b0);
  load8<kDtype>(b_rows, rb, 256 + lane * 8, b1);
  float dot = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dot += a0[i] * b0[i] + a1[i] * b1[i];
    sa += a0[i] * a0[i] + a1[i] * a1[i];
    sb += b0[i] * b0[i] + b1[i] * b1[i];
  }
  dot = warp_sum(dot);
  if (normalize) {
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    dot = dot / (fmaxf(sqrtf(sa), eps) * fmaxf(sqrtf(sb), eps));
  }
  if (lane == 0) scores[pr] = dot;
}

template <int kIn>
int launch_prep_in(jegal_ctx* ctx, int out_dtype, dim3 grid, cudaStream_t stream, const void* emb,
                   const int32_t* cu, int32_t n_clips, int normalize_rows, float row_eps,
                   float mean_eps, void* out, float* inv_meannorm, void* mean_rows) {
  if (out_dtype == JEGAL_BF16) {
    prep_kernel<kIn, JEGAL_BF16><<<grid, kPrepWarps * 32, 0, stream>>>(
        emb, cu, n_clips, normalize_rows, row_eps, mean_eps, out, inv_meannorm, mean_rows);
  } else if (out_dtype == JEGAL_F16) {
    prep_kernel<kIn, JEGAL_F16><<<grid, kPrepWarps * 32, 0, stream>>>(
        emb, cu, n_clips, normalize_rows, row_eps, mean_eps, out, inv_meannorm, mean_rows);
  } else if (out_dtype == JEGAL_F32 && out == nullptr) {  // fp32 is a mean-row format only (clip_means)
    prep_kernel<kIn, JEGAL_F32><<<grid, kPrepWarps * 32, 0, stream>>>(
        emb, cu, n_clips, normalize_rows, row_eps, mean_eps, out, inv_meannorm, mean_rows);
  } else {
    return set_err(ctx, JEGAL_ERR_ARG, "prep: out_dtype must be JEGAL_BF16 or JEGAL_F16");
  }
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // namespace

int launch_prep(jegal_ctx* ctx, const jegal_layout* layout, const void* emb, int in_dtype,
                int normalize_rows, float row_eps, float mean_eps, int out_dtype, void* out_rows,
                float* inv_meannorm, void* mean_rows, cudaStream_t stream) {
  if (layout->n_clips == 0) return JEGAL_OK;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 64;
  const dim3 grid(static_cast<unsigned>(layout->n_clips < cap ? layout->n_clips : cap));
  switch (in_dtype) {
    case JEGAL_F32:
      return launch_prep_in<JEGAL_F32>(ctx, out_dtype, grid, stream, emb, layout->cu_dev, layout->n_clips,
                                       normalize_rows, row_eps, mean_eps, out_rows, inv_meannorm, mean_rows);
    case JEGAL_F16:
      return launch_prep_in<JEGAL_F16>(ctx, out_dtype, grid, stream, emb, layout->cu_dev, layout->n_clips,
                                       normalize_rows, row_eps, mean_eps, out_rows, inv_meannorm, mean_rows);
    case JEGAL_BF16:
      return launch_prep_in<JEGAL_BF16>(ctx, out_dtype, grid, stream, emb, layout->cu_dev, layout->n_clips,
                                        normalize_rows, row_eps, mean_eps, out_rows, inv_meannorm, mean_rows);
    default:
      return set_err(ctx, JEGAL_ERR_ARG, "prep: bad in_dtype");
  }
}

int launch_pair_cosine(jegal_ctx* ctx, const void* a_rows, const void* b_rows, int dtype, const int32_t* pair_a,
                       const int32_t* pair_b, int32_t n_pairs, int normalize, float eps, float* scores,
                       cudaStream_t stream) {
  if (n_pairs <= 0) return JEGAL_OK;
  const unsigned grid = static_cast<unsigned>((n_pairs + 7) / 8);
  switch (dtype) {
    case JEGAL_F32:
      pair_cosine_kernel<JEGAL_F32><<<grid, 256, 0, stream>>>(a_rows, b_rows, pair_a, pair_b, n_pairs, normalize, eps, scores);
      break;
    case JEGAL_F16:
      pair_cosine_kernel<JEGAL_F16><<<grid, 256, 0, stream>>>(a_rows, b_rows, pair_a, pair_b, n_pairs, normalize, eps, scores);
      break;
    case JEGAL_BF16:
      pair_cosine_kernel<JEGAL_BF16><<<grid, 256, 0, stream>>>(a_rows, b_rows, pair_a, pair_b, n_pairs, normalize, eps, scores);
      break;
    default:
      return set_err(ctx, JEGAL_ERR_ARG, "pair_cosine: bad dtype");
  }
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

int launch_rowinfo(jegal_ctx* ctx, const int32_t* cu_dev, int32_t n_clips, int64_t rows,
                   int4* rowinfo_dev, cudaStream_t stream) {
  if (rows <= 0) return JEGAL_OK;
  const int threads = 256;
  int64_t blocks = (rows + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 32;
  if (blocks > cap) blocks = cap;
  rowinfo_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(cu_dev, n_clips, rows, rowinfo_dev);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // namespace jegal

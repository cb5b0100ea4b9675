// K5 — word-level mean pooling of the content branch: the reference's Python
// loops models/jegal.py:141-198 (sub-word tokens -> word, :173-179; audio frames -> word,
// :188-196) and :218-245 take, for every word, the mean of a contiguous range of feature rows
// (`x[b, start:end+1].mean(dim=0)`, ranges may overlap or leave gaps) and stack them.  Here one
// launch does every word of every clip of a batch: one warp per word, 16-byte loads along the
// feature dimension, fp32 accumulation, one rounding to the output type.  The output row stride
// and column offset are free, so the audio and text halves can be written straight into the
// concatenated [n_words, 512] fusion input (models/jegal.py:405-406) without a torch.cat.
// HBM-bound: sum(len) * D * b_in bytes read, n_words * D * b_out written.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace jegal {
namespace {

template <int kDtype>
__device__ __forceinline__ void ld8(const void* base, int64_t elem, float (&x)[8]) {
  if constexpr (kDtype == JEGAL_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + elem));
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f;
      if constexpr (kDtype == JEGAL_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      x[2 * i] = f.x; x[2 * i + 1] = f.y;
    }
  }
}

template <int kDtype>
__device__ __forceinline__ void st8(void* base, int64_t elem, const float (&x)[8]) {
  if constexpr (kDtype == JEGAL_F32) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + elem);
    p[0] = make_float4(x[0], x[1], x[2], x[3]);
    p[1] = make_float4(x[4], x[5], x[6], x[7]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (kDtype == JEGAL_F16) {
        const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
      } else {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + elem) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// One warp per segment; a lane owns 8 consecutive features per 256-feature step.
template <int kInDtype, int kOutDtype>
__global__ void __launch_bounds__(256)
segmean_kernel(const void* __restrict__ x, int32_t dim, const int32_t* __restrict__ seg_begin,
               const int32_t* __restrict__ seg_end, int32_t n_seg, void* __restrict__ out, int64_t ld_out,
               int32_t col_off) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t s = warp0; s < n_seg; s += nwarps) {
    const int32_t r0 = __ldg(seg_begin + s), r1 = __ldg(seg_end + s);
    const float inv = 1.0f / static_cast<float>(r1 - r0);
    for (int32_t c = lane * 8; c < dim; c += 256) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int32_t r = r0;
      for (; r + 3 < r1; r += 4) {  // four rows (16-byte loads) in flight per lane
        float a[8], b[8], c2[8], d[8];
        ld8<kInDtype>(x, static_cast<int64_t>(r) * dim + c, a);
        ld8<kInDtype>(x, static_cast<int64_t>(r + 1) * dim + c, b);
        ld8<kInDtype>(x, static_cast<int64_t>(r + 2) * dim + c, c2);
        ld8<kInDtype>(x, static_cast<int64_t>(r + 3) * dim + c, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += (a[i] + b[i]) + (c2[i] + d[i]);
      }
      for (; r < r1; ++r) {
        float a[8];
        ld8<kInDtype>(x, static_cast<int64_t>(r) * dim + c, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += a[i];
      }
      if (r1 - r0 > 1) {  // a one-row word is the row itself (models/jegal.py:176-179)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= inv;
      }
      st8<kOutDtype>(out, s * ld_out + col_off + c, acc);
    }
  }
}

template <int kIn>
int launch_out(jegal_ctx* ctx, int out_dtype, const void* x, int32_t dim, const int32_t* sb, const int32_t* se,
               int32_t n_seg, void* out, int64_t ld_out, int32_t col_off, cudaStream_t stream) {
  const int64_t want = (static_cast<int64_t>(n_seg) + 7) / 8;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(want, static_cast<int64_t>(ctx->sm_count) * 32));
  if (out_dtype == JEGAL_F32)
    segmean_kernel<kIn, JEGAL_F32><<<blocks, 256, 0, stream>>>(x, dim, sb, se, n_seg, out, ld_out, col_off);
  else if (out_dtype == JEGAL_F16)
    segmean_kernel<kIn, JEGAL_F16><<<blocks, 256, 0, stream>>>(x, dim, sb, se, n_seg, out, ld_out, col_off);
  else
    segmean_kernel<kIn, JEGAL_BF16><<<blocks, 256, 0, stream>>>(x, dim, sb, se, n_seg, out, ld_out, col_off);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // namespace

int launch_segmean(jegal_ctx* ctx, const void* x, int in_dtype, int32_t dim, const int32_t* seg_begin,
                   const int32_t* seg_end, int32_t n_seg, void* out, int out_dtype, int64_t ld_out,
                   int32_t col_off, cudaStream_t stream) {
  if (n_seg <= 0) return JEGAL_OK;
  if (in_dtype == JEGAL_F32)
    return launch_out<JEGAL_F32>(ctx, out_dtype, x, dim, seg_begin, seg_end, n_seg, out, ld_out, col_off, stream);
  if (in_dtype == JEGAL_F16)
    return launch_out<JEGAL_F16>(ctx, out_dtype, x, dim, seg_begin, seg_end, n_seg, out, ld_out, col_off, stream);
  return launch_out<JEGAL_BF16>(ctx, out_dtype, x, dim, seg_begin, seg_end, n_seg, out, ld_out, col_off, stream);
}

}  // namespace jegal

extern "C" int jegal_segment_mean(jegal_ctx* ctx, const void* x_dev, int in_dtype, int64_t rows, int32_t dim,
                                  const int32_t* seg_begin_dev, const int32_t* seg_end_dev, int32_t n_seg,
                                  void* out_dev, int out_dtype, int64_t ld_out, int32_t col_off, void* stream) {
  using namespace jegal;
  JEGAL_NVTX("jegal_segment_mean (K5)");
  if (!ctx) return JEGAL_ERR_ARG;
  if (n_seg < 0 || rows < 0) return set_err(ctx, JEGAL_ERR_ARG, "segment_mean: negative size");
  if (n_seg == 0) return JEGAL_OK;
  if (!x_dev || !seg_begin_dev || !seg_end_dev || !out_dev)
    return set_err(ctx, JEGAL_ERR_ARG, "segment_mean: null argument");
  if (in_dtype < JEGAL_F32 || in_dtype > JEGAL_BF16 || out_dtype < JEGAL_F32 || out_dtype > JEGAL_BF16)
    return set_err(ctx, JEGAL_ERR_ARG, "segment_mean: bad dtype");
  if (dim < 8 || (dim & 7) || (ld_out & 7) || (col_off & 7) || col_off < 0 || ld_out < static_cast<int64_t>(col_off) + dim)
    return set_err(ctx, JEGAL_ERR_ARG, "segment_mean: dim, ld_out and col_off must be multiples of 8 with col_off + dim <= ld_out");
  if ((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "segment_mean: buffers must be 16-byte aligned");
  return launch_segmean(ctx, x_dev, in_dtype, dim, seg_begin_dev, seg_end_dev, n_seg, out_dev, out_dtype, ld_out,
                        col_off, static_cast<cudaStream_t>(stream));
}

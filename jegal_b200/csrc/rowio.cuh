// Row I/O shared by K0 (prep.cu) and the fused normalise + NVLink all-gather (exchange.cu): 8 consecutive elements
// of a 512-wide row as fp32 (from fp32 / fp16 / bf16 storage), 8 fp32 values packed to the 16-bit operand types.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "internal.h"

namespace jegal {
namespace rowio {

template <int kInDtype>
__device__ __forceinline__ void load8(const void* base, int64_t row, int col, float (&x)[8]) {
  if constexpr (kInDtype == JEGAL_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + row * kD + col);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + row * kD + col));
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (kInDtype == JEGAL_F16) {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        const float2 f = __half22float2(h);
        x[2 * i] = f.x; x[2 * i + 1] = f.y;
      } else {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        const float2 f = __bfloat1622float2(h);
        x[2 * i] = f.x; x[2 * i + 1] = f.y;
      }
    }
  }
}

template <int kOutDtype>
__device__ __forceinline__ void store8(void* base, int64_t row, int col, const float (&x)[8]) {
  if constexpr (kOutDtype == JEGAL_F32) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + row * kD + col);
    p[0] = make_float4(x[0], x[1], x[2], x[3]);
    p[1] = make_float4(x[4], x[5], x[6], x[7]);
    return;
  }
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kOutDtype == JEGAL_F16) {
      const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + row * kD + col) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// 8 fp32 values -> one 16-byte vector of the 16-bit operand type
template <int kOutDtype>
__device__ __forceinline__ uint4 pack8(const float (&x)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kOutDtype == JEGAL_F16) {
      const __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace rowio
}  // namespace jegal

// K1 — all-pairs fused similarity + pooling for sm_100a.
//
//   S = R . C^T  (R: "row operand" [rowsR, 512], C: "column operand" [rowsC, 512],
//   both 16-bit, unit-norm rows, K-major) is formed tile by tile in tensor memory
//   and pooled in the epilogue; S itself never reaches shared or global memory.
//
//   out[rclip * ld_r + cclip * ld_c] =
//        rscale * cscale * ROWOP_{r in rclip} COLOP_{c in cclip} S[r, c]
//
// The host picks which of (gesture, content) is R and which is C so that the
// FIRST pooling reduction runs along columns: TMEM lane = row = one thread, so a
// column reduction is register-local (3-input FMNMX / FADD) and only the already
// reduced value crosses lanes (segmented warp shuffles).
//
// Structure (one CTA per SM, or one CTA pair per two SMs with cta_group::2):
//   - the column tile (128 rows of C per CTA, full K = 128 KB) is STATIONARY in
//     shared memory; row tiles of R stream through a kStages-deep TMA ring
//     (16 KB per k-block), so every MMA reads A and B from smem and L2->smem
//     traffic per MMA tile is halved with respect to streaming both operands;
//   - warp 0: TMA producer, warp 1: tcgen05.mma issuer, warp 2: TMEM allocator,
//     warps 4-7 / 8-11: two epilogue warpgroups (tcgen05.ld -> pooling -> global), so every
//     SM sub-partition has two epilogue warps whose TMEM-load / shuffle latencies interleave.
//     Fused modes: one warpgroup per TMEM accumulator buffer (every other tile each).
//     Two-pass mode: both warpgroups drain every tile, one half of the columns each;
//   - two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs
//     of tiles i+1 and i+2; stationary-tile k-blocks are released one by one during the
//     last row tile of a unit so the next column tile's load overlaps too;
//   - work is split over clusters by "row-tile steps" inside L2-sized phases of
//     R so that all CTAs stream the same slice of R at the same time;
//   - ragged column tiles: the MMA N of a unit is roundup16(columns its clips occupy), so
//     greedy whole-clip packing wastes no tensor time on the unused part of a 256-wide tile;
//   - two-pass mode (kRowOp == OP_NONE, chosen by the host for column sides made of many short
//     clips or of clips longer than a tile): the kernel only pools along columns and stores the
//     per-row values as M[column segment][row]; rowreduce_kernel (bottom of this file) combines
//     the pieces of a clip, reduces over rows and applies the scales.
#include <cstdio>

#include "internal.h"
#include "ptx.cuh"

namespace jegal {

using namespace ptx;

namespace {

constexpr uint32_t kTileBytes = kTileRows * kBlockK * 2;  // 16384
constexpr int kEpiWarp0 = 4;
constexpr int kEpiGroups = 2;
constexpr int kThreads = 32 * (kEpiWarp0 + 4 * kEpiGroups);  // 384

template <int kOp>
__device__ __forceinline__ float op_ident() {
  return kOp == OP_MAX ? -INFINITY : 0.0f;
}
template <int kOp>
__device__ __forceinline__ float op_apply(float a, float b) {
  static_assert(kOp == OP_MAX || kOp == OP_SUM, "OP_NONE has no combine step");
  return kOp == OP_MAX ? fmaxf(a, b) : a + b;
}

template <int kOp>
__device__ __forceinline__ float reduce8(const uint32_t* v) {
  if constexpr (kOp == OP_MAX) {
    float m1 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    float m2 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
    return fmax3(m1, m2, fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])));
  } else {
    float s0 = __uint_as_float(v[0]) + __uint_as_float(v[1]);
    float s1 = __uint_as_float(v[2]) + __uint_as_float(v[3]);
    float s2 = __uint_as_float(v[4]) + __uint_as_float(v[5]);
    float s3 = __uint_as_float(v[6]) + __uint_as_float(v[7]);
    return (s0 + s1) + (s2 + s3);
  }
}

// Work iterator shared by all warp roles: every role walks the same sequence of
// units (column tile ct, row tiles [rt0, rt0 + nrt)).
struct UnitIter {
  int32_t n_ctiles, n_rtiles, chunk;
  uint32_t cid, ncl;
  int32_t m0, cm;
  int64_t s, hi;
  __device__ UnitIter(const SimpoolParams& p, uint32_t cluster, uint32_t nclusters)
      : n_ctiles(p.n_ctiles), n_rtiles(p.n_rtiles), chunk(p.chunk_rtiles), cid(cluster),
        ncl(nclusters), m0(-p.chunk_rtiles), cm(1), s(0), hi(0) {}
  __device__ bool next(int32_t& ct, int32_t& rt0, int32_t& nrt) {
    while (s >= hi) {
      m0 += chunk;
      if (m0 >= n_rtiles) return false;
      cm = min(chunk, n_rtiles - m0);
      const int64_t steps = static_cast<int64_t>(n_ctiles) * cm;
      s = steps * cid / ncl;
      hi = steps * (cid + 1) / ncl;
    }
    ct = static_cast<int32_t>(s / cm);
    const int32_t mt = static_cast<int32_t>(s - static_cast<int64_t>(ct) * cm);
    const int64_t run_end = min(hi, static_cast<int64_t>(ct + 1) * cm);
    rt0 = m0 + mt;
    nrt = static_cast<int32_t>(run_end - s);
    s = run_end;
    return true;
  }
};

// cycle accounting for JEGAL_K1_TRACE=1 (p.trace != nullptr): slot += cycles spent in `stmt`
#define JEGAL_TRACED(slot, stmt)                         \
  do {                                                   \
    if (p.trace) {                                       \
      const long long t0__ = clock64();                  \
      stmt;                                              \
      tr[slot] += static_cast<unsigned long long>(clock64() - t0__); \
    } else {                                             \
      stmt;                                              \
    }                                                    \
  } while (0)

// Per-thread row context of one tile.  flags: bit k (k < 5): lane + 2^k is in the same row segment;
// kRcHead: first lane of its segment within the warp; kRcDirect: the whole clip lies inside this warp's
// 32 rows and the column tile is not a piece of a split clip (plain store instead of an atomic);
// kRcValid: the row exists.
struct RowCtx {
  int32_t rclip;
  float rscale;
  uint32_t flags;
  float* out_row;  // p.out + rclip * ld_r
};
constexpr uint32_t kRcHead = 1u << 5, kRcDirect = 1u << 6, kRcValid = 1u << 7;

// Reduce one pooled column-clip value over the rows of each row segment and write / combine it.
// Deliberately NOT inlined: there are ~40 emit sites in the unrolled pooling code and the epilogue
// was instruction-fetch bound (ncu: stall_no_inst + branch_resolving = 37 % of its samples with
// the body inlined, 7.7 k SASS instructions = 123 KB against a 32 KB L1.5 instruction cache).
#ifndef JEGAL_EMIT_INLINE
#define JEGAL_EMIT_ATTR __noinline__
#else
#define JEGAL_EMIT_ATTR __forceinline__
#endif
template <int kColOp, int kRowOp>
__device__ JEGAL_EMIT_ATTR void emit(float acc, int32_t cclip, float rscale, uint32_t flags, float* out_row,
                                  const float* __restrict__ cscale, const int32_t* __restrict__ cu_C,
                                  int32_t ld_c) {
  // (an integer REDUX for warps that lie inside one clip was measured: no gain, the shuffles are not the limit)
  float v = (flags & kRcValid) ? acc * rscale : op_ident<kRowOp>();
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float o = __shfl_down_sync(0xffffffffu, v, 1u << k);
    if ((flags >> k) & 1u) v = op_apply<kRowOp>(v, o);
  }
  if (flags & kRcHead) {
    float sc = cscale ? __ldg(cscale + cclip) : 1.0f;
    if constexpr (kColOp == OP_SUM) {
      const int32_t len = __ldg(cu_C + cclip + 1) - __ldg(cu_C + cclip);
      sc *= 1.0f / static_cast<float>(len);
    }
    const float val = v * sc;
    float* dst = out_row + static_cast<int64_t>(cclip) * ld_c;
    if (flags & kRcDirect) {
      st_global_cs_f32(dst, val);  // streaming store: the score matrix is written once, never re-read here
    } else if constexpr (kRowOp == OP_MAX) {
      red_global_max_f32(dst, val);
    } else {
      red_global_add_f32(dst, val);
    }
  }
}
// Two-pass mode (kRowOp == OP_NONE): no reduction over rows here.  The value goes to M[cclip][row] --
// consecutive lanes are consecutive rows, so a warp writes one 128-byte line per column clip -- and
// launch_rowreduce finishes the pooling.  m_ptr walks down M one column clip per emit (every emit is
// followed by ++cclip), so a segment costs a predicated store and a 64-bit add instead of ~50 instructions.
#define JEGAL_EMIT()                                                                                   \
  do {                                                                                                 \
    if constexpr (kRowOp == OP_NONE) {                                                                 \
      if (rc.flags & kRcValid) st_global_cs_f32(m_ptr, acc);                                           \
      m_ptr += p.ld_c;                                                                                 \
    } else {                                                                                           \
      emit<kColOp, kRowOp>(acc, cclip, rc.rscale, rc.flags, rc.out_row, p.cscale, p.cu_C,              \
                           static_cast<int32_t>(p.ld_c));                                              \
    }                                                                                                  \
  } while (0)

// One-row clips on BOTH sides (clip-level embeddings, evaluate_retrieval.py:38-48): nothing to pool,
// this is a plain GEMM epilogue — every thread stores the cosines of its row for the chunk's columns.
__device__ __forceinline__ void store_chunk_dense(const uint32_t (&v)[32], int32_t ncols, int32_t cclip,
                                                  const RowCtx& rc, const SimpoolParams& p) {
  if (!(rc.flags & kRcValid)) return;
  float* dst = rc.out_row + static_cast<int64_t>(cclip) * p.ld_c;
  const float rs = rc.rscale;
  if (p.ld_c == 1 && ncols >= 32 && !p.cscale && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0u) {
    // a full chunk of this thread's output row, 32-byte aligned: four 256-bit stores, each one whole sector
    // (16-byte stores filled a sector in two instructions: the epilogue was bound by LSU transactions, 31 % of HBM)
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[j + i]) * rs;
      st_global_cs_v8_f32(dst + j, o);
    }
    return;
  }
  if (p.ld_c == 1 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(p.cscale + cclip)) & 15u) == 0u) {
    // this thread's 32 cosines are contiguous in the output row: 16-byte stores (4x fewer store instructions)
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j + 3 < ncols) {
        float4 o;
        if (p.cscale) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(p.cscale + cclip + j));
          o = make_float4(__uint_as_float(v[j]) * rs * sc.x, __uint_as_float(v[j + 1]) * rs * sc.y,
                          __uint_as_float(v[j + 2]) * rs * sc.z, __uint_as_float(v[j + 3]) * rs * sc.w);
        } else {
          o = make_float4(__uint_as_float(v[j]) * rs, __uint_as_float(v[j + 1]) * rs,
                          __uint_as_float(v[j + 2]) * rs, __uint_as_float(v[j + 3]) * rs);
        }
        __stcs(reinterpret_cast<float4*>(dst + j), o);
      } else {
#pragma unroll
        for (int i = j; i < j + 4; ++i) {
          if (i < ncols) {
            const float sc = p.cscale ? __ldg(p.cscale + cclip + i) : 1.0f;
            dst[i] = __uint_as_float(v[i]) * rs * sc;
          }
        }
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < ncols) {
      const float sc = p.cscale ? __ldg(p.cscale + cclip + j) : 1.0f;
      dst[static_cast<int64_t>(j) * p.ld_c] = __uint_as_float(v[j]) * rs * sc;
    }
  }
}

template <int kOp>
__device__ __forceinline__ float op3(float a, float b, float c) {
  if constexpr (kOp == OP_MAX) return fmax3(a, b, c);
  return (a + b) + c;
}

// An 8-column sub-block with exactly one interior segment end after column e (0..6):
// lo = columns 0..e, hi = columns e+1..7.  e is warp-uniform, so this is one uniform jump into a
// 2-4 instruction case (the generic run loop costs ~30 instructions per run).
template <int kOp>
__device__ __forceinline__ void split8(const uint32_t* v, uint32_t e, float& lo, float& hi) {
  const float x0 = __uint_as_float(v[0]), x1 = __uint_as_float(v[1]), x2 = __uint_as_float(v[2]),
              x3 = __uint_as_float(v[3]), x4 = __uint_as_float(v[4]), x5 = __uint_as_float(v[5]),
              x6 = __uint_as_float(v[6]), x7 = __uint_as_float(v[7]);
  switch (e) {
    case 0: lo = x0; hi = op3<kOp>(x1, op3<kOp>(x2, x3, x4), op3<kOp>(x5, x6, x7)); break;
    case 1: lo = op_apply<kOp>(x0, x1); hi = op3<kOp>(op3<kOp>(x2, x3, x4), op_apply<kOp>(x5, x6), x7); break;
    case 2: lo = op3<kOp>(x0, x1, x2); hi = op3<kOp>(op3<kOp>(x3, x4, x5), x6, x7); break;
    case 3: lo = op_apply<kOp>(op_apply<kOp>(x0, x1), op_apply<kOp>(x2, x3));
            hi = op_apply<kOp>(op_apply<kOp>(x4, x5), op_apply<kOp>(x6, x7)); break;
    case 4: lo = op3<kOp>(op3<kOp>(x0, x1, x2), x3, x4); hi = op3<kOp>(x5, x6, x7); break;
    case 5: lo = op3<kOp>(op3<kOp>(x0, x1, x2), op_apply<kOp>(x3, x4), x5); hi = op_apply<kOp>(x6, x7); break;
    default: lo = op3<kOp>(x0, op3<kOp>(x1, x2, x3), op3<kOp>(x4, x5, x6)); hi = x7; break;
  }
}

// Pool one 32-column chunk of this thread's row: `em` marks the columns that end a segment and is
// warp-uniform BY CONSTRUCTION (the caller rebuilds it with a ballot, so the compiler keeps it in
// uniform registers and every branch below is a uniform branch without BSSY/BSYNC pairs).
// Works in 8-column sub-blocks: no interior end -> a 4-instruction tree; one interior end -> split8;
// two or more (clips shorter than 8 columns) -> runs with 8-wide masks.
template <int kColOp, int kRowOp>
__device__ __forceinline__ void pool_chunk(const uint32_t (&v)[32], uint32_t em, float& acc, int32_t& cclip,
                                           float*& m_ptr, const RowCtx& rc, const SimpoolParams& p) {
  if ((em & 0x7f7f7f7fu) == 0u) {
    // whole chunk on the fast path (config 5: every chunk): no per-sub-block tests
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = op_apply<kColOp>(acc, reduce8<kColOp>(v + 8 * j));
      if ((em >> (8 * j + 7)) & 1u) {
        JEGAL_EMIT();
        ++cclip;
        acc = op_ident<kColOp>();
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t em8 = (em >> (8 * j)) & 0xffu;
    const uint32_t inner = em8 & 0x7fu;
    if (inner == 0u) {
      acc = op_apply<kColOp>(acc, reduce8<kColOp>(v + 8 * j));
    } else if ((inner & (inner - 1u)) == 0u) {
      float lo, hi;
      split8<kColOp>(v + 8 * j, static_cast<uint32_t>(__ffs(inner) - 1), lo, hi);
      acc = op_apply<kColOp>(acc, lo);
      JEGAL_EMIT();
      ++cclip;
      acc = hi;
    } else {
      uint32_t rem = inner;
      uint32_t start = 0;
      while (true) {
        const uint32_t e = rem ? static_cast<uint32_t>(__ffs(rem) - 1) : 7u;
        const uint32_t m = ((2u << e) - 1u) & ~((1u << start) - 1u);  // columns start..e of the sub-block
        float r = op_ident<kColOp>();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = ((m >> i) & 1u) ? __uint_as_float(v[8 * j + i]) : op_ident<kColOp>();
          r = op_apply<kColOp>(r, x);
        }
        acc = op_apply<kColOp>(acc, r);
        if (!rem) break;
        JEGAL_EMIT();
        ++cclip;
        acc = op_ident<kColOp>();
        rem &= rem - 1u;
        start = e + 1u;
      }
    }
    if (em8 & 0x80u) {
      JEGAL_EMIT();
      ++cclip;
      acc = op_ident<kColOp>();
    }
  }
}

template <int kCG, int kStages, int kColOp, int kRowOp, bool kDense = false>
__global__ void __launch_bounds__(kThreads, 1)
simpool_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC,
               const SimpoolParams p) {
  constexpr int UMMA_M = kTileRows * kCG;
  constexpr int UMMA_N = kTileRows * kCG;
  constexpr uint32_t kTmemCols = 2 * UMMA_N;
  constexpr int NKB = kNumKBlocks;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t smC = base;
  const uint32_t smR = base + NKB * kTileBytes;
  const uint32_t bars = smR + kStages * kTileBytes;
  auto r_full = [&](int i) { return bars + 8u * i; };
  auto r_empty = [&](int i) { return bars + 8u * (kStages + i); };
  auto c_full = [&](int k) { return bars + 8u * (2 * kStages + k); };
  auto c_empty = [&](int k) { return bars + 8u * (2 * kStages + NKB + k); };
  auto t_full = [&](int b) { return bars + 8u * (2 * kStages + 2 * NKB + b); };
  auto t_empty = [&](int b) { return bars + 8u * (2 * kStages + 2 * NKB + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 2 * NKB + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_addr));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kCG == 2 ? cluster_ctarank() : 0u;
  const uint32_t cluster = kCG == 2 ? cluster_id_x() : blockIdx.x;
  const uint32_t nclusters = kCG == 2 ? num_clusters_x() : gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmR);
    prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(r_full(i), 1);
      mbar_init(r_empty(i), 1);
    }
    for (int k = 0; k < NKB; ++k) {
      mbar_init(c_full(k), 1);
      mbar_init(c_empty(k), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full(b), 1);
      // one arrival per epilogue warp that drains the buffer, in every CTA (two-pass mode: both warpgroups)
      mbar_init(t_empty(b), (kRowOp == OP_NONE ? 4 * kEpiGroups : 4) * kCG);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kCG>(tmem_slot, kTmemCols);
  tc_fence_before();
  if constexpr (kCG == 2) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      // R chunk is re-streamed by every unit (r_policy, same encoding as c_policy; default evict_last)
      const uint64_t pol_R = p.r_policy == 1 ? policy_evict_last() : p.r_policy == 2 ? policy_evict_first() : policy_evict_normal();
      // a column tile is fetched once per phase of R by exactly one cluster: no reuse inside a phase,
      // so it must not displace the R chunk every cluster re-streams (c_policy: 0 normal, 1 last, 2 first)
      const uint64_t pol_C = p.c_policy == 1 ? policy_evict_last() : p.c_policy == 2 ? policy_evict_first() : policy_evict_normal();
      UnitIter it(p, cluster, nclusters);
      int32_t ct, rt0, nrt;
      uint32_t rs = 0, rphase = 0, unit = 0;
      unsigned long long tr[4] = {0, 0, 0, 0};
      const long long t_begin = clock64();
      while (it.next(ct, rt0, nrt)) {
        // the MMA of this unit is only as wide as its clips (N = roundup16(n_valid)); in a CTA pair
        // each CTA supplies N/2 rows of the column operand
        const int32_t n_unit = (__ldg(&p.ctiles[ct].n_valid) + 15) & ~15;
        const int32_t c_row = __ldg(&p.ctiles[ct].row0) + static_cast<int32_t>(rank) * (n_unit / kCG);
        for (int32_t t = 0; t < nrt; ++t) {
          const int32_t r_row = (rt0 + t) * UMMA_M + static_cast<int32_t>(rank) * kTileRows;
          for (int kb = 0; kb < NKB; ++kb) {
            if (t == 0) {
              // (re)load stationary k-block kb once the previous unit's MMAs released it
              if (unit > 0) JEGAL_TRACED(1, mbar_wait(c_empty(kb), (unit - 1) & 1u));
              if constexpr (kCG == 2) {
                if (rank == 0) mbar_arrive_expect_tx(c_full(kb), kTileBytes * 2);
                tma_load_2d_pair(&tmC, c_full(kb) & kPeerBitMask, smC + kb * kTileBytes,
                                 kb * kBlockK, c_row, pol_C);
              } else {
                mbar_arrive_expect_tx(c_full(kb), kTileBytes);
                tma_load_2d(&tmC, c_full(kb), smC + kb * kTileBytes, kb * kBlockK, c_row, pol_C);
              }
            }
            JEGAL_TRACED(0, mbar_wait(r_empty(rs), rphase ^ 1u));
            if constexpr (kCG == 2) {
              if (rank == 0) mbar_arrive_expect_tx(r_full(rs), kTileBytes * 2);
              tma_load_2d_pair(&tmR, r_full(rs) & kPeerBitMask, smR + rs * kTileBytes, kb * kBlockK,
                               r_row, pol_R);
            } else {
              mbar_arrive_expect_tx(r_full(rs), kTileBytes);
              tma_load_2d(&tmR, r_full(rs), smR + rs * kTileBytes, kb * kBlockK, r_row, pol_R);
            }
            if (++rs == kStages) {
              rs = 0;
              rphase ^= 1u;
            }
          }
        }
        ++unit;
      }
      if (p.trace) {
        unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16;
        g[0] = tr[0]; g[1] = tr[1]; g[2] = static_cast<unsigned long long>(clock64() - t_begin);
      }
      // producer tail: do not leave while tcgen05.commit arrivals may still be in
      // flight towards this CTA's barriers (the peer CTA of a pair must stay alive).
      if (unit > 0) {
        for (int i = 0; i < kStages; ++i) {
          mbar_wait(r_empty(rs), rphase ^ 1u);
          if (++rs == kStages) {
            rs = 0;
            rphase ^= 1u;
          }
        }
        for (int kb = 0; kb < NKB; ++kb) mbar_wait(c_empty(kb), (unit - 1) & 1u);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (rank == 0 && elect_one()) {
      UnitIter it(p, cluster, nclusters);
      int32_t ct, rt0, nrt;
      uint32_t rs = 0, rphase = 0, unit = 0, tile = 0;
      unsigned long long tr[4] = {0, 0, 0, 0};
      const long long t_begin = clock64();
      const uint64_t descC0 = make_smem_desc_sw128(smC);
      const uint64_t descR0 = make_smem_desc_sw128(smR);
      while (it.next(ct, rt0, nrt)) {
        const uint32_t n_unit = (static_cast<uint32_t>(__ldg(&p.ctiles[ct].n_valid)) + 15u) & ~15u;
        const uint32_t idesc = (p.idesc & ~(0x3fu << 17)) | ((n_unit >> 3) << 17);  // N field of the descriptor
        for (int32_t t = 0; t < nrt; ++t, ++tile) {
          const uint32_t buf = tile & 1u;
          JEGAL_TRACED(0, mbar_wait(t_empty(buf), ((tile >> 1) & 1u) ^ 1u));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * UMMA_N;
          for (int kb = 0; kb < NKB; ++kb) {
            if (t == 0) JEGAL_TRACED(1, mbar_wait(c_full(kb), unit & 1u));
            JEGAL_TRACED(2, mbar_wait(r_full(rs), rphase));
            tc_fence_after();
            const uint64_t dR = descR0 + static_cast<uint64_t>((rs * kTileBytes) >> 4);
            const uint64_t dC = descC0 + static_cast<uint64_t>((kb * kTileBytes) >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // +32 bytes (2 x 16 B) per 16-element K step inside the 128 B swizzle row
              umma_f16<kCG>(d_tmem, dR + 2u * k, dC + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit<kCG>(r_empty(rs));
            if (t == nrt - 1) umma_commit<kCG>(c_empty(kb));
            if (++rs == kStages) {
              rs = 0;
              rphase ^= 1u;
            }
          }
          umma_commit<kCG>(t_full(buf));
        }
        ++unit;
      }
      if (p.trace) {
        unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16;
        g[4] = tr[0]; g[5] = tr[1]; g[6] = tr[2]; g[7] = static_cast<unsigned long long>(clock64() - t_begin);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue
    // Fused modes: warpgroup g drains accumulator buffer g (every other tile).  Then tile i+2 can only be
    // computed once tile i is pooled, so the tensor pipe stays busy iff T_epilogue(tile) <= T_mma(tile).
    // Two-pass mode (kRowOp == OP_NONE): BOTH warpgroups drain EVERY tile, warpgroup g the columns
    // [g * N/2, (g + 1) * N/2) -- a buffer is held for half the time (T_epilogue / 2 <= T_mma).  The plan cuts
    // every segment at N/2 (kPlanCutHalves), so the halves are independent: no carry, no synchronisation.
    constexpr bool kSplit = kRowOp == OP_NONE;
    constexpr int kHalfChunks = UMMA_N / 64;       // 32-column chunks per half tile
    const int q = (warp - kEpiWarp0) & 3;          // TMEM lane quarter == warp % 4
    const uint32_t grp = (warp - kEpiWarp0) >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    UnitIter it(p, cluster, nclusters);
    int32_t ct, rt0, nrt;
    uint32_t tile = 0;
    unsigned long long tr[4] = {0, 0, 0, 0};
    const long long t_begin = clock64();
    while (it.next(ct, rt0, nrt)) {
      const CTile* ctile = p.ctiles + ct;
      const int32_t n_valid = __ldg(&ctile->n_valid);
      const int32_t clip0 = __ldg(&ctile->clip0);
      const int32_t part = __ldg(&ctile->partial);
      const bool col_partial = (part & 1) != 0;
      const uint32_t my_em = lane < 8 ? __ldg(&ctile->endmask[lane]) : 0u;
      const int nchunks = (n_valid + 31) >> 5;
      const int ch_lo = kSplit ? static_cast<int>(grp) * kHalfChunks : 0;
      const int ch_hi = kSplit ? min(nchunks, ch_lo + kHalfChunks) : nchunks;
      // two-pass mode: M row of this warpgroup's first segment (segments left of the split belong to group 0)
      int32_t seg0 = clip0 + (part >> 1);
      if (kSplit && grp != 0) seg0 += __reduce_add_sync(0xffffffffu, lane < kHalfChunks ? __popc(my_em) : 0);
      int4 ri_next = make_int4(-1, 0, 1, 0);
      bool have_next = false;
      for (int32_t t = 0; t < nrt; ++t, ++tile) {
        if (!kSplit && (tile & 1u) != grp) continue;  // fused modes: the other warpgroup owns this tile
        // per-row context for this tile.  Ragged row side: the {clip, begin, end} record of this
        // thread's row was requested while the previous own tile was being pooled (software
        // pipelining: the L2 round trip was 23 % of the epilogue's time when issued here).
        RowCtx rc;
        if constexpr (kRowOp == OP_NONE) {
          const int32_t row = (rt0 + t) * UMMA_M + static_cast<int32_t>(rank) * kTileRows + q * 32 + lane;
          rc.rclip = row;
          rc.rscale = 1.0f;
          rc.flags = row < p.n_rows_R ? kRcValid : 0u;
          rc.out_row = p.out + row;
        } else {
          const int32_t row = (rt0 + t) * UMMA_M + static_cast<int32_t>(rank) * kTileRows + q * 32 + lane;
          int32_t sb = 0, se = 1;
          rc.rclip = -1;
          if (p.uni_len_R > 0) {  // all row clips have the same length: no lookup at all
            if (row < p.n_rows_R) {
              rc.rclip = row / p.uni_len_R;
              sb = rc.rclip * p.uni_len_R;
              se = sb + p.uni_len_R;
            }
          } else {
            int4 ri = ri_next;
            if (!have_next) ri = row < p.n_rows_R ? __ldg(p.rowinfo_R + row) : make_int4(-1, 0, 1, 0);
            have_next = t + 2 < nrt;
            if (have_next) {
              const int32_t nrow = row + 2 * UMMA_M;
              ri_next = nrow < p.n_rows_R ? __ldg(p.rowinfo_R + nrow) : make_int4(-1, 0, 1, 0);
            }
            rc.rclip = ri.x;
            sb = ri.y;
            se = ri.z;
          }
          float rs_ = (rc.rclip >= 0 && p.rscale) ? __ldg(p.rscale + rc.rclip) : 1.0f;
          if constexpr (kRowOp == OP_SUM) rs_ *= 1.0f / static_cast<float>(se - sb);
          rc.rscale = rs_;
          uint32_t fl = rc.rclip >= 0 ? kRcValid : 0u;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int32_t o = __shfl_down_sync(0xffffffffu, rc.rclip, 1u << k);
            if (lane + (1 << k) < 32 && o == rc.rclip) fl |= 1u << k;
          }
          const int32_t prev = __shfl_up_sync(0xffffffffu, rc.rclip, 1);
          if (rc.rclip >= 0 && (lane == 0 || prev != rc.rclip)) fl |= kRcHead;
          const int32_t wrow0 = row - lane;
          if (sb >= wrow0 && se <= wrow0 + 32 && !col_partial) fl |= kRcDirect;
          rc.flags = fl;
          rc.out_row = p.out + static_cast<int64_t>(rc.rclip) * p.ld_r;
        }
        const uint32_t buf = tile & 1u;
        JEGAL_TRACED(0, mbar_wait(t_full(buf), (tile >> 1) & 1u));
        tc_fence_after();
        const long long t_pool0 = p.trace ? clock64() : 0;
        const uint32_t t_addr = tmem_base + lane_off + buf * UMMA_N;

        // hand the TMEM buffer back to the MMA warp once its last chunk sits in registers
        auto release = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCG == 2) {
              mbar_arrive_cluster(t_empty(buf) & kPeerBitMask);
            } else {
              mbar_arrive(t_empty(buf));
            }
          }
        };

        float acc = op_ident<kColOp>();
        int32_t cclip = clip0;
        float* m_ptr = rc.out_row + static_cast<int64_t>(seg0) * p.ld_c;  // two-pass mode: &M[seg0][row]
        if (ch_lo >= ch_hi) {  // two-pass mode, narrow tile: nothing in the right half
          release();
          continue;
        }
        uint32_t va[32], vb[32];  // two chunks in flight: load c+1 while chunk c is pooled
        tmem_ld_32x32(t_addr + ch_lo * 32, va);
        for (int ch = ch_lo; ch < ch_hi; ch += 2) {
          // every lane holds the same end mask; the ballot tells the compiler so (uniform registers)
          const uint32_t em_a = __ballot_sync(0xffffffffu, (__shfl_sync(0xffffffffu, my_em, ch) >> lane) & 1u);
          const uint32_t em_b = __ballot_sync(0xffffffffu, (__shfl_sync(0xffffffffu, my_em, ch + 1) >> lane) & 1u);
          tmem_ld_wait();
          if (ch + 1 < ch_hi) tmem_ld_32x32(t_addr + (ch + 1) * 32, vb);
          else release();
          if constexpr (kDense) {
            store_chunk_dense(va, n_valid - ch * 32, clip0 + ch * 32, rc, p);
          } else {
            pool_chunk<kColOp, kRowOp>(va, em_a, acc, cclip, m_ptr, rc, p);
          }
          if (ch + 1 < ch_hi) {
            tmem_ld_wait();
            if (ch + 2 < ch_hi) tmem_ld_32x32(t_addr + (ch + 2) * 32, va);
            else release();
            if constexpr (kDense) {
              store_chunk_dense(vb, n_valid - (ch + 1) * 32, clip0 + (ch + 1) * 32, rc, p);
            } else {
              pool_chunk<kColOp, kRowOp>(vb, em_b, acc, cclip, m_ptr, rc, p);
            }
          }
        }
        if (p.trace) tr[1] += static_cast<unsigned long long>(clock64() - t_pool0);
      }
    }
    if (p.trace && q == 0 && lane == 0) {
      unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16 + 8 + grp * 4;
      g[0] = tr[0]; g[1] = tr[1]; g[2] = static_cast<unsigned long long>(clock64() - t_begin); g[3] = tile;
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  if constexpr (kCG == 2) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (warp == 2) tmem_dealloc<kCG>(tmem_base, kTmemCols);
}

template <int kCG, int kStages>
constexpr size_t simpool_smem_bytes() {
  return 1024 /*alignment slack*/ + static_cast<size_t>(kNumKBlocks + kStages) * kTileBytes +
         8 * (2 * kStages + 2 * kNumKBlocks + 4) + 16;
}

constexpr int kStagesDefault = 6;

template <int kCG, int kColOp, int kRowOp, bool kDense = false>
int launch_simpool_t(jegal_ctx* ctx, const CUtensorMap& tmR, const CUtensorMap& tmC,
                     const SimpoolParams& p, cudaStream_t stream) {
  auto kern = simpool_kernel<kCG, kStagesDefault, kColOp, kRowOp, kDense>;
  constexpr size_t smem = simpool_smem_bytes<kCG, kStagesDefault>();
  constexpr uint32_t bit = 1u << ((kCG - 1) * 8 + (kDense ? 7 : kColOp * 3 + kRowOp));
  if (!(ctx->smem_configured & bit)) {
    JEGAL_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
    ctx->smem_configured |= bit;
  }
  const int64_t steps = static_cast<int64_t>(p.n_ctiles) * p.n_rtiles;
  int nclusters = ctx->sm_count / kCG;
  if (steps < nclusters) nclusters = static_cast<int>(steps > 0 ? steps : 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(nclusters * kCG));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  JEGAL_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, tmR, tmC, p));
  ctx->launches++;
  return JEGAL_OK;
}

template <int kCG>
int launch_simpool_cg(jegal_ctx* ctx, int col_op, int row_op, const CUtensorMap& tmR,
                      const CUtensorMap& tmC, const SimpoolParams& p, cudaStream_t stream) {
  if (p.dense) return launch_simpool_t<kCG, OP_SUM, OP_SUM, true>(ctx, tmR, tmC, p, stream);
  if (col_op == OP_SUM && row_op == OP_SUM)
    return launch_simpool_t<kCG, OP_SUM, OP_SUM>(ctx, tmR, tmC, p, stream);
  if (col_op == OP_MAX && row_op == OP_SUM)
    return launch_simpool_t<kCG, OP_MAX, OP_SUM>(ctx, tmR, tmC, p, stream);
  if (col_op == OP_MAX && row_op == OP_MAX)
    return launch_simpool_t<kCG, OP_MAX, OP_MAX>(ctx, tmR, tmC, p, stream);
  if (col_op == OP_MAX && row_op == OP_NONE)
    return launch_simpool_t<kCG, OP_MAX, OP_NONE>(ctx, tmR, tmC, p, stream);
  if (col_op == OP_SUM && row_op == OP_NONE)
    return launch_simpool_t<kCG, OP_SUM, OP_NONE>(ctx, tmR, tmC, p, stream);
  return set_err(ctx, JEGAL_ERR_ARG, "simpool: unsupported (col_op,row_op)");
}

__global__ void fill_f32_kernel(float* __restrict__ dst, int64_t n, float value) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4 v4 = make_float4(value, value, value, value);
  for (int64_t i = i0; i < n4; i += stride) d4[i] = v4;
  for (int64_t i = (n4 << 2) + i0; i < n; i += stride) dst[i] = value;
}

// Pass 2 of the two-pass mode.  A block takes ONE column clip c and kRrClips = 64 consecutive row clips, a
// warp 8 consecutive row clips: together the 8 warps stream one contiguous stretch of M[c] (row clips are
// contiguous row ranges), so DRAM sees long sequential reads.  The warp walks its 8 clips in lock step --
// step k loads rows begin_j + 32k + lane of every clip j, 8 independent coalesced loads in flight per lane --
// then 8 interleaved butterflies finish the reductions and lanes 0..7 write the scores.  The pieces of a
// split / half-cut column clip (rows seg_C[c] .. seg_C[c+1] of M) are combined per row with the column
// operation first.
constexpr int kRrClips = 64;
template <int kColOp, int kRowOp>
__global__ void __launch_bounds__(256)
rowreduce_kernel(const float* __restrict__ M, int64_t ldm, const int32_t* __restrict__ cu_R, int32_t n_rclips,
                 const int32_t* __restrict__ cu_C, const int32_t* __restrict__ seg_C, int32_t n_cclips,
                 int32_t rblocks, const float* __restrict__ rscale, const float* __restrict__ cscale,
                 float* __restrict__ out, int64_t ld_r, int64_t ld_c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t c = static_cast<int32_t>(blockIdx.x / rblocks);
  const int32_t ra = static_cast<int32_t>(blockIdx.x - static_cast<int64_t>(c) * rblocks) * kRrClips + warp * 8;
  if (ra >= n_rclips) return;
  const int32_t m0 = seg_C ? __ldg(seg_C + c) : c;
  const int32_t np = seg_C ? __ldg(seg_C + c + 1) - m0 : 1;
  const float* src = M + static_cast<int64_t>(m0) * ldm;
  // clip boundaries of this warp: lane l holds cu_R[min(ra + l, n_rclips)], broadcast into registers
  const int32_t my_b = __ldg(cu_R + min(ra + min(lane, 8), n_rclips));
  int32_t bnd[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) bnd[j] = __shfl_sync(0xffffffffu, my_b, j);
  int32_t longest = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) longest = max(longest, bnd[j + 1] - bnd[j]);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = op_ident<kRowOp>();
  if (np == 1) {
    // A clip's rows [r0, r1) are split into an aligned interior [a, b) read with unmasked 16-byte loads (one
    // per lane covers 128 rows per step; 8 clips in lock step = 4 KB in flight per warp and round trip) and
    // at most 3 + 3 edge rows read by lanes 0-2 / 4-6 with one scalar load each.  (Masking every element of
    // every 16-byte load cost 24 instructions per load and kept the kernel half issue-bound at 2.6 TB/s.)
    const float4* p4[8];
    int32_t n4[8];  // 16-byte loads this lane still has to do for clip j
    float edge[8];
    int32_t longest4 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int32_t r0 = bnd[j], r1 = bnd[j + 1];
      const int32_t a = min((r0 + 3) & ~3, r1);  // head rows [r0, a)
      const int32_t b = max(r1 & ~3, a);         // tail rows [b, r1)
      const int32_t idx = lane < 4 ? r0 + lane : b + lane - 4;
      const bool ok = lane < 4 ? idx < a : (lane < 8 && idx < r1);
      edge[j] = op_ident<kRowOp>();
      if (ok) edge[j] = __ldcs(src + idx);
      p4[j] = reinterpret_cast<const float4*>(src + a) + lane;
      n4[j] = ((b - a) >> 2) - lane;
      longest4 = max(longest4, (b - a) >> 2);
    }
    for (int32_t k = 0; k < longest4; k += 32) {
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float id = op_ident<kRowOp>();
        x[j] = make_float4(id, id, id, id);
        if (n4[j] > 0) x[j] = __ldcs(p4[j]);
        p4[j] += 32;
        n4[j] -= 32;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = op_apply<kRowOp>(v[j], op_apply<kRowOp>(op_apply<kRowOp>(x[j].x, x[j].y), op_apply<kRowOp>(x[j].z, x[j].w)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = op_apply<kRowOp>(v[j], edge[j]);
  } else {
    for (int32_t k = lane; k < longest; k += 32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int32_t i = bnd[j] + k;
        if (i < bnd[j + 1]) {
          float x = __ldcs(src + i);
          for (int32_t m = 1; m < np; ++m) x = op_apply<kColOp>(x, __ldcs(src + static_cast<int64_t>(m) * ldm + i));
          v[j] = op_apply<kRowOp>(v[j], x);
        }
      }
    }
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = op_apply<kRowOp>(v[j], __shfl_xor_sync(0xffffffffu, v[j], k));
  }
  float mine = v[0];
  int32_t len = bnd[1] - bnd[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    mine = lane == j ? v[j] : mine;
    len = lane == j ? bnd[j + 1] - bnd[j] : len;
  }
  const int32_t r = ra + lane;
  if (lane < 8 && r < n_rclips) {
    float sc = (cscale ? __ldg(cscale + c) : 1.0f) * (rscale ? __ldg(rscale + r) : 1.0f);
    if constexpr (kColOp == OP_SUM) sc *= 1.0f / static_cast<float>(__ldg(cu_C + c + 1) - __ldg(cu_C + c));
    if constexpr (kRowOp == OP_SUM) sc *= 1.0f / static_cast<float>(len);
    out[static_cast<int64_t>(r) * ld_r + static_cast<int64_t>(c) * ld_c] = mine * sc;
  }
}

}  // namespace

int launch_rowreduce(jegal_ctx* ctx, const float* M, int64_t ldm, const int32_t* cu_R, int32_t n_rclips,
                     const int32_t* cu_C, const int32_t* seg_C, int32_t n_cclips, int col_op, int row_op,
                     const float* rscale, const float* cscale, float* out, int64_t ld_r, int64_t ld_c,
                     cudaStream_t stream) {
  const int32_t rblocks = (n_rclips + kRrClips - 1) / kRrClips;  // blocks per column clip
  const int64_t blocks = static_cast<int64_t>(n_cclips) * rblocks;
  if (blocks <= 0) return JEGAL_OK;
  if (blocks > 0x7fffffff) return set_err(ctx, JEGAL_ERR_UNSUPPORTED, "rowreduce: more than 2^31 blocks");
  const unsigned g = static_cast<unsigned>(blocks);
  if (col_op == OP_MAX && row_op == OP_SUM) {
    rowreduce_kernel<OP_MAX, OP_SUM><<<g, 256, 0, stream>>>(M, ldm, cu_R, n_rclips, cu_C, seg_C, n_cclips, rblocks,
                                                            rscale, cscale, out, ld_r, ld_c);
  } else if (col_op == OP_MAX && row_op == OP_MAX) {
    rowreduce_kernel<OP_MAX, OP_MAX><<<g, 256, 0, stream>>>(M, ldm, cu_R, n_rclips, cu_C, seg_C, n_cclips, rblocks,
                                                            rscale, cscale, out, ld_r, ld_c);
  } else if (col_op == OP_SUM && row_op == OP_SUM) {
    rowreduce_kernel<OP_SUM, OP_SUM><<<g, 256, 0, stream>>>(M, ldm, cu_R, n_rclips, cu_C, seg_C, n_cclips, rblocks,
                                                            rscale, cscale, out, ld_r, ld_c);
  } else {
    return set_err(ctx, JEGAL_ERR_ARG, "rowreduce: unsupported (col_op,row_op)");
  }
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

int launch_simpool(jegal_ctx* ctx, int cta_group, int col_op, int row_op, const CUtensorMap& tmR,
                   const CUtensorMap& tmC, const SimpoolParams& p, cudaStream_t stream) {
  if (cta_group == 2) return launch_simpool_cg<2>(ctx, col_op, row_op, tmR, tmC, p, stream);
  if (cta_group == 1) return launch_simpool_cg<1>(ctx, col_op, row_op, tmR, tmC, p, stream);
  return set_err(ctx, JEGAL_ERR_ARG, "simpool: cta_group must be 1 or 2");
}

int launch_fill_f32(jegal_ctx* ctx, float* dst, int64_t n, float value, cudaStream_t stream) {
  if (n <= 0) return JEGAL_OK;
  const int threads = 256;
  int64_t blocks = (n / 4 + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  fill_f32_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(dst, n, value);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // namespace jegal

// K3 (word spotting) and K4 (listed clip pairs, e.g. active-speaker groups) for sm_100a.
//
// Both score LISTED (gesture clip, content clip) pairs instead of all pairs, so the
// work per byte is tiny (T x W x 512 MACs for (T + W) x 1 KB of operands): these
// kernels are HBM-bound and are built to stream, not to saturate the tensor pipe.
//
//   item = one pair; its T frames are the M side (TMEM lane = frame = one epilogue
//   thread), its W <= 64 words the N side (columns), so everything the reference
//   does "per frame over words" (softmax over words, evaluate_spotting.py:52-54;
//   max over words) is register-local, and only per-frame scalars cross lanes.
//
//   warp 0  TMA producer: frames 32 rows per box so short clips do not drag 128 rows through
//           L2, words 16 rows per box.  Stages live in a byte ring and take exactly the
//           bytes they load (a 69-frame x 8-word k-block is 14 KB, not a fixed 24 KB), so
//           ~200 KB of payload — two to three clips — is in flight per SM, across items
//   warp 1  tcgen05.mma issuer: M = 128, N = roundup16(W), 32 MMAs per row tile
//   warp 2  TMEM allocator (4 accumulator buffers of 64 columns)
//   warps 4-7 epilogue: tcgen05.ld -> softmax / pooling -> coalesced stores
//   warps 8-11 (kFuse only) a second epilogue warpgroup (the two groups take the items alternately)
//   warps 12-19 (kFuse only) row norms: the L2 normalisation of the reference
//           (F.normalize at evaluate_spotting.py:49-50, CosineSimilarity's clamped norms at
//           evaluate_asd.py:45-47) is FUSED INTO THE LOAD: the operands are the rows exactly as
//           the .pkl stores them (fp16, or bf16), TMA stages them once, the tensor core contracts
//           the raw rows, and these warps re-read the staged k-blocks from shared memory (a row's
//           128 bytes stay inside its own swizzled line, so a sum of squares needs no
//           un-swizzling): norm warp w owns k-block w of every row tile, stores its partial
//           sum(x^2) of every frame / word into per-k-block tables, and the epilogue scales lane
//           (frame) and column (word) by rsqrt(max(sum, eps^2)) = 1 / max(||row||, eps).
//           No normalised copy of the operands ever exists in HBM: bytes per clip are the
//           (T + W) x 1 KB of the stored rows, read once.
//           (One warp per k-block, not one warp per row range of every k-block: a warp walks the
//           stage sequence serially -- wait, load, arrive -- at ~700 cycles per iteration whatever
//           the payload, measured with JEGAL_GROUPED_TRACE; eight warps take one stage in eight.
//           640 threads: setmaxnreg moves registers from the producer / norm warpgroups to the epilogue's.)
//
// Items are dealt round-robin to a persistent grid of one CTA per SM.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"

namespace jegal {

using namespace ptx;

namespace {

constexpr int kGThreads = 256;
constexpr int kGThreadsFuse = 640;                // + a second epilogue warpgroup and 8 row-norm warps
constexpr int kNormWarp0 = 12;                    // kFuse: warps 4-7 / 8-11 epilogue groups, 12-15 row norms
constexpr int kNormWarps = 8;                     // norm warp w owns k-block w of every row tile
constexpr int kNormTables = 2;                    // tile parity
constexpr int kNormRows = 192;                    // per accumulator buffer: 128 frame slots + 64 word slots
constexpr int kGStages = 32;                     // barrier slots; smem is a BYTE ring (see RingAlloc)
constexpr int kNMax = 64;                        // words per clip on the N side
constexpr uint32_t kGRingBytes = 206 * 1024;     // operand ring: a stage takes only the bytes it loads
constexpr int kNumAcc = 4;
constexpr uint32_t kGTmemCols = kNumAcc * kNMax;  // 256
constexpr int kBoxG = 32, kBoxC = 16;
// k-blocks per pipeline stage.  The single producer thread spends ~400 cycles of ring / barrier bookkeeping per stage
// next to ~40 cycles per TMA issue (JEGAL_GROUPED_TRACE: it was busy 87 % of the fused K3 and every other role waited
// for it); two k-blocks per stage halve that cost -- and the MMA issuer's and the norm warps' barrier traffic.
// Config 3, fused: 1 -> 0.305 ms, 2 -> 0.255-0.260 ms, 4 -> 0.285 ms (stages of up to 96 KB: the 206 KB ring holds too few).
#ifndef JEGAL_KPER
#define JEGAL_KPER 2
#endif
constexpr uint32_t kKPerStage = JEGAL_KPER;
constexpr int kStagesPerTile = kNumKBlocks / kKPerStage;

enum GroupedEpi : int { EPI_SPOT = 0, EPI_POOL = 1 };

struct GroupedParams {
  int32_t n_items;
  const int32_t* item_g;  // nullable: item i -> gesture clip (default i)
  const int32_t* item_c;  // nullable: item i -> content clip (default i)
  const int32_t* cu_T;
  const int32_t* cu_W;
  uint32_t idesc_base;    // instruction descriptor with N = 0
  // spotting
  const int32_t* word_idx;
  float inv_tau;
  float* heat;
  float* full_heat;
  const int64_t* full_off;
  int32_t* pred_frame;
  float* pred_score;
  const int32_t* win_lo;
  const int32_t* win_hi;
  float thresh;
  uint8_t* correct;
  // pooling
  float row_eps;          // kFuse: 1 / max(||row||, row_eps)
  unsigned long long* trace;  // nullable debug buffer: 16 cycle counters per CTA (JEGAL_GROUPED_TRACE=1)
  int32_t pool_mode;
  const float* gscale;
  const float* cscale;
  float* scores;
};

// One tensor map per box height, so a stage is always exactly two TMA operations: frames in
// boxes of 32/64/96/128 rows, words in boxes of 16/32/48/64 rows.
struct alignas(64) GroupedMaps {
  CUtensorMap g[4];
  CUtensorMap c[4];
};

struct Item {
  int32_t g, c, g_row0, T, c_row0, W, n_rt, n16;
};

__device__ __forceinline__ Item load_item(const GroupedParams& p, int32_t i) {
  Item it;
  it.g = p.item_g ? __ldg(p.item_g + i) : i;
  it.c = p.item_c ? __ldg(p.item_c + i) : i;
  it.g_row0 = __ldg(p.cu_T + it.g);
  it.T = __ldg(p.cu_T + it.g + 1) - it.g_row0;
  it.c_row0 = __ldg(p.cu_W + it.c);
  it.W = __ldg(p.cu_W + it.c + 1) - it.c_row0;
  it.n_rt = (it.T + 127) >> 7;
  it.n16 = (it.W + 15) >> 4;
  return it;
}

// Deterministic byte-ring placement shared by the producer and the MMA issuer: both walk the same
// item sequence, so both compute the same offsets.  A stage that would cross the end wraps to 0.
struct RingAlloc {
  uint32_t w = 0;
  // returns the offset of a stage of `sz` bytes; `need` = bytes consumed including the wrap skip
  // `kb_bytes` = bytes of ONE k-block (frames + words); a stage holds kKPerStage of them back to back
  __device__ __forceinline__ uint32_t place(uint32_t kb_bytes, uint32_t& need) {
    uint32_t skip = 0;
    const uint32_t sz = kKPerStage * kb_bytes;
    // the MMA always reads 128 frame rows (16 KB) from the start of every k-block, whatever was loaded:
    // keep that window inside the ring (rows past the clip feed accumulator lanes nobody reads)
    const uint32_t win = (kKPerStage - 1) * kb_bytes + 16384u;
    const uint32_t span = sz > win ? sz : win;
    if (w + span > kGRingBytes) {
      skip = kGRingBytes - w;
      w = 0;
    }
    const uint32_t off = w;
    w += sz;
    need = skip + sz;
    return off;
  }
};

// cycle accounting, compiled in with -DJEGAL_GTRACE (make EXTRA=-DJEGAL_GTRACE) and switched on with
// JEGAL_GROUPED_TRACE=1: tr[slot] += cycles spent in `stmt`.  Off by default: the duplicated wait sites cost
// instruction-cache space the kernel does not have (ncu: stall_no_inst was the top stall reason).
#ifdef JEGAL_GTRACE
#define JEGAL_GTRACED(slot, stmt)                                        \
  do {                                                                   \
    if (JEGAL_GTRACE_ON(p)) {                                                       \
      const long long t0__ = clock64();                                  \
      stmt;                                                              \
      tr[slot] += static_cast<unsigned long long>(clock64() - t0__);     \
    } else {                                                             \
      stmt;                                                              \
    }                                                                    \
  } while (0)
#define JEGAL_GTRACE_ON(p) ((p).trace != nullptr)
#else
#define JEGAL_GTRACED(slot, stmt) do { stmt; } while (0)
#define JEGAL_GTRACE_ON(p) false
#endif

__device__ __forceinline__ void bar_sync_epi(uint32_t grp) { asm volatile("bar.sync %0, 128;" ::"r"(1u + grp) : "memory"); }

// sum of squares of the 8 16-bit values of one 16-byte chunk, accumulated in fp32.
// kImpl 0: widen to fp32 (HADD2.F32 / a shift for bf16) + FFMA; kImpl 1: the mixed-precision fma of sm_100
// (FHFMA takes the half of a packed register directly: half the instructions -- measured A/B)
template <bool kBf16, int kImpl>
__device__ __forceinline__ float sumsq8(uint4 r, float acc) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  if constexpr (kImpl == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned short lo = static_cast<unsigned short>(w[i] & 0xffffu), hi = static_cast<unsigned short>(w[i] >> 16);
      if constexpr (kBf16) {
        asm("fma.rn.f32.bf16 %0, %1, %1, %0;" : "+f"(acc) : "h"(lo));
        asm("fma.rn.f32.bf16 %0, %1, %1, %0;" : "+f"(acc) : "h"(hi));
      } else {
        asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(acc) : "h"(lo));
        asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(acc) : "h"(hi));
      }
    }
    return acc;
  } else {
    float a0 = acc, a1 = 0.f;  // two chains per chunk
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x, y;
      if constexpr (kBf16) {
        x = __uint_as_float(w[i] << 16);
        y = __uint_as_float(w[i] & 0xffff0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        x = f.x;
        y = f.y;
      }
      a0 = fmaf(x, x, a0);
      a1 = fmaf(y, y, a1);
    }
    return a0 + a1;
  }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

template <int kEpi, int kFuse>
__global__ void __launch_bounds__(kFuse ? kGThreadsFuse : kGThreads, 1)
grouped_kernel(const __grid_constant__ GroupedMaps maps, const GroupedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t bars = base + kGRingBytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kGStages + s); };
  auto t_full = [&](int b) { return bars + 8u * (2 * kGStages + b); };
  auto t_empty = [&](int b) { return bars + 8u * (2 * kGStages + kNumAcc + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * kGStages + 2 * kNumAcc);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_addr));
  // epilogue scratch (after the tmem slot), per epilogue warpgroup, double-buffered by item parity so one named
  // barrier per item suffices: per (group, parity) 4 warps x 64 floats + 4 floats + 4 ints
  constexpr int kScratchF = 4 * kNMax + 4;
  float* epi_f0 = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - raw_addr));
  int32_t* epi_i0 = reinterpret_cast<int32_t*>(epi_f0 + 4 * kScratchF);
  volatile uint32_t* need_smem = reinterpret_cast<volatile uint32_t*>(epi_i0 + 16);  // [kGStages], producer-private
  // kFuse: partial sums of squares of a tile's rows, one table per (tile parity, k-block): [tb][kb][0..127] frames,
  // [tb][kb][128..191] words (plain stores: fp32 shared-memory atomics are a CAS loop, 14 % of the kernel's stall
  // samples when the eight k-block warps added into one table); inv_w: per epilogue warp, the 64 inverse word
  // norms of the tile it is draining
  float* part = reinterpret_cast<float*>(const_cast<uint32_t*>(need_smem) + kGStages);
  float* inv_w = part + kNormTables * kNumKBlocks * kNormRows;
  const uint32_t n_full0 = smem_u32(inv_w + 8 * kNMax);
  auto n_full = [&](int b) { return n_full0 + 8u * b; };                   // all 8 partial tables of a tile are written
  auto n_empty = [&](int b) { return n_full0 + 8u * (kNormTables + b); };  // the 4 epilogue warps have read them

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane < 8) prefetch_tmap(lane < 4 ? &maps.g[lane] : &maps.c[lane - 4]);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), kFuse ? 1 + kKPerStage : 1);  // the MMAs' commit (+ the row-norm warps that own its k-blocks)
    }
    for (int b = 0; b < kNumAcc; ++b) {
      mbar_init(t_full(b), 1);
      mbar_init(t_empty(b), 4);
      if constexpr (kFuse != 0) {
        if (b < kNormTables) {
          // every lane that writes / reads a table entry arrives itself (16 writers per k-block warp, 32 readers per
          // epilogue warp): the ordering does not lean on a warp-level sync in front of a single arrival
          mbar_init(n_full(b), kNumKBlocks * 16);
          mbar_init(n_empty(b), 4 * 32);
        }
      }
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, kGTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Role dispatch by warpgroup.  kFuse: 640 threads, i.e. 96 registers per thread, which is why the epilogue works on
  // one 16-column group of the accumulator at a time instead of holding all 64 columns.  (setmaxnreg was tried to move
  // registers from the producer / row-norm warpgroups to the epilogue's: this ptxas caps the whole kernel at the
  // SMALLEST setmaxnreg value it sees, whatever dominates what, so it only made things worse.)
  if (warp < 4) {
  if (warp == 0) {
    // (Two producer threads -- frame boxes from warp 0, word boxes from warp 3, barrier count 2 -- were measured twice,
    // before and after the epilogue / norm warps stopped being the limit: 0.33 vs 0.31 ms on config 3 both times.
    // Letting the whole warp walk the loop with lane 0 issuing, so that the bookkeeping could live in uniform registers,
    // does not help either: ptxas does not treat the LDG-derived operands as uniform and wraps every TMA in an
    // ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- more instructions per stage, not fewer.)
    if (elect_one()) {
      const uint64_t pol = policy_evict_first();  // every operand byte is used once
      // stage n uses barrier slot n % kGStages; `inflight` = ring bytes of stages not yet released.
      // The bookkeeping of released bytes goes through a small smem array (need_smem).
      uint32_t slot = 0, old_slot = 0, old_phase = 0, n_out = 0, inflight = 0;
      unsigned long long tr[4] = {0, 0, 0, 0};
      const long long t_begin = JEGAL_GTRACE_ON(p) ? clock64() : 0;
      (void)tr;
      (void)t_begin;
      RingAlloc ring;
      Item nxt = load_item(p, min(static_cast<int32_t>(blockIdx.x), p.n_items - 1));
      for (int32_t i = blockIdx.x; i < p.n_items; i += gridDim.x) {
        const Item it = nxt;
        // descriptor of the next item: its loads fly while this item's TMA is issued
        nxt = load_item(p, min(i + static_cast<int32_t>(gridDim.x), p.n_items - 1));
        const CUtensorMap* tmC = &maps.c[it.n16 - 1];
        for (int32_t rt = 0; rt < it.n_rt; ++rt) {
          const int32_t rows = min(128, it.T - rt * 128);
          const int32_t nb = (rows + kBoxG - 1) / kBoxG;
          const CUtensorMap* tmG = &maps.g[nb - 1];
          const uint32_t bytes_g = nb * (kBoxG * 128);
          const uint32_t bytes = bytes_g + it.n16 * (kBoxC * 128);
          const int32_t g_row = it.g_row0 + rt * 128;
#pragma unroll 1
          for (int st = 0; st < kStagesPerTile; ++st) {
            uint32_t need;
            const uint32_t off = base + ring.place(bytes, need);
            // free ring space / a barrier slot by retiring the oldest stages (the MMAs release in order)
            while (inflight + need > kGRingBytes || n_out >= kGStages) {
              JEGAL_GTRACED(0, mbar_wait_lean(empty(old_slot), old_phase));
              inflight -= need_smem[old_slot];
              --n_out;
              if (++old_slot == kGStages) {
                old_slot = 0;
                old_phase ^= 1u;
              }
            }
            need_smem[slot] = need;
            inflight += need;
            ++n_out;
            const uint32_t fb = full(slot);
            mbar_arrive_expect_tx(fb, kKPerStage * bytes);
#pragma unroll
            for (uint32_t h = 0; h < kKPerStage; ++h) {
              const int32_t kcol = (st * static_cast<int32_t>(kKPerStage) + static_cast<int32_t>(h)) * kBlockK;
              tma_load_2d(tmG, fb, off + h * bytes, kcol, g_row, pol);
              tma_load_2d(tmC, fb, off + h * bytes + bytes_g, kcol, it.c_row0, pol);
            }
            if (++slot == kGStages) slot = 0;
          }
        }
      }
      if (JEGAL_GTRACE_ON(p)) {
        unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16;
        g[0] = tr[0];
        g[1] = static_cast<unsigned long long>(clock64() - t_begin);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      uint32_t slot = 0, phase = 0, tile = 0;
      unsigned long long tr[4] = {0, 0, 0, 0};
      const long long t_begin = JEGAL_GTRACE_ON(p) ? clock64() : 0;
      (void)tr;
      (void)t_begin;
      RingAlloc ring;
      Item nxt = load_item(p, min(static_cast<int32_t>(blockIdx.x), p.n_items - 1));
      for (int32_t i = blockIdx.x; i < p.n_items; i += gridDim.x) {
        const Item it = nxt;
        nxt = load_item(p, min(i + static_cast<int32_t>(gridDim.x), p.n_items - 1));
        const uint32_t idesc = p.idesc_base | (static_cast<uint32_t>(it.n16 * 16) >> 3) << 17;
        for (int32_t rt = 0; rt < it.n_rt; ++rt, ++tile) {
          const int32_t rows = min(128, it.T - rt * 128);
          const uint32_t bytes_g = ((rows + kBoxG - 1) / kBoxG) * (kBoxG * 128);
          const uint32_t bytes = bytes_g + it.n16 * (kBoxC * 128);
          const uint32_t buf = tile % kNumAcc;
          JEGAL_GTRACED(0, mbar_wait_lean(t_empty(buf), ((tile / kNumAcc) & 1u) ^ 1u));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kNMax;
          // a row tile of <= 64 frames is an M = 64 MMA: its A window is 64 rows (8 KB per k-block instead of 16 KB of
          // shared-memory reads); the accumulator then sits on lanes 0-15 of every 32-lane quarter (rows 16 q .. 16 q + 15)
          const uint32_t idesc_t = rows <= 64 ? ((idesc & ~(0x1fu << 24)) | (4u << 24)) : idesc;
#pragma unroll 1
          for (int st = 0; st < kStagesPerTile; ++st) {
            uint32_t need;
            const uint32_t off = base + ring.place(bytes, need);
            JEGAL_GTRACED(1, mbar_wait_lean(full(slot), phase));
            tc_fence_after();
#pragma unroll
            for (uint32_t h = 0; h < kKPerStage; ++h) {
              const uint64_t dG = make_smem_desc_sw128(off + h * bytes);
              const uint64_t dC = make_smem_desc_sw128(off + h * bytes + bytes_g);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_f16<1>(d_tmem, dG + 2u * k, dC + 2u * k, idesc_t, (st | static_cast<int>(h) | k) != 0 ? 1u : 0u);
            }
            umma_commit<1>(empty(slot));
            if (++slot == kGStages) {
              slot = 0;
              phase ^= 1u;
            }
          }
          umma_commit<1>(t_full(buf));
        }
      }
      if (JEGAL_GTRACE_ON(p)) {
        unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16;
        g[2] = tr[0];
        g[3] = tr[1];
        g[4] = static_cast<unsigned long long>(clock64() - t_begin);
      }
    }
  }
  } else if (warp < (kFuse != 0 ? kNormWarp0 : 8)) {
  {
    // Epilogue.  kFuse: TWO warpgroups (warps 4-7 and 8-11) take the items alternately -- with the scaling by the
    // row norms in front of the softmax one warpgroup needed ~4400 cycles per clip of dependent instructions against
    // the ~3400 cycles a clip may take at the HBM roofline (JEGAL_GROUPED_TRACE: the epilogue never waited).
    const int q = (warp - 4) & 3;
    const uint32_t grp = static_cast<uint32_t>(warp - 4) >> 2;
    constexpr uint32_t kGroups = kFuse != 0 ? 2u : 1u;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t tile = 0, seq = 0;
    unsigned long long tr[4] = {0, 0, 0, 0};
    const long long t_begin = JEGAL_GTRACE_ON(p) ? clock64() : 0;
    (void)tr;
    (void)t_begin;
    Item nxt = load_item(p, min(static_cast<int32_t>(blockIdx.x), p.n_items - 1));
    for (int32_t i = blockIdx.x; i < p.n_items; i += gridDim.x, ++seq) {
      const Item it = nxt;
      nxt = load_item(p, min(i + static_cast<int32_t>(gridDim.x), p.n_items - 1));
      if (kGroups > 1 && (seq % kGroups) != grp) {  // the other warpgroup's item
        tile += static_cast<uint32_t>(it.n_rt);
        continue;
      }
      const uint32_t parity = (seq / kGroups) & 1u;
      float* epi_f = epi_f0 + (grp * 2 + parity) * kScratchF;
      int32_t* epi_i = epi_i0 + (grp * 2 + parity) * 4;
      // per-item state; every scalar the finalisation needs is fetched now, not after the barrier
      float best_v = -1.0f;
      int32_t best_t = 0x7fffffff;
      float racc = 0.f;                   // POOL: running reduction over frames
      float wmax0 = -INFINITY, wmax1 = -INFINITY;  // POOL max_t_mean_w: running max of words lane, lane+32
      int32_t target = 0, wlo = 0, whi = 0;
      float gs = 1.0f, cs = 1.0f;
      float* fh = nullptr;
      if constexpr (kEpi == EPI_SPOT) {
        target = __ldg(p.word_idx + i);
        if (p.full_heat) fh = p.full_heat + __ldg(p.full_off + i);
        if (p.correct) {
          wlo = __ldg(p.win_lo + i);
          whi = __ldg(p.win_hi + i);
        }
      } else {
        if (p.pool_mode == JEGAL_POOL_MAX_MAX) racc = -INFINITY;
        if (p.gscale) gs = __ldg(p.gscale + it.g);
        if (p.cscale) cs = __ldg(p.cscale + it.c);
      }
      for (int32_t rt = 0; rt < it.n_rt; ++rt, ++tile) {
        const uint32_t buf = tile % kNumAcc;
        // frame of this thread within the row tile: lane of a full (M = 128) tile, lanes 0-15 of every quarter for an
        // M = 64 tile (the issuer's rule: <= 64 frames)
        const bool m64 = it.T - rt * 128 <= 64;
        const int et = m64 ? q * 16 + (lane & 15) : q * 32 + lane;
        JEGAL_GTRACED(0, mbar_wait_lean(t_full(buf), (tile / kNumAcc) & 1u));
        tc_fence_after();
        const uint32_t t_addr = tmem_base + lane_off + buf * kNMax;
        float ig = 1.0f;
        const float* iw = nullptr;
        if constexpr (kFuse != 0) {
          // cos[t, w] = (g_t . c_w) / (max(||g_t||, eps) max(||c_w||, eps)): the row-norm warps summed the squares of
          // the very bytes the MMAs consumed; 1 / max(sqrt(s), eps) = rsqrt(max(s, eps^2))
          const uint32_t tb = tile % kNormTables;
          JEGAL_GTRACED(1, mbar_wait_lean(n_full(tb), (tile / kNormTables) & 1u));
          const float* pt_ = part + tb * (kNumKBlocks * kNormRows);
          const float eps2 = p.row_eps * p.row_eps;
          float sf = 0.f, sw0 = 0.f, sw1 = 0.f;
#pragma unroll
          for (int k = 0; k < kNumKBlocks; ++k) {
            sf += pt_[k * kNormRows + et];
            sw0 += pt_[k * kNormRows + 128 + lane];
            sw1 += pt_[k * kNormRows + 160 + lane];
          }
          ig = rsqrtf(fmaxf(sf, eps2));
          float* iw_w = inv_w + (grp * 4 + q) * kNMax;
          iw_w[lane] = rsqrtf(fmaxf(sw0, eps2));
          iw_w[32 + lane] = rsqrtf(fmaxf(sw1, eps2));
          mbar_arrive(n_empty(tb));  // this lane has read its entries of the partial tables
          __syncwarp();              // inv_w is complete before anyone reads it
          iw = iw_w;
        }
        // One 16-column group of the accumulator at a time (load, scale by the row norms): 16 live values per thread
        // instead of 64, which is what lets 640 threads (two epilogue warpgroups + eight norm warps) fit the register
        // file; clips of <= 16 words -- the usual case -- still see a single TMEM load.
        auto load_group = [&](int g, float (&x)[16]) {
          uint32_t r[16];
          tmem_ld_32x16(t_addr + g * 16, r);
          tmem_ld_wait();
          if constexpr (kFuse != 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 ic = *reinterpret_cast<const float4*>(iw + g * 16 + j);
              x[j] = __uint_as_float(r[j]) * ig * ic.x;
              x[j + 1] = __uint_as_float(r[j + 1]) * ig * ic.y;
              x[j + 2] = __uint_as_float(r[j + 2]) * ig * ic.z;
              x[j + 3] = __uint_as_float(r[j + 3]) * ig * ic.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(r[j]);
          }
        };
        auto release = [&]() {  // the accumulator buffer goes back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty(buf));
        };

        const int32_t t = rt * 128 + et;
        const bool valid = t < it.T && !(m64 && lane >= 16);
        if constexpr (kEpi == EPI_SPOT) {
          // softmax over words of s / tau (evaluate_spotting.py:52-54), per frame
          float pt = 0.f;
          if (it.n16 == 1) {  // W <= 16: everything in registers
            float x[16];
            load_group(0, x);
            release();
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < it.W) m = fmaxf(m, x[j]);
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j < it.W) {
                x[j] = __expf((x[j] - m) * p.inv_tau);
                den += x[j];
                if (j == target) pt = x[j];
              }
            }
            const float inv_den = 1.0f / den;
            pt *= inv_den;
            if (valid && fh) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < it.W) fh[static_cast<int64_t>(j) * it.T + t] = x[j] * inv_den;
            }
          } else {  // up to 64 words: three passes over the accumulator's column groups (TMEM reads are cheap)
            float m = -INFINITY;
#pragma unroll 1
            for (int g = 0; g < it.n16; ++g) {
              float x[16];
              load_group(g, x);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (g * 16 + j < it.W) m = fmaxf(m, x[j]);
            }
            float den = 0.f;
#pragma unroll 1
            for (int g = 0; g < it.n16; ++g) {
              float x[16];
              load_group(g, x);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int w = g * 16 + j;
                if (w < it.W) {
                  const float e = __expf((x[j] - m) * p.inv_tau);
                  den += e;
                  if (w == target) pt = e;
                }
              }
            }
            const float inv_den = 1.0f / den;
            pt *= inv_den;
            if (fh) {
#pragma unroll 1
              for (int g = 0; g < it.n16; ++g) {
                float x[16];
                load_group(g, x);
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int w = g * 16 + j;
                    if (w < it.W) fh[static_cast<int64_t>(w) * it.T + t] = __expf((x[j] - m) * p.inv_tau) * inv_den;
                  }
                }
              }
            }
            release();
          }
          if (valid) {
            if (p.heat) p.heat[it.g_row0 + t] = pt;
            if (pt > best_v) {  // frames arrive in increasing order: strict > keeps the first maximum
              best_v = pt;
              best_t = t;
            }
          }
        } else {
          float c = p.pool_mode == JEGAL_POOL_MEAN_MEAN ? 0.f : -INFINITY;
#pragma unroll 1
          for (int g = 0; g < it.n16; ++g) {
            float x[16];
            load_group(g, x);
            if (g == it.n16 - 1) release();
            if (p.pool_mode == JEGAL_POOL_MAX_T_MEAN_W) {
              // max over frames first: per word, reduce across the warp's valid lanes
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int w = g * 16 + j;
                if (w < it.W) {
                  float y = valid ? x[j] : -INFINITY;
#pragma unroll
                  for (int o = 16; o > 0; o >>= 1) y = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, o));
                  if (lane == (w & 31)) {
                    if (w < 32) wmax0 = fmaxf(wmax0, y); else wmax1 = fmaxf(wmax1, y);
                  }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (g * 16 + j < it.W) c = p.pool_mode == JEGAL_POOL_MEAN_MEAN ? c + x[j] : fmaxf(c, x[j]);
              }
            }
          }
          if (p.pool_mode != JEGAL_POOL_MAX_T_MEAN_W && valid) racc = p.pool_mode == JEGAL_POOL_MAX_MAX ? fmaxf(racc, c) : racc + c;
        }
      }
      // ---- per-item finalisation across the 4 epilogue warps
      if constexpr (kEpi == EPI_SPOT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
          const int32_t ot = __shfl_xor_sync(0xffffffffu, best_t, o);
          if (ov > best_v || (ov == best_v && ot < best_t)) {
            best_v = ov;
            best_t = ot;
          }
        }
        if (lane == 0) {
          epi_f[4 * kNMax + q] = best_v;
          epi_i[q] = best_t;
        }
        bar_sync_epi(grp);
        if (q == 0 && lane == 0) {
          float bv = epi_f[4 * kNMax];
          int32_t bt = epi_i[0];
          for (int w = 1; w < 4; ++w) {
            const float ov = epi_f[4 * kNMax + w];
            const int32_t ot = epi_i[w];
            if (ov > bv || (ov == bv && ot < bt)) {
              bv = ov;
              bt = ot;
            }
          }
          if (p.pred_frame) p.pred_frame[i] = bt;
          if (p.pred_score) p.pred_score[i] = bv;
          if (p.correct) {
            const bool ok = bt >= wlo && bt <= whi && bv >= p.thresh;
            p.correct[i] = ok ? 1 : 0;
          }
        }
      } else {
        if (p.pool_mode == JEGAL_POOL_MAX_T_MEAN_W) {
          epi_f[q * kNMax + lane] = wmax0;
          epi_f[q * kNMax + 32 + lane] = wmax1;
          bar_sync_epi(grp);
          if (q == 0) {
            float s = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int w = h * 32 + lane;
              if (w < it.W) {
                float m = epi_f[w];
#pragma unroll
                for (int ww = 1; ww < 4; ++ww) m = fmaxf(m, epi_f[ww * kNMax + w]);
                s += m;
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) p.scores[i] = s / static_cast<float>(it.W) * gs * cs;
          }
        } else {
          const bool is_max = p.pool_mode == JEGAL_POOL_MAX_MAX;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, racc, o);
            racc = is_max ? fmaxf(racc, ov) : racc + ov;
          }
          if (lane == 0) epi_f[4 * kNMax + q] = racc;
          bar_sync_epi(grp);
          if (q == 0 && lane == 0) {
            float r = epi_f[4 * kNMax];
            for (int w = 1; w < 4; ++w) r = is_max ? fmaxf(r, epi_f[4 * kNMax + w]) : r + epi_f[4 * kNMax + w];
            float sc = gs * cs;
            if (p.pool_mode == JEGAL_POOL_MEAN_MEAN) sc /= static_cast<float>(it.T) * static_cast<float>(it.W);
            if (p.pool_mode == JEGAL_POOL_MAX_W_MEAN_T) sc /= static_cast<float>(it.T);
            p.scores[i] = r * sc;
          }
        }
      }
    }
    if (JEGAL_GTRACE_ON(p) && q == 0 && lane == 0 && grp == 0) {
      unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16;
      g[5] = tr[0];
      g[6] = tr[1];
      g[7] = static_cast<unsigned long long>(clock64() - t_begin);
    }
  }
  } else {
  if constexpr (kFuse != 0) {
    {
      // ---- row norms, fused into the load.  Norm warp kb owns k-block kb of every row tile: stage 8 * tile + kb.
      // A stage is [frames: nb x 32 rows][words: n16 x 16 rows] of 128-byte swizzled lines, walked in units of
      // 16 rows: lanes l and l + 16 share row l & 15, one takes the row's 16-byte chunks 0-3, the other 4-7
      // (which one alternates with the row's parity), each in an order rotated by the row number, so the eight
      // lanes of every shared-memory phase touch eight different bank groups.
      constexpr bool kBf16 = kFuse == 2;
      constexpr int kImpl = kFuse == 3 ? 0 : 1;
      const int nw = warp - kNormWarp0;
      const int r16 = lane & 15;
      const int half = (r16 & 1) ^ (lane >> 4);
      uint32_t lane_off[4];  // byte offset of this lane's four chunks inside a unit
#pragma unroll
      for (int c = 0; c < 4; ++c)
        lane_off[c] = static_cast<uint32_t>(r16) * 128u + static_cast<uint32_t>(half * 4 + ((c + (r16 >> 1)) & 3)) * 16u;
      uint32_t tile = 0;
      unsigned long long tr[4] = {0, 0, 0, 0};
      const long long t_begin = JEGAL_GTRACE_ON(p) ? clock64() : 0;
      (void)tr;
      (void)t_begin;
      RingAlloc ring;
      Item nxt = load_item(p, min(static_cast<int32_t>(blockIdx.x), p.n_items - 1));
      for (int32_t i = blockIdx.x; i < p.n_items; i += gridDim.x) {
        const Item it = nxt;
        nxt = load_item(p, min(i + static_cast<int32_t>(gridDim.x), p.n_items - 1));
        for (int32_t rt = 0; rt < it.n_rt; ++rt, ++tile) {
          const int32_t rows = min(128, it.T - rt * 128);
          const int32_t rows_g = ((rows + kBoxG - 1) / kBoxG) * kBoxG;
          const int32_t nunits = (rows_g >> 4) + it.n16;  // <= 12
          const uint32_t bytes = static_cast<uint32_t>(nunits) * 2048u;
          // every role replays the same ring placement; this warp only touches its own two stages
          constexpr int kPasses = kNumKBlocks / kNormWarps;
          uint32_t my_off[kPasses];
#pragma unroll
          for (int st = 0; st < kStagesPerTile; ++st) {
            uint32_t need;
            const uint32_t o = ring.place(bytes, need);
#pragma unroll
            for (int ps = 0; ps < kPasses; ++ps) {
              const int kb = nw + ps * kNormWarps;
              if (st == kb / static_cast<int>(kKPerStage)) my_off[ps] = o + static_cast<uint32_t>(kb % static_cast<int>(kKPerStage)) * bytes;
            }
          }
          const uint32_t tb = tile % kNormTables;
          // this tile's tables were last used two tiles back: all four warps of the epilogue group that drained
          // that tile have read them
          JEGAL_GTRACED(1, mbar_wait_lean(n_empty(tb), ((tile / kNormTables) & 1u) ^ 1u));
          // frame units that lie entirely past the clip's last frame (the box is rounded up to 32 rows) are skipped
          const int32_t units_g = (rows + 15) >> 4;
          const int32_t nwork = units_g + it.n16;
#pragma unroll 1
          for (int pass = 0; pass < kPasses; ++pass) {
            const int kb = nw + pass * kNormWarps;
            const uint32_t stage = tile * kStagesPerTile + static_cast<uint32_t>(kb) / kKPerStage;
            const uint32_t slot = stage % kGStages, phase = (stage / kGStages) & 1u;
            const uint32_t a0 = base + my_off[pass];
            float* d = part + (tb * kNumKBlocks + kb) * kNormRows + r16;
            JEGAL_GTRACED(0, mbar_wait_lean(full(slot), phase));
            // a ROLLED loop over the stage's 16-row units (two in flight): the body is ~50 instructions and the
            // kernel's hot code has to fit the instruction cache
#pragma unroll 2
            for (int32_t u = 0; u < nwork; ++u) {
              const bool is_word = u >= units_g;
              const int32_t row0 = is_word ? rows_g + (u - units_g) * 16 : u * 16;  // first row of the unit in the stage
              const uint32_t au = a0 + static_cast<uint32_t>(row0) * 128u;
              const uint4 x0 = lds128(au + lane_off[0]), x1 = lds128(au + lane_off[1]);
              const uint4 x2 = lds128(au + lane_off[2]), x3 = lds128(au + lane_off[3]);
              const float s0 = sumsq8<kBf16, kImpl>(x0, 0.f), s1 = sumsq8<kBf16, kImpl>(x1, 0.f);
              const float s2 = sumsq8<kBf16, kImpl>(x2, 0.f), s3 = sumsq8<kBf16, kImpl>(x3, 0.f);
              float ss = (s0 + s1) + (s2 + s3);
              ss += __shfl_xor_sync(0xffffffffu, ss, 16);
              if (lane < 16) d[is_word ? 128 + (u - units_g) * 16 : row0] = ss;
            }
            if (lane < 16) mbar_arrive(n_full(tb));  // this lane's table entries of k-block kb are written
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(slot));  // every lane's shared-memory reads of the stage have returned
          }
        }
      }
      if (JEGAL_GTRACE_ON(p) && lane == 0) {
        unsigned long long* g = p.trace + static_cast<size_t>(blockIdx.x) * 16;
        if (nw == 0) {
          g[8] = tr[0];
          g[9] = tr[1];
          g[10] = static_cast<unsigned long long>(clock64() - t_begin);
        } else if (nw == kNormWarps - 1) {
          g[11] = tr[0];
          g[12] = tr[1];
        }
      }
    }
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, kGTmemCols);
}

constexpr size_t grouped_smem_bytes() {
  return 1024 + static_cast<size_t>(kGRingBytes) + 8 * (2 * kGStages + 2 * kNumAcc) + 16 +
         4 * (sizeof(float) * (4 * kNMax + 4) + sizeof(int32_t) * 4) + sizeof(uint32_t) * kGStages + 16 +
         sizeof(float) * (kNormTables * kNumKBlocks * kNormRows + 8 * kNMax) + 8 * 2 * kNormTables;  // kFuse: partial squared norms, inverse word norms, barriers
}

// one thread per group: softmax(scores / tau) within the group + first argmax
// (evaluate_asd.py:47-51,99)
__global__ void group_softmax_kernel(const float* __restrict__ scores, int32_t n_groups, int32_t gsz,
                                     int64_t stride, float inv_tau, float* __restrict__ probs,
                                     int32_t* __restrict__ argmax) {
  const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const float* s = scores + static_cast<int64_t>(g) * stride;
  float m = -INFINITY;
  int32_t am = 0;
  for (int32_t k = 0; k < gsz; ++k) {
    const float x = __ldg(s + k);
    if (x > m) {
      m = x;
      am = k;
    }
  }
  if (argmax) argmax[g] = am;
  if (probs) {
    float den = 0.f;
    for (int32_t k = 0; k < gsz; ++k) den += expf((__ldg(s + k) - m) * inv_tau);
    const float inv = 1.0f / den;
    for (int32_t k = 0; k < gsz; ++k)
      probs[static_cast<int64_t>(g) * gsz + k] = expf((__ldg(s + k) - m) * inv_tau) * inv;
  }
}

// ---- spotting from a dense cosine matrix (clips with more words than the grouped kernel's 64 columns).
// cos is [rows of the gesture layout, rows of the content layout] (K1's plain-GEMM epilogue over the packed frames x
// words of a group of clips); only the block-diagonal is read: one warp per frame row takes the softmax over its
// clip's words, exactly the arithmetic of the K3 epilogue.
__global__ void __launch_bounds__(256)
spot_dense_rows_kernel(const float* __restrict__ cos, int64_t ld, const int4* __restrict__ rowinfo, int64_t rows,
                       const int32_t* __restrict__ cu_W, const int32_t* __restrict__ word_idx, float inv_tau,
                       float* __restrict__ heat, float* __restrict__ full, const int64_t* __restrict__ full_off) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int4 ri = __ldg(rowinfo + r);
  const int32_t clip = ri.x, t = static_cast<int32_t>(r) - ri.y, T = ri.z - ri.y;
  const int32_t c0 = __ldg(cu_W + clip), W = __ldg(cu_W + clip + 1) - c0;
  const int32_t target = __ldg(word_idx + clip);
  const float* x = cos + r * ld + c0;
  float m = -INFINITY;
  for (int32_t w = lane; w < W; w += 32) m = fmaxf(m, __ldg(x + w));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float den = 0.f;
  for (int32_t w = lane; w < W; w += 32) den += __expf((__ldg(x + w) - m) * inv_tau);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  const float inv_den = 1.0f / den;
  float* fh = full ? full + __ldg(full_off + clip) : nullptr;
  for (int32_t w = lane; w < W; w += 32) {
    const float pr = __expf((__ldg(x + w) - m) * inv_tau) * inv_den;
    if (fh) fh[static_cast<int64_t>(w) * T + t] = pr;
    if (w == target) heat[r] = pr;
  }
}

// one warp per clip: first maximum of the clip's heat-map row, window + threshold decision (evaluate_spotting.py:72-82)
__global__ void __launch_bounds__(256)
spot_dense_clips_kernel(const float* __restrict__ heat, const int32_t* __restrict__ cu_T, int32_t n_clips,
                        int32_t* __restrict__ pred_frame, float* __restrict__ pred_score, const int32_t* __restrict__ win_lo,
                        const int32_t* __restrict__ win_hi, float thresh, uint8_t* __restrict__ correct) {
  const int lane = threadIdx.x & 31;
  const int32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_clips) return;
  const int32_t r0 = __ldg(cu_T + i), r1 = __ldg(cu_T + i + 1);
  float bv = -1.0f;
  int32_t bt = 0x7fffffff;
  for (int32_t r = r0 + lane; r < r1; r += 32) {
    const float v = __ldg(heat + r);
    if (v > bv) {  // a lane sees its frames in increasing order: strict > keeps its first maximum
      bv = v;
      bt = r - r0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int32_t ot = __shfl_xor_sync(0xffffffffu, bt, o);
    if (ov > bv || (ov == bv && ot < bt)) {
      bv = ov;
      bt = ot;
    }
  }
  if (lane == 0) {
    if (pred_frame) pred_frame[i] = bt;
    if (pred_score) pred_score[i] = bv;
    if (correct) correct[i] = (bt >= __ldg(win_lo + i) && bt <= __ldg(win_hi + i) && bv >= thresh) ? 1 : 0;
  }
}

int make_grouped_maps(jegal_ctx* ctx, GroupedMaps* m, const void* gest_rows, int64_t gest_n, const void* cont_rows,
                      int64_t cont_n, int op_dtype) {
  for (int i = 0; i < 4; ++i) {
    int rc = make_box_tmap_impl(ctx, &m->g[i], gest_rows, gest_n, op_dtype, kBoxG * (i + 1));
    if (rc != JEGAL_OK) return rc;
    rc = make_box_tmap_impl(ctx, &m->c[i], cont_rows, cont_n, op_dtype, kBoxC * (i + 1));
    if (rc != JEGAL_OK) return rc;
  }
  return JEGAL_OK;
}

template <int kEpi, int kFuse>
int launch_grouped(jegal_ctx* ctx, const GroupedMaps& maps, const GroupedParams& p, cudaStream_t stream) {
  auto kern = grouped_kernel<kEpi, kFuse>;
  constexpr size_t smem = grouped_smem_bytes();
  constexpr uint32_t bit = 1u << (16 + kEpi + 2 * kFuse);
  if (!(ctx->smem_configured & bit)) {
    JEGAL_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ctx->smem_configured |= bit;
  }
  int grid = ctx->sm_count;
  if (p.n_items < grid) grid = p.n_items;
  const char* tr_env = std::getenv("JEGAL_GROUPED_TRACE");
  if (tr_env && *tr_env == '1') {  // debug: per-role cycle accounting, printed to stderr (synchronises)
    GroupedParams pt = p;
    JEGAL_CUDA_OK(ctx, cudaMalloc(&pt.trace, sizeof(unsigned long long) * 16 * grid));
    JEGAL_CUDA_OK(ctx, cudaMemsetAsync(pt.trace, 0, sizeof(unsigned long long) * 16 * grid, stream));
    kern<<<grid, kFuse ? kGThreadsFuse : kGThreads, smem, stream>>>(maps, pt);
    JEGAL_CUDA_OK(ctx, cudaStreamSynchronize(stream));
    std::vector<unsigned long long> h(static_cast<size_t>(16) * grid);
    cudaMemcpy(h.data(), pt.trace, sizeof(unsigned long long) * 16 * grid, cudaMemcpyDeviceToHost);
    cudaFree(pt.trace);
    double a[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 16; ++k) a[k] += static_cast<double>(h[static_cast<size_t>(16) * b + k]) / grid;
    std::fprintf(stderr,
                 "[grouped trace%s, mean cycles over %d CTAs, %d items] producer: wait empty %.0f of %.0f | mma: wait t_empty %.0f, "
                 "wait full %.0f of %.0f | epilogue: wait t_full %.0f, wait norms %.0f of %.0f | norm warp 0: wait full %.0f, "
                 "wait table %.0f of %.0f; norm warp 7: wait full %.0f, wait table %.0f\n",
                 kFuse ? " (fused norms)" : "", grid, p.n_items, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
    ctx->launches++;
    return JEGAL_OK;
  }
  kern<<<grid, kFuse ? kGThreadsFuse : kGThreads, smem, stream>>>(maps, p);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

// kFuse: 0 operands are used as they are; 1 / 2 stored fp16 / bf16 rows, normalisation fused into the load
// (3: fp16 with widen + FFMA arithmetic instead of FHFMA, an A/B knob: JEGAL_NORM_IMPL=0)
template <int kEpi>
int launch_grouped_any(jegal_ctx* ctx, int normalize_rows, int op_dtype, const GroupedMaps& maps, const GroupedParams& p,
                       cudaStream_t stream) {
  if (!normalize_rows) return launch_grouped<kEpi, 0>(ctx, maps, p, stream);
  if (op_dtype == JEGAL_BF16) return launch_grouped<kEpi, 2>(ctx, maps, p, stream);
  const char* e = std::getenv("JEGAL_NORM_IMPL");
  if (e && *e == '0') return launch_grouped<kEpi, 3>(ctx, maps, p, stream);
  return launch_grouped<kEpi, 1>(ctx, maps, p, stream);
}

}  // namespace

}  // namespace jegal

using namespace jegal;

namespace {

int check_w(jegal_ctx* ctx, const jegal_layout* cont, const char* who) {
  if (cont->max_len > kNMax)
    return set_err(ctx, JEGAL_ERR_UNSUPPORTED, std::string(who) + ": a content clip has " +
                                                   std::to_string(cont->max_len) + " words; at most " +
                                                   std::to_string(kNMax) + " are supported");
  return JEGAL_OK;
}

}  // namespace

extern "C" {

int jegal_spot(jegal_ctx* ctx, const jegal_layout* gest_layout, const void* gest_rows_dev,
               const jegal_layout* cont_layout, const void* cont_rows_dev, int op_dtype,
               int normalize_rows, float row_eps, const int32_t* word_idx_dev, float tau, float* heat_dev, float* full_heat_dev,
               const int64_t* full_off_dev, int32_t* pred_frame_dev, float* pred_score_dev,
               const int32_t* win_lo_dev, const int32_t* win_hi_dev, float thresh, uint8_t* correct_dev,
               void* stream_) {
  JEGAL_NVTX("jegal_spot (K3)");
  if (!ctx || !gest_layout || !cont_layout || !gest_rows_dev || !cont_rows_dev || !word_idx_dev)
    return set_err(ctx, JEGAL_ERR_ARG, "spot: null argument");
  if (op_dtype != JEGAL_BF16 && op_dtype != JEGAL_F16) return set_err(ctx, JEGAL_ERR_ARG, "spot: bad op_dtype");
  if (gest_layout->n_clips != cont_layout->n_clips)
    return set_err(ctx, JEGAL_ERR_ARG, "spot: gesture and content layouts must hold the same clips");
  if (!(tau > 0.f)) return set_err(ctx, JEGAL_ERR_ARG, "spot: tau must be > 0");
  if (normalize_rows && !(row_eps > 0.f)) return set_err(ctx, JEGAL_ERR_ARG, "spot: row_eps must be > 0");
  if ((reinterpret_cast<uintptr_t>(gest_rows_dev) | reinterpret_cast<uintptr_t>(cont_rows_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "spot: operand rows must be 16-byte aligned");
  if (full_heat_dev && !full_off_dev) return set_err(ctx, JEGAL_ERR_ARG, "spot: full_heat needs full_off");
  if (correct_dev && (!win_lo_dev || !win_hi_dev)) return set_err(ctx, JEGAL_ERR_ARG, "spot: correct needs win_lo/win_hi");
  if (gest_layout->n_clips == 0) return JEGAL_OK;
  int rc = check_w(ctx, cont_layout, "spot");
  if (rc != JEGAL_OK) return rc;
  GroupedParams p{};
  p.n_items = gest_layout->n_clips;
  p.cu_T = gest_layout->cu_dev;
  p.cu_W = cont_layout->cu_dev;
  p.idesc_base = ptx::make_idesc_f16(op_dtype == JEGAL_BF16 ? 1u : 0u, 128, 0);
  p.word_idx = word_idx_dev;
  p.inv_tau = 1.0f / tau;
  p.heat = heat_dev;
  p.full_heat = full_heat_dev;
  p.full_off = full_off_dev;
  p.pred_frame = pred_frame_dev;
  p.pred_score = pred_score_dev;
  p.win_lo = win_lo_dev;
  p.win_hi = win_hi_dev;
  p.thresh = thresh;
  p.correct = correct_dev;
  p.row_eps = row_eps;
  GroupedMaps maps;
  rc = make_grouped_maps(ctx, &maps, gest_rows_dev, gest_layout->rows, cont_rows_dev, cont_layout->rows, op_dtype);
  if (rc != JEGAL_OK) return rc;
  return launch_grouped_any<EPI_SPOT>(ctx, normalize_rows, op_dtype, maps, p, static_cast<cudaStream_t>(stream_));
}

int jegal_spot_dense(jegal_ctx* ctx, const float* cos_dev, int64_t ld, const jegal_layout* gest_layout,
                     const jegal_layout* cont_layout, const int32_t* word_idx_dev, float tau, float* heat_dev,
                     float* full_heat_dev, const int64_t* full_off_dev, int32_t* pred_frame_dev, float* pred_score_dev,
                     const int32_t* win_lo_dev, const int32_t* win_hi_dev, float thresh, uint8_t* correct_dev, void* stream_) {
  JEGAL_NVTX("jegal_spot_dense (K3 for clips of more than 64 words)");
  if (!ctx || !gest_layout || !cont_layout) return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: null argument");
  if (gest_layout->n_clips != cont_layout->n_clips)
    return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: gesture and content layouts must hold the same clips");
  if (gest_layout->n_clips == 0) return JEGAL_OK;
  if (!cos_dev || !word_idx_dev || !heat_dev) return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: null argument");
  if (ld < cont_layout->rows) return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: ld is smaller than the content rows");
  if (!(tau > 0.f)) return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: tau must be > 0");
  if (full_heat_dev && !full_off_dev) return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: full_heat needs full_off");
  if (correct_dev && (!win_lo_dev || !win_hi_dev)) return set_err(ctx, JEGAL_ERR_ARG, "spot_dense: correct needs win_lo/win_hi");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t rows = gest_layout->rows;
  spot_dense_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      cos_dev, ld, gest_layout->rowinfo_dev, rows, cont_layout->cu_dev, word_idx_dev, 1.0f / tau, heat_dev, full_heat_dev,
      full_off_dev);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  spot_dense_clips_kernel<<<static_cast<unsigned>((gest_layout->n_clips + 7) / 8), 256, 0, stream>>>(
      heat_dev, gest_layout->cu_dev, gest_layout->n_clips, pred_frame_dev, pred_score_dev, win_lo_dev, win_hi_dev, thresh,
      correct_dev);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

int jegal_simpool_pairs(jegal_ctx* ctx, const jegal_layout* gest_layout, const void* gest_rows_dev,
                        const jegal_layout* cont_layout, const void* cont_rows_dev, int op_dtype,
                        int normalize_rows, float row_eps, int pool_mode, const float* gscale_dev, const float* cscale_dev,
                        const int32_t* pair_gest_dev, const int32_t* pair_cont_dev, int32_t n_pairs,
                        int32_t group_size, float tau, float* scores_dev, float* probs_dev,
                        int32_t* argmax_dev, void* stream_) {
  JEGAL_NVTX("jegal_simpool_pairs (K4)");
  if (!ctx || !gest_layout || !cont_layout || !gest_rows_dev || !cont_rows_dev || !scores_dev)
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: null argument");
  if (op_dtype != JEGAL_BF16 && op_dtype != JEGAL_F16) return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: bad op_dtype");
  if (pool_mode < 0 || pool_mode > 3) return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: bad pool_mode");
  if (n_pairs < 0) return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: negative n_pairs");
  if (normalize_rows && !(row_eps > 0.f)) return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: row_eps must be > 0");
  if ((reinterpret_cast<uintptr_t>(gest_rows_dev) | reinterpret_cast<uintptr_t>(cont_rows_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: operand rows must be 16-byte aligned");
  if (!pair_gest_dev && n_pairs > gest_layout->n_clips) return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: n_pairs exceeds clips");
  if (!pair_cont_dev && n_pairs > cont_layout->n_clips) return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: n_pairs exceeds clips");
  const bool want_groups = probs_dev || argmax_dev;
  if (want_groups && (group_size < 1 || n_pairs % group_size != 0 || !(tau > 0.f)))
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_pairs: group_size must divide n_pairs and tau must be > 0");
  if (n_pairs == 0) return JEGAL_OK;
  int rc = check_w(ctx, cont_layout, "simpool_pairs");
  if (rc != JEGAL_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GroupedParams p{};
  p.n_items = n_pairs;
  p.item_g = pair_gest_dev;
  p.item_c = pair_cont_dev;
  p.cu_T = gest_layout->cu_dev;
  p.cu_W = cont_layout->cu_dev;
  p.idesc_base = ptx::make_idesc_f16(op_dtype == JEGAL_BF16 ? 1u : 0u, 128, 0);
  p.pool_mode = pool_mode;
  p.gscale = gscale_dev;
  p.cscale = cscale_dev;
  p.scores = scores_dev;
  p.row_eps = row_eps;
  GroupedMaps maps;
  rc = make_grouped_maps(ctx, &maps, gest_rows_dev, gest_layout->rows, cont_rows_dev, cont_layout->rows, op_dtype);
  if (rc != JEGAL_OK) return rc;
  rc = launch_grouped_any<EPI_POOL>(ctx, normalize_rows, op_dtype, maps, p, stream);
  if (rc != JEGAL_OK) return rc;
  if (want_groups) {
    const int32_t n_groups = n_pairs / group_size;
    group_softmax_kernel<<<(n_groups + 127) / 128, 128, 0, stream>>>(scores_dev, n_groups, group_size, group_size,
                                                                    1.0f / tau, probs_dev, argmax_dev);
    JEGAL_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches++;
  }
  return JEGAL_OK;
}

int jegal_group_softmax(jegal_ctx* ctx, const float* scores_dev, int32_t n_groups, int32_t group_size,
                        int64_t stride, float tau, float* probs_dev, int32_t* argmax_dev, void* stream_) {
  JEGAL_NVTX("jegal_group_softmax");
  if (!ctx || !scores_dev) return set_err(ctx, JEGAL_ERR_ARG, "group_softmax: null argument");
  if (n_groups < 0 || group_size < 1 || stride < group_size || !(tau > 0.f))
    return set_err(ctx, JEGAL_ERR_ARG, "group_softmax: bad shape or tau");
  if (n_groups == 0) return JEGAL_OK;
  group_softmax_kernel<<<(n_groups + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
      scores_dev, n_groups, group_size, stride, 1.0f / tau, probs_dev, argmax_dev);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // extern "C"

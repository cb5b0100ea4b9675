// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/jegal_b200.h"

// NVTX range around the host side of a C-ABI entry point (header-only NVTX 3: a no-op unless a profiler
// injects its library), so a timeline shows which call enqueued which kernels.
struct JegalNvtxRange {
  explicit JegalNvtxRange(const char* name) { nvtxRangePushA(name); }
  ~JegalNvtxRange() { nvtxRangePop(); }
  JegalNvtxRange(const JegalNvtxRange&) = delete;
  JegalNvtxRange& operator=(const JegalNvtxRange&) = delete;
};
#define JEGAL_NVTX(name) JegalNvtxRange jegal_nvtx_range__(name)

namespace jegal {

constexpr int kD = JEGAL_EMB_DIM;  // 512
constexpr int kBlockK = 64;        // one 128-byte swizzle atom of 16-bit elements
constexpr int kNumKBlocks = kD / kBlockK;
constexpr int kTileRows = 128;     // operand rows per CTA per TMA box

// One column tile of the all-pairs kernel: up to `width` consecutive rows of the
// column operand holding whole clips (or one piece of an over-long clip).
struct CTile {
  int32_t row0;     // first operand row of the tile
  int32_t n_valid;  // columns that belong to this tile's clips
  int32_t clip0;    // first clip index
  int32_t partial;  // bit 0: the tile holds a piece of a split clip (fused mode: combined atomically);
                    // bits 1..: extra pieces before this tile -> its first segment is number clip0 + (partial >> 1)
  uint32_t endmask[8];  // bit j of word c: column 32c+j is the last column of a segment
};
static_assert(sizeof(CTile) == 48, "CTile layout");

struct SimpoolParams {
  const CTile* ctiles;
  int32_t n_ctiles;
  int32_t n_rtiles;      // row tiles of (128 * cta_group) rows
  int32_t chunk_rtiles;  // row tiles per L2-resident phase
  int32_t n_rows_R;
  int32_t uni_len_R;        // > 0: every row-side clip has this many rows (no per-row lookup)
  const int4* rowinfo_R;    // per row {clip, clip's first row, clip's end row, 0}
  const int32_t* cu_R;
  const int32_t* cu_C;
  const float* rscale;  // nullable
  const float* cscale;  // nullable
  float* out;
  int64_t ld_r;
  int64_t ld_c;
  uint32_t idesc;
  int32_t c_policy;  // L2 eviction hint for column-operand tiles: 0 normal, 1 evict_last, 2 evict_first
  int32_t r_policy;  // the same for row-operand tiles (default 1)
  unsigned long long* trace;  // nullable debug buffer: 16 cycle counters per CTA (JEGAL_K1_TRACE=1)
  int32_t dense;  // 1: every clip on both sides is one row -> plain GEMM epilogue (no pooling)
};

// OP_NONE (row side only): no reduction over rows in K1 -- the per-row values go to a workspace matrix
// and launch_rowreduce finishes the pooling (two-pass mode for column sides made of many short clips).
enum Op : int { OP_SUM = 0, OP_MAX = 1, OP_NONE = 2 };

}  // namespace jegal

struct jegal_ctx {
  int device = -1;
  int sm_count = 0;
  std::string err;
  int64_t launches = 0;
  void* encode_tiled = nullptr;  // PFN_cuTensorMapEncodeTiled
  // K2 workspace for rows cut into column slices: per (row, slice) a 32-entry list + a ticket per row
  float* topk_ws_val = nullptr;
  int32_t* topk_ws_idx = nullptr;
  uint32_t* topk_ws_ticket = nullptr;
  size_t topk_ws_elems = 0, topk_ws_rows = 0;
  float* rowmat_ws = nullptr;  // two-pass K1 workspace [n column clips][row stride] fp32, grown on demand
  size_t rowmat_ws_elems = 0;
  uint32_t smem_configured = 0;  // bit per kernel instantiation whose max dynamic smem was raised on this device
};

struct jegal_layout {
  jegal_ctx* ctx = nullptr;
  int32_t n_clips = 0;
  int64_t rows = 0;
  int32_t max_len = 0;
  bool warp_aligned = false;  // every clip lies inside one aligned 32-row window
  std::vector<int32_t> cu_host;
  int32_t* cu_dev = nullptr;
  int4* rowinfo_dev = nullptr;  // per row {clip, first row, end row, 0}
  int32_t uniform_len = 0;      // > 0 when all clips have the same length
  struct CTileSet {
    int width = 0;
    int flags = 0;  // kPlanAllowSplit | kPlanCutHalves
    bool any_partial = false;
    int32_t extra_pieces = 0;      // segments - clips (pieces of split clips, half-tile cuts)
    std::vector<int32_t> seg_host; // [n_clips + 1] first segment number of every clip
    int32_t* seg_dev = nullptr;    // uploaded when extra_pieces > 0
    int n = 0;
    std::vector<jegal::CTile> host;
    jegal::CTile* dev = nullptr;
  };
  std::vector<CTileSet*> ctile_sets;
};

namespace jegal {

int set_err(jegal_ctx* ctx, int code, const std::string& msg);
#define JEGAL_CUDA_OK(ctx, expr)                                                        \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return ::jegal::set_err(ctx, JEGAL_ERR_CUDA,                                      \
                              std::string(#expr) + ": " + cudaGetErrorString(e__));    \
  } while (0)

// host-side launchers implemented in the kernel files
int launch_simpool(jegal_ctx* ctx, int cta_group, int col_op, int row_op, const CUtensorMap& tmR,
                   const CUtensorMap& tmC, const SimpoolParams& p, cudaStream_t stream);
// pass 2 of the two-pass mode: out[r * ld_r + c * ld_c] = rscale[r] * cscale[c] * ROWOP_{row in clip r} M[c * ldm + row]
// (row mean: the sum is divided by the clip's rows; col_op == OP_SUM: also by the column clip's rows)
// seg_C (nullable): column clip c owns the rows [seg_C[c], seg_C[c+1]) of M (pieces of a split clip, combined
// with col_op); null: row c.
int launch_rowreduce(jegal_ctx* ctx, const float* M, int64_t ldm, const int32_t* cu_R, int32_t n_rclips,
                     const int32_t* cu_C, const int32_t* seg_C, int32_t n_cclips, int col_op, int row_op, const float* rscale,
                     const float* cscale, float* out, int64_t ld_r, int64_t ld_c, cudaStream_t stream);
int launch_fill_f32(jegal_ctx* ctx, float* dst, int64_t n, float value, cudaStream_t stream);
int launch_rowinfo(jegal_ctx* ctx, const int32_t* cu_dev, int32_t n_clips, int64_t rows,
                   int4* rowinfo_dev, cudaStream_t stream);
int launch_prep(jegal_ctx* ctx, const jegal_layout* layout, const void* emb, int in_dtype,
                int normalize_rows, float row_eps, float mean_eps, int out_dtype, void* out_rows,
                float* inv_meannorm, void* mean_rows, cudaStream_t stream);
int launch_pair_cosine(jegal_ctx* ctx, const void* a_rows, const void* b_rows, int dtype, const int32_t* pair_a,
                       const int32_t* pair_b, int32_t n_pairs, int normalize, float eps, float* scores,
                       cudaStream_t stream);
int launch_segmean(jegal_ctx* ctx, const void* x, int in_dtype, int32_t dim, const int32_t* seg_begin,
                   const int32_t* seg_end, int32_t n_seg, void* out, int out_dtype, int64_t ld_out,
                   int32_t col_off, cudaStream_t stream);
int launch_topk(jegal_ctx* ctx, const float* scores, int32_t n_q, int32_t n_g, int64_t ld, int32_t k,
                int32_t idx_offset, float* out_val, int32_t* out_idx, cudaStream_t stream);
int launch_topk_merge(jegal_ctx* ctx, const float* vals, const int32_t* idxs, int32_t n_lists,
                      int32_t n_q, int32_t k, float* out_val, int32_t* out_idx, cudaStream_t stream);
int launch_rank_of_positive(jegal_ctx* ctx, const float* scores, int32_t n_q, int32_t n_g,
                            int64_t ld_row, int64_t ld_col, const int32_t* gt, int32_t* n_greater,
                            int32_t* n_equal, cudaStream_t stream);

// tensor map over [n_rows, 512] 16-bit rows, box = 64 elements x box_rows rows, 128 B swizzle
int make_box_tmap_impl(jegal_ctx* ctx, CUtensorMap* out, const void* rows_dev, int64_t n_rows, int op_dtype,
                       int box_rows);
int make_operand_tmap(jegal_ctx* ctx, CUtensorMap* out, const void* rows_dev, int64_t n_rows,
                      int op_dtype);

// Pure host logic (unit-tested without a GPU): pack whole clips greedily into column tiles of
// `width` rows; clips longer than `width` are cut into partial tiles when `allow_split`.
// Returns JEGAL_OK, or JEGAL_ERR_UNSUPPORTED with *bad_clip set.
// kPlanCutHalves: additionally no segment crosses column width/2 (a clip that would is cut there).
// seg_of_clip (nullable): [n_clips + 1] number of every clip's first segment.
constexpr int kPlanAllowSplit = 1, kPlanCutHalves = 2;
int plan_column_tiles(const int32_t* cu, int32_t n_clips, int width, int flags, std::vector<CTile>* out,
                      bool* any_partial, int32_t* bad_clip, std::vector<int32_t>* seg_of_clip);

}  // namespace jegal

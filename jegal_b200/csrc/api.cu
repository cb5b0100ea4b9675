// C ABI (include/jegal_b200.h): context, ragged layouts, tile planning and the
// host side of every entry point.  No torch types, no exceptions across the ABI.
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "internal.h"
#include "ptx.cuh"

namespace jegal {

int set_err(jegal_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int make_box_tmap_impl(jegal_ctx* ctx, CUtensorMap* out, const void* rows_dev, int64_t n_rows, int op_dtype,
                       int box_rows) {
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ctx->encode_tiled);
  const CUtensorMapDataType dt =
      op_dtype == JEGAL_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kD), static_cast<cuuint64_t>(n_rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kD) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, dt, 2, const_cast<void*>(rows_dev), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(ctx, JEGAL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return JEGAL_OK;
}

int make_operand_tmap(jegal_ctx* ctx, CUtensorMap* out, const void* rows_dev, int64_t n_rows,
                      int op_dtype) {
  return make_box_tmap_impl(ctx, out, rows_dev, n_rows, op_dtype, kTileRows);
}

int plan_column_tiles(const int32_t* cu, int32_t n_clips, int width, int flags, std::vector<CTile>* out,
                      bool* any_partial, int32_t* bad_clip, std::vector<int32_t>* seg_of_clip) {
  const bool allow_split = (flags & kPlanAllowSplit) != 0;
  const bool cut_halves = (flags & kPlanCutHalves) != 0;
  const int32_t h = width / 2;
  out->clear();
  *any_partial = false;
  if (seg_of_clip) seg_of_clip->assign(static_cast<size_t>(n_clips) + 1, 0);
  int32_t i = 0;
  int32_t extra = 0;  // segments beyond one per clip, so far (pieces of split clips, half-tile cuts)
  while (i < n_clips) {
    const int32_t len = cu[i + 1] - cu[i];
    if (len > width) {
      if (!allow_split) {
        *bad_clip = i;
        return JEGAL_ERR_UNSUPPORTED;
      }
      if (seg_of_clip) (*seg_of_clip)[i] = i + extra;
      int32_t segs_done = 0;
      for (int32_t off = 0; off < len; off += width) {
        CTile t{};
        t.row0 = cu[i] + off;
        t.n_valid = std::min(width, len - off);
        t.clip0 = i;
        t.partial = 1 | ((extra + segs_done) << 1);
        const int32_t e = t.n_valid - 1;
        t.endmask[e >> 5] |= 1u << (e & 31);
        ++segs_done;
        if (cut_halves && t.n_valid > h) {
          t.endmask[(h - 1) >> 5] |= 1u << ((h - 1) & 31);
          ++segs_done;
        }
        out->push_back(t);
        *any_partial = true;
      }
      extra += segs_done - 1;
      ++i;
      continue;
    }
    CTile t{};
    t.row0 = cu[i];
    t.clip0 = i;
    t.partial = extra << 1;
    const int32_t first = i;
    int32_t used = 0;
    while (i < n_clips) {
      const int32_t l = cu[i + 1] - cu[i];
      if (used + l > width) break;
      used += l;
      const int32_t e = used - 1;
      t.endmask[e >> 5] |= 1u << (e & 31);
      ++i;
    }
    t.n_valid = used;
    // two-pass mode with both warpgroups on every tile: no segment may cross the middle of the tile, so a
    // clip that does is cut there (its two pieces are combined by the second pass like a split clip's)
    const bool cut = cut_halves && used > h && !((t.endmask[(h - 1) >> 5] >> ((h - 1) & 31)) & 1u);
    if (seg_of_clip)
      for (int32_t c = first; c < i; ++c)
        (*seg_of_clip)[c] = c + extra + ((cut && cu[c] - cu[first] >= h) ? 1 : 0);
    if (cut) {
      t.endmask[(h - 1) >> 5] |= 1u << ((h - 1) & 31);
      ++extra;
    }
    out->push_back(t);
  }
  if (seg_of_clip) (*seg_of_clip)[n_clips] = n_clips + extra;
  return JEGAL_OK;
}

namespace {

int build_ctiles(jegal_ctx* ctx, const jegal_layout* L, int width, int flags, jegal_layout::CTileSet* set) {
  set->width = width;
  set->flags = flags;
  int32_t bad = -1;
  const int rc = plan_column_tiles(L->cu_host.data(), L->n_clips, width, flags, &set->host, &set->any_partial, &bad,
                                   &set->seg_host);
  if (rc != JEGAL_OK)
    return set_err(ctx, rc, "clip " + std::to_string(bad) + " has " +
                                std::to_string(L->cu_host[bad + 1] - L->cu_host[bad]) +
                                " rows on the column side; a max-then-mean pooling needs <= " + std::to_string(width));
  set->n = static_cast<int>(set->host.size());
  set->extra_pieces = set->seg_host[L->n_clips] - L->n_clips;
  return JEGAL_OK;
}

int get_ctiles(jegal_ctx* ctx, jegal_layout* L, int width, int flags, cudaStream_t stream,
               jegal_layout::CTileSet** out) {
  for (auto* s : L->ctile_sets) {
    // a no-split set without partial tiles also serves callers that would allow splitting
    if (s->width == width && (s->flags & kPlanCutHalves) == (flags & kPlanCutHalves) &&
        ((s->flags & kPlanAllowSplit) == (flags & kPlanAllowSplit) || !s->any_partial)) {
      *out = s;
      return JEGAL_OK;
    }
  }
  auto* set = new (std::nothrow) jegal_layout::CTileSet();
  if (!set) return set_err(ctx, JEGAL_ERR_NOMEM, "out of host memory");
  int rc = build_ctiles(ctx, L, width, flags, set);
  if (rc != JEGAL_OK) {
    delete set;
    return rc;
  }
  if (set->n > 0) {
    cudaError_t e = cudaMalloc(&set->dev, sizeof(CTile) * set->n);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(set->dev, set->host.data(), sizeof(CTile) * set->n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && set->extra_pieces > 0) {  // first segment number of every clip, for the two-pass mode
      e = cudaMalloc(&set->seg_dev, sizeof(int32_t) * set->seg_host.size());
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(set->seg_dev, set->seg_host.data(), sizeof(int32_t) * set->seg_host.size(),
                            cudaMemcpyHostToDevice, stream);
    }
    // one-time cost: later calls may run on another stream
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
      if (set->dev) cudaFree(set->dev);
      if (set->seg_dev) cudaFree(set->seg_dev);
      delete set;
      return set_err(ctx, JEGAL_ERR_CUDA, std::string("ctile upload: ") + cudaGetErrorString(e));
    }
  }
  L->ctile_sets.push_back(set);
  *out = set;
  return JEGAL_OK;
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

}  // namespace
}  // namespace jegal

using namespace jegal;

extern "C" {

const char* jegal_version(void) { return "jegal_b200 0.1 (sm_100a)"; }

int jegal_plan_column_tiles(const int32_t* cu_len_host, int32_t n_clips, int32_t width, int32_t allow_split,
                            jegal_column_tile* out, int32_t max_out, int32_t* n_out) {
  if (!cu_len_host || n_clips < 0 || width < 32 || width > 256 || (width & 31) || !n_out) return JEGAL_ERR_ARG;
  std::vector<CTile> tiles;
  bool any_partial = false;
  int32_t bad = -1;
  const int rc = plan_column_tiles(cu_len_host, n_clips, width, allow_split & 3, &tiles, &any_partial, &bad, nullptr);
  if (rc != JEGAL_OK) {
    *n_out = bad;
    return rc;
  }
  *n_out = static_cast<int32_t>(tiles.size());
  if (out) {
    static_assert(sizeof(jegal_column_tile) == sizeof(CTile), "public and internal tile layouts must agree");
    const int32_t n = std::min<int32_t>(max_out, *n_out);
    std::memcpy(out, tiles.data(), sizeof(CTile) * static_cast<size_t>(n));
  }
  return JEGAL_OK;
}

int jegal_ctx_create(int device, jegal_ctx** out) {
  if (!out) return JEGAL_ERR_ARG;
  *out = nullptr;
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JEGAL_ERR_DEVICE;
  if (prop.major != 10) return JEGAL_ERR_DEVICE;  // sm_100a only: no fallback path exists
  if (cudaSetDevice(device) != cudaSuccess) return JEGAL_ERR_DEVICE;
  auto* ctx = new (std::nothrow) jegal_ctx();
  if (!ctx) return JEGAL_ERR_NOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn) {
    delete ctx;
    return JEGAL_ERR_CUDA;
  }
  ctx->encode_tiled = fn;
  *out = ctx;
  return JEGAL_OK;
}

void jegal_ctx_destroy(jegal_ctx* ctx) {
  if (!ctx) return;
  if (ctx->topk_ws_val) cudaFree(ctx->topk_ws_val);
  if (ctx->topk_ws_idx) cudaFree(ctx->topk_ws_idx);
  if (ctx->topk_ws_ticket) cudaFree(ctx->topk_ws_ticket);
  if (ctx->rowmat_ws) cudaFree(ctx->rowmat_ws);
  delete ctx;
}

const char* jegal_last_error(const jegal_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int64_t jegal_launch_count(const jegal_ctx* ctx) { return ctx ? ctx->launches : 0; }

int jegal_layout_create(jegal_ctx* ctx, const int32_t* cu_len_host, int32_t n_clips, void* stream_,
                        jegal_layout** out) {
  JEGAL_NVTX("jegal_layout_create");
  if (!ctx || !out || !cu_len_host || n_clips < 0) return set_err(ctx, JEGAL_ERR_ARG, "layout_create: null/negative argument");
  *out = nullptr;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cu_len_host[0] != 0) return set_err(ctx, JEGAL_ERR_ARG, "layout_create: cu_len[0] must be 0");
  bool aligned = true;
  int32_t max_len = 0, min_len = 0x7fffffff;
  for (int32_t i = 0; i < n_clips; ++i) {
    const int32_t b = cu_len_host[i], e = cu_len_host[i + 1];
    if (e <= b)
      return set_err(ctx, JEGAL_ERR_ARG, "layout_create: clip " + std::to_string(i) + " is empty or offsets decrease");
    max_len = std::max(max_len, e - b);
    min_len = std::min(min_len, e - b);
    if ((b >> 5) != ((e - 1) >> 5)) aligned = false;
  }
  auto* L = new (std::nothrow) jegal_layout();
  if (!L) return set_err(ctx, JEGAL_ERR_NOMEM, "out of host memory");
  L->ctx = ctx;
  L->n_clips = n_clips;
  L->rows = cu_len_host[n_clips];
  L->max_len = max_len;
  L->warp_aligned = aligned;
  L->uniform_len = (n_clips > 0 && min_len == max_len) ? max_len : 0;
  L->cu_host.assign(cu_len_host, cu_len_host + n_clips + 1);
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaMalloc(&L->cu_dev, sizeof(int32_t) * (n_clips + 1));
  if (e == cudaSuccess) e = cudaMalloc(&L->rowinfo_dev, sizeof(int4) * std::max<int64_t>(L->rows, 1));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(L->cu_dev, L->cu_host.data(), sizeof(int32_t) * (n_clips + 1), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) {
    jegal_layout_destroy(L);
    return set_err(ctx, JEGAL_ERR_CUDA, std::string("layout_create: ") + cudaGetErrorString(e));
  }
  int rc = launch_rowinfo(ctx, L->cu_dev, n_clips, L->rows, L->rowinfo_dev, stream);
  if (rc == JEGAL_OK && cudaStreamSynchronize(stream) != cudaSuccess)  // one-time: layout is usable on any stream
    rc = set_err(ctx, JEGAL_ERR_CUDA, "layout_create: synchronize failed");
  if (rc != JEGAL_OK) {
    jegal_layout_destroy(L);
    return rc;
  }
  *out = L;
  return JEGAL_OK;
}

void jegal_layout_destroy(jegal_layout* L) {
  if (!L) return;
  if (L->cu_dev) cudaFree(L->cu_dev);
  if (L->rowinfo_dev) cudaFree(L->rowinfo_dev);
  for (auto* s : L->ctile_sets) {
    if (s->dev) cudaFree(s->dev);
    if (s->seg_dev) cudaFree(s->seg_dev);
    delete s;
  }
  delete L;
}

int64_t jegal_layout_rows(const jegal_layout* L) { return L ? L->rows : 0; }
int32_t jegal_layout_clips(const jegal_layout* L) { return L ? L->n_clips : 0; }

int jegal_prep(jegal_ctx* ctx, const jegal_layout* layout, const void* emb_dev, int in_dtype,
               int normalize_rows, float row_eps, float mean_eps, int out_dtype, void* out_rows_dev,
               float* inv_meannorm_dev, void* mean_rows_dev, void* stream) {
  JEGAL_NVTX("jegal_prep (K0)");
  if (!ctx || !layout) return set_err(ctx, JEGAL_ERR_ARG, "prep: null argument");
  if (layout->n_clips == 0) return JEGAL_OK;
  if (!emb_dev || !out_rows_dev) return set_err(ctx, JEGAL_ERR_ARG, "prep: null argument");
  if ((reinterpret_cast<uintptr_t>(emb_dev) | reinterpret_cast<uintptr_t>(out_rows_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "prep: buffers must be 16-byte aligned");
  return launch_prep(ctx, layout, emb_dev, in_dtype, normalize_rows, row_eps, mean_eps, out_dtype,
                     out_rows_dev, inv_meannorm_dev, mean_rows_dev, static_cast<cudaStream_t>(stream));
}

int jegal_clip_means(jegal_ctx* ctx, const jegal_layout* layout, const void* emb_dev, int in_dtype, float mean_eps,
                     int out_dtype, void* mean_rows_dev, float* inv_meannorm_dev, void* stream) {
  JEGAL_NVTX("jegal_clip_means (K0 means)");
  if (!ctx || !layout) return set_err(ctx, JEGAL_ERR_ARG, "clip_means: null argument");
  if (layout->n_clips == 0) return JEGAL_OK;
  if (!emb_dev || (!mean_rows_dev && !inv_meannorm_dev)) return set_err(ctx, JEGAL_ERR_ARG, "clip_means: null argument");
  if ((reinterpret_cast<uintptr_t>(emb_dev) | reinterpret_cast<uintptr_t>(mean_rows_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "clip_means: buffers must be 16-byte aligned");
  if (out_dtype != JEGAL_F32 && out_dtype != JEGAL_F16 && out_dtype != JEGAL_BF16)
    return set_err(ctx, JEGAL_ERR_ARG, "clip_means: bad out_dtype");
  return launch_prep(ctx, layout, emb_dev, in_dtype, 0, 1e-12f, mean_eps, out_dtype, nullptr, inv_meannorm_dev,
                     mean_rows_dev, static_cast<cudaStream_t>(stream));
}

int jegal_pair_cosine(jegal_ctx* ctx, const void* a_rows_dev, int64_t n_a, const void* b_rows_dev, int64_t n_b,
                      int dtype, const int32_t* pair_a_dev, const int32_t* pair_b_dev, int32_t n_pairs,
                      int normalize, float eps, float* scores_dev, void* stream) {
  JEGAL_NVTX("jegal_pair_cosine");
  if (!ctx) return JEGAL_ERR_ARG;
  if (n_pairs < 0 || n_a < 0 || n_b < 0) return set_err(ctx, JEGAL_ERR_ARG, "pair_cosine: negative size");
  if (n_pairs == 0) return JEGAL_OK;
  if (!a_rows_dev || !b_rows_dev || !scores_dev) return set_err(ctx, JEGAL_ERR_ARG, "pair_cosine: null argument");
  if ((!pair_a_dev && n_pairs > n_a) || (!pair_b_dev && n_pairs > n_b))
    return set_err(ctx, JEGAL_ERR_ARG, "pair_cosine: n_pairs exceeds the rows of an unlisted side");
  if ((reinterpret_cast<uintptr_t>(a_rows_dev) | reinterpret_cast<uintptr_t>(b_rows_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "pair_cosine: rows must be 16-byte aligned");
  if (normalize && !(eps > 0.f)) return set_err(ctx, JEGAL_ERR_ARG, "pair_cosine: eps must be > 0");
  return launch_pair_cosine(ctx, a_rows_dev, b_rows_dev, dtype, pair_a_dev, pair_b_dev, n_pairs, normalize, eps,
                            scores_dev, static_cast<cudaStream_t>(stream));
}

int jegal_simpool_allpairs(jegal_ctx* ctx, const jegal_layout* gest_layout, const void* gest_rows_dev,
                           const jegal_layout* cont_layout, const void* cont_rows_dev, int op_dtype,
                           int pool_mode, const float* gscale_dev, const float* cscale_dev,
                           float* scores_dev, int64_t ld_g, int64_t ld_c, void* stream_) {
  JEGAL_NVTX("jegal_simpool_allpairs (K1)");
  if (!ctx || !gest_layout || !cont_layout) return set_err(ctx, JEGAL_ERR_ARG, "simpool_allpairs: null argument");
  if (op_dtype != JEGAL_BF16 && op_dtype != JEGAL_F16)
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_allpairs: op_dtype must be JEGAL_BF16 or JEGAL_F16");
  const int32_t nG = gest_layout->n_clips, nC = cont_layout->n_clips;
  if (nG == 0 || nC == 0) return JEGAL_OK;  // an empty side: empty score matrix, nothing to launch
  if (!gest_rows_dev || !cont_rows_dev || !scores_dev)
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_allpairs: null argument");
  if (!((ld_g == nC && ld_c == 1) || (ld_g == 1 && ld_c == nG)))
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_allpairs: (ld_g, ld_c) must be (n_cont, 1) or (1, n_gest)");
  if ((reinterpret_cast<uintptr_t>(gest_rows_dev) | reinterpret_cast<uintptr_t>(cont_rows_dev)) & 15u)
    return set_err(ctx, JEGAL_ERR_ARG, "simpool_allpairs: operand rows must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  // Orientation: the first reduction must run along columns (in-thread).
  bool cols_are_gest;
  int col_op, row_op;
  switch (pool_mode) {
    case JEGAL_POOL_MEAN_MEAN: cols_are_gest = false; col_op = OP_SUM; row_op = OP_SUM; break;
    case JEGAL_POOL_MAX_T_MEAN_W: cols_are_gest = true; col_op = OP_MAX; row_op = OP_SUM; break;
    case JEGAL_POOL_MAX_W_MEAN_T: cols_are_gest = false; col_op = OP_MAX; row_op = OP_SUM; break;
    case JEGAL_POOL_MAX_MAX: cols_are_gest = false; col_op = OP_MAX; row_op = OP_MAX; break;
    default: return set_err(ctx, JEGAL_ERR_ARG, "simpool_allpairs: bad pool_mode");
  }
  if (col_op == row_op) {
    // either orientation is exact: put the longer clips on the column side (fewer, longer
    // in-register reductions; fewer cross-lane combines per tile)
    cols_are_gest = gest_layout->rows * static_cast<int64_t>(nC) >= cont_layout->rows * static_cast<int64_t>(nG);
    // one-row clips on both sides (plain GEMM epilogue): the columns are the side whose index is contiguous in
    // the output, so every epilogue thread owns a contiguous stretch of an output row (256-bit stores)
    if (gest_layout->uniform_len == 1 && cont_layout->uniform_len == 1) cols_are_gest = ld_g == 1;
    const int force = env_int("JEGAL_SIMPOOL_COLS", -1);  // testing knob: 0 = content, 1 = gesture
    if (force >= 0) cols_are_gest = force != 0;
  }

  jegal_layout* LC = const_cast<jegal_layout*>(cols_are_gest ? gest_layout : cont_layout);
  const jegal_layout* LR = cols_are_gest ? cont_layout : gest_layout;
  const void* rowsC = cols_are_gest ? gest_rows_dev : cont_rows_dev;
  const void* rowsR = cols_are_gest ? cont_rows_dev : gest_rows_dev;

  int cg = env_int("JEGAL_CTA_GROUP", 2);
  if (cg != 1 && cg != 2) cg = 2;
  if (ctx->sm_count < 2) cg = 1;
  const int width = kTileRows * cg;

  // Two-pass mode.  The fused epilogue pays ~50 dependent instructions (segmented shuffles + an atomic) per
  // column clip per tile; with many short column clips (word clips on the column side: ~12 per tile) it
  // takes longer than the tile's MMAs (measured: 57 % of tensor peak on config 2, max_w_mean_t).  There K1
  // only pools along columns and stores the per-row values as M[column segment][row] (one coalesced 128-byte
  // store per warp and segment), and launch_rowreduce finishes over rows: 2 x rows x segments x 4 bytes of
  // extra HBM traffic, no atomics, no output initialisation, bitwise reproducible.  Its plan cuts segments
  // at the middle of every tile, so each of the two epilogue warpgroups pools one half of EVERY tile.
  // JEGAL_ROWMAT: -1 auto (default), 0 never, 1 always; JEGAL_ROWMAT_MAX_MB bounds the workspace (default 1024).
  const int rowmat_mode = env_int("JEGAL_ROWMAT", -1);
  const bool dense = LR->uniform_len == 1 && LC->uniform_len == 1;  // 1 x 1 tiles: all pooling modes coincide
  // auto: column clips shorter than 48 rows on average (>= ~5.3 segments per 256-column tile)
  const bool many_short = LC->rows < static_cast<int64_t>(48) * LC->n_clips;
  const int64_t ldm = (LR->rows + 31) & ~static_cast<int64_t>(31);

  jegal_layout::CTileSet* cts = nullptr;
  int rc = get_ctiles(ctx, LC, width, col_op == row_op ? kPlanAllowSplit : 0, stream, &cts);
  // A max-then-mean pooling cannot combine the pieces of a column clip longer than the tile after the
  // row reduction; the two-pass mode can (its second pass sees the per-row values of every piece).
  const bool need_two_pass = rc == JEGAL_ERR_UNSUPPORTED && col_op != row_op;
  const std::string unsupported_why = need_two_pass ? ctx->err : std::string();
  if (rc != JEGAL_OK && !need_two_pass) return rc;

  bool two_pass = false;
  int64_t ws_elems = 0;
  if (!dense && rowmat_mode != 0 && (rowmat_mode == 1 || need_two_pass || many_short)) {
    jegal_layout::CTileSet* cts2 = nullptr;
    rc = get_ctiles(ctx, LC, width, kPlanAllowSplit | kPlanCutHalves, stream, &cts2);
    if (rc != JEGAL_OK) return rc;
    ws_elems = ldm * (static_cast<int64_t>(LC->n_clips) + cts2->extra_pieces);  // one row per column segment
    const int64_t max_bytes = static_cast<int64_t>(env_int("JEGAL_ROWMAT_MAX_MB", 1024)) << 20;
    two_pass = ws_elems * 4 <= max_bytes && ldm < (1ll << 31);
    if (two_pass && ctx->rowmat_ws_elems < static_cast<size_t>(ws_elems)) {
      if (ctx->rowmat_ws) {
        JEGAL_CUDA_OK(ctx, cudaStreamSynchronize(stream));
        cudaFree(ctx->rowmat_ws);
        ctx->rowmat_ws = nullptr;
        ctx->rowmat_ws_elems = 0;
      }
      if (cudaMalloc(&ctx->rowmat_ws, static_cast<size_t>(ws_elems) * 4) != cudaSuccess) {
        cudaGetLastError();
        ctx->rowmat_ws = nullptr;
        two_pass = false;  // not enough device memory for the workspace: fused single pass
      } else {
        ctx->rowmat_ws_elems = static_cast<size_t>(ws_elems);
      }
    }
    if (two_pass) cts = cts2;
  }
  if (need_two_pass && !two_pass)
    return set_err(ctx, JEGAL_ERR_UNSUPPORTED,
                   unsupported_why + " in one pass; the two-pass mode needs a workspace of " +
                       std::to_string((ldm * (static_cast<int64_t>(LC->n_clips) + 8) * 4) >> 20) +
                       "+ MB (JEGAL_ROWMAT_MAX_MB, JEGAL_ROWMAT)");

  SimpoolParams p{};
  p.ctiles = cts->dev;
  p.n_ctiles = cts->n;
  p.n_rtiles = static_cast<int32_t>((LR->rows + width - 1) / width);
  // L2 plan: DRAM traffic of a launch = R once + (#phases) x C, because a column tile has no reuse
  // inside a phase; so phases of R are made as large as stays L2-resident while every cluster
  // re-streams them, and column tiles are fetched with evict_first so they do not displace R.
  p.c_policy = env_int("JEGAL_C_POLICY", 2);
  p.r_policy = env_int("JEGAL_R_POLICY", 1);
  const int64_t chunk_bytes = static_cast<int64_t>(env_int("JEGAL_CHUNK_MB", 24)) << 20;
  p.chunk_rtiles = static_cast<int32_t>(std::max<int64_t>(1, chunk_bytes / (static_cast<int64_t>(width) * kD * 2)));
  p.n_rows_R = static_cast<int32_t>(LR->rows);
  p.rowinfo_R = LR->rowinfo_dev;
  p.uni_len_R = LR->uniform_len;
  p.dense = dense ? 1 : 0;
  p.cu_R = LR->cu_dev;
  p.cu_C = LC->cu_dev;
  p.rscale = cols_are_gest ? cscale_dev : gscale_dev;
  p.cscale = cols_are_gest ? gscale_dev : cscale_dev;
  p.out = scores_dev;
  p.ld_r = cols_are_gest ? ld_c : ld_g;
  p.ld_c = cols_are_gest ? ld_g : ld_c;
  p.idesc = ptx::make_idesc_f16(op_dtype == JEGAL_BF16 ? 1u : 0u, static_cast<uint32_t>(width),
                                static_cast<uint32_t>(width));

  const int final_row_op = row_op;
  if (two_pass) {
    p.out = ctx->rowmat_ws;
    p.ld_r = 1;
    p.ld_c = ldm;
    p.rscale = nullptr;
    p.cscale = nullptr;
    row_op = OP_NONE;
  }

  // Results of clips that straddle warps / row tiles / column tiles are combined
  // with atomics and need an initialised output.
  if (!two_pass && (!LR->warp_aligned || cts->any_partial)) {
    rc = launch_fill_f32(ctx, scores_dev, static_cast<int64_t>(nG) * nC, row_op == OP_MAX ? -INFINITY : 0.0f, stream);
    if (rc != JEGAL_OK) return rc;
  }
  auto finish = [&]() -> int {
    if (!two_pass) return JEGAL_OK;
    return launch_rowreduce(ctx, ctx->rowmat_ws, ldm, LR->cu_dev, LR->n_clips, LC->cu_dev, cts->seg_dev, LC->n_clips,
                            col_op, final_row_op, cols_are_gest ? cscale_dev : gscale_dev,
                            cols_are_gest ? gscale_dev : cscale_dev, scores_dev, cols_are_gest ? ld_c : ld_g,
                            cols_are_gest ? ld_g : ld_c, stream);
  };

  CUtensorMap tmR, tmC;
  rc = make_operand_tmap(ctx, &tmR, rowsR, LR->rows, op_dtype);
  if (rc != JEGAL_OK) return rc;
  rc = make_operand_tmap(ctx, &tmC, rowsC, LC->rows, op_dtype);
  if (rc != JEGAL_OK) return rc;
  if (env_int("JEGAL_K1_TRACE", 0)) {  // debug: per-role cycle accounting, printed to stderr (synchronises)
    const int ncta = ctx->sm_count;
    unsigned long long* tbuf = nullptr;
    JEGAL_CUDA_OK(ctx, cudaMalloc(&tbuf, sizeof(unsigned long long) * 16 * ncta));
    JEGAL_CUDA_OK(ctx, cudaMemsetAsync(tbuf, 0, sizeof(unsigned long long) * 16 * ncta, stream));
    p.trace = tbuf;
    rc = launch_simpool(ctx, cg, col_op, row_op, tmR, tmC, p, stream);
    JEGAL_CUDA_OK(ctx, cudaStreamSynchronize(stream));
    std::vector<unsigned long long> h(16 * ncta);
    cudaMemcpy(h.data(), tbuf, sizeof(unsigned long long) * 16 * ncta, cudaMemcpyDeviceToHost);
    cudaFree(tbuf);
    double acc[16] = {0};
    int n_lead = 0;
    for (int b = 0; b < ncta; ++b)
      if (h[16 * b + 7] > 0) { ++n_lead; for (int k = 0; k < 16; ++k) acc[k] += static_cast<double>(h[16 * b + k]); }
    if (n_lead > 0) {
      for (int k = 0; k < 16; ++k) acc[k] /= n_lead;
      std::fprintf(stderr,
                   "%s[K1 trace, mean cycles over %d leader CTAs] producer: wait r_empty %.0f, wait c_empty %.0f, total %.0f | "
                   "mma: wait t_empty %.0f, wait c_full %.0f, wait r_full %.0f, total %.0f | "
                   "epi0: wait t_full %.0f, pool %.0f, total %.0f, tiles %.0f | epi1: wait t_full %.0f, pool %.0f, total %.0f\n",
                   two_pass ? "(two-pass) " : "", n_lead, acc[0], acc[1], acc[2], acc[4], acc[5], acc[6], acc[7], acc[8], acc[9], acc[10], acc[11], acc[12],
                   acc[13], acc[14]);
    }
    return rc != JEGAL_OK ? rc : finish();
  }
  rc = launch_simpool(ctx, cg, col_op, row_op, tmR, tmC, p, stream);
  return rc != JEGAL_OK ? rc : finish();
}

int jegal_topk(jegal_ctx* ctx, const float* scores_dev, int32_t n_q, int32_t n_g, int64_t ld, int32_t k,
               int32_t idx_offset, float* topk_val_dev, int32_t* topk_idx_dev, void* stream) {
  JEGAL_NVTX("jegal_topk (K2)");
  if (!ctx || (!scores_dev && n_g > 0 && n_q > 0) || !topk_val_dev || !topk_idx_dev)
    return set_err(ctx, JEGAL_ERR_ARG, "topk: null argument");
  if (k < 1 || k > 32) return set_err(ctx, JEGAL_ERR_UNSUPPORTED, "topk: k must be in [1, 32]");
  if (n_q < 0 || n_g < 0 || ld < n_g) return set_err(ctx, JEGAL_ERR_ARG, "topk: bad shape");
  return launch_topk(ctx, scores_dev, n_q, n_g, ld, k, idx_offset, topk_val_dev, topk_idx_dev,
                     static_cast<cudaStream_t>(stream));
}

int jegal_topk_merge(jegal_ctx* ctx, const float* vals_dev, const int32_t* idxs_dev, int32_t n_lists,
                     int32_t n_q, int32_t k, float* out_val_dev, int32_t* out_idx_dev, void* stream) {
  JEGAL_NVTX("jegal_topk_merge (K2)");
  if (!ctx || !vals_dev || !idxs_dev || !out_val_dev || !out_idx_dev)
    return set_err(ctx, JEGAL_ERR_ARG, "topk_merge: null argument");
  if (k < 1 || k > 32) return set_err(ctx, JEGAL_ERR_UNSUPPORTED, "topk_merge: k must be in [1, 32]");
  if (n_lists < 1 || n_q < 0) return set_err(ctx, JEGAL_ERR_ARG, "topk_merge: bad shape");
  return launch_topk_merge(ctx, vals_dev, idxs_dev, n_lists, n_q, k, out_val_dev, out_idx_dev,
                           static_cast<cudaStream_t>(stream));
}

int jegal_rank_of_positive(jegal_ctx* ctx, const float* scores_dev, int32_t n_q, int32_t n_g,
                           int64_t ld_row, int64_t ld_col, const int32_t* gt_dev, int32_t* n_greater_dev,
                           int32_t* n_equal_dev, void* stream) {
  JEGAL_NVTX("jegal_rank_of_positive (K2)");
  if (!ctx || !scores_dev || !n_greater_dev) return set_err(ctx, JEGAL_ERR_ARG, "rank_of_positive: null argument");
  if (n_q < 0 || n_g < 0) return set_err(ctx, JEGAL_ERR_ARG, "rank_of_positive: bad shape");
  if (!gt_dev && n_q > n_g) return set_err(ctx, JEGAL_ERR_ARG, "rank_of_positive: diagonal needs n_q <= n_g");
  return launch_rank_of_positive(ctx, scores_dev, n_q, n_g, ld_row, ld_col, gt_dev, n_greater_dev,
                                 n_equal_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

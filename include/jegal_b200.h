/*
 * jegal_b200 — C ABI of the B200-native (sm_100a) JEGAL cross-modal scoring path.
 *
 * The reference (Sindhu-Hegde/jegal) has no FFI: its scoring "API" is a set of
 * module-level Python functions that call torch-CPU / numpy.  Each entry point
 * below names the reference code it replaces (paths relative to the reference
 * root).  The Python mirror of those functions (same names, same arguments)
 * lives in jegal_b200/scoring.py and binds this header through ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - return 0 on success, a negative jegal_status otherwise; no C++ exception
 *     crosses the boundary; jegal_last_error(ctx) describes the last failure.
 *   - every "dev" pointer is caller-owned device memory on the ctx's device,
 *     16-byte aligned and contiguous; "host" pointers are ordinary host memory
 *     that is consumed before the call returns.
 *   - all kernels are enqueued on `stream` (a cudaStream_t passed as void*) and
 *     the call returns without synchronising.  The *_host convenience entry
 *     points (suffix _host) are the exception: they copy in, run, copy out and
 *     synchronise, because that is what a drop-in of a CPU function must do.
 *   - a ctx is bound to one device and is not thread-safe.
 *   - there is no CPU fallback: a device that is not sm_100 yields
 *     JEGAL_ERR_DEVICE.
 *   - D (embedding width) is fixed at 512 (models/jegal.py:18,71-76).
 */
#ifndef JEGAL_B200_H_
#define JEGAL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JEGAL_EMB_DIM 512

typedef enum {
  JEGAL_OK = 0,
  JEGAL_ERR_ARG = -1,         /* bad argument (null pointer, negative size, bad enum) */
  JEGAL_ERR_DEVICE = -2,      /* not an sm_100 device / device mismatch */
  JEGAL_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed */
  JEGAL_ERR_UNSUPPORTED = -4, /* shape outside what the kernels cover (see message) */
  JEGAL_ERR_NOMEM = -5
} jegal_status;

typedef enum { JEGAL_F32 = 0, JEGAL_F16 = 1, JEGAL_BF16 = 2 } jegal_dtype;

/* How the T x W cosine tile of one (gesture clip, content clip) pair is pooled.
 * MEAN_MEAN with the per-clip scales returned by jegal_prep(.., inv_meannorm)
 * reproduces evaluation/evaluate_retrieval.py:30-31,38-48 and
 * evaluation/evaluate_asd.py:31-36,43-47 exactly. */
typedef enum {
  JEGAL_POOL_MEAN_MEAN = 0,    /* mean over frames and words            */
  JEGAL_POOL_MAX_T_MEAN_W = 1, /* max over frames, then mean over words */
  JEGAL_POOL_MAX_W_MEAN_T = 2, /* max over words, then mean over frames */
  JEGAL_POOL_MAX_MAX = 3       /* max over both                         */
} jegal_pool_mode;

typedef struct jegal_ctx jegal_ctx;
/* Ragged layout of one packed operand: n_clips clips, clip i owns rows
 * [cu_len[i], cu_len[i+1]) of a [cu_len[n_clips], 512] row-major matrix. */
typedef struct jegal_layout jegal_layout;

const char* jegal_version(void);

int jegal_ctx_create(int device, jegal_ctx** out);
void jegal_ctx_destroy(jegal_ctx* ctx);
const char* jegal_last_error(const jegal_ctx* ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
int64_t jegal_launch_count(const jegal_ctx* ctx);

/* cu_len_host: n_clips + 1 non-decreasing int32 offsets starting at 0.
 * Uploads the offsets and builds the row->clip map on the device (async on stream). */
int jegal_layout_create(jegal_ctx* ctx, const int32_t* cu_len_host, int32_t n_clips, void* stream,
                        jegal_layout** out);
void jegal_layout_destroy(jegal_layout* layout);
int64_t jegal_layout_rows(const jegal_layout* layout);
int32_t jegal_layout_clips(const jegal_layout* layout);

/* K0 — normalise + cast (+ per-clip mean-vector norm), one pass over the rows.
 * Replaces F.normalize(x, p=2, dim=-1) at inference_embs.py:630-636,
 * evaluation/evaluate_spotting.py:49-50 and the numpy .mean(axis=0) +
 * F.normalize / CosineSimilarity norms at evaluate_retrieval.py:30-31,41,44 and
 * evaluate_asd.py:32,36,45-47.
 *   emb_dev        [rows, 512] in `in_dtype`
 *   normalize_rows 1: out row = row / max(||row||, row_eps) (fp32 math), 0: cast only
 *   out_rows_dev   [rows, 512] in `out_dtype` (JEGAL_BF16 or JEGAL_F16)
 *   inv_meannorm_dev (nullable) [n_clips] fp32: 1 / max(||mean of the clip's INPUT rows||, mean_eps)
 *   mean_rows_dev  (nullable) [n_clips, 512] in `out_dtype`: mean row / max(||mean row||, mean_eps),
 *                  i.e. the clip-level embedding evaluate_retrieval.py:30-31,41 feeds to its matmul
 */
int jegal_prep(jegal_ctx* ctx, const jegal_layout* layout, const void* emb_dev, int in_dtype,
               int normalize_rows, float row_eps, float mean_eps, int out_dtype, void* out_rows_dev,
               float* inv_meannorm_dev, void* mean_rows_dev, void* stream);

/* K0 without the row output: one read-only pass that yields, per clip, the unit-norm mean row
 * (mean over the clip's INPUT rows, divided by max(||mean||, mean_eps)) in `out_dtype` (JEGAL_F32, JEGAL_F16 or
 * JEGAL_BF16; nullable) and/or 1 / max(||mean||, mean_eps) (nullable).  This is load_feats' temporal mean of
 * evaluate_retrieval.py:30-31 / evaluate_asd.py:31-36 (numpy's fp16 result rounding is mirrored for fp16
 * inputs) followed by the norm of F.normalize / CosineSimilarity; bytes: rows x 512 x b_in read, n_clips rows written. */
int jegal_clip_means(jegal_ctx* ctx, const jegal_layout* layout, const void* emb_dev, int in_dtype, float mean_eps,
                     int out_dtype, void* mean_rows_dev, float* inv_meannorm_dev, void* stream);

/* Cosine (normalize = 1) or dot product (0) of LISTED pairs of 512-wide rows, one warp per pair:
 *   scores[p] = a[pair_a[p]] . b[pair_b[p]] / (max(||a||, eps) max(||b||, eps))
 * — nn.CosineSimilarity(dim=1, eps=1e-8) of evaluate_asd.py:45-47 on clip-level embeddings (the outputs of
 * jegal_clip_means, or get_similarity_cos' (1, 512) / (P, 512) arguments).  Both matrices are [n, 512] in
 * `dtype` (JEGAL_F32 / F16 / BF16); pair_a / pair_b nullable (=> p). */
int jegal_pair_cosine(jegal_ctx* ctx, const void* a_rows_dev, int64_t n_a, const void* b_rows_dev, int64_t n_b,
                      int dtype, const int32_t* pair_a_dev, const int32_t* pair_b_dev, int32_t n_pairs,
                      int normalize, float eps, float* scores_dev, void* stream);

/* K1 — all-pairs fused similarity + pooling (tcgen05 / TMEM / TMA).
 * scores[g * ld_g + c * ld_c] = gscale[g] * cscale[c] * pool_{t,w}(G_g C_c^T).
 * Generalises get_similarity_matrix (evaluation/evaluate_retrieval.py:38-48) from
 * mean-pooled vectors to the frame x word tile of every clip pair; the T x W
 * similarity tile lives only in tensor memory.
 *   gest_rows_dev / cont_rows_dev : outputs of jegal_prep, both in `op_dtype`
 *   gscale_dev / cscale_dev       : nullable per-clip fp32 multipliers (> 0)
 *   (ld_g, ld_c) must be (n_cont, 1) or (1, n_gest): a dense matrix or its transpose.
 */
int jegal_simpool_allpairs(jegal_ctx* ctx, const jegal_layout* gest_layout,
                           const void* gest_rows_dev, const jegal_layout* cont_layout,
                           const void* cont_rows_dev, int op_dtype, int pool_mode,
                           const float* gscale_dev, const float* cscale_dev, float* scores_dev,
                           int64_t ld_g, int64_t ld_c, void* stream);

/* Host-only planning step of K1 (no device, no ctx): how the clips of the column-side operand are
 * packed into tiles of `width` (128 or 256) columns.  Whole clips are packed greedily; a clip
 * longer than `width` is cut into pieces, one tile each, when allow_split (the fused single-pass
 * kernel can combine pieces only when both pooling reductions are the same operation; the two-pass
 * mode always can), otherwise JEGAL_ERR_UNSUPPORTED is returned with *n_out = the offending clip.
 * partial: bit 0 = the tile is a piece of a split clip; bits 1.. = pieces beyond the first of all
 * split clips before this tile, i.e. the tile's first column segment is number clip0 + (partial >> 1).
 * Bit j of endmask[c] marks column 32c+j as the last column of a segment.
 * out may be NULL to query the count; at most max_out tiles are written. */
typedef struct {
  int32_t row0, n_valid, clip0, partial;
  uint32_t endmask[8];
} jegal_column_tile;
int jegal_plan_column_tiles(const int32_t* cu_len_host, int32_t n_clips, int32_t width, int32_t allow_split,
                            jegal_column_tile* out, int32_t max_out, int32_t* n_out);

/* K2 — per-query top-k of a dense [n_q, n_g] fp32 score matrix (row stride ld).
 * Descending by score, ties broken towards the lower index; indices are
 * returned + idx_offset (the shard's first global clip).  1 <= k <= 32.
 * Replaces the full np.sort in compute_metrics (evaluate_retrieval.py:52). */
int jegal_topk(jegal_ctx* ctx, const float* scores_dev, int32_t n_q, int32_t n_g, int64_t ld,
               int32_t k, int32_t idx_offset, float* topk_val_dev, int32_t* topk_idx_dev,
               void* stream);

/* Merge n_lists sorted top-k lists per query ([n_lists, n_q, k] val/idx, e.g. the
 * all-gathered per-shard results) into one [n_q, k] list with the same ordering
 * rule as jegal_topk. */
int jegal_topk_merge(jegal_ctx* ctx, const float* vals_dev, const int32_t* idxs_dev,
                     int32_t n_lists, int32_t n_q, int32_t k, float* out_val_dev,
                     int32_t* out_idx_dev, void* stream);

/* Rank of the ground-truth column of every row of a dense [n_q, n_g] matrix:
 * gt_dev[i] (nullable => i, the diagonal) is the positive's column.
 *   n_greater[i] = #{j : x[i,j] >  x[i,gt]}   (the reference's `ind`, evaluate_retrieval.py:52-57)
 *   n_equal[i]   = #{j : x[i,j] == x[i,gt]}   (>= 1; the reference over-counts rows with ties)
 * Replaces np.sort + np.where in compute_metrics (evaluate_retrieval.py:51-65). */
int jegal_rank_of_positive(jegal_ctx* ctx, const float* scores_dev, int32_t n_q, int32_t n_g,
                           int64_t ld_row, int64_t ld_col, const int32_t* gt_dev,
                           int32_t* n_greater_dev, int32_t* n_equal_dev, void* stream);

/* K3 — word spotting (evaluation/evaluate_spotting.py:39-90, utils/plot_heatmap.py:34-59).
 * For clip i: A = softmax_w((G_i C_i^T) / tau) (softmax over words per frame).
 *   normalize_rows    1: gest_rows_dev / cont_rows_dev are the rows AS STORED in the .pkl files (op_dtype =
 *                     JEGAL_F16, or JEGAL_BF16) and the L2 normalisation of evaluate_spotting.py:49-50 is fused
 *                     into the load: the kernel derives 1 / max(||row||, row_eps) of every frame and word from
 *                     the staged operand bytes and scales the accumulator — no normalised copy exists in HBM.
 *                     0: the rows are used as they are (outputs of jegal_prep, or plot_heatmap.py:51-57 which
 *                     does not re-normalise).
 *   word_idx_dev[i]   the target word's row of A^T
 *   heat_dev          (nullable) [rows of gest_layout] fp32: A[:, word_idx] for every frame
 *   full_heat_dev     (nullable) [sum_i T_i * W_i] fp32: the whole (W_i x T_i) matrix of clip i,
 *                     word-major as the reference returns it, at offset full_off_dev[i]
 *   pred_frame_dev[i] argmax_t A[t, word_idx] (first maximum); pred_score_dev[i] its value
 *   correct_dev[i]    (nullable; needs win_lo/win_hi) 1 iff win_lo[i] <= pred <= win_hi[i] and
 *                     pred_score >= thresh  (evaluate_spotting.py:75-82)
 */
int jegal_spot(jegal_ctx* ctx, const jegal_layout* gest_layout, const void* gest_rows_dev,
               const jegal_layout* cont_layout, const void* cont_rows_dev, int op_dtype,
               int normalize_rows, float row_eps, const int32_t* word_idx_dev, float tau, float* heat_dev, float* full_heat_dev,
               const int64_t* full_off_dev, int32_t* pred_frame_dev, float* pred_score_dev,
               const int32_t* win_lo_dev, const int32_t* win_hi_dev, float thresh,
               uint8_t* correct_dev, void* stream);

/* K3 from a DENSE cosine matrix, for clips with more than 64 words (long transcripts): cos_dev is the
 * [rows of gest_layout, rows of cont_layout] frame x word cosine matrix of a GROUP of clips (row stride ld; e.g.
 * jegal_simpool_allpairs over one-row layouts of the packed, normalised frames and words), of which only the
 * block-diagonal (clip i's frames x clip i's words) is read.  Same outputs and arithmetic as jegal_spot
 * (evaluate_spotting.py:52-54,70-82); heat_dev is required. */
int jegal_spot_dense(jegal_ctx* ctx, const float* cos_dev, int64_t ld, const jegal_layout* gest_layout,
                     const jegal_layout* cont_layout, const int32_t* word_idx_dev, float tau, float* heat_dev,
                     float* full_heat_dev, const int64_t* full_off_dev, int32_t* pred_frame_dev, float* pred_score_dev,
                     const int32_t* win_lo_dev, const int32_t* win_hi_dev, float thresh, uint8_t* correct_dev, void* stream);

/* K4 — grouped scoring (evaluation/evaluate_asd.py:43-51,94-100): n_pairs listed
 * (gesture clip, content clip) pairs in groups of `group_size` consecutive pairs.
 *   normalize_rows / row_eps: as in jegal_spot (row normalisation fused into the load)
 *   scores_dev[p]  pooled score of pair p (x gscale x cscale as in K1)
 *   probs_dev      (nullable) softmax(scores / tau) within each group
 *   argmax_dev     (nullable) [n_pairs / group_size] first index of the group maximum
 */
int jegal_simpool_pairs(jegal_ctx* ctx, const jegal_layout* gest_layout, const void* gest_rows_dev,
                        const jegal_layout* cont_layout, const void* cont_rows_dev, int op_dtype,
                        int normalize_rows, float row_eps, int pool_mode, const float* gscale_dev,
                        const float* cscale_dev, const int32_t* pair_gest_dev, const int32_t* pair_cont_dev, int32_t n_pairs,
                        int32_t group_size, float tau, float* scores_dev, float* probs_dev,
                        int32_t* argmax_dev, void* stream);

/* C1 — top-k exchange between the GPUs of one box over NVLink peer memory, fused into K2
 * (no reference counterpart: the reference has no multi-GPU code).
 * Every rank creates an exchange block, publishes its 64-byte CUDA IPC handle, receives the handles
 * of all ranks (rank-major, world * 64 bytes; any transport — torch.distributed in jegal_b200/ops.py)
 * and connects.  jegal_topk_exchange then runs per-query top-k on this rank's score shard
 * [n_q, n_g], stores the k (value, global index) pairs into every peer's block with plain global
 * stores on the IPC-mapped pointers, publishes a sequence flag (system-scope release), waits for the
 * flags of all ranks and merges the world lists locally: out = the global top-k, identical on
 * every rank and identical to a single-GPU jegal_topk over the concatenated shards. */
typedef struct jegal_exchange jegal_exchange;
int jegal_exchange_create(jegal_ctx* ctx, int32_t rank, int32_t world, int32_t n_q, int32_t k,
                          jegal_exchange** out);
int jegal_exchange_ipc_handle(const jegal_exchange* ex, void* handle_out_64B);
int jegal_exchange_connect(jegal_exchange* ex, const void* all_handles);
void jegal_exchange_destroy(jegal_exchange* ex);
int jegal_topk_exchange(jegal_ctx* ctx, jegal_exchange* ex, const float* scores_dev, int32_t n_g, int64_t ld,
                        int32_t idx_offset, float* out_val_dev, int32_t* out_idx_dev, void* stream);

/* C2 — normalise + cast + all-gather of the REPLICATED operand of a sharded run in one kernel, over NVLink peer
 * memory (no reference counterpart).  Every rank creates a gather block for an operand of `rows` x 512 16-bit values,
 * exchanges the 64-byte IPC handles like jegal_exchange and connects.  jegal_prep_gather then takes this rank's slice
 * (rows [row0, row0 + n_rows) of the raw matrix, any rank-disjoint cover of [0, rows)), applies K0's row arithmetic
 * (x / max(||x||, row_eps) in fp32, one rounding to out_dtype) and stores the result rows into EVERY rank's block;
 * the stream continues (a small wait kernel) once all ranks' slices of this call have landed locally.
 * *result_dev: the complete [rows, 512] operand in this rank's memory, valid until the call after next. */
typedef struct jegal_qgather jegal_qgather;
int jegal_qgather_create(jegal_ctx* ctx, int32_t rank, int32_t world, int64_t rows, jegal_qgather** out);
int jegal_qgather_ipc_handle(const jegal_qgather* qg, void* handle_out_64B);
int jegal_qgather_connect(jegal_qgather* qg, const void* all_handles);
void jegal_qgather_destroy(jegal_qgather* qg);
int jegal_prep_gather(jegal_ctx* ctx, jegal_qgather* qg, const void* emb_slice_dev, int in_dtype, int64_t row0,
                      int32_t n_rows, int normalize_rows, float row_eps, int out_dtype, void** result_dev, void* stream);

/* softmax(scores / tau) and first argmax inside groups: group g covers
 * scores[g * stride .. g * stride + group_size).  stride >= group_size lets a caller score
 * prefixes of a wider candidate list (evaluate_asd.py:94-100 scores the first 2, 4, 6).
 * probs_dev (nullable) is [n_groups, group_size] dense; argmax_dev (nullable) [n_groups]. */
int jegal_group_softmax(jegal_ctx* ctx, const float* scores_dev, int32_t n_groups, int32_t group_size,
                        int64_t stride, float tau, float* probs_dev, int32_t* argmax_dev, void* stream);

/* K5 — word-level mean pooling of the content branch (the producer side of the scoring path):
 * out[s, col_off : col_off + dim] = mean of the feature rows x[seg_begin[s] : seg_end[s]] (end exclusive).
 * Replaces the per-word Python loops of JEGAL.get_word_level_embs (models/jegal.py:141-198: sub-word
 * tokens of a word :173-179, audio frames of a word :188-196) and get_audio_word_level_embs (:218-245);
 * ranges may overlap (the reference's frame ranges are end-inclusive) or leave gaps.  A one-row range
 * copies the row.  x is [rows, dim] dense in `in_dtype`; out has row stride ld_out (elements) in
 * `out_dtype`, so the audio and text halves can be written into the concatenated fusion input
 * (models/jegal.py:405-406) directly.  dim, ld_out and col_off must be multiples of 8; seg_* are
 * device int32 arrays the caller validated (0 <= begin < end <= rows). */
int jegal_segment_mean(jegal_ctx* ctx, const void* x_dev, int in_dtype, int64_t rows, int32_t dim,
                       const int32_t* seg_begin_dev, const int32_t* seg_end_dev, int32_t n_seg, void* out_dev,
                       int out_dtype, int64_t ld_out, int32_t col_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JEGAL_B200_H_ */

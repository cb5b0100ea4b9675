// C1 — top-k exchange between the GPUs of one box over NVLink peer memory, fused into K2.
//
// The gallery is sharded by clip; after K1 every rank holds scores [Q, G_local].  Instead of
// "K2, then an NCCL all-gather, then a merge", ONE pair of kernels does the exchange itself:
//
//   topk_exchange_kernel   per-query top-k of the local shard (same code path as K2); the final
//                          k (value, global index) pairs are stored straight into EVERY peer's
//                          exchange block (slot = this rank) with ordinary st.global on
//                          IPC-mapped peer pointers — the stores travel over NVLink/NVSwitch;
//                          the last block to finish publishes a per-rank sequence flag to every
//                          peer (system-scope release);
//   topk_merge_wait_kernel waits (system-scope acquire) until every rank's flag carries this
//                          step's sequence number, then merges the world lists that sit in LOCAL
//                          memory with K2's ordering rule (score desc, ties to lower global index).
//
// Two block parities alternate, which makes the protocol self-synchronising: a rank can only
// finish step s+1 after every peer published s+1, i.e. after every peer finished reading step s.
// The payload is tiny (Q*k*8 B per rank: 80 KB for config 5) — this path is latency-, not
// bandwidth-bound, and saves two host-launched collectives per step.
#include <cstring>
#include <new>

#include "internal.h"
#include "rowio.cuh"
#include "warplist.cuh"

// C2 — the replicated (query-side) operand of a sharded run, normalised and all-gathered in ONE kernel: rank r
// owns rows [r0, r1) of the raw query matrix (e.g. the slice it copied from host memory over its own PCIe link),
// normalises + casts them (K0's row arithmetic) and stores every output row into EVERY rank's operand buffer over
// NVLink peer memory; a per-rank sequence flag (system-scope release / acquire) tells the consumers when all N
// slices have landed.  It replaces "copy all queries on one rank, NCCL broadcast / all-gather, K0 on every rank":
// on the 8-GPU box the NCCL all-gather of the 65 MB query set alone took 1.3 ms of an 8 ms end-to-end step.
struct jegal_qgather {
  jegal_ctx* ctx = nullptr;
  int32_t rank = 0, world = 1;
  int64_t rows = 0;
  uint8_t* local = nullptr;  // header | operand[2 parities][rows][512] 16-bit
  size_t bytes = 0;
  uint8_t* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool opened[8] = {false, false, false, false, false, false, false, false};
  uint32_t seq = 0;
};

struct jegal_exchange {
  jegal_ctx* ctx = nullptr;
  int32_t rank = 0, world = 1, n_q = 0, k = 0;
  uint8_t* local = nullptr;  // cudaMalloc'ed exchange block of this rank
  size_t bytes = 0;
  uint8_t* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool opened[8] = {false, false, false, false, false, false, false, false};
  uint32_t seq = 0;
};

namespace jegal {
namespace {

constexpr int kMaxWorld = 8;
constexpr size_t kHdrBytes = 256;  // [0,64): flags[parity][world] u32   [128]: block counter

struct ExView {
  uint8_t* base[kMaxWorld];
};

__host__ __device__ inline size_t ex_list_elems(int world, int n_q, int k) {
  return static_cast<size_t>(world) * n_q * k;
}
// block layout: header | vals[2][world][n_q][k] f32 | idxs[2][world][n_q][k] i32
__host__ __device__ inline size_t ex_vals_off(int parity, int world, int n_q, int k) {
  return kHdrBytes + parity * ex_list_elems(world, n_q, k) * 4;
}
__host__ __device__ inline size_t ex_idxs_off(int parity, int world, int n_q, int k) {
  return kHdrBytes + 2 * ex_list_elems(world, n_q, k) * 4 + parity * ex_list_elems(world, n_q, k) * 4;
}

using namespace k2;

constexpr int kXWarps = kTopkWarps;  // the row scan is K2's (warplist.cuh)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one block per query row; the result goes to every peer's block (slot = rank)
__global__ void __launch_bounds__(kXWarps * 32)
topk_exchange_kernel(const float* __restrict__ scores, int32_t n_g, int64_t ld, int32_t k, int32_t idx_offset,
                     ExView peers, int32_t rank, int32_t world, int32_t n_q, uint32_t seq) {
  __shared__ float sv[kXWarps][32];
  __shared__ int32_t si[kXWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = blockIdx.x;
  const float* row = scores + q * ld;
  WarpList L;
  L.init();
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
  const int32_t n4 = vec_ok ? (n_g >> 2) : 0;
  const float4* row4 = reinterpret_cast<const float4*>(row);
  scan_row4(L, row4, 0, n4, k, warp, lane);  // the same pipelined scan as K2
  scan_row_tail(L, row, n4 * 4, n_g, k, warp, lane);
  sv[warp][lane] = L.v;
  si[warp][lane] = L.i;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < kXWarps; ++w) L.offer(sv[w][lane], si[w][lane], lane < k, k, lane);
    const int parity = seq & 1u;
    if (lane < k) {
      const bool real = L.i != 0x7fffffff;
      const int32_t gi = real ? L.i + idx_offset : -1;
      const size_t slot = (static_cast<size_t>(rank) * n_q + q) * k + lane;
      for (int p = 0; p < world; ++p) {  // NVLink stores into every peer's exchange block
        float* pv = reinterpret_cast<float*>(peers.base[p] + ex_vals_off(parity, world, n_q, k));
        int32_t* pi = reinterpret_cast<int32_t*>(peers.base[p] + ex_idxs_off(parity, world, n_q, k));
        pv[slot] = L.v;
        pi[slot] = gi;
      }
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
      uint32_t* counter = reinterpret_cast<uint32_t*>(peers.base[rank] + 128);
      const uint32_t done = atomicAdd(counter, 1u);
      if (done == gridDim.x - 1) {  // every block's lists are out: publish this rank's flag everywhere
        *counter = 0;
        __threadfence_system();
        for (int p = 0; p < world; ++p)
          st_release_sys(reinterpret_cast<uint32_t*>(peers.base[p]) + parity * kMaxWorld + rank, seq);
      }
    }
  }
}

// C2, kernel 1: one warp per row of this rank's slice.  normalise (fp32) + cast + store to all `world` operand buffers.
template <int kIn, int kOut>
__global__ void __launch_bounds__(256)
prep_gather_kernel(const void* __restrict__ emb, int32_t n_rows, int64_t row0, int normalize, float eps, ExView peers,
                   size_t buf_off, int32_t world, int32_t rank, uint32_t seq) {
  using namespace rowio;
  const int lane = threadIdx.x & 31;
  const int32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r < n_rows) {
    float a[8], b[8];
    load8<kIn>(emb, r, lane * 8, a);
    load8<kIn>(emb, r, 256 + lane * 8, b);
    if (normalize) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += a[i] * a[i] + b[i] * b[i];
      ss = warp_sum(ss);
      const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] *= inv;
        b[i] *= inv;
      }
    }
    const uint4 va = pack8<kOut>(a), vb = pack8<kOut>(b);
    const size_t off = buf_off + static_cast<size_t>(row0 + r) * (kD * 2) + static_cast<size_t>(lane) * 16;
    for (int p = 0; p < world; ++p) {  // NVLink stores into every rank's operand buffer (own copy included)
      *reinterpret_cast<uint4*>(peers.base[p] + off) = va;
      *reinterpret_cast<uint4*>(peers.base[p] + off + 512) = vb;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* counter = reinterpret_cast<uint32_t*>(peers.base[rank] + 128);
    const uint32_t done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {  // every block's rows are out: publish this rank's flag everywhere
      *counter = 0;
      __threadfence_system();
      const int parity = seq & 1u;
      for (int p = 0; p < world; ++p)
        st_release_sys(reinterpret_cast<uint32_t*>(peers.base[p]) + parity * kMaxWorld + rank, seq);
    }
  }
}

// C2, kernel 2: the stream continues once every rank's slice of this step has landed in LOCAL memory
__global__ void gather_wait_kernel(const uint8_t* __restrict__ local, int32_t world, uint32_t seq) {
  if (static_cast<int32_t>(threadIdx.x) < world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(local) + (seq & 1u) * kMaxWorld + threadIdx.x;
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flag) - seq) < 0) {
      if (++spins > (1u << 26)) {
        printf("jegal: gather watchdog: rank flag %d never reached seq %u\n", (int)threadIdx.x, seq);
        __trap();
      }
    }
  }
}

// one warp per query: wait for all ranks' flags, then merge the lists in local memory
__global__ void __launch_bounds__(128)
topk_merge_wait_kernel(const uint8_t* __restrict__ local, int32_t world, int32_t n_q, int32_t k, uint32_t seq,
                       float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  const int parity = seq & 1u;
  if (threadIdx.x < world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(local) + parity * kMaxWorld + threadIdx.x;
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flag) - seq) < 0) {
      if (++spins > (1u << 26)) {
        printf("jegal: exchange watchdog: rank flag %d never reached seq %u\n", (int)threadIdx.x, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_q) return;
  const float* vals = reinterpret_cast<const float*>(local + ex_vals_off(parity, world, n_q, k));
  const int32_t* idxs = reinterpret_cast<const int32_t*>(local + ex_idxs_off(parity, world, n_q, k));
  WarpList L;
  L.init();
  for (int32_t l = 0; l < world; ++l) {
    const size_t off = (static_cast<size_t>(l) * n_q + q) * k;
    const bool valid = lane < k;
    const float cv = valid ? __ldcg(vals + off + lane) : 0.f;  // written by peers: bypass L1
    const int32_t ci = valid ? __ldcg(idxs + off + lane) : 0;
    L.offer(cv, ci, valid && ci >= 0, k, lane);
  }
  if (lane < k) {
    const bool real = L.i != 0x7fffffff;
    out_val[static_cast<int64_t>(q) * k + lane] = L.v;
    out_idx[static_cast<int64_t>(q) * k + lane] = real ? L.i : -1;
  }
}

}  // namespace
}  // namespace jegal

using namespace jegal;

extern "C" {

int jegal_exchange_create(jegal_ctx* ctx, int32_t rank, int32_t world, int32_t n_q, int32_t k,
                          jegal_exchange** out) {
  if (!ctx || !out) return set_err(ctx, JEGAL_ERR_ARG, "exchange_create: null argument");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n_q < 1 || k < 1 || k > 32)
    return set_err(ctx, JEGAL_ERR_ARG, "exchange_create: need 1 <= world <= 8, 0 <= rank < world, n_q >= 1, 1 <= k <= 32");
  auto* ex = new (std::nothrow) jegal_exchange();
  if (!ex) return set_err(ctx, JEGAL_ERR_NOMEM, "out of host memory");
  ex->ctx = ctx;
  ex->rank = rank;
  ex->world = world;
  ex->n_q = n_q;
  ex->k = k;
  ex->bytes = kHdrBytes + 4 * ex_list_elems(world, n_q, k) * 4;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaMalloc(&ex->local, ex->bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->bytes);
  if (e != cudaSuccess) {
    if (ex->local) cudaFree(ex->local);
    delete ex;
    return set_err(ctx, JEGAL_ERR_CUDA, std::string("exchange_create: ") + cudaGetErrorString(e));
  }
  ex->peer[rank] = ex->local;
  *out = ex;
  return JEGAL_OK;
}

int jegal_exchange_ipc_handle(const jegal_exchange* ex, void* handle_out_64B) {
  if (!ex || !handle_out_64B) return JEGAL_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, ex->local);
  if (e != cudaSuccess) return set_err(ex->ctx, JEGAL_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  std::memcpy(handle_out_64B, &h, 64);
  return JEGAL_OK;
}

int jegal_exchange_connect(jegal_exchange* ex, const void* all_handles) {
  if (!ex || !all_handles) return JEGAL_ERR_ARG;
  cudaSetDevice(ex->ctx->device);
  for (int p = 0; p < ex->world; ++p) {
    if (p == ex->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const uint8_t*>(all_handles) + 64 * p, 64);
    void* ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return set_err(ex->ctx, JEGAL_ERR_CUDA, "cudaIpcOpenMemHandle(rank " + std::to_string(p) + "): " + cudaGetErrorString(e));
    ex->peer[p] = static_cast<uint8_t*>(ptr);
    ex->opened[p] = true;
  }
  return JEGAL_OK;
}

void jegal_exchange_destroy(jegal_exchange* ex) {
  if (!ex) return;
  for (int p = 0; p < kMaxWorld; ++p)
    if (ex->opened[p]) cudaIpcCloseMemHandle(ex->peer[p]);
  if (ex->local) cudaFree(ex->local);
  delete ex;
}

int jegal_qgather_create(jegal_ctx* ctx, int32_t rank, int32_t world, int64_t rows, jegal_qgather** out) {
  if (!ctx || !out) return set_err(ctx, JEGAL_ERR_ARG, "qgather_create: null argument");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || rows < 1)
    return set_err(ctx, JEGAL_ERR_ARG, "qgather_create: need 1 <= world <= 8, 0 <= rank < world, rows >= 1");
  auto* qg = new (std::nothrow) jegal_qgather();
  if (!qg) return set_err(ctx, JEGAL_ERR_NOMEM, "out of host memory");
  qg->ctx = ctx;
  qg->rank = rank;
  qg->world = world;
  qg->rows = rows;
  qg->bytes = kHdrBytes + 2 * static_cast<size_t>(rows) * kD * 2;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaMalloc(&qg->local, qg->bytes);
  if (e == cudaSuccess) e = cudaMemset(qg->local, 0, kHdrBytes);
  if (e != cudaSuccess) {
    if (qg->local) cudaFree(qg->local);
    delete qg;
    return set_err(ctx, JEGAL_ERR_CUDA, std::string("qgather_create: ") + cudaGetErrorString(e));
  }
  qg->peer[rank] = qg->local;
  *out = qg;
  return JEGAL_OK;
}

int jegal_qgather_ipc_handle(const jegal_qgather* qg, void* handle_out_64B) {
  if (!qg || !handle_out_64B) return JEGAL_ERR_ARG;
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, qg->local);
  if (e != cudaSuccess) return set_err(qg->ctx, JEGAL_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  std::memcpy(handle_out_64B, &h, 64);
  return JEGAL_OK;
}

int jegal_qgather_connect(jegal_qgather* qg, const void* all_handles) {
  if (!qg || !all_handles) return JEGAL_ERR_ARG;
  cudaSetDevice(qg->ctx->device);
  for (int p = 0; p < qg->world; ++p) {
    if (p == qg->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const uint8_t*>(all_handles) + 64 * p, 64);
    void* ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return set_err(qg->ctx, JEGAL_ERR_CUDA, "cudaIpcOpenMemHandle(rank " + std::to_string(p) + "): " + cudaGetErrorString(e));
    qg->peer[p] = static_cast<uint8_t*>(ptr);
    qg->opened[p] = true;
  }
  return JEGAL_OK;
}

void jegal_qgather_destroy(jegal_qgather* qg) {
  if (!qg) return;
  for (int p = 0; p < kMaxWorld; ++p)
    if (qg->opened[p]) cudaIpcCloseMemHandle(qg->peer[p]);
  if (qg->local) cudaFree(qg->local);
  delete qg;
}

int jegal_prep_gather(jegal_ctx* ctx, jegal_qgather* qg, const void* emb_slice_dev, int in_dtype, int64_t row0,
                      int32_t n_rows, int normalize_rows, float row_eps, int out_dtype, void** result_dev, void* stream_) {
  JEGAL_NVTX("jegal_prep_gather (C2: K0 + NVLink all-gather)");
  if (!ctx || !qg || !result_dev) return set_err(ctx, JEGAL_ERR_ARG, "prep_gather: null argument");
  if (n_rows < 0 || row0 < 0 || row0 + n_rows > qg->rows) return set_err(ctx, JEGAL_ERR_ARG, "prep_gather: slice outside the operand");
  if (n_rows > 0 && (!emb_slice_dev || (reinterpret_cast<uintptr_t>(emb_slice_dev) & 15u)))
    return set_err(ctx, JEGAL_ERR_ARG, "prep_gather: the slice must be a 16-byte aligned device pointer");
  if (out_dtype != JEGAL_BF16 && out_dtype != JEGAL_F16) return set_err(ctx, JEGAL_ERR_ARG, "prep_gather: out_dtype must be JEGAL_BF16 or JEGAL_F16");
  for (int p = 0; p < qg->world; ++p)
    if (!qg->peer[p]) return set_err(ctx, JEGAL_ERR_ARG, "prep_gather: not connected (jegal_qgather_connect)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  qg->seq += 1;
  const size_t buf_off = kHdrBytes + (qg->seq & 1u) * static_cast<size_t>(qg->rows) * kD * 2;
  ExView view;
  for (int p = 0; p < kMaxWorld; ++p) view.base[p] = qg->peer[p];
  const unsigned grid = static_cast<unsigned>(n_rows > 0 ? (n_rows + 7) / 8 : 1);  // an empty slice still publishes its flag
#define JEGAL_PG(IN, OUT)                                                                                              \
  prep_gather_kernel<IN, OUT><<<grid, 256, 0, stream>>>(emb_slice_dev, n_rows, row0, normalize_rows, row_eps, view, buf_off, \
                                                        qg->world, qg->rank, qg->seq)
  const bool bf = out_dtype == JEGAL_BF16;
  switch (in_dtype) {
    case JEGAL_F32: if (bf) JEGAL_PG(JEGAL_F32, JEGAL_BF16); else JEGAL_PG(JEGAL_F32, JEGAL_F16); break;
    case JEGAL_F16: if (bf) JEGAL_PG(JEGAL_F16, JEGAL_BF16); else JEGAL_PG(JEGAL_F16, JEGAL_F16); break;
    case JEGAL_BF16: if (bf) JEGAL_PG(JEGAL_BF16, JEGAL_BF16); else JEGAL_PG(JEGAL_BF16, JEGAL_F16); break;
    default: return set_err(ctx, JEGAL_ERR_ARG, "prep_gather: bad in_dtype");
  }
#undef JEGAL_PG
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  gather_wait_kernel<<<1, 32, 0, stream>>>(qg->local, qg->world, qg->seq);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  *result_dev = qg->local + buf_off;
  return JEGAL_OK;
}

int jegal_topk_exchange(jegal_ctx* ctx, jegal_exchange* ex, const float* scores_dev, int32_t n_g, int64_t ld,
                        int32_t idx_offset, float* out_val_dev, int32_t* out_idx_dev, void* stream_) {
  JEGAL_NVTX("jegal_topk_exchange (C1)");
  if (!ctx || !ex || !out_val_dev || !out_idx_dev || (!scores_dev && n_g > 0))
    return set_err(ctx, JEGAL_ERR_ARG, "topk_exchange: null argument");
  if (n_g < 0 || ld < n_g) return set_err(ctx, JEGAL_ERR_ARG, "topk_exchange: bad shape");
  for (int p = 0; p < ex->world; ++p)
    if (!ex->peer[p]) return set_err(ctx, JEGAL_ERR_ARG, "topk_exchange: exchange not connected (jegal_exchange_connect)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ex->seq += 1;
  ExView view;
  for (int p = 0; p < kMaxWorld; ++p) view.base[p] = ex->peer[p];
  topk_exchange_kernel<<<static_cast<unsigned>(ex->n_q), kXWarps * 32, 0, stream>>>(
      scores_dev, n_g, ld, ex->k, idx_offset, view, ex->rank, ex->world, ex->n_q, ex->seq);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  const int warps = 4;
  topk_merge_wait_kernel<<<static_cast<unsigned>((ex->n_q + warps - 1) / warps), warps * 32, 0, stream>>>(
      ex->local, ex->world, ex->n_q, ex->k, ex->seq, out_val_dev, out_idx_dev);
  JEGAL_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return JEGAL_OK;
}

}  // extern "C"

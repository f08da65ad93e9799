// hpf_engine.cu -- host side of libhpf_b200.so: the hpf_ctx object and the C ABI
// declared in include/hpf_cuda.h.  Everything numerical runs in the kernels of
// hpf_kernels.cuh on the ctx's own stream; there is no CPU compute path.
//
// HBM layout per parameter side (theta: R = local users, beta: R = items), all
// fp32.  Kp = K rounded up to 4 floats (whole float4s hold data; the pad lanes
// hold 0 / -inf); the row stride ld is Kp rounded up so that a row never
// straddles more 128-byte lines than it must (a multiple of 32 floats once
// Kp >= 32, the next power of two below that) -- the gather of one K=100 row
// then costs 4 L1 wavefronts instead of 7:
//   A      [R x ld]  exp(Elog - rowmax)      sweep input (gathered by the other side)
//   shape  [R x ld]  Gamma shape             written every iteration, with A and the two rate terms
//   Elog   [R x ld]  expected log            materialised on demand (derive_kernel): hpf_get_state, ELBO, exact fallback
//   Ev     [R x ld]  expectation             materialised on demand: held-out ll, top-N, item ranks
//   rate   [R x ld] (hier) or [Kp]           materialised on demand: hpf_get_state / checkpoints
//   T      [R x ld]  sweep output sum (y/Z) * A_other   (+ Tpart for split rows)
//   Tdirect[R x ld]  exact-fallback accumulator (all zero in normal operation)
// plus per-row vectors (shift, xi/eta GPArray, bias GPMatrix, aux) and the
// ratings in both orientations (CSR for the user pass, CSC for the item pass).
#include "../../include/hpf_cuda.h"
#include "hpf_kernels.cuh"
#include "hpf_topn.cuh"
#include "hpf_head.cuh"
#include "hpf_elbo.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cudaTypedefs.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace hpf;

namespace {

thread_local std::string g_create_error = "";

// HPF_TRACE=1: wall-clock milliseconds of the set-up phases on stderr (debug aid)
struct Trace {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  explicit Trace(cudaStream_t s) : on(getenv("HPF_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what)
  {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[hpf trace] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// ---- NCCL through dlopen: single-GPU users never need the library ----------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load(std::string &err)
  {
    if (handle) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) {
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    ReduceScatter = (decltype(ReduceScatter))dlsym(handle, "ncclReduceScatter");
    AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
    GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "libnccl lacks expected symbols"; return false; }
    return true;
  }
};
NcclApi g_nccl;

// grow-only scratch (device memory or pinned host memory), carved linearly per call
struct Arena {
  char *base = nullptr;
  size_t cap = 0, off = 0;
  bool pinned_host = false;
  void reset() { off = 0; }
  template <class T> T *get(size_t count)
  {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > cap) return nullptr;
    T *p = reinterpret_cast<T *>(base + off);
    off += bytes;
    return p;
  }
};

struct DensePlan {      // dense head of the sweep on tcgen05 (hpf_head.cuh)
  bool on = false;
  uint32_t *Yw = nullptr; size_t Yw_cap = 0;          // dense head ratings, as 32-bit words
  uint32_t *head_ids = nullptr; size_t head_ids_cap = 0;
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  size_t a_hi_cap = 0, a_lo_cap = 0, b_hi_cap = 0, b_lo_cap = 0;
  float *dB_part = nullptr; size_t dB_part_cap = 0;   // per-CTA partial sums of the head items' T_beta rows
  uint32_t ntiles = 0, nhead = 0, nblocks = 0; // nhead items in nblocks blocks of 128 (by descending popularity)
  uint64_t head_nnz = 0;
  bool a_dirty = true; // the split copy of A does not match A (after hpf_set_state / a new plan); update_kernel keeps it fresh
  CUtensorMap map_a_hi, map_a_lo, map_b_hi[4], map_b_lo[4];
};
constexpr uint32_t kMaxHeadBlocks = 4;
constexpr uint32_t kElboLaunches = 7; // nnz, theta, beta, xi, eta, theta bias, beta bias

constexpr uint32_t kMaxChunks = 16;
constexpr uint32_t kRowPad = HPF_MAX_DEVICES; // spare rows of the arrays the sharded beta update scatters / gathers in place
struct WorkList {       // segments of one orientation: (chunk of rows, L2 tile, descending length)
  uint4 *seg = nullptr;
  uint32_t *seg_out = nullptr;
  uint32_t nsegs = 0, npartial = 0, nmulti = 0;
  uint32_t *multi_row = nullptr, *multi_first = nullptr, *multi_cnt = nullptr;
  const uint32_t *idx = nullptr; // device, per nonzero
  const uint8_t *y = nullptr;
  bool packed = false;           // idx carries the rating in its top byte (gathered side < 2^24 rows)
  size_t seg_cap = 0, seg_out_cap = 0, multi_cap[3] = { 0, 0, 0 };
  // chunks of rows, launched one after the other (all-reduce of chunk c under the sweep of chunk c + 1)
  uint32_t nchunks = 1, chunk_rows = 0;
  uint32_t chunk_seg[kMaxChunks + 1] = { 0 }, chunk_multi[kMaxChunks + 1] = { 0 };
};

struct Side {
  uint32_t R = 0;
  float *A = nullptr, *Elog = nullptr, *Ev = nullptr, *shape = nullptr, *rate = nullptr;
  float *T = nullptr, *Tpart = nullptr, *Tdirect = nullptr, *shift = nullptr;
  float *pr_shape = nullptr, *pr_rate = nullptr, *pr_Ev = nullptr;           // GPArray (hier)
  float *pr_shape_prev = nullptr, *pr_rate_prev = nullptr; // HPF_LOGL: the GPArray as the last iteration's set_prior_rate saw it
  float *b_shape = nullptr, *b_rate = nullptr, *b_Ev = nullptr, *b_Elog = nullptr; // bias GPMatrix
  float *Tb = nullptr, *Tbpart = nullptr, *Tbdirect = nullptr;
  float2 *aux = nullptr;
  float *colsum = nullptr;         // [Kp] sum over rows of Ev (this side)
  float *colsum_partial = nullptr; // [update_grid x Kp]
  // the two terms of the rate the last update used: rate_uk = rate_row[u] + rate_col[k] (hier), or the GR
  // rate vector in rate_col.  Ev / rate are materialised from them on demand (derived_valid).
  float *rate_row = nullptr, *rate_col = nullptr;
  bool derived_valid = false;
  uint32_t *direct_flag = nullptr;
  uint32_t update_grid = 0;
  size_t tpart_cap = 0, tbpart_cap = 0;
  WorkList wl;
  double prior_shape = 0.3, prior_rate = 0.3, pr_prior_shape = 0.3, pr_prior_rate = 0.3;
  double bias_prior_shape = 0.3, bias_prior_rate = 0.3;
  bool have_state = false, have_pr = false, have_bias = false;
};

} // namespace

struct hpf_ctx {
  hpf_config cfg;
  uint32_t K = 0, Kp = 0, K4 = 0, ld = 0;
  bool hier = false, bias = false, binary = false, jacobi = false;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // hpf_set_ratings_csr: the ratings' upload runs under the first device sort
  cudaEvent_t ev_col = nullptr, ev_y = nullptr;
  bool y_pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t pev[16] = { nullptr };
  bool profiling = false;
  std::string err;
  std::vector<std::pair<void *, size_t>> allocs;
  uint64_t device_bytes = 0;
  Side th, be; // theta (users), beta (items)
  // ratings
  uint64_t nnz = 0;
  uint32_t *csr_idx = nullptr, *csc_idx = nullptr, *upass_idx = nullptr; // upass_*: CSR regrouped by item tile (or null)
  uint8_t *csr_y = nullptr, *csc_y = nullptr, *upass_y = nullptr;
  size_t csr_idx_cap = 0, csr_y_cap = 0, csc_idx_cap = 0, csc_y_cap = 0, upass_idx_cap = 0, upass_y_cap = 0;
  Arena dev_arena, pin_arena;    // grow-only device / pinned-host scratch of hpf_set_ratings_csr
  DensePlan dense;               // the most popular items as a dense block on the tensor cores
  double dense_block_share = -1.0; // HPF_DENSE_BLOCK_SHARE: minimum share of the nonzeros for a 2nd..4th head block (< 0: the cost model in hpf_set_ratings_csr)
  int head_variant = 7;          // HPF_HEAD_VARIANT: epilogue organisation of head_kernel (hpf_head.cuh); 7 measured fastest (profiles/r02b_exp_head_variants.log)
  int dense_head_mode = -1;      // HPF_DENSE_HEAD: -1 auto (on when the head carries >= 15 % of the nonzeros), 0 off, 1 forced
  uint32_t *tail_idx = nullptr; uint8_t *tail_y = nullptr; size_t tail_idx_cap = 0, tail_y_cap = 0; // user-pass tail CSR
  uint32_t th_tiles = 1, be_tiles = 1;
  uint64_t l2_tile_bytes = 32ull << 20; // factor rows of one gather tile (0: no tiling)
  bool l2_tile_forced = false;          // HPF_L2_TILE_KB (tests): ignore the run-length cap
  uint32_t *scratch_u32 = nullptr;
  uint32_t seg_len = 512;
  bool pack_ok = true;           // HPF_PACK=0 (tests / A-B timing): keep index and rating in separate streams
  int sweep_g = 0, sweep_v = 0;
  bool aux_dirty = true, ratings_set = false, th_colsum_global = false;
  // item-side reduce block [T_beta | Tb_beta | colsum_theta | fallback flag]: T_beta is all-reduced chunk by
  // chunk, the tail [Tb_beta | colsum_theta | flag] in one piece; redblock2 = [Tdirect_beta | Tbdirect_beta]
  float *redblock = nullptr, *red_tail = nullptr, *red_flag = nullptr, *redblock2 = nullptr;
  size_t red_count = 0, red_tail_count = 0, red2_count = 0;
  float *mg_fired = nullptr;         // device: sum of the all-reduced fallback flags since it was last cleared
  float *colsum_theta_old = nullptr; // -novb
  unsigned long long *slow_count = nullptr;
  double *logfact = nullptr, *ll_blocks = nullptr, *ll_out = nullptr;
  // HPF_LOGL (-logl): what hpf_elbo needs beyond the hot path's own state
  bool logl = false, pr_prev_valid = false, csr_has_y = false;
  uint64_t *csr_rowptr = nullptr; size_t csr_rowptr_cap = 0; // device copy of the CSR row pointer
  double *elbo_blocks = nullptr;                             // [kElboLaunches x sm_count x 8] per-block partial sums
  // multi-GPU: every collective of this ctx is issued on comm_stream, in the same order on all ranks; events
  // carry the dependencies to and from the compute stream (one_iteration)
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_chunk[kMaxChunks] = { nullptr }, ev_theta = nullptr, ev_comm = nullptr;
  bool ar_defer = false;          // HPF_AR_DEFER (experiments): chunked item pass, but no all-reduce before its last chunk
  int ar_chunks = -1;             // HPF_AR_CHUNKS: chunks of the item pass (unset: one; see hpf_set_ratings_csr)
  // A nonzero that takes the exact fallback adds to the rank-local Tdirect buffers; on the item side those would
  // have to be summed over the ranks too.  That never happens in a real fit, so the reduction is
  // optimistic: the ranks agree on a "fallback fired" flag that rides with the column sums, and hpf_iterate re-runs
  // its window from a snapshot with the fallback buffers inside the all-reduce (mg_exact, sticky) when it is set.
  bool mg_exact = false;
  // Sharded beta update (multi-GPU, many items): T_beta is reduce-SCATTERED, each rank updates its slice of
  // ceil(m / N) items, and only A_beta -- all the sweeps read -- is all-gathered every iteration; the rest of beta's
  // state (shape, xi/eta terms, shift) is gathered when the window of hpf_iterate ends, so every other entry point
  // sees whole arrays.  Inside a window the exact fallback would read stale rows of them: ANY fallback (user or item
  // side) raises the flag and the window re-runs unsharded in mg_exact mode.
  // Measured on 8 x B200 at MSD scale (profiles/r02t_bench_n8_msd_{replicated,sharded}.json): 2.26 ms against 2.33 ms
  // replicated -- the all-gather of A_beta (0.55 ms) cannot hide under anything, while the replicated form's all-reduce
  // hides under the user pass and the theta update; with the faster update kernel that followed the replicated form
  // is ahead, and at 2 GPUs it always was (6.40 vs 6.51 ms).  So: opt-in.
  int shard_mode = 0;             // HPF_SHARD_BETA: 0 off (default), 1 on, -1 on when the payload >= kShardMinBytes
  bool shard_now = false;         // this window runs sharded
  bool last_sharded = false;      // ... the last one did (hpf_stats)
  uint32_t slice_rows = 0;        // ceil(m / N)
  cudaEvent_t ev_beta = nullptr;
  float *snap = nullptr; size_t snap_cap = 0;
  // stats
  uint64_t launches = 0, iterations = 0;
  float last_ms = 0.f, last_topn_ms = 0.f;
  // group ctx (hpf_config.n_devices > 1): owns one child ctx per device and only routes; every field above is unused
  bool is_group = false;
  std::vector<hpf_ctx *> kids;
  std::vector<uint32_t> bounds; // first user of each child's range, then n_users
};

namespace {

int fail(hpf_ctx *c, int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(c, e_ == cudaErrorMemoryAllocation ? HPF_ENOMEM : HPF_ECUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                               \
  } while (0)

template <class T> int dalloc(hpf_ctx *c, T **p, size_t count, bool zero = true)
{
  *p = nullptr;
  if (count == 0) count = 1;
  void *q = nullptr;
  CU(cudaMalloc(&q, count * sizeof(T)));
  c->allocs.emplace_back(q, count * sizeof(T));
  c->device_bytes += count * sizeof(T);
  if (zero) CU(cudaMemsetAsync(q, 0, count * sizeof(T), c->stream));
  *p = (T *)q;
  return 0;
}

int dfree(hpf_ctx *c, void *p)
{
  if (!p) return 0;
  for (auto it = c->allocs.begin(); it != c->allocs.end(); ++it)
    if (it->first == p) {
      c->device_bytes -= it->second;
      c->allocs.erase(it);
      break;
    }
  cudaFree(p);
  return 0;
}

#define TRY(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != 0) return rc_; \
  } while (0)

uint32_t row_grid(const hpf_ctx *c, uint32_t R)
{
  uint32_t need = (R + kUpdateWarps - 1) / kUpdateWarps;
  uint32_t cap = (uint32_t)c->sm_count * 8u;
  return std::max(1u, std::min(need, cap));
}

// grid of update_kernel: one resident wave (the kernel strides over the rows and keeps column sums in registers)
template <int V> int update_occupancy()
{
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, update_kernel<V, false>, kUpdateWarps * 32, 0) != cudaSuccess || occ < 1) occ = 1;
  return occ;
}
uint32_t update_grid_for(const hpf_ctx *c, uint32_t R)
{
  int occ = 1;
  switch ((c->K4 + 31) / 32) {
  case 1: occ = update_occupancy<1>(); break;
  case 2: occ = update_occupancy<2>(); break;
  case 3: occ = update_occupancy<3>(); break;
  case 4: occ = update_occupancy<4>(); break;
  case 5: occ = update_occupancy<5>(); break;
  case 6: occ = update_occupancy<6>(); break;
  case 7: occ = update_occupancy<7>(); break;
  default: occ = update_occupancy<8>(); break;
  }
  const uint32_t need = (R + kUpdateWarps - 1) / kUpdateWarps;
  return std::max(1u, std::min(need, (uint32_t)c->sm_count * (uint32_t)occ));
}

// pick lanes-per-nonzero G and float4-per-lane V: smallest G with V <= 4 (measured best at K=100:
// G=8/V=4 beats G=4/V=7 -- fewer registers, twice the resident warps)
void pick_sweep_shape(hpf_ctx *c)
{
  int g = 1;
  while (g < 32 && (int)((c->K4 + g - 1) / g) > 4) g *= 2;
  if (const char *e = getenv("HPF_SWEEP_G")) {
    int eg = atoi(e);
    if (eg == 1 || eg == 2 || eg == 4 || eg == 8 || eg == 16 || eg == 32)
      if ((c->K4 + eg - 1) / eg <= 8) g = eg;
  }
  c->sweep_g = g;
  c->sweep_v = (int)((c->K4 + g - 1) / g);
}

// own_direct: the side allocates its fallback buffers itself (the item side's live in the second reduce block)
int alloc_side(hpf_ctx *c, Side &s, uint32_t R, bool own_direct)
{
  s.R = R;
  const size_t rk = (size_t)R * c->ld;
  // what the sharded beta update gathers in place carries kRowPad spare rows: N equal slices of ceil(R / N) rows
  const size_t Rg = (size_t)R + kRowPad, rkg = Rg * c->ld;
  TRY(dalloc(c, &s.A, rkg));
  TRY(dalloc(c, &s.Elog, rk));
  TRY(dalloc(c, &s.Ev, rk));
  TRY(dalloc(c, &s.shape, rkg));
  TRY(dalloc(c, &s.rate, c->hier ? rk : (size_t)c->Kp));
  TRY(dalloc(c, &s.rate_col, c->Kp));
  if (c->hier) TRY(dalloc(c, &s.rate_row, Rg));
  if (own_direct) TRY(dalloc(c, &s.Tdirect, rk));
  TRY(dalloc(c, &s.shift, Rg));
  TRY(dalloc(c, &s.direct_flag, 1));
  if (c->hier) {
    TRY(dalloc(c, &s.pr_shape, Rg));
    TRY(dalloc(c, &s.pr_rate, Rg));
    TRY(dalloc(c, &s.pr_Ev, Rg));
    if (c->logl) {
      TRY(dalloc(c, &s.pr_shape_prev, R));
      TRY(dalloc(c, &s.pr_rate_prev, R));
    }
  }
  if (c->bias) {
    TRY(dalloc(c, &s.b_shape, R));
    TRY(dalloc(c, &s.b_rate, R));
    TRY(dalloc(c, &s.b_Ev, R));
    TRY(dalloc(c, &s.b_Elog, R));
    if (own_direct) TRY(dalloc(c, &s.Tbdirect, R));
    TRY(dalloc(c, &s.aux, R));
  }
  s.update_grid = update_grid_for(c, R);
  TRY(dalloc(c, &s.colsum_partial, (size_t)std::max(s.update_grid, row_grid(c, R)) * c->Kp));
  return 0;
}

// ---- sweep dispatch ----------------------------------------------------------
template <int G, int V> int launch_sweep_gv(hpf_ctx *c, const SweepArgs &a)
{
  constexpr uint32_t groups_per_block = kSweepThreads / G;
  const uint32_t grid = (a.nsegs + groups_per_block - 1) / groups_per_block;
  if (grid == 0) return 0;
  if (c->bias) sweep_kernel<G, V, true><<<grid, kSweepThreads, 0, c->stream>>>(a);
  else sweep_kernel<G, V, false><<<grid, kSweepThreads, 0, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int G> int launch_sweep_g(hpf_ctx *c, const SweepArgs &a)
{
  switch (c->sweep_v) {
  case 1: return launch_sweep_gv<G, 1>(c, a);
  case 2: return launch_sweep_gv<G, 2>(c, a);
  case 3: return launch_sweep_gv<G, 3>(c, a);
  case 4: return launch_sweep_gv<G, 4>(c, a);
  case 5: return launch_sweep_gv<G, 5>(c, a);
  case 6: return launch_sweep_gv<G, 6>(c, a);
  case 7: return launch_sweep_gv<G, 7>(c, a);
  case 8: return launch_sweep_gv<G, 8>(c, a);
  }
  return fail(c, HPF_EINVAL, "unsupported sweep shape G=%d V=%d", G, c->sweep_v);
}

ElogSrc elog_src(const hpf_ctx *c, const Side &s)
{
  ElogSrc e;
  e.Elog = s.Elog; e.shape = s.shape; e.rate_row = s.rate_row; e.rate_col = s.rate_col;
  e.valid = s.derived_valid ? 1 : 0; e.hier = c->hier ? 1 : 0;
  return e;
}

// sweep over the segments of chunk `chunk` of the row side's work list (chunk < 0: all of them)
int launch_sweep(hpf_ctx *c, Side &rowside, Side &colside, int chunk = -1)
{
  SweepArgs a;
  memset(&a, 0, sizeof a);
  const WorkList &w = rowside.wl;
  const uint32_t s0 = chunk < 0 ? 0u : w.chunk_seg[chunk], s1 = chunk < 0 ? w.nsegs : w.chunk_seg[chunk + 1];
  a.seg = w.seg + s0; a.seg_out = w.seg_out + s0; a.nsegs = s1 - s0; a.R = rowside.R;
  a.idx = w.idx; a.y = w.y; a.packed = w.packed ? 1u : 0u;
  a.Arow = rowside.A; a.Acol = colside.A;
  a.T = rowside.T; a.Tpart = rowside.Tpart;
  a.row_aux = rowside.aux; a.col_aux = colside.aux;
  a.Tb = rowside.Tb; a.Tbpart = rowside.Tbpart;
  a.ElogRow = elog_src(c, rowside); a.ElogCol = elog_src(c, colside);
  a.ElogbRow = rowside.b_Elog; a.ElogbCol = colside.b_Elog;
  a.Tdirect = rowside.Tdirect; a.Tbdirect = rowside.Tbdirect;
  a.direct_flag = rowside.direct_flag; a.slow_count = c->slow_count;
  a.K = c->K; a.K4 = c->K4; a.ld = c->ld; a.ld4 = c->ld / 4;
  switch (c->sweep_g) {
  case 1: return launch_sweep_g<1>(c, a);
  case 2: return launch_sweep_g<2>(c, a);
  case 4: return launch_sweep_g<4>(c, a);
  case 8: return launch_sweep_g<8>(c, a);
  case 16: return launch_sweep_g<16>(c, a);
  case 32: return launch_sweep_g<32>(c, a);
  }
  return fail(c, HPF_EINVAL, "unsupported sweep group %d", c->sweep_g);
}

int launch_combine(hpf_ctx *c, Side &s, int chunk = -1)
{
  const uint32_t m0 = chunk < 0 ? 0u : s.wl.chunk_multi[chunk], m1 = chunk < 0 ? s.wl.nmulti : s.wl.chunk_multi[chunk + 1];
  if (m1 == m0) return 0;
  CombineArgs a;
  a.multi_row = s.wl.multi_row + m0; a.multi_first = s.wl.multi_first + m0; a.multi_cnt = s.wl.multi_cnt + m0;
  a.nmulti = m1 - m0; a.Kp = c->Kp; a.ld = c->ld; a.Tpart = s.Tpart; a.T = s.T;
  a.Tbpart = c->bias ? s.Tbpart : nullptr; a.Tb = s.Tb;
  const size_t sm = ((size_t)kUpdateWarps * c->Kp + kUpdateWarps) * sizeof(float);
  combine_kernel<<<m1 - m0, kUpdateWarps * 32, sm, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int V> int launch_update_v(hpf_ctx *c, const UpdateArgs &a, uint32_t grid)
{
  const size_t sm = (size_t)kUpdateWarps * c->Kp * sizeof(float);
  if (a.K == a.Kp) update_kernel<V, true><<<grid, kUpdateWarps * 32, sm, c->stream>>>(a);
  else update_kernel<V, false><<<grid, kUpdateWarps * 32, sm, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

// dense update of one side.  colsum_other: the other side's column sums of E[v]; bias_count: what the bias rate adds
// (m for users, n_global for items).  The theta update also publishes the item side's fallback flag next to its
// column sums (tail of the reduce block); the beta update accumulates the all-reduced flag (multi-GPU).
int launch_update(hpf_ctx *c, Side &s, const float *colsum_other, double bias_count, uint32_t row0 = 0, uint32_t rows = UINT32_MAX)
{
  const bool theta = &s == &c->th;
  if (rows == UINT32_MAX) rows = s.R;
  const bool slice = rows != s.R; // sharded beta update: rows [row0, row0 + rows) only (never with -bias, never theta)
  const uint32_t grid = slice ? update_grid_for(c, std::max(rows, 1u)) : s.update_grid;
  const size_t o = (size_t)row0 * c->ld;
  UpdateArgs a;
  memset(&a, 0, sizeof a);
  a.R = rows; a.K = c->K; a.Kp = c->Kp; a.K4 = c->K4; a.ld4 = c->ld / 4;
  a.T = reinterpret_cast<const float4 *>(s.T + o); a.Tdirect = reinterpret_cast<float4 *>(s.Tdirect + o); a.direct_flag = s.direct_flag;
  a.direct_flag_all = (!theta && c->mg_exact && c->nranks > 1) ? c->red_flag : nullptr;
  a.A = reinterpret_cast<float4 *>(s.A + o); a.shape = reinterpret_cast<float4 *>(s.shape + o);
  a.shift = s.shift + row0;
  a.hier = c->hier; a.colsum_other = colsum_other;
  a.rate_vec = s.rate_col; a.rate_row = s.rate_row ? s.rate_row + row0 : nullptr;
  a.prior_shape = (float)s.prior_shape; a.prior_rate = (float)s.prior_rate;
  a.pr_shape = s.pr_shape ? s.pr_shape + row0 : nullptr; a.pr_rate = s.pr_rate ? s.pr_rate + row0 : nullptr;
  a.pr_Ev = s.pr_Ev ? s.pr_Ev + row0 : nullptr;
  a.pr_prior_shape = (float)s.pr_prior_shape; a.pr_prior_rate = (float)s.pr_prior_rate;
  a.bias = c->bias; a.Tb = s.Tb; a.Tbdirect = s.Tbdirect;
  a.b_shape = s.b_shape; a.b_rate = s.b_rate; a.b_Ev = s.b_Ev; a.b_Elog = s.b_Elog; a.aux = s.aux;
  a.bias_prior_shape = (float)s.bias_prior_shape;
  a.bias_rate_total = (float)(s.bias_prior_rate + bias_count);
  a.colsum_partial = s.colsum_partial;
  if (theta && c->dense.on && !c->dense.a_dirty) { // keep the dense head's operand copy of A in step
    a.split_hi = c->dense.a_hi; a.split_lo = c->dense.a_lo; a.split_ld = head::kFact;
  }
  s.derived_valid = false;
  if (rows > 0) switch ((c->K4 + 31) / 32) {
  case 1: TRY(launch_update_v<1>(c, a, grid)); break;
  case 2: TRY(launch_update_v<2>(c, a, grid)); break;
  case 3: TRY(launch_update_v<3>(c, a, grid)); break;
  case 4: TRY(launch_update_v<4>(c, a, grid)); break;
  case 5: TRY(launch_update_v<5>(c, a, grid)); break;
  case 6: TRY(launch_update_v<6>(c, a, grid)); break;
  case 7: TRY(launch_update_v<7>(c, a, grid)); break;
  case 8: TRY(launch_update_v<8>(c, a, grid)); break;
  default: return fail(c, HPF_EINVAL, "unsupported factor count %u", c->K);
  }
  const bool mg = c->nranks > 1;
  colsum_finalize_kernel<<<(c->Kp + 31) / 32, 32 * kFinalizeGroups, 0, c->stream>>>(
      s.colsum_partial, rows > 0 ? grid : 0, c->Kp, s.colsum, s.direct_flag,
      theta && mg ? c->be.direct_flag : nullptr, theta && mg && c->shard_now ? c->th.direct_flag : nullptr,
      theta && mg ? c->red_flag : nullptr, !theta && mg ? c->red_flag : nullptr, !theta && mg ? c->mg_fired : nullptr);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

// E[v], E[log v] and the rate matrix of a side, materialised from shape and the rate terms of the last update
// (report-window consumers: held-out ll, top-N, ELBO, hpf_get_state; afterwards the exact fallback reads the array too)
int ensure_derived(hpf_ctx *c, Side &s)
{
  if (s.derived_valid) return 0;
  DeriveArgs a;
  a.R = s.R; a.K = c->K; a.K4 = c->K4; a.ld4 = c->ld / 4;
  a.shape = reinterpret_cast<const float4 *>(s.shape); a.Ev = reinterpret_cast<float4 *>(s.Ev);
  a.Elog = reinterpret_cast<float4 *>(s.Elog);
  a.rate = c->hier ? reinterpret_cast<float4 *>(s.rate) : nullptr;
  a.hier = c->hier; a.rate_row = s.rate_row; a.rate_col = s.rate_col;
  const uint64_t total = (uint64_t)s.R * c->K4;
  const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 16));
  derive_kernel<<<grid, 256, 0, c->stream>>>(a);
  c->launches++;
  CU(cudaGetLastError());
  if (!c->hier) CU(cudaMemcpyAsync(s.rate, s.rate_col, sizeof(float) * c->Kp, cudaMemcpyDeviceToDevice, c->stream));
  s.derived_valid = true;
  return 0;
}

int refresh_colsum(hpf_ctx *c, Side &s)
{
  const size_t sm = (size_t)kUpdateWarps * c->Kp * sizeof(float);
  colsum_partial_kernel<<<s.update_grid, kUpdateWarps * 32, sm, c->stream>>>(s.Ev, s.R, c->Kp, c->ld, s.colsum_partial);
  c->launches++;
  CU(cudaGetLastError());
  colsum_finalize_kernel<<<(c->Kp + 31) / 32, 32 * kFinalizeGroups, 0, c->stream>>>(s.colsum_partial, s.update_grid, c->Kp, s.colsum,
                                                                     nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

// ---- set-up memory: grow-only arenas, so a repeated hpf_set_ratings_csr allocates nothing ----
// (cudaMalloc / cudaFree synchronise the device and cost milliseconds at GB sizes)
int arena_reserve(hpf_ctx *c, Arena &a, size_t bytes)
{
  a.reset();
  if (a.cap >= bytes) return 0;
  bytes += bytes / 8 + (1u << 20);
  if (a.base) {
    CU(cudaStreamSynchronize(c->stream));
    if (a.pinned_host) cudaFreeHost(a.base); else { cudaFree(a.base); c->device_bytes -= a.cap; }
    a.base = nullptr; a.cap = 0;
  }
  void *q = nullptr;
  if (a.pinned_host) CU(cudaMallocHost(&q, bytes));
  else { CU(cudaMalloc(&q, bytes)); c->device_bytes += bytes; }
  a.base = (char *)q; a.cap = bytes;
  return 0;
}

template <class T> int ensure(hpf_ctx *c, T **p, size_t *cap, size_t count)
{
  if (*p != nullptr && *cap >= count) return 0;
  dfree(c, *p);
  *p = nullptr; *cap = 0;
  const size_t want = count + count / 16 + 64;
  TRY(dalloc(c, p, want, false));
  *cap = want;
  return 0;
}

size_t pad256(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

int bits_for(uint64_t nvalues)
{
  int b = 1;
  while (b < 32 && (1ull << b) < nvalues) ++b;
  return b;
}

// How many tiles the gathered side (C rows of ld floats) is cut into.  Tiling keeps the gathers of one
// pass inside an L2-sized window, but every (row, tile) run becomes a segment with its own partial
// row: it only pays while runs stay long (measured: Netflix-scale item pass 6.2 -> 3.8 ms with 8 tiles
// of ~5.7K-nonzero rows; MSD scale, ~4 nonzeros per run, 13.7 -> 30.7 ms).  So the tile count is also
// capped by an average of >= 64 nonzeros per run.
uint32_t tiles_for(const hpf_ctx *c, uint32_t C, uint32_t R, uint64_t nnz)
{
  if (c->l2_tile_bytes == 0) return 1;
  const uint64_t bytes = (uint64_t)C * c->ld * sizeof(float);
  if (bytes <= 2 * c->l2_tile_bytes) return 1;
  uint64_t t = (bytes + c->l2_tile_bytes - 1) / c->l2_tile_bytes;
  if (!c->l2_tile_forced) t = std::min<uint64_t>(t, std::max<uint64_t>(1, nnz / ((uint64_t)std::max(R, 1u) * 64)));
  while (t > 1 && t * (uint64_t)R >= 0xfffffff0ull) --t; // the composite key is 32 bits
  return (uint32_t)std::min<uint64_t>(t, C);
}

// scratch device memory of one call (hpf_topn); freed on scope exit
struct Scratch {
  std::vector<void *> v;
  ~Scratch() { for (void *p : v) cudaFree(p); }
  template <class T> cudaError_t get(T **p, size_t count)
  {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count * sizeof(T), 16));
    if (e == cudaSuccess) v.push_back(q);
    *p = (T *)q;
    return e;
  }
};

// One orientation of the ratings on the device: the nonzeros ordered by (tile of col, row), the
// gathered-side index and the rating permuted accordingly, and the run pointers (run (t, r) at
// d_run[t * R + r]).  d_row / d_col / d_y are per nonzero in the caller's order.  presorted: the input is
// already ordered by row and d_rowptr is its row pointer -- with one tile that IS the orientation.
struct Orientation {
  uint32_t ntiles = 1, tile_cols = 0;
  const uint64_t *d_run = nullptr; // device, ntiles * R + 1
  const uint32_t *d_idx = nullptr;
  const uint8_t *d_y = nullptr;
  bool packed = false;             // d_idx = index | rating << 24, d_y unused
};

// ---- work lists, built on the device (kernels and the ordering: hpf_kernels.cuh, "work lists") -------------
// Everything is enqueued on the ctx stream; the counts (segments, partial slots, multi-segment rows, chunk
// boundaries) land in pinned memory and are read by finish_worklist after the caller's next synchronisation.
struct WlPending {
  uint32_t *h_info = nullptr; // pinned {nsegs, nslots, nmulti, 0, seg bound[nchunks + 1], multi bound[nchunks + 1]}
  uint32_t nchunks = 1, chunk_rows = 0;
};

uint64_t worklist_seg_bound(uint64_t nnz, uint32_t R, uint32_t ntiles, uint32_t L)
{
  return nnz / L + std::min<uint64_t>(nnz, (uint64_t)ntiles * R) + R + 1;
}

size_t worklist_dev_bytes(uint64_t nnz, uint32_t R, uint32_t ntiles, uint32_t L, size_t cub_bytes)
{
  const uint64_t b = worklist_seg_bound(nnz, R, ntiles, L);
  return 6 * pad256((size_t)R * 4) + pad256(b * 16) + 5 * pad256(b * 4) + 2 * pad256(cub_bytes) + pad256(4 * (6 + 2 * kMaxChunks)) + 8192;
}

int build_worklist_device(hpf_ctx *c, Arena &dev, Arena &pin, Side &s, const uint64_t *d_run, uint32_t ntiles, uint64_t nnz,
                          const uint32_t *d_skip_slot, uint32_t nchunks, const Orientation &o, size_t cub_bytes, WlPending *pend)
{
  const uint32_t R = s.R, L = c->seg_len;
  WorkList &w = s.wl;
  w.idx = o.d_idx; w.y = o.d_y; w.packed = o.packed;
  nchunks = std::max(1u, std::min(std::min(nchunks, kMaxChunks), R));
  const uint32_t chunk_rows = (R + nchunks - 1) / nchunks;
  nchunks = (R + chunk_rows - 1) / chunk_rows;
  if ((uint64_t)nchunks * ntiles * (L + 1) >= 0xffffffffull) return fail(c, HPF_EINVAL, "work-list keys do not fit 32 bits (%u chunks x %u tiles)", nchunks, ntiles);
  const uint64_t bound64 = worklist_seg_bound(nnz, R, ntiles, L);
  if (bound64 >= 0xfffffff0ull) return fail(c, HPF_EINVAL, "too many work segments (%llu)", (unsigned long long)bound64);
  const uint32_t bound = (uint32_t)bound64;
  TRY(ensure(c, &w.seg, &w.seg_cap, bound));
  TRY(ensure(c, &w.seg_out, &w.seg_out_cap, bound));
  const size_t multi_bound = std::min<size_t>(R, bound / 2 + 1);
  TRY(ensure(c, &w.multi_row, &w.multi_cap[0], multi_bound));
  TRY(ensure(c, &w.multi_first, &w.multi_cap[1], multi_bound));
  TRY(ensure(c, &w.multi_cnt, &w.multi_cap[2], multi_bound));
  WlArgs a;
  a.run_ptr = d_run; a.R = R; a.ntiles = ntiles; a.L = L; a.chunk_rows = chunk_rows; a.skip_slot = d_skip_slot;
  a.segcnt = dev.get<uint32_t>(R); a.multicnt = dev.get<uint32_t>(R); a.ismulti = dev.get<uint32_t>(R);
  a.seg_off = dev.get<uint32_t>(R); a.first_off = dev.get<uint32_t>(R); a.multi_off = dev.get<uint32_t>(R);
  a.seg_u = dev.get<uint4>(bound); a.out_u = dev.get<uint32_t>(bound); a.key_u = dev.get<uint32_t>(bound);
  a.multi_row = w.multi_row; a.multi_first = w.multi_first; a.multi_cnt = w.multi_cnt;
  uint32_t *key_s = dev.get<uint32_t>(bound), *perm_in = dev.get<uint32_t>(bound), *perm_out = dev.get<uint32_t>(bound);
  void *d_tmp = dev.get<char>(cub_bytes);
  const uint32_t ninfo = 4 + 2 * (nchunks + 1);
  uint32_t *d_info = dev.get<uint32_t>(ninfo);
  pend->h_info = pin.get<uint32_t>(ninfo);
  pend->nchunks = nchunks; pend->chunk_rows = chunk_rows;
  if (!a.segcnt || !a.multicnt || !a.ismulti || !a.seg_off || !a.first_off || !a.multi_off || !a.seg_u || !a.out_u || !a.key_u ||
      !key_s || !perm_in || !perm_out || !d_tmp || !d_info || !pend->h_info)
    return fail(c, HPF_ENOMEM, "set-up arena too small (work list)");
  const unsigned rb = (R + 255) / 256, sb = (bound + 255) / 256;
  CU(cudaMemsetAsync(a.key_u, 0xff, (size_t)bound * 4, c->stream)); // unused slots sort behind every segment
  wl_count_kernel<<<rb, 256, 0, c->stream>>>(a);
  size_t tb = cub_bytes;
  CU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, (const uint32_t *)a.segcnt, a.seg_off, (int64_t)R, c->stream));
  tb = cub_bytes;
  CU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, (const uint32_t *)a.multicnt, a.first_off, (int64_t)R, c->stream));
  tb = cub_bytes;
  CU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, (const uint32_t *)a.ismulti, a.multi_off, (int64_t)R, c->stream));
  wl_emit_kernel<<<rb, 256, 0, c->stream>>>(a);
  iota_kernel<<<sb, 256, 0, c->stream>>>(perm_in, bound);
  tb = cub_bytes;
  CU(cub::DeviceRadixSort::SortPairs(d_tmp, tb, (const uint32_t *)a.key_u, key_s, (const uint32_t *)perm_in, perm_out, (int64_t)bound, 0, 32, c->stream));
  wl_info_kernel<<<1, 32, 0, c->stream>>>(a, key_s, nchunks, d_info);
  wl_gather_kernel<<<sb, 256, 0, c->stream>>>(perm_out, a.seg_u, a.out_u, d_info, w.seg, w.seg_out);
  c->launches += 5;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(pend->h_info, d_info, (size_t)ninfo * 4, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

// after the stream has been synchronised: counts into the side's work list, partial-sum slots sized
int finish_worklist(hpf_ctx *c, Side &s, const WlPending &p)
{
  WorkList &w = s.wl;
  w.nsegs = p.h_info[0]; w.npartial = p.h_info[1]; w.nmulti = p.h_info[2];
  w.nchunks = p.nchunks; w.chunk_rows = p.chunk_rows;
  for (uint32_t q = 0; q <= p.nchunks; ++q) {
    w.chunk_seg[q] = p.h_info[4 + q];
    w.chunk_multi[q] = p.h_info[4 + p.nchunks + 1 + q];
  }
  TRY(ensure(c, &s.Tpart, &s.tpart_cap, (size_t)w.npartial * c->ld));
  if (c->bias) TRY(ensure(c, &s.Tbpart, &s.tbpart_cap, w.npartial));
  return 0;
}

size_t orientation_dev_bytes(uint64_t nnz, uint32_t R, uint32_t ntiles, size_t cub_bytes)
{
  return 4 * pad256(nnz * 4) + pad256(((size_t)ntiles * R + 1) * 8) + pad256(cub_bytes) + 4096;
}

// the ratings travel host -> device on copy_stream, after the indices; the first kernel that reads them waits here
int wait_for_ratings(hpf_ctx *c)
{
  if (c->y_pending) {
    CU(cudaStreamWaitEvent(c->stream, c->ev_y, 0));
    c->y_pending = false;
  }
  return 0;
}

int orient_device(hpf_ctx *c, Arena &dev, uint64_t nnz, const uint32_t *d_row, const uint32_t *d_col, const uint8_t *d_y,
                  bool presorted, const uint64_t *d_rowptr, uint32_t R, uint32_t C, uint32_t **own_idx, size_t *own_idx_cap,
                  uint8_t **own_y, size_t *own_y_cap, size_t cub_bytes, Orientation *o)
{
  o->ntiles = tiles_for(c, C, R, nnz);
  o->tile_cols = (uint32_t)(((uint64_t)C + o->ntiles - 1) / o->ntiles);
  o->packed = C < (1u << 24) && c->pack_ok; // the sweep then reads ONE word per nonzero and broadcasts one value
  if (presorted && o->ntiles == 1) {
    o->d_run = d_rowptr; o->d_idx = d_col; o->d_y = d_y;
    if (o->packed && nnz > 0) {
      TRY(ensure(c, own_idx, own_idx_cap, nnz));
      TRY(wait_for_ratings(c));
      pack_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c->stream>>>(d_col, d_y, nnz, *own_idx);
      c->launches++;
      o->d_idx = *own_idx; o->d_y = nullptr;
    } else o->packed = false;
    return 0;
  }
  const size_t nruns = (size_t)o->ntiles * R;
  TRY(ensure(c, own_idx, own_idx_cap, nnz));
  if (d_y && !o->packed) TRY(ensure(c, own_y, own_y_cap, nnz));
  o->d_idx = *own_idx; o->d_y = (d_y && !o->packed) ? *own_y : nullptr;
  uint64_t *d_run = dev.get<uint64_t>(nruns + 1);
  if (!d_run) return fail(c, HPF_ENOMEM, "device set-up arena too small");
  o->d_run = d_run;
  if (nnz == 0) {
    CU(cudaMemsetAsync(d_run, 0, (nruns + 1) * 8, c->stream));
    return 0;
  }
  uint32_t *key = dev.get<uint32_t>(nnz), *key2 = dev.get<uint32_t>(nnz), *pos = dev.get<uint32_t>(nnz), *pos2 = dev.get<uint32_t>(nnz);
  void *d_tmp = dev.get<char>(cub_bytes);
  if (!key || !key2 || !pos || !pos2 || !d_tmp) return fail(c, HPF_ENOMEM, "device set-up arena too small");
  size_t tmp_bytes = cub_bytes;
  const unsigned nb = (unsigned)((nnz + 255) / 256);
  // one stable sort of the positions by (tile of the gathered-side index, row); the index and the rating follow through
  // the sorted positions -- the ratings may still be on their way from the host until then
  orient_key_kernel<<<nb, 256, 0, c->stream>>>(d_row, d_col, o->tile_cols, R, nnz, key, pos);
  CU(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, (const uint32_t *)key, key2, (const uint32_t *)pos, pos2, (int64_t)nnz, 0,
                                     bits_for((uint64_t)nruns), c->stream));
  run_ptr_kernel<<<(unsigned)((nnz + 256) / 256), 256, 0, c->stream>>>(key2, nnz, (uint32_t)nruns, d_run);
  TRY(wait_for_ratings(c));
  orient_gather_kernel<<<nb, 256, 0, c->stream>>>(pos2, d_col, d_y, nnz, *own_idx, o->d_y ? *own_y : nullptr, o->packed ? 1 : 0);
  c->launches += 3;
  CU(cudaGetLastError());
  return 0;
}

PFN_cuTensorMapEncodeTiled_v12000 tensormap_encoder()
{
  static PFN_cuTensorMapEncodeTiled_v12000 fnp = nullptr;
  if (!fnp) {
    cudaDriverEntryPointQueryResult qres;
    void *fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn)
      fnp = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  }
  return fnp;
}

// K-major bf16 matrix [rows x cols] -> tensor map with 128B-swizzled boxes of [64 cols x box_rows]
bool make_bf16_map(CUtensorMap *map, void *ptr, uint64_t rows, uint32_t cols, uint32_t box_rows)
{
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensormap_encoder();
  if (!encode) return false;
  cuuint64_t dims[2] = { cols, rows };
  cuuint64_t strides[1] = { (cuuint64_t)cols * 2 };
  cuuint32_t box[2] = { 64, box_rows };
  cuuint32_t estr[2] = { 1, 1 };
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// dense head of one iteration: operand splits, head rows of T_beta cleared, then the tcgen05 kernel
int launch_dense_head(hpf_ctx *c)
{
  DensePlan &d = c->dense;
  if (!d.on) return 0;
  const uint32_t n = c->th.R, n_pad = d.ntiles * head::kUsers;
  if (d.a_dirty) {
    if (c->bias)
      head::split_aux_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->th.A, c->ld, c->Kp, c->th.aux, 0, nullptr, n, n_pad, d.a_hi, d.a_lo);
    else
      topk::split_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->th.A, c->ld, c->K, nullptr, 0, nullptr, n, n_pad, head::kFact, d.a_hi, d.a_lo);
    c->launches++;
    d.a_dirty = false;
  }
  typedef void (*head_fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const head::HeadArgs);
  static const head_fn variants[8] = { head::head_kernel<0>, head::head_kernel<1>, head::head_kernel<2>, head::head_kernel<3>,
                                       head::head_kernel<4>, head::head_kernel<5>, head::head_kernel<6>, head::head_kernel<7> };
  const head_fn kernel = variants[c->head_variant & 7];
  const int head_threads = head::head_threads(c->head_variant & 7);
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head::kSmemBytes));
  const uint32_t grid = std::min<uint32_t>(d.ntiles, (uint32_t)c->sm_count);
  const size_t part_stride = (size_t)grid * head::kHead * head::kFact;
  // the operand rows of ALL head blocks in one launch (b_hi / b_lo and head_ids are contiguous over the blocks)
  if (c->bias)
    head::split_aux_kernel<<<64 * d.nblocks, 256, 0, c->stream>>>(c->be.A, c->ld, c->Kp, c->be.aux, 1, d.head_ids, d.nhead, d.nblocks * head::kHead, d.b_hi, d.b_lo);
  else
    topk::split_kernel<<<64 * d.nblocks, 256, 0, c->stream>>>(c->be.A, c->ld, c->K, nullptr, 0, d.head_ids, d.nhead, d.nblocks * head::kHead, head::kFact, d.b_hi, d.b_lo);
  c->launches++;
  for (uint32_t b = 0; b < d.nblocks; ++b) { // one pass over the users per block of 128 head items
    const uint32_t *ids = d.head_ids + (size_t)b * head::kHead;
    float *part = d.dB_part + (size_t)b * part_stride;
    head::HeadArgs a;
    a.n = n; a.ntiles = d.ntiles; a.K = c->Kp; a.ld = c->ld; a.Ktrue = c->K;
    a.Y = reinterpret_cast<const uint8_t *>(d.Yw) + (size_t)b * n_pad * head::kHead; a.head_ids = ids;
    a.T_theta = c->th.T; a.dB_part = part;
    a.ElogT = elog_src(c, c->th); a.ElogB = elog_src(c, c->be); a.TdirectT = c->th.Tdirect; a.TdirectB = c->be.Tdirect;
    a.flagT = c->th.direct_flag; a.flagB = c->be.direct_flag; a.slow_count = c->slow_count;
    a.Tb_theta = c->bias ? c->th.Tb : nullptr;
    a.ElogbT = c->th.b_Elog; a.ElogbB = c->be.b_Elog; a.TbdirectT = c->th.Tbdirect; a.TbdirectB = c->be.Tbdirect;
    kernel<<<grid, head_threads, head::kSmemBytes, c->stream>>>(d.map_a_hi, d.map_a_lo, d.map_b_hi[b], d.map_b_lo[b], a);
    c->launches++;
  }
  // the head items' T_beta rows from the CTAs' partial sums, all blocks in one launch (unused slots return at once)
  head::head_reduce_kernel<<<d.nblocks * head::kHead, head::kFact * head::kReduceSplit, 0, c->stream>>>(
      d.dB_part, part_stride, grid, d.head_ids, c->Kp, c->ld, c->be.T, c->bias ? c->be.Tb : nullptr);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

int ensure_aux(hpf_ctx *c)
{
  if (!c->bias || !c->aux_dirty) return 0;
  build_aux_kernel<<<(c->th.R + 255) / 256, 256, 0, c->stream>>>(c->th.b_Elog, c->th.shift, c->th.R, c->th.aux);
  build_aux_kernel<<<(c->be.R + 255) / 256, 256, 0, c->stream>>>(c->be.b_Elog, c->be.shift, c->be.R, c->be.aux);
  c->launches += 2;
  CU(cudaGetLastError());
  c->aux_dirty = false;
  return 0;
}

// ---- collectives: all on comm_stream, ordered against the compute stream through events ---------------------
int comm_streams(hpf_ctx *c) // after c->comm / c->nranks are set
{
  if (c->nranks > 1 && !c->comm_stream) {
    // highest priority: the collective's few CTAs must get onto the SMs while a sweep grid is still draining
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char *pe = getenv("HPF_COMM_PRIO"); // experiments: 0 = default priority
    CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, pe && atoi(pe) == 0 ? prio_lo : prio_hi));
    for (auto &e : c->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_theta, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_beta, cudaEventDisableTiming));
  }
  return 0;
}

int nccl_fail(hpf_ctx *c, int rc, const char *what)
{
  return fail(c, HPF_ENCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
}

// in-place sum over the ranks of `count` floats at p, enqueued on comm_stream (the caller has made comm_stream
// wait for the producer)
int comm_allreduce(hpf_ctx *c, float *p, size_t count)
{
  if (count == 0) return 0;
  const int rc = g_nccl.AllReduce(p, p, count, ncclFloat32, ncclSum, c->comm, c->comm_stream);
  return rc == ncclSuccess ? 0 : nccl_fail(c, rc, "ncclAllReduce");
}

// one CAVI iteration; order of hgaprec.cc:1340-1414 (hier), 928-956 (vb), 1227-1297 (vb_bias, both orderings).
// The two sweeps are independent of each other (both read the previous expectations), so their order is free:
//   no dense head:  item pass (chunk by chunk) -> user pass -> theta update -> beta update
//   dense head:     user pass -> head (adds to T_theta, writes the head items' T_beta rows) -> item pass -> ...
// Multi-GPU: the all-reduce of a chunk's T_beta rows starts on comm_stream as soon as the chunk's sweep and combine
// are done and runs under everything that follows on the compute stream; the tail [Tb_beta | sum_u E[theta] | flag]
// follows the theta update; the beta update waits for comm_stream.
// Profiling events come in pairs (begin, end) per stage: 0 user pass, 1 item pass, 2 combine (theta), 3 dense head,
// 4 theta update, 5 exposed wait for the collectives, 6 beta update, 7 whole iteration.
#define STAGE_BEGIN(i) do { if (c->profiling) CU(cudaEventRecord(c->pev[2 * (i)], c->stream)); } while (0)
#define STAGE_END(i) do { if (c->profiling) CU(cudaEventRecord(c->pev[2 * (i) + 1], c->stream)); } while (0)

int user_pass(hpf_ctx *c)
{
  STAGE_BEGIN(0);
  TRY(launch_sweep(c, c->th, c->be)); // CSR rows = users (the tail items when the head runs dense): T_theta
  STAGE_END(0);
  STAGE_BEGIN(2);
  TRY(launch_combine(c, c->th));
  STAGE_END(2);
  return 0;
}

int item_pass(hpf_ctx *c)
{
  const bool mg = c->nranks > 1;
  const WorkList &w = c->be.wl;
  STAGE_BEGIN(1);
  for (uint32_t q = 0; q < w.nchunks; ++q) {
    TRY(launch_sweep(c, c->be, c->th, (int)q)); // CSC rows = items of chunk q: T_beta (local users only)
    TRY(launch_combine(c, c->be, (int)q));
    if (mg) CU(cudaEventRecord(c->ev_chunk[q], c->stream)); // rows of chunk q are final on this rank from here on
  }
  STAGE_END(1);
  return 0;
}

// The collectives of one iteration, issued after every compute kernel up to the theta update has been enqueued, so
// the host never sits in an NCCL call while the compute stream runs dry.  On the device comm_stream waits for each
// chunk's event, so chunk q's all-reduce runs under whatever the compute stream does after that chunk.
int issue_collectives(hpf_ctx *c)
{
  const WorkList &w = c->be.wl;
  if (c->shard_now) { // T_beta: every rank keeps the sum of its own slice of items only
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_chunk[w.nchunks - 1], 0));
    const size_t cnt = (size_t)c->slice_rows * c->ld;
    const int rc = g_nccl.ReduceScatter(c->be.T, c->be.T + (size_t)c->rank * cnt, cnt, ncclFloat32, ncclSum, c->comm, c->comm_stream);
    if (rc != ncclSuccess) return nccl_fail(c, rc, "ncclReduceScatter");
  }
  for (uint32_t q = 0; q < w.nchunks && !c->shard_now; ++q) {
    const size_t r0 = (size_t)q * w.chunk_rows, r1 = std::min<size_t>((size_t)(q + 1) * w.chunk_rows, c->be.R);
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_chunk[c->ar_defer ? w.nchunks - 1 : q], 0));
    TRY(comm_allreduce(c, c->be.T + r0 * c->ld, (r1 - r0) * c->ld));
  }
  // [Tb_beta | sum_u E[theta] | fallback flag] summed over the user shards; with mg_exact also the item side's
  // fallback buffers [Tdirect_beta | Tbdirect_beta]
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_theta, 0));
  TRY(comm_allreduce(c, c->red_tail, c->red_tail_count));
  if (c->mg_exact) TRY(comm_allreduce(c, c->redblock2, c->red2_count));
  CU(cudaEventRecord(c->ev_comm, c->comm_stream));
  return 0;
}

// in-place all-gather of an array whose rank-r slice is elements [r * count, (r + 1) * count)
int comm_allgather(hpf_ctx *c, float *base, size_t count)
{
  if (count == 0) return 0;
  const int rc = g_nccl.AllGather(base + (size_t)c->rank * count, base, count, ncclFloat32, c->comm, c->comm_stream);
  return rc == ncclSuccess ? 0 : nccl_fail(c, rc, "ncclAllGather");
}

// sharded beta update (opt-in): does this window qualify?  Every input is the same on all ranks.
constexpr uint64_t kShardMinBytes = 64ull << 20;
void decide_sharding(hpf_ctx *c)
{
  c->shard_now = c->last_sharded = false;
  if (c->nranks <= 1 || c->shard_mode == 0 || c->bias || c->logl || c->jacobi || c->mg_exact) return;
  if (!g_nccl.ReduceScatter || !g_nccl.AllGather) return;
  const uint64_t payload = (uint64_t)c->be.R * c->ld * sizeof(float);
  if (c->shard_mode < 0 && payload < kShardMinBytes) return; // few items: the replicated update is cheaper than two more collectives
  c->shard_now = c->last_sharded = true;
  c->slice_rows = (c->be.R + (uint32_t)c->nranks - 1) / (uint32_t)c->nranks;
}

// end of a sharded window: the rest of beta's state, whole on every rank again
int gather_beta_state(hpf_ctx *c)
{
  if (!c->shard_now) return 0;
  CU(cudaEventRecord(c->ev_beta, c->stream));
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_beta, 0));
  const size_t rows = c->slice_rows;
  if (g_nccl.GroupStart && g_nccl.GroupEnd) g_nccl.GroupStart();
  int rc = comm_allgather(c, c->be.shape, rows * c->ld);
  if (!rc) rc = comm_allgather(c, c->be.shift, rows);
  if (!rc && c->hier) {
    rc = comm_allgather(c, c->be.rate_row, rows);
    if (!rc) rc = comm_allgather(c, c->be.pr_shape, rows);
    if (!rc) rc = comm_allgather(c, c->be.pr_rate, rows);
    if (!rc) rc = comm_allgather(c, c->be.pr_Ev, rows);
  }
  if (g_nccl.GroupStart && g_nccl.GroupEnd) {
    const int grc = g_nccl.GroupEnd();
    if (!rc && grc != ncclSuccess) rc = nccl_fail(c, grc, "ncclGroupEnd");
  }
  if (rc) return rc;
  CU(cudaEventRecord(c->ev_comm, c->comm_stream));
  CU(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
  c->shard_now = false;
  return 0;
}

int one_iteration(hpf_ctx *c)
{
  const bool mg = c->nranks > 1;
  if (c->logl && c->hier) { // logl() sees the rate priors of THIS iteration's set_prior_rate (gpbase.hh:163-173)
    CU(cudaMemcpyAsync(c->th.pr_shape_prev, c->th.pr_shape, sizeof(float) * c->th.R, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->th.pr_rate_prev, c->th.pr_rate, sizeof(float) * c->th.R, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->be.pr_shape_prev, c->be.pr_shape, sizeof(float) * c->be.R, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->be.pr_rate_prev, c->be.pr_rate, sizeof(float) * c->be.R, cudaMemcpyDeviceToDevice, c->stream));
    c->pr_prev_valid = true;
  }
  STAGE_BEGIN(7);
  if (c->dense.on) {
    TRY(user_pass(c));
    STAGE_BEGIN(3);
    TRY(launch_dense_head(c)); // the dense head block on tcgen05: T_theta +=, T_beta[head items] =
    STAGE_END(3);
    TRY(item_pass(c));
  } else {
    TRY(item_pass(c));
    TRY(user_pass(c));
  }
  const double n_glob = c->cfg.n_users_global ? (double)c->cfg.n_users_global : (double)c->cfg.n_users;
  if (c->jacobi) { // -novb: beta's rate uses the OLD (global) sum_u E[theta], hgaprec.cc:1278-1283
    if (mg && !c->th_colsum_global) { // first iteration after hpf_set_state: the local sums have not been reduced yet
      CU(cudaEventRecord(c->ev_theta, c->stream));
      CU(cudaStreamWaitEvent(c->comm_stream, c->ev_theta, 0));
      TRY(comm_allreduce(c, c->th.colsum, c->Kp));
      CU(cudaEventRecord(c->ev_comm, c->comm_stream));
      CU(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
    }
    CU(cudaMemcpyAsync(c->colsum_theta_old, c->th.colsum, sizeof(float) * c->Kp, cudaMemcpyDeviceToDevice, c->stream));
  }
  // theta: rate from the old beta column sums (Gauss-Seidel and Jacobi alike)
  STAGE_BEGIN(4);
  TRY(launch_update(c, c->th, c->be.colsum, (double)c->cfg.n_items));
  STAGE_END(4);
  STAGE_BEGIN(5);
  if (mg) {
    CU(cudaEventRecord(c->ev_theta, c->stream));
    TRY(issue_collectives(c));
    CU(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
    c->th_colsum_global = true;
  }
  STAGE_END(5);
  // beta: Gauss-Seidel uses the NEW sum_u E[theta] (hgaprec.cc:1380-1384), -novb the old one
  STAGE_BEGIN(6);
  if (c->shard_now) {
    const uint32_t r0 = std::min<uint64_t>((uint64_t)c->rank * c->slice_rows, c->be.R);
    const uint32_t r1 = std::min<uint64_t>((uint64_t)(c->rank + 1) * c->slice_rows, c->be.R);
    TRY(launch_update(c, c->be, c->th.colsum, n_glob, r0, r1 - r0));
    // sum_i E[beta] over the slices, then the rows of A_beta the other ranks updated: all the next sweeps read
    CU(cudaEventRecord(c->ev_beta, c->stream));
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_beta, 0));
    TRY(comm_allreduce(c, c->be.colsum, c->Kp));
    TRY(comm_allgather(c, c->be.A, (size_t)c->slice_rows * c->ld));
    CU(cudaEventRecord(c->ev_comm, c->comm_stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
  } else {
    TRY(launch_update(c, c->be, c->jacobi ? c->colsum_theta_old : c->th.colsum, n_glob));
  }
  STAGE_END(6);
  STAGE_END(7);
  c->iterations++;
  return 0;
}

int check_ready(hpf_ctx *c)
{
  if (!c->ratings_set) return fail(c, HPF_EINVAL, "hpf_set_ratings_csr has not been called");
  if (!c->th.have_state || !c->be.have_state) return fail(c, HPF_EINVAL, "hpf_set_state(HPF_THETA/HPF_BETA) has not been called");
  if (c->hier && (!c->th.have_pr || !c->be.have_pr))
    return fail(c, HPF_EINVAL, "hpf_set_state(HPF_THETARATE/HPF_BETARATE) has not been called");
  if (c->bias && (!c->th.have_bias || !c->be.have_bias))
    return fail(c, HPF_EINVAL, "hpf_set_state(HPF_THETABIAS/HPF_BETABIAS) has not been called");
  return 0;
}

// upload a host fp64 array into a temporary device buffer
int stage_in(hpf_ctx *c, const double *host, size_t count, double **dev)
{
  void *q = nullptr;
  CU(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(double)));
  *dev = (double *)q;
  cudaError_t e = cudaMemcpyAsync(q, host, count * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) { cudaFree(q); return fail(c, HPF_ECUDA, "H2D copy: %s", cudaGetErrorString(e)); }
  return 0;
}

// ---- group ctx (hpf_config.n_devices > 1): defined at the end of this file ----
int group_create(const hpf_config *cfg, hpf_ctx **out);
int group_set_ratings(hpf_ctx *g, const uint64_t *row_ptr, const uint32_t *col_idx, const uint8_t *y);
int group_set_state(hpf_ctx *g, int which, const double *shape, const double *rate, const double *Ev, const double *Elogv);
int group_get_state(hpf_ctx *g, int which, double *shape, double *rate, double *Ev, double *Elogv);
int group_iterate(hpf_ctx *g, uint32_t n_iters, hpf_iter_profile *prof);
int group_heldout(hpf_ctx *g, const uint32_t *u, const uint32_t *i, const uint8_t *y, uint64_t npairs, double *sum_ll);
int group_elbo(hpf_ctx *g, double *out);
int group_topn(hpf_ctx *g, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx, uint32_t topn,
               uint32_t *items_out, float *scores_out);
int group_item_ranks(hpf_ctx *g, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx,
                     const uint64_t *query_ptr, const uint32_t *query_idx, uint32_t *rank_out, float *score_out);
int group_get_stats(const hpf_ctx *g, hpf_stats *out);

} // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

void hpf_config_default(hpf_config *cfg)
{
  memset(cfg, 0, sizeof *cfg);
  cfg->abi_version = HPF_ABI_VERSION;
  cfg->theta_shape = cfg->theta_rate = cfg->beta_shape = cfg->beta_rate = 0.3;
  cfg->thetarate_shape = cfg->thetarate_rate = cfg->betarate_shape = cfg->betarate_rate = 0.3;
  cfg->thetabias_shape = cfg->thetabias_rate = cfg->betabias_shape = cfg->betabias_rate = 0.3;
}

const char *hpf_last_error(const hpf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hpf_create(const hpf_config *cfg, hpf_ctx **out)
{
  hpf_ctx *c = nullptr; // CU()/fail() report into the thread-local slot until the ctx exists
  if (!cfg || !out) return fail(c, HPF_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->abi_version != HPF_ABI_VERSION) return fail(c, HPF_EINVAL, "abi_version %u != %u", cfg->abi_version, HPF_ABI_VERSION);
  if (cfg->k == 0 || cfg->k > 1024) return fail(c, HPF_EINVAL, "k=%u out of range [1,1024]", cfg->k);
  if (cfg->n_items == 0) return fail(c, HPF_EINVAL, "n_items must be > 0");
  if (cfg->n_users == 0) return fail(c, HPF_EINVAL, "n_users must be > 0 (an empty user shard: use fewer ranks)");
  if (cfg->n_devices > 1) return group_create(cfg, out);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(c, HPF_ENODEVICE, "no CUDA device available (libhpf_b200 has no CPU path)");
  const int device = cfg->n_devices == 1 ? cfg->devices[0] : cfg->device;
  if (device < 0 || device >= ndev) return fail(c, HPF_EINVAL, "device %d not in [0,%d)", device, ndev);
  CU(cudaSetDevice(device));
  hpf_ctx *n = new hpf_ctx();
  n->cfg = *cfg;
  n->cfg.device = device;
  n->K = cfg->k; n->Kp = (cfg->k + 3u) & ~3u; n->K4 = n->Kp / 4;
  if (n->Kp >= 32) n->ld = (n->Kp + 31u) & ~31u;
  else for (n->ld = 4; n->ld < n->Kp; n->ld *= 2) {}
  n->hier = cfg->flags & HPF_HIER; n->bias = cfg->flags & HPF_BIAS; n->binary = cfg->flags & HPF_BINARY;
  n->jacobi = (cfg->flags & HPF_JACOBI) && !n->hier;
  n->logl = cfg->flags & HPF_LOGL;
  cudaDeviceGetAttribute(&n->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  if (const char *e = getenv("HPF_SEG_LEN")) { int v = atoi(e); if (v >= 8 && v <= 65536) n->seg_len = v; }
  if (const char *e = getenv("HPF_L2_TILE_MB")) { int v = atoi(e); if (v >= 0 && v <= 4096) n->l2_tile_bytes = (uint64_t)v << 20; }
  if (const char *e = getenv("HPF_L2_TILE_KB")) { int v = atoi(e); if (v >= 0) { n->l2_tile_bytes = (uint64_t)v << 10; n->l2_tile_forced = true; } } // tests
  pick_sweep_shape(n);
  if (const char *e = getenv("HPF_DENSE_HEAD")) n->dense_head_mode = atoi(e);
  if (const char *e = getenv("HPF_DENSE_BLOCK_SHARE")) n->dense_block_share = atof(e);
  if (const char *e = getenv("HPF_HEAD_VARIANT")) n->head_variant = atoi(e) & 7;
  if (const char *e = getenv("HPF_PACK")) n->pack_ok = atoi(e) != 0;
  if (const char *e = getenv("HPF_AR_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= (int)kMaxChunks) n->ar_chunks = v; }
  if (const char *e = getenv("HPF_SHARD_BETA")) n->shard_mode = atoi(e);
  if (const char *e = getenv("HPF_AR_DEFER")) n->ar_defer = atoi(e) != 0;
  if (const char *e = getenv("HPF_MG_EXACT")) n->mg_exact = atoi(e) != 0; // tests: fallback buffers inside the all-reduce from the start
  c = n;
  int rc = 0;
  do {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(nullptr, HPF_ECUDA, "cudaStreamCreate failed"); break; }
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(nullptr, HPF_ECUDA, "cudaStreamCreate failed"); break; }
    cudaEventCreateWithFlags(&c->ev_col, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_y, cudaEventDisableTiming);
    for (auto &e : c->pev) cudaEventCreate(&e);
    c->th.prior_shape = cfg->theta_shape; c->th.prior_rate = cfg->theta_rate;
    c->th.pr_prior_shape = cfg->thetarate_shape; c->th.pr_prior_rate = cfg->thetarate_rate;
    c->th.bias_prior_shape = cfg->thetabias_shape; c->th.bias_prior_rate = cfg->thetabias_rate;
    c->be.prior_shape = cfg->beta_shape; c->be.prior_rate = cfg->beta_rate;
    c->be.pr_prior_shape = cfg->betarate_shape; c->be.pr_prior_rate = cfg->betarate_rate;
    c->be.bias_prior_shape = cfg->betabias_shape; c->be.bias_prior_rate = cfg->betabias_rate;
    if ((rc = alloc_side(c, c->th, cfg->n_users, true))) break;
    if ((rc = alloc_side(c, c->be, cfg->n_items, false))) break;
    // item-side reduce blocks: [T_beta (m x ld) | Tb_beta (m, padded to 4) | colsum_theta (Kp) | flag (4)] and
    // [Tdirect_beta (m x ld) | Tbdirect_beta (m, padded to 4)]
    // (T_beta with kRowPad spare rows: the in-place reduce-scatter of the sharded beta update works on equal slices)
    const size_t mk = ((size_t)cfg->n_items + kRowPad) * c->ld, mpad = c->bias ? (((size_t)cfg->n_items + 3) & ~(size_t)3) : 0;
    c->red_count = mk + mpad + c->Kp + 4;
    if ((rc = dalloc(c, &c->redblock, c->red_count))) break;
    c->be.T = c->redblock;
    c->be.Tb = c->bias ? c->redblock + mk : nullptr;
    c->th.colsum = c->redblock + mk + mpad;
    c->red_tail = c->redblock + mk; c->red_tail_count = mpad + c->Kp + 4;
    c->red_flag = c->redblock + mk + mpad + c->Kp;
    const size_t mk2 = (size_t)cfg->n_items * c->ld;
    c->red2_count = mk2 + mpad;
    if ((rc = dalloc(c, &c->redblock2, c->red2_count))) break;
    c->be.Tdirect = c->redblock2;
    c->be.Tbdirect = c->bias ? c->redblock2 + mk2 : nullptr;
    if ((rc = dalloc(c, &c->mg_fired, 4))) break;
    if ((rc = dalloc(c, &c->be.colsum, c->Kp))) break;
    if ((rc = dalloc(c, &c->th.T, (size_t)cfg->n_users * c->ld))) break;
    if (c->bias && (rc = dalloc(c, &c->th.Tb, cfg->n_users))) break;
    if ((rc = dalloc(c, &c->colsum_theta_old, c->Kp))) break;
    if ((rc = dalloc(c, &c->slow_count, 1))) break;
    if ((rc = dalloc(c, &c->scratch_u32, 4))) break;
    if ((rc = dalloc(c, &c->logfact, 256))) break;
    if ((rc = dalloc(c, &c->ll_blocks, (size_t)c->sm_count * 8))) break;
    if ((rc = dalloc(c, &c->ll_out, 1))) break;
    if (c->logl && (rc = dalloc(c, &c->elbo_blocks, (size_t)kElboLaunches * c->sm_count * 8))) break;
    double lf[256];
    lf[0] = lf[1] = log(1.0);
    for (int v = 2; v < 256; ++v) lf[v] = lf[v - 1] + log((double)v); // log_factorial, hgaprec.cc:1563-1570
    if (cudaMemcpyAsync(c->logfact, lf, sizeof lf, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(c, HPF_ECUDA, "initial upload failed"); break; }
  } while (0);
  if (rc != 0) {
    g_create_error = c->err.empty() ? g_create_error : c->err;
    hpf_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

void hpf_destroy(hpf_ctx *c)
{
  if (!c) return;
  if (c->is_group) {
    for (hpf_ctx *k : c->kids) hpf_destroy(k);
    delete c;
    return;
  }
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->ev_col) cudaEventDestroy(c->ev_col);
  if (c->ev_y) cudaEventDestroy(c->ev_y);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (auto &e : c->ev_chunk) if (e) cudaEventDestroy(e);
  if (c->ev_theta) cudaEventDestroy(c->ev_theta);
  if (c->ev_comm) cudaEventDestroy(c->ev_comm);
  if (c->ev_beta) cudaEventDestroy(c->ev_beta);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  for (auto &p : c->allocs) cudaFree(p.first);
  if (c->dev_arena.base) cudaFree(c->dev_arena.base);
  if (c->pin_arena.base) cudaFreeHost(c->pin_arena.base);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (auto &e : c->pev) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int hpf_set_ratings_csr(hpf_ctx *c, const uint64_t *row_ptr, const uint32_t *col_idx, const uint8_t *y)
{
  if (!c || !row_ptr) return fail(c, HPF_EINVAL, "null argument");
  if (c->is_group) return group_set_ratings(c, row_ptr, col_idx, y);
  CU(cudaSetDevice(c->cfg.device));
  const uint32_t n = c->cfg.n_users, m = c->cfg.n_items;
  const uint64_t nnz = row_ptr[n];
  if (row_ptr[0] != 0) return fail(c, HPF_EINVAL, "row_ptr[0] must be 0");
  for (uint32_t r = 0; r < n; ++r)
    if (row_ptr[r + 1] < row_ptr[r]) return fail(c, HPF_EINVAL, "row_ptr not monotone at row %u", r);
  if (nnz > 0 && !col_idx) return fail(c, HPF_EINVAL, "col_idx is null");
  if (nnz >= 0xffffffffull) return fail(c, HPF_EINVAL, "nnz=%llu per ctx exceeds 2^32-1", (unsigned long long)nnz);
  Trace tr(c->stream);
  c->ratings_set = false;
  c->nnz = nnz;
  c->pin_arena.pinned_host = true;
  CU(cudaStreamSynchronize(c->copy_stream)); // a copy left over from a call that failed half-way
  c->y_pending = false;
  const uint32_t L = c->seg_len;
  const uint32_t th_t = tiles_for(c, m, n, nnz), be_t = tiles_for(c, n, m, nnz);
  const bool try_dense = c->dense_head_mode != 0 && c->Kp + (c->bias ? 2u : 0u) <= (uint32_t)head::kFact && nnz > 0;
  c->dense.on = false;
  // Chunks of the item pass (multi-GPU): with C > 1 chunks the all-reduce of a finished chunk's T_beta rows starts under
  // the next chunk's sweep.  Measured (MSD scale, 307 MB payload) that does not pay, wherever the host issues the calls
  // from: the item sweep is HBM-bound and NCCL's reduce/copy CTAs take bandwidth and SMs from it, so the sweep loses
  // about what the all-reduce gains (8 x B200: item pass 0.71 / 1.27 / 2.39 ms with 1 / 3 / 6 chunks, iteration 2.67 /
  // 2.74 / 3.69 ms, profiles/r02p_bench_n8_msd*.json; 2 x B200 with the calls deferred: 6.47 vs 6.81 ms,
  // profiles/r02q_*).  One chunk, whose all-reduce runs under the user pass and the theta update, is the default;
  // HPF_AR_CHUNKS keeps the pipelined form available (and tested).
  uint32_t item_chunks = 1;
  if (c->nranks > 1 && c->ar_chunks > 0) item_chunks = (uint32_t)c->ar_chunks;
  size_t cub_bytes = 0, scan_bytes = 0;
  const uint64_t seg_bound = std::max(worklist_seg_bound(nnz, n, th_t, L), worklist_seg_bound(nnz, m, be_t, L));
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint64_t *)nullptr,
                                  (uint64_t *)nullptr, (int64_t)std::max<uint64_t>(std::max<uint64_t>(nnz, 1), seg_bound), 0, 32, c->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                (int64_t)std::max<uint64_t>(std::max<uint64_t>(nnz + 1, m), n), c->stream);
  cub_bytes = std::max(cub_bytes, scan_bytes);
  const size_t dev_need = pad256(((size_t)n + 1) * 8) + pad256(nnz * 4) + 8 * pad256((size_t)m * 4) +
                          orientation_dev_bytes(nnz, m, be_t, cub_bytes) + orientation_dev_bytes(nnz, n, th_t, cub_bytes) +
                          worklist_dev_bytes(nnz, n, th_t, L, cub_bytes) + worklist_dev_bytes(nnz, m, be_t, L, cub_bytes) +
                          2 * pad256((nnz + 1) * 4) + pad256(nnz * 4) + 2 * pad256(((size_t)n + 1) * 8) + 2 * pad256(cub_bytes) + (1u << 16);
  const size_t pin_need = 2 * pad256((size_t)m * 4) + (1u << 16);
  TRY(arena_reserve(c, c->dev_arena, dev_need));
  TRY(arena_reserve(c, c->pin_arena, pin_need));
  TRY(ensure(c, &c->csr_idx, &c->csr_idx_cap, nnz));
  if (y) TRY(ensure(c, &c->csr_y, &c->csr_y_cap, nnz));
  const uint8_t *d_y = y ? c->csr_y : nullptr;
  c->csr_has_y = y != nullptr;
  Arena &dev = c->dev_arena, &pin = c->pin_arena;
  tr.mark("reserve");

  // ================= stage 1 (async): upload, check, item ordering, item degrees =================
  uint64_t *d_rowptr = dev.get<uint64_t>((size_t)n + 1);
  uint32_t *d_rowof = dev.get<uint32_t>(nnz);
  uint32_t *h_bad = pin.get<uint32_t>(1);
  if (!d_rowptr || !d_rowof || !h_bad) return fail(c, HPF_ENOMEM, "set-up arena too small");
  *h_bad = 0;
  CU(cudaMemsetAsync(c->scratch_u32, 0, 8, c->stream));
  const unsigned nb = (unsigned)((nnz + 255) / 256);
  CU(cudaMemcpyAsync(d_rowptr, row_ptr, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (nnz > 0) {
    CU(cudaMemcpyAsync(c->csr_idx, col_idx, nnz * 4, cudaMemcpyHostToDevice, c->stream));
    if (y) { // the ratings follow the indices on the copy stream: nothing needs them before the first sort is done
      CU(cudaEventRecord(c->ev_col, c->stream));
      CU(cudaStreamWaitEvent(c->copy_stream, c->ev_col, 0));
      CU(cudaMemcpyAsync(c->csr_y, y, nnz, cudaMemcpyHostToDevice, c->copy_stream));
      CU(cudaEventRecord(c->ev_y, c->copy_stream));
      c->y_pending = true;
    }
    check_range_kernel<<<nb, 256, 0, c->stream>>>(c->csr_idx, nnz, m, c->scratch_u32); // every item index must be < n_items
    expand_rows_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(d_rowptr, n, nnz, d_rowof);
    c->launches += 2;
  }
  if (c->logl) { // hpf_elbo walks the ratings user by user
    TRY(ensure(c, &c->csr_rowptr, &c->csr_rowptr_cap, (size_t)n + 1));
    CU(cudaMemcpyAsync(c->csr_rowptr, d_rowptr, ((size_t)n + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(h_bad, c->scratch_u32, 4, cudaMemcpyDeviceToHost, c->stream));
  // item pass ordering: rows = items, gathers user rows (users ascending inside a run)
  Orientation io, uo;
  TRY(orient_device(c, dev, nnz, c->csr_idx, d_rowof, d_y, false, nullptr, m, n, &c->csc_idx, &c->csc_idx_cap, &c->csc_y,
                    &c->csc_y_cap, cub_bytes, &io));
  // item degrees (from the item runs) and their descending order: the candidates for the dense head
  uint32_t *d_id_s = nullptr, *h_degkey = nullptr;
  if (try_dense) {
    uint32_t *d_deg = dev.get<uint32_t>(m), *d_degkey = dev.get<uint32_t>(m), *d_degkey_s = dev.get<uint32_t>(m), *d_id = dev.get<uint32_t>(m);
    d_id_s = dev.get<uint32_t>(m);
    h_degkey = pin.get<uint32_t>(m);
    void *d_tmp = dev.get<char>(cub_bytes);
    if (!d_deg || !d_degkey || !d_degkey_s || !d_id || !d_id_s || !h_degkey || !d_tmp) return fail(c, HPF_ENOMEM, "set-up arena too small (degrees)");
    degree_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(io.d_run, m, io.ntiles, d_deg);
    neg_key_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(d_deg, m, d_degkey, d_id);
    size_t tb = cub_bytes;
    CU(cub::DeviceRadixSort::SortPairs(d_tmp, tb, (const uint32_t *)d_degkey, d_degkey_s, (const uint32_t *)d_id, d_id_s, (int64_t)m, 0, 32, c->stream));
    c->launches += 2;
    const uint32_t ncand = std::min<uint32_t>(m, kMaxHeadBlocks * head::kHead);
    CU(cudaMemcpyAsync(h_degkey, d_degkey_s, (size_t)ncand * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream)); // the head decision needs the degrees on the host
    CU(cudaGetLastError());
    tr.mark("stage 1: upload, item order");
    if (*h_bad != 0) return fail(c, HPF_EINVAL, "col_idx holds item %u >= n_items=%u", *h_bad, m);
  }

  // ================= decision: dense head on the tensor cores =================
  // blocks of 128 items by descending popularity.  The first must carry >= 15 % of the nonzeros.  Every further one
  // must pay for itself: a block costs one dense pass over ALL local users whatever its density -- measured ~9.8 us
  // per wave of user tiles (one tile of 128 users per SM) plus ~30 us of fixed cost (three launches, operand split,
  // TMEM / B_head prologue) -- and saves ~6.6e-8 ms of gathers per nonzero over the two passes (Netflix scale: 5.36 /
  // 5.24 / 5.23 ms per iteration with 2 / 3 / 4 blocks, profiles/r02f_exp_block_share.log; with the users spread over 8
  // GPUs the fixed part dominates and two blocks are right).  HPF_DENSE_BLOCK_SHARE overrides the model with a
  // plain minimum share of the nonzeros.
  bool dense_head = false;
  uint32_t H = 0;
  uint64_t head_nnz = 0;
  if (try_dense) {
    uint32_t nblk = 0;
    for (uint32_t b = 0; b < kMaxHeadBlocks && b * head::kHead < m; ++b) {
      uint64_t blk = 0;
      for (uint32_t r = b * head::kHead; r < std::min<uint32_t>((b + 1) * head::kHead, m); ++r) blk += 0xffffffffu - h_degkey[r];
      const double waves = std::ceil((double)((n + head::kUsers - 1) / head::kUsers) / (double)c->sm_count);
      const bool pays = c->dense_block_share >= 0.0 ? (double)blk >= c->dense_block_share * (double)nnz
                                                    : (double)blk * 6.6e-8 > 0.030 + 0.0098 * waves;
      const bool take = b == 0 ? (c->dense_head_mode == 1 || (double)blk >= 0.15 * (double)nnz) : pays;
      if (!take) break;
      head_nnz += blk;
      nblk = b + 1;
    }
    dense_head = nblk > 0;
    if (dense_head) {
      H = std::min<uint32_t>(nblk * head::kHead, m);
      c->dense.head_nnz = head_nnz;
      c->dense.nblocks = nblk;
    }
  }

  // ================= stage 2 (async): work lists on the device; user-pass head / tail split =================
  WlPending up, ip;
  uint32_t *d_slot = nullptr; // slot_of[item] among the head items, 0xffffffff for the tail
  uint32_t *h_yovf = nullptr; // dense head: a cell of the byte matrix Y overflowed
  if (dense_head) {
    d_slot = dev.get<uint32_t>(m);
    if (!d_slot) return fail(c, HPF_ENOMEM, "set-up arena too small (head slots)");
    CU(cudaMemsetAsync(d_slot, 0xff, (size_t)m * 4, c->stream));
    head_slot_kernel<<<(H + 255) / 256, 256, 0, c->stream>>>(d_id_s, H, d_slot);
    c->launches++;
  }
  // the item pass has no rows for head items: head_kernel produces their T_beta rows
  TRY(build_worklist_device(c, dev, pin, c->be, io.d_run, io.ntiles, nnz, d_slot, item_chunks, io, cub_bytes, &ip));
  if (dense_head) {
    const uint64_t ntail = nnz - head_nnz;
    TRY(ensure(c, &c->tail_idx, &c->tail_idx_cap, nnz));
    if (y) TRY(ensure(c, &c->tail_y, &c->tail_y_cap, nnz));
    uint32_t *d_istail = dev.get<uint32_t>(nnz + 1), *d_tailpos = dev.get<uint32_t>(nnz + 1);
    uint64_t *d_tailptr = dev.get<uint64_t>((size_t)n + 1), *d_headptr = dev.get<uint64_t>((size_t)n + 1);
    void *d_tmp = dev.get<char>(cub_bytes);
    if (!d_istail || !d_tailpos || !d_tailptr || !d_headptr || !d_tmp) return fail(c, HPF_ENOMEM, "set-up arena too small (head split)");
    head_flag_kernel<<<(unsigned)((nnz + 256) / 256), 256, 0, c->stream>>>(c->csr_idx, d_slot, nnz, d_istail);
    size_t sb = cub_bytes;
    CU(cub::DeviceScan::ExclusiveSum(d_tmp, sb, (const uint32_t *)d_istail, d_tailpos, (int64_t)(nnz + 1), c->stream));
    head_split_kernel<<<nb, 256, 0, c->stream>>>(c->csr_idx, d_y, d_slot, d_tailpos, nnz, c->tail_idx, c->tail_y, nullptr, nullptr);
    split_ptr_kernel<<<(n + 256) / 256, 256, 0, c->stream>>>(d_rowptr, d_tailpos, n, d_tailptr, d_headptr);
    c->launches += 3;
    DensePlan &dp = c->dense;
    dp.ntiles = (n + head::kUsers - 1) / head::kUsers;
    dp.nhead = H;
    const size_t n_pad = (size_t)dp.ntiles * head::kUsers;
    const size_t NB = dp.nblocks;
    TRY(ensure(c, &dp.Yw, &dp.Yw_cap, NB * n_pad * head::kHead / 4));
    TRY(ensure(c, &dp.head_ids, &dp.head_ids_cap, NB * head::kHead));
    TRY(ensure(c, &dp.a_hi, &dp.a_hi_cap, n_pad * head::kFact));
    TRY(ensure(c, &dp.a_lo, &dp.a_lo_cap, n_pad * head::kFact));
    TRY(ensure(c, &dp.b_hi, &dp.b_hi_cap, NB * head::kHead * head::kFact));
    TRY(ensure(c, &dp.b_lo, &dp.b_lo_cap, NB * head::kHead * head::kFact));
    TRY(ensure(c, &dp.dB_part, &dp.dB_part_cap, NB * std::min<uint32_t>(dp.ntiles, (uint32_t)c->sm_count) * head::kHead * head::kFact));
    CU(cudaMemsetAsync(dp.Yw, 0, NB * n_pad * head::kHead, c->stream));
    CU(cudaMemsetAsync(dp.head_ids, 0xff, NB * head::kHead * 4, c->stream));
    CU(cudaMemcpyAsync(dp.head_ids, d_id_s, (size_t)H * 4, cudaMemcpyDeviceToDevice, c->stream));
    h_yovf = pin.get<uint32_t>(1);
    if (!h_yovf) return fail(c, HPF_ENOMEM, "pinned arena too small");
    *h_yovf = 0;
    head::dense_y_kernel<<<nb, 256, 0, c->stream>>>(d_rowof, c->csr_idx, d_y, d_slot, nnz, n_pad * head::kHead, dp.Yw, c->scratch_u32 + 1);
    CU(cudaMemcpyAsync(h_yovf, c->scratch_u32 + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    c->launches++;
    bool maps_ok = make_bf16_map(&dp.map_a_hi, dp.a_hi, n_pad, head::kFact, head::kUsers) &&
                   make_bf16_map(&dp.map_a_lo, dp.a_lo, n_pad, head::kFact, head::kUsers);
    for (size_t b = 0; b < NB; ++b)
      maps_ok = maps_ok && make_bf16_map(&dp.map_b_hi[b], dp.b_hi + b * head::kHead * head::kFact, head::kHead, head::kFact, head::kHead) &&
                make_bf16_map(&dp.map_b_lo[b], dp.b_lo + b * head::kHead * head::kFact, head::kHead, head::kFact, head::kHead);
    if (!maps_ok) return fail(c, HPF_ECUDA, "cuTensorMapEncodeTiled failed for the dense head operands");
    // user pass = the tail: a CSR over the same users (presorted by row); L2-tiled like the plain user pass when needed
    uint32_t *d_tailrow = nullptr;
    if (tiles_for(c, m, n, ntail) > 1 && ntail > 0) {
      d_tailrow = dev.get<uint32_t>(ntail);
      if (!d_tailrow) return fail(c, HPF_ENOMEM, "set-up arena too small (tail rows)");
      expand_rows_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(d_tailptr, n, ntail, d_tailrow);
      c->launches++;
    }
    TRY(orient_device(c, dev, ntail, d_tailrow, c->tail_idx, y ? c->tail_y : nullptr, true, d_tailptr, n, m, &c->upass_idx,
                      &c->upass_idx_cap, &c->upass_y, &c->upass_y_cap, cub_bytes, &uo));
    TRY(build_worklist_device(c, dev, pin, c->th, uo.d_run, uo.ntiles, ntail, nullptr, 1, uo, cub_bytes, &up));
  } else { // user pass = the CSR itself (L2-tiled when the item rows outgrow the budget)
    TRY(orient_device(c, dev, nnz, d_rowof, c->csr_idx, d_y, true, d_rowptr, n, m, &c->upass_idx, &c->upass_idx_cap, &c->upass_y,
                      &c->upass_y_cap, cub_bytes, &uo));
    TRY(build_worklist_device(c, dev, pin, c->th, uo.d_run, uo.ntiles, nnz, nullptr, 1, uo, cub_bytes, &up));
  }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  tr.mark("stage 2: work lists, split");
  if (*h_bad != 0) return fail(c, HPF_EINVAL, "col_idx holds item %u >= n_items=%u", *h_bad, m);
  if (dense_head && *h_yovf != 0) {
    // repeated (user, head item) lines whose ratings add up past 255 do not fit a byte of Y (the reference walks every
    // line, hgaprec.cc:1340-1366, so the sum is what counts): plan this input again without the dense head
    const int mode = c->dense_head_mode;
    c->dense_head_mode = 0;
    const int rc = hpf_set_ratings_csr(c, row_ptr, col_idx, y);
    c->dense_head_mode = mode;
    return rc;
  }
  TRY(finish_worklist(c, c->th, up));
  TRY(finish_worklist(c, c->be, ip));
  if (dense_head) {
    c->dense.on = true;
    c->dense.a_dirty = true;
  }
  c->th_tiles = uo.ntiles; c->be_tiles = io.ntiles;
  c->ratings_set = true;
  return 0;
}

int hpf_set_state(hpf_ctx *c, int which, const double *shape, const double *rate, const double *Ev, const double *Elogv)
{
  if (!c) return fail(c, HPF_EINVAL, "null ctx");
  if (c->is_group) return group_set_state(c, which, shape, rate, Ev, Elogv);
  CU(cudaSetDevice(c->cfg.device));
  const bool theta_side = which == HPF_THETA || which == HPF_THETARATE || which == HPF_THETABIAS;
  Side &s = theta_side ? c->th : c->be;
  const uint32_t R = s.R, K = c->K, Kp = c->Kp, ld = c->ld;
  double *d0 = nullptr, *d1 = nullptr, *d2 = nullptr, *d3 = nullptr;
  int rc = 0;
  switch (which) {
  case HPF_THETA:
  case HPF_BETA: {
    if (!shape || !rate || !Ev || !Elogv) return fail(c, HPF_EINVAL, "THETA/BETA need shape, rate, Ev and Elogv");
    const size_t rk = (size_t)R * K;
    const uint32_t g = s.update_grid;
    if ((rc = stage_in(c, shape, rk, &d0)) || (rc = stage_in(c, rate, c->hier ? rk : K, &d1)) ||
        (rc = stage_in(c, Ev, rk, &d2)) || (rc = stage_in(c, Elogv, rk, &d3))) break;
    import_matrix_kernel<<<g, kUpdateWarps * 32, 0, c->stream>>>(d0, R, K, Kp, ld, s.shape, 0.f);
    if (c->hier) import_matrix_kernel<<<g, kUpdateWarps * 32, 0, c->stream>>>(d1, R, K, Kp, ld, s.rate, 1.f);
    else import_matrix_kernel<<<1, kUpdateWarps * 32, 0, c->stream>>>(d1, 1, K, Kp, ld, s.rate, 1.f);
    import_matrix_kernel<<<g, kUpdateWarps * 32, 0, c->stream>>>(d2, R, K, Kp, ld, s.Ev, 0.f);
    import_elog_kernel<<<g, kUpdateWarps * 32, 0, c->stream>>>(d3, R, K, Kp, ld, s.Elog, s.A, s.shift);
    c->launches += 4;
    if ((rc = refresh_colsum(c, s))) break;
    s.have_state = true;
    s.derived_valid = true; // rate and E[v] are the caller's (not functions of shape: src/gpbase.hh:324-340)
    c->aux_dirty = true;
    if (theta_side) c->dense.a_dirty = true;
    if (theta_side) c->th_colsum_global = false;
    break;
  }
  case HPF_THETARATE:
  case HPF_BETARATE: {
    if (!c->hier) return fail(c, HPF_EINVAL, "THETARATE/BETARATE exist only with HPF_HIER");
    if (!shape || !rate || !Ev) return fail(c, HPF_EINVAL, "THETARATE/BETARATE need shape, rate and Ev");
    if ((rc = stage_in(c, shape, R, &d0)) || (rc = stage_in(c, rate, R, &d1)) || (rc = stage_in(c, Ev, R, &d2))) break;
    const unsigned nb = (R + 255) / 256;
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d0, R, s.pr_shape);
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d1, R, s.pr_rate);
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d2, R, s.pr_Ev);
    c->launches += 3;
    s.have_pr = true;
    c->pr_prev_valid = false; // no iteration has used this xi / eta as a rate prior yet
    break;
  }
  case HPF_THETABIAS:
  case HPF_BETABIAS: {
    if (!c->bias) return fail(c, HPF_EINVAL, "THETABIAS/BETABIAS exist only with HPF_BIAS");
    if (!shape || !rate || !Ev || !Elogv) return fail(c, HPF_EINVAL, "bias sets need shape, rate, Ev and Elogv");
    if ((rc = stage_in(c, shape, R, &d0)) || (rc = stage_in(c, rate, R, &d1)) || (rc = stage_in(c, Ev, R, &d2)) ||
        (rc = stage_in(c, Elogv, R, &d3))) break;
    const unsigned nb = (R + 255) / 256;
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d0, R, s.b_shape);
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d1, R, s.b_rate);
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d2, R, s.b_Ev);
    import_vector_kernel<<<nb, 256, 0, c->stream>>>(d3, R, s.b_Elog);
    c->launches += 4;
    s.have_bias = true;
    c->aux_dirty = true;
    if (theta_side) c->dense.a_dirty = true; // the operand copy of A carries the user-bias columns
    break;
  }
  default:
    return fail(c, HPF_EINVAL, "unknown parameter id %d", which);
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(d0); cudaFree(d1); cudaFree(d2); cudaFree(d3);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(c, HPF_ECUDA, "hpf_set_state: %s", cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, HPF_ECUDA, "hpf_set_state: %s", cudaGetErrorString(e));
  return 0;
}

int hpf_get_state(hpf_ctx *c, int which, double *shape, double *rate, double *Ev, double *Elogv)
{
  if (!c) return fail(c, HPF_EINVAL, "null ctx");
  if (c->is_group) return group_get_state(c, which, shape, rate, Ev, Elogv);
  CU(cudaSetDevice(c->cfg.device));
  const bool theta_side = which == HPF_THETA || which == HPF_THETARATE || which == HPF_THETABIAS;
  Side &s = theta_side ? c->th : c->be;
  const uint32_t R = s.R, K = c->K, ld = c->ld;
  double *stage = nullptr;
  const size_t rk = (size_t)R * K;
  CU(cudaMalloc((void **)&stage, std::max<size_t>(rk, 16) * sizeof(double)));
  cudaError_t e = cudaSuccess;
  auto out_matrix = [&](const float *src, double *dst, uint32_t rows) {
    if (!dst || e != cudaSuccess) return;
    export_matrix_kernel<<<row_grid(c, rows), kUpdateWarps * 32, 0, c->stream>>>(src, rows, K, ld, stage);
    c->launches++;
    e = cudaMemcpyAsync(dst, stage, (size_t)rows * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  };
  auto out_vector = [&](const float *src, double *dst) {
    if (!dst || e != cudaSuccess) return;
    export_vector_kernel<<<(R + 255) / 256, 256, 0, c->stream>>>(src, R, stage);
    c->launches++;
    e = cudaMemcpyAsync(dst, stage, (size_t)R * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  };
  int rc = 0;
  if ((which == HPF_THETA || which == HPF_BETA) && (rate || Ev || Elogv)) {
    rc = ensure_derived(c, s);
    if (rc) { cudaFree(stage); return rc; }
  }
  switch (which) {
  case HPF_THETA:
  case HPF_BETA:
    out_matrix(s.shape, shape, R);
    out_matrix(s.rate, rate, c->hier ? R : 1);
    out_matrix(s.Ev, Ev, R);
    out_matrix(s.Elog, Elogv, R);
    break;
  case HPF_THETARATE:
  case HPF_BETARATE:
    if (!c->hier) { rc = fail(c, HPF_EINVAL, "THETARATE/BETARATE exist only with HPF_HIER"); break; }
    out_vector(s.pr_shape, shape);
    out_vector(s.pr_rate, rate);
    out_vector(s.pr_Ev, Ev);
    if (Elogv && e == cudaSuccess) {
      export_gparray_elog_kernel<<<(R + 255) / 256, 256, 0, c->stream>>>(s.pr_shape, s.pr_rate, R, stage);
      c->launches++;
      e = cudaMemcpyAsync(Elogv, stage, (size_t)R * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    break;
  case HPF_THETABIAS:
  case HPF_BETABIAS:
    if (!c->bias) { rc = fail(c, HPF_EINVAL, "THETABIAS/BETABIAS exist only with HPF_BIAS"); break; }
    out_vector(s.b_shape, shape);
    out_vector(s.b_rate, rate);
    out_vector(s.b_Ev, Ev);
    out_vector(s.b_Elog, Elogv);
    break;
  default:
    rc = fail(c, HPF_EINVAL, "unknown parameter id %d", which);
  }
  cudaFree(stage);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(c, HPF_ECUDA, "hpf_get_state: %s", cudaGetErrorString(e));
  return 0;
}

// ---- multi-GPU: optimistic handling of the exact fallback (see hpf_ctx::mg_exact) ---------------------------------
// every array an iteration changes and a later one reads
void state_arrays(hpf_ctx *c, std::vector<std::pair<void *, size_t>> *v)
{
  for (int side = 0; side < 2; ++side) {
    Side &s = side == 0 ? c->th : c->be;
    const size_t rk = (size_t)s.R * c->ld * sizeof(float), rv = (size_t)s.R * sizeof(float);
    v->emplace_back(s.A, rk); v->emplace_back(s.shape, rk); // rate, E[v], E[log v]: not written inside a window
    v->emplace_back(s.shift, rv); v->emplace_back(s.rate_col, c->Kp * sizeof(float));
    if (c->hier) {
      v->emplace_back(s.rate_row, rv); v->emplace_back(s.pr_shape, rv); v->emplace_back(s.pr_rate, rv); v->emplace_back(s.pr_Ev, rv);
      if (c->logl) { v->emplace_back(s.pr_shape_prev, rv); v->emplace_back(s.pr_rate_prev, rv); }
    }
    if (c->bias) {
      v->emplace_back(s.b_shape, rv); v->emplace_back(s.b_rate, rv); v->emplace_back(s.b_Ev, rv); v->emplace_back(s.b_Elog, rv);
      v->emplace_back(s.aux, (size_t)s.R * sizeof(float2));
    }
    v->emplace_back(s.colsum, c->Kp * sizeof(float));
  }
  v->emplace_back(c->colsum_theta_old, c->Kp * sizeof(float));
}

int snapshot_state(hpf_ctx *c, bool restore)
{
  std::vector<std::pair<void *, size_t>> v;
  state_arrays(c, &v);
  size_t total = 0;
  for (auto &e : v) total += pad256(e.second);
  if (!restore) TRY(ensure(c, &c->snap, &c->snap_cap, total / sizeof(float)));
  char *p = reinterpret_cast<char *>(c->snap);
  for (auto &e : v) {
    if (restore) CU(cudaMemcpyAsync(e.first, p, e.second, cudaMemcpyDeviceToDevice, c->stream));
    else CU(cudaMemcpyAsync(p, e.first, e.second, cudaMemcpyDeviceToDevice, c->stream));
    p += pad256(e.second);
  }
  return 0;
}

int run_window(hpf_ctx *c, uint32_t n_iters)
{
  CU(cudaEventRecord(c->ev0, c->stream));
  decide_sharding(c);
  for (uint32_t it = 0; it < n_iters; ++it) TRY(one_iteration(c));
  TRY(gather_beta_state(c));
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  return 0;
}

} // namespace

extern "C" {

int hpf_iterate(hpf_ctx *c, uint32_t n_iters)
{
  if (!c) return fail(c, HPF_EINVAL, "null ctx");
  if (c->is_group) return group_iterate(c, n_iters, nullptr);
  CU(cudaSetDevice(c->cfg.device));
  TRY(check_ready(c));
  TRY(ensure_aux(c));
  if (n_iters == 0) return 0;
  const bool optimistic = c->nranks > 1 && !c->mg_exact;
  struct Saved { bool th_valid, be_valid, pr_prev_valid, th_global, a_dirty; uint64_t iterations; } sv;
  if (optimistic) { // keep what a re-run of this window needs
    sv.th_valid = c->th.derived_valid; sv.be_valid = c->be.derived_valid; sv.pr_prev_valid = c->pr_prev_valid;
    sv.th_global = c->th_colsum_global; sv.a_dirty = c->dense.a_dirty; sv.iterations = c->iterations;
    CU(cudaMemsetAsync(c->mg_fired, 0, sizeof(float), c->stream));
    TRY(snapshot_state(c, false));
  }
  TRY(run_window(c, n_iters));
  if (optimistic) {
    float fired = 0.f; // all-reduced: the same value on every rank, so every rank takes the same branch
    CU(cudaMemcpy(&fired, c->mg_fired, sizeof(float), cudaMemcpyDeviceToHost));
    if (fired != 0.f) {
      // an item-side nonzero took the exact fallback on some rank: its contribution sits in that rank's Tdirect_beta
      // only.  Re-run the window from the snapshot with the fallback buffers inside the all-reduce, and stay there.
      c->mg_exact = true;
      TRY(snapshot_state(c, true));
      // a sharded window consumed (and cleared) only each rank's own slice of the item side's fallback sums
      CU(cudaMemsetAsync(c->redblock2, 0, c->red2_count * sizeof(float), c->stream));
      // rate, E[v] and E[log v] are never written inside a window (only derive_kernel and hpf_set_state write them): if
      // they were current when the window began they still are -- and after hpf_set_state they are NOT functions of
      // the shape (src/gpbase.hh:324-340), so the fallback of the first iteration must read them, not recompute them
      c->th.derived_valid = sv.th_valid; c->be.derived_valid = sv.be_valid;
      c->pr_prev_valid = sv.pr_prev_valid; c->th_colsum_global = sv.th_global; c->iterations = sv.iterations;
      c->dense.a_dirty = true;
      const float first_ms = c->last_ms;
      TRY(run_window(c, n_iters));
      c->last_ms += first_ms;
    }
  }
  return 0;
}

int hpf_iterate_profiled(hpf_ctx *c, uint32_t n_iters, hpf_iter_profile *out)
{
  if (!c || !out) return fail(c, HPF_EINVAL, "null argument");
  memset(out, 0, sizeof *out);
  if (n_iters == 0) return 0;
  if (c->is_group) return group_iterate(c, n_iters, out);
  CU(cudaSetDevice(c->cfg.device));
  TRY(check_ready(c));
  TRY(ensure_aux(c));
  float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  decide_sharding(c);
  for (uint32_t it = 0; it < n_iters; ++it) {
    c->profiling = true;
    int rc = one_iteration(c);
    c->profiling = false;
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 8; ++i) {
      if (i == 3 && !c->dense.on) continue; // stage not run: its events were never recorded
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, c->pev[2 * i], c->pev[2 * i + 1]));
      acc[i] += ms;
    }
  }
  TRY(gather_beta_state(c));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  const float inv = 1.f / (float)n_iters;
  out->sweep_user_ms = (acc[0] + acc[3]) * inv; out->sweep_item_ms = acc[1] * inv; out->combine_ms = acc[2] * inv;
  out->update_theta_ms = acc[4] * inv; out->allreduce_ms = acc[5] * inv; out->update_beta_ms = acc[6] * inv;
  out->total_ms = acc[7] * inv;
  out->sweep_user_head_ms = acc[3] * inv;
  return 0;
}

int hpf_heldout_loglik(hpf_ctx *c, const uint32_t *u, const uint32_t *i, const uint8_t *y, uint64_t npairs, double *sum_ll)
{
  if (!c || !sum_ll) return fail(c, HPF_EINVAL, "null argument");
  *sum_ll = 0.0;
  if (npairs == 0) return 0;
  if (!u || !i || !y) return fail(c, HPF_EINVAL, "null pair arrays");
  if (c->is_group) return group_heldout(c, u, i, y, npairs, sum_ll);
  CU(cudaSetDevice(c->cfg.device));
  if (!c->th.have_state || !c->be.have_state) return fail(c, HPF_EINVAL, "state has not been set");
  for (uint64_t p = 0; p < npairs; ++p)
    if (u[p] >= c->th.R || i[p] >= c->be.R) return fail(c, HPF_EINVAL, "pair %llu out of range", (unsigned long long)p);
  TRY(ensure_derived(c, c->th));
  TRY(ensure_derived(c, c->be));
  // pairs go through the grow-only set-up arena: no allocation in the steady state
  TRY(arena_reserve(c, c->dev_arena, 2 * pad256(npairs * 4) + pad256(npairs) + 4096));
  uint32_t *du = c->dev_arena.get<uint32_t>(npairs), *di = c->dev_arena.get<uint32_t>(npairs);
  uint8_t *dy = c->dev_arena.get<uint8_t>(npairs);
  if (!du || !di || !dy) return fail(c, HPF_ENOMEM, "device arena too small");
  cudaError_t e = cudaMemcpyAsync(du, u, npairs * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(di, i, npairs * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, npairs, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    HeldoutArgs a;
    a.u = du; a.i = di; a.y = dy; a.npairs = npairs;
    a.Et = c->th.Ev; a.Eb = c->be.Ev;
    a.Etb = c->bias ? c->th.b_Ev : nullptr; a.Ebb = c->bias ? c->be.b_Ev : nullptr;
    a.K4 = c->K4; a.ld4 = c->ld / 4; a.binary = c->binary; a.logfact = c->logfact; a.block_sums = c->ll_blocks;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)c->sm_count * 8, (npairs * 8 + kSweepThreads - 1) / kSweepThreads);
    // fixed mapping: 8 lanes per pair, up to 8 float4 per lane covers K <= 256; wider K uses 32 lanes
    if (c->K4 <= 64) heldout_kernel<8, 8><<<grid, kSweepThreads, 0, c->stream>>>(a);
    else heldout_kernel<32, 8><<<grid, kSweepThreads, 0, c->stream>>>(a);
    sum_blocks_kernel<<<1, 32, 0, c->stream>>>(c->ll_blocks, grid, c->ll_out);
    c->launches += 2;
    e = cudaMemcpyAsync(sum_ll, c->ll_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  }
  if (e != cudaSuccess) return fail(c, HPF_ECUDA, "hpf_heldout_loglik: %s", cudaGetErrorString(e));
  return 0;
}

// HGAPRec::logl(), hgaprec.cc:2160-2255 (kernels and the algebra: hpf_elbo.cuh)
int hpf_elbo(hpf_ctx *c, double *elbo_out)
{
  if (!c || !elbo_out) return fail(c, HPF_EINVAL, "null argument");
  *elbo_out = 0.0;
  if (c->is_group) return group_elbo(c, elbo_out);
  if (!c->logl) return fail(c, HPF_EINVAL, "hpf_elbo needs a ctx created with HPF_LOGL");
  TRY(check_ready(c));
  if (c->hier && !c->pr_prev_valid)
    return fail(c, HPF_EINVAL, "hpf_elbo with HPF_HIER needs at least one hpf_iterate since THETARATE/BETARATE were set "
                               "(logl() uses the rate priors of the last iteration, gpbase.hh:163-173)");
  CU(cudaSetDevice(c->cfg.device));
  TRY(ensure_derived(c, c->th));
  TRY(ensure_derived(c, c->be));
  const uint32_t per = (uint32_t)c->sm_count * 8u; // partial sums per launch
  CU(cudaMemsetAsync(c->elbo_blocks, 0, sizeof(double) * kElboLaunches * per, c->stream));
  uint32_t slot = 0;
  auto grid_for = [&](uint64_t threads) { return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(per, (threads + elbo::kThreads - 1) / elbo::kThreads)); };
  // the nonzero terms of this ctx's users
  if (c->nnz > 0) {
    elbo::NnzArgs a;
    a.row_ptr = c->csr_rowptr; a.idx = c->csr_idx; a.y = c->csr_has_y ? c->csr_y : nullptr;
    a.n = c->th.R; a.K = c->K; a.ld = c->ld;
    a.ElogT = c->th.Elog; a.ElogB = c->be.Elog; a.EvT = c->th.Ev; a.EvB = c->be.Ev;
    a.ElogbT = c->bias ? c->th.b_Elog : nullptr; a.ElogbB = c->bias ? c->be.b_Elog : nullptr;
    a.EvbT = c->bias ? c->th.b_Ev : nullptr; a.EvbB = c->bias ? c->be.b_Ev : nullptr;
    a.block_sums = c->elbo_blocks + (size_t)slot * per;
    elbo::nnz_kernel<<<grid_for((uint64_t)c->th.R * 32), elbo::kThreads, 0, c->stream>>>(a);
    c->launches++;
  }
  ++slot;
  // Gamma terms: the user side always, the item side on rank 0 only (it is replicated), so that the sum of
  // hpf_elbo over the ranks is the ELBO of the whole problem
  for (int side = 0; side < 2; ++side) {
    Side &s = side == 0 ? c->th : c->be;
    const bool mine = side == 0 || c->rank == 0;
    if (mine && s.R > 0) {
      elbo::GammaArgs g;
      g.R = s.R; g.K = c->K; g.ld = c->ld;
      g.shape = s.shape; g.rate = s.rate; g.Ev = s.Ev; g.Elog = s.Elog;
      g.rate_is_vector = c->hier ? 0 : 1;
      g.sprior = s.prior_shape; g.rprior = s.prior_rate; g.lg_sprior = lgamma(s.prior_shape);
      g.pri_shape = c->hier ? s.pr_shape_prev : nullptr; g.pri_rate = c->hier ? s.pr_rate_prev : nullptr;
      g.block_sums = c->elbo_blocks + (size_t)(slot + side) * per;
      elbo::gamma_matrix_kernel<<<grid_for((uint64_t)s.R * c->K), elbo::kThreads, 0, c->stream>>>(g);
      c->launches++;
      if (c->hier) { // xi / eta (GPArray)
        elbo::gamma_array_kernel<<<grid_for(s.R), elbo::kThreads, 0, c->stream>>>(
            s.pr_shape, s.pr_rate, s.R, s.pr_prior_shape, s.pr_prior_rate, lgamma(s.pr_prior_shape),
            c->elbo_blocks + (size_t)(slot + 2 + side) * per);
        c->launches++;
      }
      if (c->bias) { // rows x 1 GPMatrix that never saw set_prior_rate: constant prior
        elbo::GammaArgs b;
        b.R = s.R; b.K = 1; b.ld = 1;
        b.shape = s.b_shape; b.rate = s.b_rate; b.Ev = s.b_Ev; b.Elog = s.b_Elog;
        b.rate_is_vector = 0;
        b.sprior = s.bias_prior_shape; b.rprior = s.bias_prior_rate; b.lg_sprior = lgamma(s.bias_prior_shape);
        b.pri_shape = nullptr; b.pri_rate = nullptr;
        b.block_sums = c->elbo_blocks + (size_t)(slot + 4 + side) * per;
        elbo::gamma_matrix_kernel<<<grid_for(s.R), elbo::kThreads, 0, c->stream>>>(b);
        c->launches++;
      }
    }
  }
  CU(cudaGetLastError());
  std::vector<double> part((size_t)kElboLaunches * per);
  CU(cudaMemcpyAsync(part.data(), c->elbo_blocks, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  for (double v : part) tot += v; // fixed order: launch by launch, block by block
  *elbo_out = tot;
  return 0;
}

int hpf_topn(hpf_ctx *c, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx,
             uint32_t topn, uint32_t *items_out, float *scores_out)
{
  if (!c) return fail(c, HPF_EINVAL, "null ctx");
  if (nu == 0) return 0;
  if (!users || !items_out || !scores_out) return fail(c, HPF_EINVAL, "null argument");
  if (c->is_group) return group_topn(c, users, nu, excl_ptr, excl_idx, topn, items_out, scores_out);
  if (topn == 0 || topn > (uint32_t)topk::kMaxTopN) return fail(c, HPF_EINVAL, "topn=%u out of range [1,%d]", topn, topk::kMaxTopN);
  if (!c->th.have_state || !c->be.have_state) return fail(c, HPF_EINVAL, "state has not been set");
  if (c->bias && (!c->th.have_bias || !c->be.have_bias)) return fail(c, HPF_EINVAL, "bias state has not been set");
  CU(cudaSetDevice(c->cfg.device));
  TRY(ensure_derived(c, c->th));
  TRY(ensure_derived(c, c->be));
  const uint32_t n = c->th.R, m = c->be.R;
  for (uint32_t a = 0; a < nu; ++a)
    if (users[a] >= n) return fail(c, HPF_EINVAL, "users[%u]=%u >= n_users=%u", a, users[a], n);
  const uint64_t nex = excl_ptr ? excl_ptr[nu] : 0;
  if (excl_ptr) {
    if (excl_ptr[0] != 0) return fail(c, HPF_EINVAL, "excl_ptr[0] must be 0");
    for (uint32_t a = 0; a < nu; ++a)
      if (excl_ptr[a + 1] < excl_ptr[a]) return fail(c, HPF_EINVAL, "excl_ptr not monotone at %u", a);
    if (nex > 0 && !excl_idx) return fail(c, HPF_EINVAL, "excl_idx is null");
  }
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensormap_encoder();
  if (!encode) return fail(c, HPF_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const uint32_t Kext = c->K + (c->bias ? 2u : 0u);
  const uint32_t Kpad = (Kext + topk::kBlockK - 1) / topk::kBlockK * topk::kBlockK;
  const uint32_t m_pad = (m + topk::kTileN - 1) / topk::kTileN * topk::kTileN;
  const uint32_t chunk_users = 2048u * topk::kTileM; // bounds the candidate scratch to 1 GiB per launch
  const uint32_t nu_cap = std::min<uint64_t>(((uint64_t)nu + topk::kTileM - 1) / topk::kTileM * topk::kTileM, chunk_users);

  Scratch tmp;
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  uint32_t *d_users = nullptr, *d_exidx = nullptr, *d_rowof = nullptr, *d_items = nullptr;
  uint64_t *d_exptr = nullptr, *d_key = nullptr, *d_key2 = nullptr;
  unsigned long long *d_cand = nullptr;
  float *d_scores = nullptr;
  CU(tmp.get(&a_hi, (size_t)nu_cap * Kpad)); CU(tmp.get(&a_lo, (size_t)nu_cap * Kpad));
  CU(tmp.get(&b_hi, (size_t)m_pad * Kpad)); CU(tmp.get(&b_lo, (size_t)m_pad * Kpad));
  CU(tmp.get(&d_users, nu)); CU(tmp.get(&d_exptr, (size_t)nu + 1));
  CU(tmp.get(&d_cand, (size_t)nu_cap * topk::kCap));
  CU(tmp.get(&d_items, (size_t)nu_cap * topn)); CU(tmp.get(&d_scores, (size_t)nu_cap * topn));
  CU(cudaMemcpyAsync(d_users, users, (size_t)nu * 4, cudaMemcpyHostToDevice, c->stream));
  if (excl_ptr) CU(cudaMemcpyAsync(d_exptr, excl_ptr, ((size_t)nu + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  else CU(cudaMemsetAsync(d_exptr, 0, ((size_t)nu + 1) * 8, c->stream));
  if (nex > 0) {
    // exclusion lists sorted by (user position, item) so the epilogue can binary-search them
    CU(tmp.get(&d_exidx, nex)); CU(tmp.get(&d_rowof, nex)); CU(tmp.get(&d_key, nex)); CU(tmp.get(&d_key2, nex));
    CU(cudaMemcpyAsync(d_exidx, excl_idx, nex * 4, cudaMemcpyHostToDevice, c->stream));
    const unsigned nb = (unsigned)((nex + 255) / 256);
    CU(cudaMemsetAsync(c->scratch_u32, 0, 4, c->stream));
    check_range_kernel<<<nb, 256, 0, c->stream>>>(d_exidx, nex, m, c->scratch_u32);
    expand_rows_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(d_exptr, nu, nex, d_rowof);
    topk::excl_key_kernel<<<nb, 256, 0, c->stream>>>(d_rowof, d_exidx, nex, d_key);
    c->launches += 3;
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, c->scratch_u32, 4, cudaMemcpyDeviceToHost, c->stream));
    size_t tmp_bytes = 0;
    void *d_tmp = nullptr;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)nex, 0, 64, c->stream);
    CU(tmp.get((char **)&d_tmp, tmp_bytes));
    CU(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, (const uint64_t *)d_key, d_key2, (int64_t)nex, 0, 32 + bits_for(nu), c->stream));
    topk::excl_unkey_kernel<<<nb, 256, 0, c->stream>>>(d_key2, nex, d_exidx);
    c->launches += 1;
    CU(cudaStreamSynchronize(c->stream));
    if (bad != 0) return fail(c, HPF_EINVAL, "excl_idx holds item %u >= n_items=%u", bad, m);
  }
  // item-side operands: hi / lo split of E[beta] (+ {1, E[betabias]})
  topk::split_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->be.Ev, c->ld, c->K, c->bias ? c->be.b_Ev : nullptr, 0, nullptr, m,
                                                         m_pad, Kpad, b_hi, b_lo);
  c->launches++;
  CU(cudaGetLastError());
  auto make_map = [&](CUtensorMap *map, void *ptr, uint64_t rows, uint32_t box_rows) -> bool {
    cuuint64_t dims[2] = { Kpad, rows };
    cuuint64_t strides[1] = { (cuuint64_t)Kpad * 2 };
    cuuint32_t box[2] = { (cuuint32_t)topk::kBlockK, box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;
  if (!make_map(&map_b_hi, b_hi, m_pad, topk::kTileN) || !make_map(&map_b_lo, b_lo, m_pad, topk::kTileN))
    return fail(c, HPF_ECUDA, "cuTensorMapEncodeTiled failed for the item operand");
  CU(cudaFuncSetAttribute(topk::topn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topk::kSmemBytes));
  c->last_topn_ms = 0.f;
  for (uint32_t u0 = 0; u0 < nu; u0 += chunk_users) {
    const uint32_t cu = std::min(chunk_users, nu - u0);
    const uint32_t cu_pad = (cu + topk::kTileM - 1) / topk::kTileM * topk::kTileM;
    topk::split_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->th.Ev, c->ld, c->K, c->bias ? c->th.b_Ev : nullptr, 1, d_users + u0,
                                                           cu, cu_pad, Kpad, a_hi, a_lo);
    if (!make_map(&map_a_hi, a_hi, cu_pad, topk::kTileM) || !make_map(&map_a_lo, a_lo, cu_pad, topk::kTileM))
      return fail(c, HPF_ECUDA, "cuTensorMapEncodeTiled failed for the user operand");
    topk::TopnArgs a;
    a.nu = cu; a.m = m; a.nkb = Kpad / topk::kBlockK; a.ntiles_n = m_pad / topk::kTileN; a.topn = topn;
    a.excl_ptr = d_exptr + u0; a.excl_sorted = d_exidx; a.cand = d_cand; a.items_out = d_items; a.scores_out = d_scores;
    CU(cudaEventRecord(c->ev0, c->stream));
    topk::topn_kernel<<<cu_pad / topk::kTileM, topk::kThreads, topk::kSmemBytes, c->stream>>>(map_a_hi, map_a_lo, map_b_hi, map_b_lo, a);
    CU(cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(items_out + (size_t)u0 * topn, d_items, (size_t)cu * topn * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(scores_out + (size_t)u0 * topn, d_scores, (size_t)cu * topn * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_topn_ms += ms;
  }
  return 0;
}

int hpf_item_ranks(hpf_ctx *c, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx,
                   const uint64_t *query_ptr, const uint32_t *query_idx, uint32_t *rank_out, float *score_out)
{
  if (!c) return fail(c, HPF_EINVAL, "null ctx");
  if (nu == 0) return 0;
  if (!users || !query_ptr || !rank_out || !score_out) return fail(c, HPF_EINVAL, "null argument");
  if (c->is_group) return group_item_ranks(c, users, nu, excl_ptr, excl_idx, query_ptr, query_idx, rank_out, score_out);
  if (!c->th.have_state || !c->be.have_state) return fail(c, HPF_EINVAL, "state has not been set");
  if (c->bias && (!c->th.have_bias || !c->be.have_bias)) return fail(c, HPF_EINVAL, "bias state has not been set");
  CU(cudaSetDevice(c->cfg.device));
  TRY(ensure_derived(c, c->th));
  TRY(ensure_derived(c, c->be));
  const uint32_t n = c->th.R, m = c->be.R;
  for (uint32_t a = 0; a < nu; ++a)
    if (users[a] >= n) return fail(c, HPF_EINVAL, "users[%u]=%u >= n_users=%u", a, users[a], n);
  const uint64_t nq = query_ptr[nu], nex = excl_ptr ? excl_ptr[nu] : 0;
  if (query_ptr[0] != 0 || (excl_ptr && excl_ptr[0] != 0)) return fail(c, HPF_EINVAL, "pointer arrays must start at 0");
  for (uint32_t a = 0; a < nu; ++a)
    if (query_ptr[a + 1] < query_ptr[a] || (excl_ptr && excl_ptr[a + 1] < excl_ptr[a])) return fail(c, HPF_EINVAL, "pointer array not monotone at %u", a);
  if (nq == 0) return 0;
  if (!query_idx || (nex > 0 && !excl_idx)) return fail(c, HPF_EINVAL, "null index array");
  for (uint64_t q = 0; q < nq; ++q)
    if (query_idx[q] >= m) return fail(c, HPF_EINVAL, "query_idx[%llu]=%u >= n_items=%u", (unsigned long long)q, query_idx[q], m);
  Scratch tmp;
  uint32_t *d_users = nullptr, *d_exidx = nullptr, *d_rowof = nullptr, *d_qidx = nullptr, *d_rank = nullptr;
  uint64_t *d_exptr = nullptr, *d_qptr = nullptr, *d_key = nullptr, *d_key2 = nullptr;
  unsigned long long *d_qkey = nullptr;
  float *d_score = nullptr;
  CU(tmp.get(&d_users, nu)); CU(tmp.get(&d_exptr, (size_t)nu + 1)); CU(tmp.get(&d_qptr, (size_t)nu + 1));
  CU(tmp.get(&d_qidx, nq)); CU(tmp.get(&d_rank, nq)); CU(tmp.get(&d_score, nq)); CU(tmp.get(&d_qkey, nq));
  CU(cudaMemcpyAsync(d_users, users, (size_t)nu * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_qptr, query_ptr, ((size_t)nu + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_qidx, query_idx, nq * 4, cudaMemcpyHostToDevice, c->stream));
  if (excl_ptr) CU(cudaMemcpyAsync(d_exptr, excl_ptr, ((size_t)nu + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  else CU(cudaMemsetAsync(d_exptr, 0, ((size_t)nu + 1) * 8, c->stream));
  if (nex > 0) { // exclusion lists sorted by (user position, item), as in hpf_topn
    CU(tmp.get(&d_exidx, nex)); CU(tmp.get(&d_rowof, nex)); CU(tmp.get(&d_key, nex)); CU(tmp.get(&d_key2, nex));
    CU(cudaMemcpyAsync(d_exidx, excl_idx, nex * 4, cudaMemcpyHostToDevice, c->stream));
    const unsigned nb = (unsigned)((nex + 255) / 256);
    CU(cudaMemsetAsync(c->scratch_u32, 0, 4, c->stream));
    check_range_kernel<<<nb, 256, 0, c->stream>>>(d_exidx, nex, m, c->scratch_u32);
    expand_rows_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(d_exptr, nu, nex, d_rowof);
    topk::excl_key_kernel<<<nb, 256, 0, c->stream>>>(d_rowof, d_exidx, nex, d_key);
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, c->scratch_u32, 4, cudaMemcpyDeviceToHost, c->stream));
    size_t tmp_bytes = 0;
    void *d_tmp = nullptr;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)nex, 0, 64, c->stream);
    CU(tmp.get((char **)&d_tmp, tmp_bytes));
    CU(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, (const uint64_t *)d_key, d_key2, (int64_t)nex, 0, 32 + bits_for(nu), c->stream));
    topk::excl_unkey_kernel<<<nb, 256, 0, c->stream>>>(d_key2, nex, d_exidx);
    c->launches += 4;
    CU(cudaStreamSynchronize(c->stream));
    if (bad != 0) return fail(c, HPF_EINVAL, "excl_idx holds item %u >= n_items=%u", bad, m);
  }
  // operands of the tensor-core scoring pass, as in hpf_topn: hi / lo split of E[beta] (+ {1, E[betabias]}) and of the
  // listed users' E[theta] rows (+ {E[thetabias], 1})
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensormap_encoder();
  if (!encode) return fail(c, HPF_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const uint32_t Kext = c->K + (c->bias ? 2u : 0u);
  const uint32_t Kpad = (Kext + topk::kBlockK - 1) / topk::kBlockK * topk::kBlockK;
  const uint32_t m_pad = (m + topk::kTileN - 1) / topk::kTileN * topk::kTileN;
  const uint32_t nu_pad = (nu + topk::kTileM - 1) / topk::kTileM * topk::kTileM;
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  CU(tmp.get(&a_hi, (size_t)nu_pad * Kpad)); CU(tmp.get(&a_lo, (size_t)nu_pad * Kpad));
  CU(tmp.get(&b_hi, (size_t)m_pad * Kpad)); CU(tmp.get(&b_lo, (size_t)m_pad * Kpad));
  topk::split_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->be.Ev, c->ld, c->K, c->bias ? c->be.b_Ev : nullptr, 0, nullptr, m, m_pad,
                                                         Kpad, b_hi, b_lo);
  topk::split_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->th.Ev, c->ld, c->K, c->bias ? c->th.b_Ev : nullptr, 1, d_users, nu, nu_pad,
                                                         Kpad, a_hi, a_lo);
  c->launches += 2;
  auto make_map = [&](CUtensorMap *map, void *ptr, uint64_t rows, uint32_t box_rows) -> bool {
    cuuint64_t dims[2] = { Kpad, rows };
    cuuint64_t strides[1] = { (cuuint64_t)Kpad * 2 };
    cuuint32_t box[2] = { (cuuint32_t)topk::kBlockK, box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;
  if (!make_map(&map_a_hi, a_hi, nu_pad, topk::kTileM) || !make_map(&map_a_lo, a_lo, nu_pad, topk::kTileM) ||
      !make_map(&map_b_hi, b_hi, m_pad, topk::kTileN) || !make_map(&map_b_lo, b_lo, m_pad, topk::kTileN))
    return fail(c, HPF_ECUDA, "cuTensorMapEncodeTiled failed for the ranking operands");
  topk::RankArgs a;
  a.nu = nu; a.m = m; a.K = c->K; a.ld = c->ld;
  a.nkb = Kpad / topk::kBlockK; a.ntiles_n = m_pad / topk::kTileN;
  a.users = d_users; a.Et = c->th.Ev; a.Eb = c->be.Ev;
  a.Etb = c->bias ? c->th.b_Ev : nullptr; a.Ebb = c->bias ? c->be.b_Ev : nullptr;
  a.excl_ptr = d_exptr; a.excl_sorted = d_exidx; a.q_ptr = d_qptr; a.q_idx = d_qidx; a.q_key = d_qkey;
  a.rank_out = d_rank; a.score_out = d_score;
  // the queries' own keys (fp32 dot products: the scores handed back), then the counting pass on the tensor cores
  topk::rank_prep_kernel<<<(nu + topk::kRankWarps - 1) / topk::kRankWarps, topk::kRankWarps * 32, (size_t)topk::kRankWarps * c->K * 4, c->stream>>>(a);
  CU(cudaFuncSetAttribute(topk::rank_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topk::kRankSmemBytes));
  CU(cudaEventRecord(c->ev0, c->stream));
  topk::rank_mma_kernel<<<nu_pad / topk::kTileM, topk::kRankThreads, topk::kRankSmemBytes, c->stream>>>(map_a_hi, map_a_lo, map_b_hi, map_b_lo, a);
  CU(cudaEventRecord(c->ev1, c->stream));
  c->launches += 2;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(rank_out, d_rank, nq * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(score_out, d_score, nq * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->last_topn_ms, c->ev0, c->ev1));
  return 0;
}

int hpf_partition_users(const uint64_t *row_ptr, uint32_t n_users, uint32_t nranks, uint32_t *first)
{
  if (!row_ptr || !first || nranks == 0) return fail(nullptr, HPF_EINVAL, "bad partition arguments");
  const uint64_t nnz = row_ptr[n_users];
  first[0] = 0;
  for (uint32_t r = 1; r < nranks; ++r) {
    // smallest u whose prefix holds at least r/nranks of the nonzeros, never before the previous cut
    const uint64_t want = (uint64_t)(((unsigned __int128)nnz * r) / nranks);
    const uint64_t *p = std::lower_bound(row_ptr + first[r - 1], row_ptr + n_users + 1, want);
    first[r] = (uint32_t)std::min<uint64_t>((uint64_t)(p - row_ptr), n_users);
  }
  first[nranks] = n_users;
  return 0;
}

int hpf_comm_unique_id(void *id_out, size_t id_bytes)
{
  std::string err;
  if (!id_out || id_bytes < sizeof(ncclUniqueId)) return fail(nullptr, HPF_EINVAL, "id buffer must hold %zu bytes", sizeof(ncclUniqueId));
  if (!g_nccl.load(err)) return fail(nullptr, HPF_ENCCL, "%s", err.c_str());
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return fail(nullptr, HPF_ENCCL, "ncclGetUniqueId failed");
  memcpy(id_out, &id, sizeof id);
  return 0;
}

int hpf_comm_init(hpf_ctx *c, int rank, int nranks, const void *id, size_t id_bytes)
{
  if (!c || !id || id_bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks)
    return fail(c, HPF_EINVAL, "bad communicator arguments");
  if (c->is_group) return fail(c, HPF_EINVAL, "a ctx with n_devices > 1 owns its communicator (one box); hpf_comm_init is for one-device ctxs");
  CU(cudaSetDevice(c->cfg.device));
  std::string err;
  if (!g_nccl.load(err)) return fail(c, HPF_ENCCL, "%s", err.c_str());
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  int rc = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
  if (rc != ncclSuccess) return fail(c, HPF_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  c->rank = rank; c->nranks = nranks;
  // ratings planned before the communicator existed have one chunk: still correct, the all-reduce just overlaps less
  return comm_streams(c);
}

int hpf_get_stats(const hpf_ctx *c, hpf_stats *out)
{
  if (!c || !out) return HPF_EINVAL;
  memset(out, 0, sizeof *out);
  if (c->is_group) return group_get_stats(c, out);
  out->kernel_launches = c->launches;
  out->iterations = c->iterations;
  unsigned long long sc = 0;
  cudaSetDevice(c->cfg.device);
  cudaMemcpy(&sc, c->slow_count, sizeof sc, cudaMemcpyDeviceToHost);
  out->slow_path_nnz = sc;
  out->nnz = c->nnz;
  out->device_bytes = c->device_bytes;
  out->last_iterate_ms = c->last_ms;
  out->sweep_group = c->sweep_g;
  out->sweep_vec = c->sweep_v;
  out->last_topn_ms = c->last_topn_ms;
  out->user_l2_tiles = c->th_tiles;
  out->item_l2_tiles = c->be_tiles;
  out->head_nnz = c->dense.on ? c->dense.head_nnz : 0;
  out->item_chunks = c->be.wl.nchunks;
  out->mg_exact = c->mg_exact ? 1u : 0u;
  out->n_devices = 1;
  out->beta_sharded = c->last_sharded ? 1u : 0u;
  return 0;
}

} // extern "C"

// =============================================================================
// group ctx: ONE ctx, several GPUs of one box (hpf_config.n_devices > 1).  The reference is one process
// (SURVEY.md 8b "threading"); this is the form its command line uses (`hgaprec -gpus N`).  The group owns one
// ordinary one-device ctx per GPU ("kid"), shards the users over them with hpf_partition_users at the first
// hpf_set_ratings_csr, joins them with ncclCommInitAll and drives each kid from its own worker thread inside a call
// -- exactly the configuration the 2-GPU tests exercise with explicit threads.  Every argument and result of the ABI
// keeps GLOBAL user numbers.
// =============================================================================
namespace {

template <class F> int for_each_kid(hpf_ctx *g, F f)
{
  const size_t nk = g->kids.size();
  std::vector<int> rc(nk, 0);
  std::vector<std::thread> th;
  for (size_t i = 0; i < nk; ++i) th.emplace_back([&, i]() { rc[i] = f(i, g->kids[i]); });
  for (auto &t : th) t.join();
  for (size_t i = 0; i < nk; ++i)
    if (rc[i] != 0) {
      g->err = "device " + std::to_string(g->kids[i]->cfg.device) + ": " + g->kids[i]->err;
      return rc[i];
    }
  return 0;
}

// one ordinary ctx per device over the user ranges b[0..nd], joined by ncclCommInitAll
int group_make_kids(hpf_ctx *g, const std::vector<uint32_t> &b)
{
  const uint32_t nd = g->cfg.n_devices;
  for (uint32_t i = 0; i < nd; ++i)
    if (b[i + 1] <= b[i]) return fail(g, HPF_EINVAL, "the user partition leaves device %u without users; use fewer devices", i);
  std::vector<ncclComm_t> comms(nd);
  std::vector<int> devs(g->cfg.devices, g->cfg.devices + nd);
  const int rc = g_nccl.CommInitAll(comms.data(), (int)nd, devs.data());
  if (rc != ncclSuccess) return fail(g, HPF_ENCCL, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  for (uint32_t i = 0; i < nd; ++i) {
    hpf_config kc = g->cfg;
    kc.n_devices = 0; kc.device = devs[i];
    kc.n_users = b[i + 1] - b[i];
    kc.n_users_global = g->cfg.n_users_global ? g->cfg.n_users_global : g->cfg.n_users;
    hpf_ctx *k = nullptr;
    const int krc = hpf_create(&kc, &k);
    if (krc != 0) {
      g->err = "device " + std::to_string(devs[i]) + ": " + g_create_error;
      for (hpf_ctx *q : g->kids) hpf_destroy(q);
      g->kids.clear();
      for (uint32_t j = i; j < nd; ++j) g_nccl.CommDestroy(comms[j]);
      return krc;
    }
    k->comm = comms[i]; k->rank = (int)i; k->nranks = (int)nd;
    g->kids.push_back(k);
  }
  g->bounds = b;
  for (hpf_ctx *k : g->kids) {
    cudaSetDevice(k->cfg.device);
    TRY(comm_streams(k));
  }
  return 0;
}

// calls that can come before any ratings (hpf_set_state ahead of hpf_topn in `-gen-ranking`): equal user counts
int group_need_kids(hpf_ctx *g)
{
  if (!g->kids.empty()) return 0;
  const uint32_t nd = g->cfg.n_devices;
  std::vector<uint32_t> b(nd + 1);
  for (uint32_t i = 0; i <= nd; ++i) b[i] = (uint32_t)(((uint64_t)g->cfg.n_users * i) / nd);
  return group_make_kids(g, b);
}

uint32_t owner_of(const hpf_ctx *g, uint32_t user)
{
  return (uint32_t)(std::upper_bound(g->bounds.begin(), g->bounds.end(), user) - g->bounds.begin()) - 1u;
}

int group_create(const hpf_config *cfg, hpf_ctx **out)
{
  hpf_ctx *c = nullptr;
  if (cfg->n_devices > HPF_MAX_DEVICES) return fail(c, HPF_EINVAL, "n_devices=%u > %d", cfg->n_devices, HPF_MAX_DEVICES);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(c, HPF_ENODEVICE, "no CUDA device available (libhpf_b200 has no CPU path)");
  for (uint32_t i = 0; i < cfg->n_devices; ++i) {
    if (cfg->devices[i] < 0 || cfg->devices[i] >= ndev) return fail(c, HPF_EINVAL, "devices[%u]=%d not in [0,%d)", i, cfg->devices[i], ndev);
    for (uint32_t j = 0; j < i; ++j)
      if (cfg->devices[j] == cfg->devices[i]) return fail(c, HPF_EINVAL, "device %d listed twice", cfg->devices[i]);
  }
  if (cfg->n_users < cfg->n_devices) return fail(c, HPF_EINVAL, "fewer users (%u) than devices (%u)", cfg->n_users, cfg->n_devices);
  std::string err;
  if (!g_nccl.load(err)) return fail(c, HPF_ENCCL, "%s", err.c_str());
  if (!g_nccl.CommInitAll) return fail(c, HPF_ENCCL, "libnccl lacks ncclCommInitAll");
  hpf_ctx *g = new hpf_ctx();
  g->cfg = *cfg;
  g->is_group = true;
  g->K = cfg->k;
  g->hier = cfg->flags & HPF_HIER; g->bias = cfg->flags & HPF_BIAS;
  *out = g;
  return 0;
}

int group_set_ratings(hpf_ctx *g, const uint64_t *row_ptr, const uint32_t *col_idx, const uint8_t *y)
{
  const uint32_t n = g->cfg.n_users, nd = g->cfg.n_devices;
  if (g->kids.empty()) { // the first ratings fix the partition: contiguous user ranges balanced by nonzeros
    std::vector<uint32_t> b(nd + 1);
    TRY(hpf_partition_users(row_ptr, n, nd, b.data()));
    TRY(group_make_kids(g, b));
  }
  // every kid gets its slice of the CSR (row pointer rebased), all uploads and device set-up side by side
  return for_each_kid(g, [&](size_t i, hpf_ctx *k) {
    const uint32_t lo = g->bounds[i], hi = g->bounds[i + 1];
    std::vector<uint64_t> rp((size_t)hi - lo + 1);
    for (uint32_t u = lo; u <= hi; ++u) rp[u - lo] = row_ptr[u] - row_ptr[lo];
    return hpf_set_ratings_csr(k, rp.data(), col_idx ? col_idx + row_ptr[lo] : nullptr, y ? y + row_ptr[lo] : nullptr);
  });
}

bool user_side(int which) { return which == HPF_THETA || which == HPF_THETARATE || which == HPF_THETABIAS; }

int group_set_state(hpf_ctx *g, int which, const double *shape, const double *rate, const double *Ev, const double *Elogv)
{
  TRY(group_need_kids(g));
  const size_t K = g->K;
  return for_each_kid(g, [&](size_t i, hpf_ctx *k) {
    if (!user_side(which)) return hpf_set_state(k, which, shape, rate, Ev, Elogv); // replicated
    const size_t lo = g->bounds[i];
    const size_t w = which == HPF_THETA ? K : 1;                                    // doubles per row
    const size_t wr = which == HPF_THETA && !g->hier ? 0 : w;                        // GR rate: one K-vector for all rows
    auto at = [&](const double *p, size_t width) { return p ? p + lo * width : nullptr; };
    return hpf_set_state(k, which, at(shape, w), at(rate, wr), at(Ev, w), at(Elogv, w));
  });
}

int group_get_state(hpf_ctx *g, int which, double *shape, double *rate, double *Ev, double *Elogv)
{
  TRY(group_need_kids(g));
  if (!user_side(which)) { // replicated, bitwise identical on every device
    const int rc = hpf_get_state(g->kids[0], which, shape, rate, Ev, Elogv);
    if (rc) g->err = g->kids[0]->err;
    return rc;
  }
  const size_t K = g->K;
  return for_each_kid(g, [&](size_t i, hpf_ctx *k) {
    const size_t lo = g->bounds[i];
    const size_t w = which == HPF_THETA ? K : 1;
    const size_t wr = which == HPF_THETA && !g->hier ? 0 : w;
    auto at = [&](double *p, size_t width) { return p ? p + lo * width : nullptr; };
    // the GR rate vector is the same on every kid: only kid 0 writes it
    return hpf_get_state(k, which, at(shape, w), (wr == 0 && i != 0) ? nullptr : at(rate, wr), at(Ev, w), at(Elogv, w));
  });
}

int group_iterate(hpf_ctx *g, uint32_t n_iters, hpf_iter_profile *prof)
{
  TRY(group_need_kids(g));
  std::vector<hpf_iter_profile> pr(g->kids.size());
  TRY(for_each_kid(g, [&](size_t i, hpf_ctx *k) { return prof ? hpf_iterate_profiled(k, n_iters, &pr[i]) : hpf_iterate(k, n_iters); }));
  g->last_ms = 0.f;
  for (hpf_ctx *k : g->kids) g->last_ms = std::max(g->last_ms, k->last_ms);
  if (prof) { // per stage: the slowest device
    *prof = pr[0];
    for (auto &p : pr) {
      float *a = reinterpret_cast<float *>(prof);
      const float *b = reinterpret_cast<const float *>(&p);
      for (size_t q = 0; q < sizeof(hpf_iter_profile) / sizeof(float); ++q) a[q] = std::max(a[q], b[q]);
    }
  }
  return 0;
}

int group_heldout(hpf_ctx *g, const uint32_t *u, const uint32_t *i, const uint8_t *y, uint64_t npairs, double *sum_ll)
{
  TRY(group_need_kids(g));
  const size_t nk = g->kids.size();
  std::vector<std::vector<uint32_t>> ku(nk), ki(nk);
  std::vector<std::vector<uint8_t>> ky(nk);
  for (uint64_t p = 0; p < npairs; ++p) {
    if (u[p] >= g->cfg.n_users) return fail(g, HPF_EINVAL, "pair %llu out of range", (unsigned long long)p);
    const uint32_t o = owner_of(g, u[p]);
    ku[o].push_back(u[p] - g->bounds[o]); ki[o].push_back(i[p]); ky[o].push_back(y[p]);
  }
  std::vector<double> part(nk, 0.0);
  TRY(for_each_kid(g, [&](size_t q, hpf_ctx *k) { return hpf_heldout_loglik(k, ku[q].data(), ki[q].data(), ky[q].data(), ku[q].size(), &part[q]); }));
  double s = 0.0;
  for (double v : part) s += v; // device order: repeatable
  *sum_ll = s;
  return 0;
}

int group_elbo(hpf_ctx *g, double *out)
{
  TRY(group_need_kids(g));
  std::vector<double> part(g->kids.size(), 0.0);
  TRY(for_each_kid(g, [&](size_t q, hpf_ctx *k) { return hpf_elbo(k, &part[q]); }));
  double s = 0.0;
  for (double v : part) s += v; // a kid returns its users' part (+ the item-side terms on rank 0): they add up
  *out = s;
  return 0;
}

// users of a query routed to their owners: per kid the positions in the caller's arrays, local user numbers and the
// rebased pointer array(s)
struct Routed {
  std::vector<uint32_t> pos, users;
  std::vector<uint64_t> eptr, qptr;
  std::vector<uint32_t> eidx, qidx;
};
int route_users(hpf_ctx *g, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx,
                const uint64_t *query_ptr, const uint32_t *query_idx, std::vector<Routed> *out)
{
  out->assign(g->kids.size(), Routed());
  for (auto &r : *out) { r.eptr.push_back(0); r.qptr.push_back(0); }
  for (uint32_t a = 0; a < nu; ++a) {
    if (users[a] >= g->cfg.n_users) return fail(g, HPF_EINVAL, "users[%u]=%u >= n_users=%u", a, users[a], g->cfg.n_users);
    Routed &r = (*out)[owner_of(g, users[a])];
    r.pos.push_back(a);
    r.users.push_back(users[a] - g->bounds[owner_of(g, users[a])]);
    if (excl_ptr) {
      if (excl_ptr[a + 1] < excl_ptr[a]) return fail(g, HPF_EINVAL, "excl_ptr not monotone at %u", a);
      r.eidx.insert(r.eidx.end(), excl_idx + excl_ptr[a], excl_idx + excl_ptr[a + 1]);
    }
    r.eptr.push_back(r.eidx.size());
    if (query_ptr) {
      if (query_ptr[a + 1] < query_ptr[a]) return fail(g, HPF_EINVAL, "query_ptr not monotone at %u", a);
      r.qidx.insert(r.qidx.end(), query_idx + query_ptr[a], query_idx + query_ptr[a + 1]);
      r.qptr.push_back(r.qidx.size());
    }
  }
  return 0;
}

int group_topn(hpf_ctx *g, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx, uint32_t topn,
               uint32_t *items_out, float *scores_out)
{
  TRY(group_need_kids(g));
  if (excl_ptr && excl_ptr[nu] > 0 && !excl_idx) return fail(g, HPF_EINVAL, "excl_idx is null");
  std::vector<Routed> rt;
  TRY(route_users(g, users, nu, excl_ptr, excl_idx, nullptr, nullptr, &rt));
  TRY(for_each_kid(g, [&](size_t q, hpf_ctx *k) {
    Routed &r = rt[q];
    if (r.users.empty()) return 0;
    std::vector<uint32_t> items(r.users.size() * (size_t)topn);
    std::vector<float> scores(r.users.size() * (size_t)topn);
    const int rc = hpf_topn(k, r.users.data(), (uint32_t)r.users.size(), r.eptr.data(), r.eidx.data(), topn, items.data(), scores.data());
    if (rc) return rc;
    for (size_t a = 0; a < r.users.size(); ++a) {
      memcpy(items_out + (size_t)r.pos[a] * topn, items.data() + a * topn, sizeof(uint32_t) * topn);
      memcpy(scores_out + (size_t)r.pos[a] * topn, scores.data() + a * topn, sizeof(float) * topn);
    }
    return 0;
  }));
  g->last_topn_ms = 0.f;
  for (hpf_ctx *k : g->kids) g->last_topn_ms = std::max(g->last_topn_ms, k->last_topn_ms);
  return 0;
}

int group_item_ranks(hpf_ctx *g, const uint32_t *users, uint32_t nu, const uint64_t *excl_ptr, const uint32_t *excl_idx,
                     const uint64_t *query_ptr, const uint32_t *query_idx, uint32_t *rank_out, float *score_out)
{
  TRY(group_need_kids(g));
  if (query_ptr[nu] > 0 && !query_idx) return fail(g, HPF_EINVAL, "null index array");
  if (excl_ptr && excl_ptr[nu] > 0 && !excl_idx) return fail(g, HPF_EINVAL, "null index array");
  std::vector<Routed> rt;
  TRY(route_users(g, users, nu, excl_ptr, excl_idx, query_ptr, query_idx, &rt));
  return for_each_kid(g, [&](size_t q, hpf_ctx *k) {
    Routed &r = rt[q];
    if (r.users.empty() || r.qidx.empty()) return 0;
    std::vector<uint32_t> rank(r.qidx.size());
    std::vector<float> score(r.qidx.size());
    const int rc = hpf_item_ranks(k, r.users.data(), (uint32_t)r.users.size(), r.eptr.data(), r.eidx.data(), r.qptr.data(), r.qidx.data(),
                                  rank.data(), score.data());
    if (rc) return rc;
    for (size_t a = 0; a < r.users.size(); ++a) { // the caller's queries of user pos[a] start at query_ptr[pos[a]]
      const uint64_t dst = query_ptr[r.pos[a]], src = r.qptr[a], cnt = r.qptr[a + 1] - r.qptr[a];
      memcpy(rank_out + dst, rank.data() + src, sizeof(uint32_t) * cnt);
      memcpy(score_out + dst, score.data() + src, sizeof(float) * cnt);
    }
    return 0;
  });
}

int group_get_stats(const hpf_ctx *g, hpf_stats *out)
{
  out->n_devices = g->cfg.n_devices;
  out->last_iterate_ms = g->last_ms;
  out->last_topn_ms = g->last_topn_ms;
  for (size_t i = 0; i < g->kids.size(); ++i) {
    hpf_stats s;
    hpf_get_stats(g->kids[i], &s);
    out->kernel_launches += s.kernel_launches; out->slow_path_nnz += s.slow_path_nnz; out->nnz += s.nnz;
    out->device_bytes += s.device_bytes; out->head_nnz += s.head_nnz;
    if (i == 0) {
      out->iterations = s.iterations; out->sweep_group = s.sweep_group; out->sweep_vec = s.sweep_vec;
      out->user_l2_tiles = s.user_l2_tiles; out->item_l2_tiles = s.item_l2_tiles; out->item_chunks = s.item_chunks;
    }
    out->mg_exact |= s.mg_exact;
    out->beta_sharded |= s.beta_sharded;
  }
  return 0;
}

} // namespace

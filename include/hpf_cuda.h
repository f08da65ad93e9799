/* hpf_cuda.h -- C ABI of libhpf_b200.so: the B200 (sm_100a) CAVI engine for
 * (Hierarchical) Poisson Factorization that replaces the hot path of
 * premgopalan/hgaprec.
 *
 * The reference has no plugin / FFI interface: its boundary is
 * the three loop bodies HGAPRec::vb_hier / vb / vb_bias and the GPMatrix /
 * GPMatrixGR / GPArray accessors every downstream consumer reads.  This header
 * is the boundary a maintainer binds instead; INTEGRATION.md shows the stub.
 * Paths below are relative to the reference tree (/root/reference).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - every host matrix is contiguous row-major fp64 (the reference's
 *     Matrix = D2Array<double>, src/env.hh:23, flattened row by row).
 *   - the caller owns all host buffers; the library copies during the call and
 *     never keeps a host pointer.  The library owns all device memory, its
 *     stream and its NCCL communicator inside hpf_ctx.
 *   - return 0 on success, a negative HPF_E* code on failure; the message is
 *     available from hpf_last_error().  Nothing exits or throws across the ABI
 *     (the reference exit(-1)s / asserts, e.g. src/hgaprec.cc:42-45); the host
 *     wrapper maps nonzero to its own lerr()+exit(-1).
 *   - one host thread per ctx; a ctx is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device hpf_create fails.
 *   - multi-GPU comes in two forms with identical numerics: (a) hpf_config.n_devices > 1 -- ONE ctx in one
 *     process drives that many GPUs of the box: the library shards the users (hpf_partition_users), keeps one
 *     stream, one NCCL rank and one worker thread per device, and every call below takes GLOBAL user numbers
 *     (this is what the `hgaprec -gpus N` command line uses); (b) one ctx per GPU in separate processes or
 *     threads joined with hpf_comm_init (what bench.py uses under torchrun), where the caller shards.
 */
#ifndef HPF_CUDA_H
#define HPF_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPF_ABI_VERSION 2
#define HPF_MAX_DEVICES 16

#if defined(__GNUC__)
#define HPF_API __attribute__((visibility("default")))
#else
#define HPF_API
#endif

/* error codes */
#define HPF_OK          0
#define HPF_EINVAL     (-1) /* bad argument / call order                      */
#define HPF_ECUDA      (-2) /* CUDA runtime error                              */
#define HPF_ENOMEM     (-3) /* host or device allocation failed                */
#define HPF_ENCCL      (-4) /* NCCL missing or failed                          */
#define HPF_ENODEVICE  (-5) /* no usable CUDA device (there is no CPU path)    */

/* hpf_config.flags -- mirror the reference's CLI switches (src/main.cc:99-232) */
#define HPF_HIER    1u /* -hier        : vb_hier(), src/hgaprec.cc:1321-1436          */
#define HPF_BIAS    2u /* -bias        : K+2 wide phi, src/hgaprec.cc:222-239         */
#define HPF_BINARY  4u /* -binary-data : likelihood form, src/hgaprec.cc:1557-1558    */
#define HPF_JACOBI  8u /* -novb        : vb_bias() else-branch, src/hgaprec.cc:1276-1297
                          (ignored with HPF_HIER, as in the reference)                */
#define HPF_LOGL   16u /* -logl        : keep what hpf_elbo() needs (src/main.cc:128-130,
                          src/hgaprec.cc:1426-1427): a device copy of the CSR row pointer and,
                          with HPF_HIER, the xi / eta sets as they were before each iteration.
                          Changes no result of any other call.                         */

/* which parameter set a get/set call addresses; names follow HGAPRec's members
 * (src/hgaprec.hh:105-115).  With HPF_HIER, THETA/BETA are _htheta/_hbeta
 * (GPMatrix, rate n x k); without, _theta/_beta (GPMatrixGR, rate is a k-vector). */
enum hpf_param {
  HPF_THETA = 0,     /* n x k                                   */
  HPF_BETA = 1,      /* m x k                                   */
  HPF_THETARATE = 2, /* n     (GPArray _thetarate, hier only)   */
  HPF_BETARATE = 3,  /* m     (GPArray _betarate,  hier only)   */
  HPF_THETABIAS = 4, /* n x 1 (GPMatrix _thetabias, bias only)  */
  HPF_BETABIAS = 5   /* m x 1 (GPMatrix _betabias,  bias only)  */
};

typedef struct hpf_ctx hpf_ctx;

typedef struct hpf_config {
  uint32_t abi_version;   /* HPF_ABI_VERSION */
  uint32_t n_users;       /* users held by THIS ctx (the local shard; all users when
                             n_devices > 1)                                        */
  uint32_t n_items;       /* m: all items (the beta side is replicated per rank)   */
  uint32_t k;             /* factors                                               */
  uint32_t flags;         /* HPF_HIER | HPF_BIAS | HPF_BINARY | HPF_JACOBI          */
  int32_t  device;        /* CUDA device ordinal                                   */
  uint64_t n_users_global;/* all users over all ranks (0 == n_users); the item-bias
                             rate adds it, src/hgaprec.cc:1394                      */
  /* (shape, rate) priors.  The reference hard-codes 0.3 for every one of them
   * (src/hgaprec.cc:13-20; -a/-b/-c/-d are parsed but never used).             */
  double theta_shape, theta_rate;         /* _theta / _htheta                    */
  double beta_shape, beta_rate;           /* _beta / _hbeta                      */
  double thetarate_shape, thetarate_rate; /* _thetarate (xi)                     */
  double betarate_shape, betarate_rate;   /* _betarate (eta)                     */
  double thetabias_shape, thetabias_rate; /* _thetabias                          */
  double betabias_shape, betabias_rate;   /* _betabias                           */
  /* n_devices > 1: this ctx shards its users over devices[0 .. n_devices) of this box and sums the item
   * side over NCCL; `device` is ignored.  The user ranges are contiguous and fixed by the first call that
   * needs them: balanced by nonzeros when that is hpf_set_ratings_csr (training), by user count when the
   * state is uploaded without ratings (-gen-ranking).  0: one GPU, `device`; 1: one GPU, devices[0].   */
  uint32_t n_devices;
  int32_t  devices[HPF_MAX_DEVICES];
} hpf_config;

/* counters readable after any call (all monotone since hpf_create) */
typedef struct hpf_stats {
  uint64_t kernel_launches;  /* kernels of this library launched so far            */
  uint64_t iterations;       /* CAVI iterations completed                          */
  uint64_t slow_path_nnz;    /* nonzeros that took the exact log-domain fallback   */
  uint64_t nnz;              /* training nonzeros held                             */
  uint64_t device_bytes;     /* device memory currently allocated by the ctx       */
  float    last_iterate_ms;  /* device time of the last hpf_iterate (CUDA events
                                on the library's own stream)                       */
  uint32_t sweep_group;      /* lanes cooperating on one nonzero in the sweep      */
  uint32_t sweep_vec;        /* float4 values per lane                             */
  uint32_t user_l2_tiles;    /* L2 tiles of the item rows the user pass is grouped by (1: none) */
  uint32_t item_l2_tiles;    /* L2 tiles of the user rows the item pass is grouped by (1: none) */
  uint64_t head_nnz;         /* nonzeros of the most popular items served by the dense
                                tcgen05 head instead of the gather kernel (0: head off) */
  uint32_t item_chunks;      /* chunks of the item pass (multi-GPU: one all-reduce each,
                                overlapped with the sweeps that follow)             */
  uint32_t mg_exact;         /* 1: the item side's fallback buffers ride in the all-reduce
                                (set after a fallback fired on a multi-GPU run)     */
  float    last_topn_ms;     /* device time of the scoring + selection kernel(s) of
                                the last hpf_topn                                   */
  uint32_t n_devices;        /* GPUs this ctx drives                                */
  uint32_t beta_sharded;     /* 1: the last hpf_iterate ran the item side sharded over the ranks
                                (reduce-scatter, each rank updates m/N items, all-gather) */
} hpf_stats;

/* Fill *cfg with the reference's defaults (all priors 0.3, device 0). */
HPF_API void hpf_config_default(hpf_config *cfg);

HPF_API int hpf_create(const hpf_config *cfg, hpf_ctx **out);
HPF_API void hpf_destroy(hpf_ctx *ctx);

/* Last error text of this ctx (ctx == NULL: of the last failed hpf_create on
 * this thread).  Never NULL. */
HPF_API const char *hpf_last_error(const hpf_ctx *ctx);

/* Training matrix as CSR over the ctx's users, in the order the reference's
 * loop walks it: row u lists Ratings::get_movies(u) in file order and y holds
 * Ratings::r(u, i) (src/hgaprec.cc:1340-1345, src/ratings.hh:153-181).
 * row_ptr: n_users+1 entries; col_idx, y: row_ptr[n_users] entries;
 * y == NULL means every rating is 1 (-binary-data).  Ratings are >= 1: the
 * reference's reader drops class-0 lines (src/ratings.hh:191-197), and a value
 * that wrapped to 0 in its uint8 is walked by its loop as a 1 (only y > 1 scales,
 * src/hgaprec.cc:1355-1356) -- pass 1 for it, as hgaprec_b200/host does; an entry
 * with y == 0 contributes nothing here.  A (user, item) pair may be listed more
 * than once; every entry is processed, as the reference walks every line.
 * Replaces the Ratings adjacency-list iterator (src/env.hh:36-37).  May be called
 * again to replace the matrix. */
HPF_API int hpf_set_ratings_csr(hpf_ctx *ctx, const uint64_t *row_ptr,
                        const uint32_t *col_idx, const uint8_t *y);

/* Upload / download one parameter set: shape_curr(), rate_curr(), expected_v(),
 * expected_logv() of the matching GP* object (src/gpbase.hh:81-93, 463-475,
 * 807-819).  Host arrays are fp64 row-major (rows x k, or rows for the
 * GPArray / bias sets; the RATE of THETA/BETA without HPF_HIER is k long).
 * In hpf_get_state any pointer may be NULL.  hpf_set_state needs all four for
 * THETA/BETA/THETABIAS/BETABIAS (the reference's initial expectations are NOT
 * a function of its initial rates, src/gpbase.hh:324-340); THETARATE/BETARATE
 * need shape, rate and Ev. */
HPF_API int hpf_set_state(hpf_ctx *ctx, int which, const double *shape, const double *rate,
                  const double *Ev, const double *Elogv);
HPF_API int hpf_get_state(hpf_ctx *ctx, int which, double *shape, double *rate,
                  double *Ev, double *Elogv);

/* THE HOT PATH: n_iters full CAVI iterations (one pass of the loop bodies at
 * src/hgaprec.cc:1336-1435 / 927-979 / 1226-1318 without the report block).
 * Returns after the work has completed on the device. */
HPF_API int hpf_iterate(hpf_ctx *ctx, uint32_t n_iters);

/* compute_likelihood's sum (src/hgaprec.cc:1439-1470, 1503-1570) over held-out
 * (user, item, y) triples; users are LOCAL row numbers of this ctx.
 * *sum_ll receives the sum (the reference divides by npairs itself). */
HPF_API int hpf_heldout_loglik(hpf_ctx *ctx, const uint32_t *u, const uint32_t *i,
                       const uint8_t *y, uint64_t npairs, double *sum_ll);

/* HGAPRec::logl() (src/hgaprec.cc:2160-2255): the variational lower bound the reference appends
 * to logl.txt in every report window under -logl -- the per-nonzero terms over the ctx's training
 * matrix plus compute_elbo_term() of every parameter set (src/gpbase.hh:360-387, 717-741,
 * 951-969).  Needs a ctx created with HPF_LOGL.  With HPF_HIER the Gamma terms of theta / beta
 * use the rate priors the last iteration's set_prior_rate stored (src/gpbase.hh:163-173), so at
 * least one hpf_iterate must have run since THETARATE / BETARATE were set (the reference only
 * calls logl() inside the loop).  Multi-GPU: the value is this rank's part -- its users'
 * nonzeros and user-side sets, plus the (replicated) item-side sets on rank 0 only -- so the sum
 * over the ranks is the ELBO of the whole problem.
 * Known deviation: a rating that wrapped to 0 in the reference's uint8 (256, 512, ...) is a 1 in
 * the CSR (see hpf_set_ratings_csr).  The reference's logl() multiplies such an entry's entropy
 * terms by the raw 0, so it contributes only -E[theta].E[beta] there, and the full y = 1 term
 * here; the sweeps, the held-out likelihood and the rankings are unaffected. */
HPF_API int hpf_elbo(hpf_ctx *ctx, double *elbo_out);

/* compute_precision's scoring pass (src/hgaprec.cc:1703-1763, 1969-1991): for
 * each listed local user score every item with E[theta_u].E[beta_i] (+biases),
 * force the items of the user's exclusion list (train U validation) to 0.0 --
 * they stay candidates, as in the reference -- and return the topn best
 * (descending score, ties by ascending item).  excl_ptr: nu+1 entries.
 * items_out / scores_out: nu x topn. */
HPF_API int hpf_topn(hpf_ctx *ctx, const uint32_t *users, uint32_t nu,
             const uint64_t *excl_ptr, const uint32_t *excl_idx, uint32_t topn,
             uint32_t *items_out, float *scores_out);

/* compute_itemrank's ranking (src/hgaprec.cc:1607-1701): for each listed local
 * user and each of its query items (query_ptr: nu+1 entries into query_idx; the
 * reference asks for the user's test items), the 0-based position of the item in
 * the user's full descending score list -- same scores, same exclusion rule
 * (excluded items score 0.0) and same tie order (ascending item) as hpf_topn --
 * and its score.  rank_out / score_out: one entry per query. */
HPF_API int hpf_item_ranks(hpf_ctx *ctx, const uint32_t *users, uint32_t nu,
                   const uint64_t *excl_ptr, const uint32_t *excl_idx,
                   const uint64_t *query_ptr, const uint32_t *query_idx,
                   uint32_t *rank_out, float *score_out);

/* Contiguous user ranges balanced by NONZEROS (not by user count) for nranks
 * shards: first_user_out[r] .. first_user_out[r+1] is rank r's range
 * (nranks + 1 entries, first 0, last n_users).  Pure host arithmetic on the CSR
 * row pointer; needs no device and no ctx.  The reference has no counterpart
 * (it is single-threaded); this is the partition of SURVEY.md 8e. */
HPF_API int hpf_partition_users(const uint64_t *row_ptr, uint32_t n_users, uint32_t nranks,
                        uint32_t *first_user_out);

/* Multi-GPU (one process per GPU, users sharded, beta replicated): join an NCCL
 * communicator.  Rank 0 obtains an id with hpf_comm_unique_id and hands it to
 * the other ranks by any means (bench.py uses torch.distributed).  After this,
 * hpf_iterate all-reduces the item-side accumulators once per iteration and
 * hpf_heldout_loglik stays rank-local. */
#define HPF_COMM_ID_BYTES 128
HPF_API int hpf_comm_unique_id(void *id_out, size_t id_bytes);
HPF_API int hpf_comm_init(hpf_ctx *ctx, int rank, int nranks, const void *id, size_t id_bytes);

HPF_API int hpf_get_stats(const hpf_ctx *ctx, hpf_stats *out);

/* Measurement aid: run n_iters iterations with CUDA events recorded on the
 * library's stream around every kernel class and return the AVERAGE device
 * time per iteration of each (milliseconds).  Same arithmetic as hpf_iterate. */
typedef struct hpf_iter_profile {
  float sweep_user_ms;   /* phi sweep over the CSR (rows = users): gather kernel
                            (tail items) + tile kernel (head items)            */
  float sweep_item_ms;   /* phi sweep over the CSC (rows = items): tile kernel
                            over user blocks, or the gather kernel             */
  float combine_ms;      /* partial-sum combine of split rows (both sides)     */
  float update_theta_ms; /* dense theta row update + column sums               */
  float allreduce_ms;    /* NCCL all-reduce of the item-side block (0 if 1 GPU) */
  float update_beta_ms;  /* dense beta row update + column sums                */
  float total_ms;        /* first kernel start to last kernel end              */
  float sweep_user_head_ms; /* part of sweep_user_ms spent in the shared-memory
                               tile sweep over the most popular items           */
} hpf_iter_profile;
HPF_API int hpf_iterate_profiled(hpf_ctx *ctx, uint32_t n_iters, hpf_iter_profile *out);

#ifdef __cplusplus
}
#endif
#endif /* HPF_CUDA_H */

// hpf_topn.cuh -- K7: -gen-ranking scoring + mask + per-user top-N on sm_100a.
//
// Replaces the scoring loop of HGAPRec::compute_precision (hgaprec.cc:1725-1763)
// and prediction_score[_hier] (1850-1880, 1969-1991): for every listed user the
// score of EVERY item, E[theta_u] . E[beta_i] (+ E[thetabias_u] + E[betabias_i]),
// items of the user's exclusion list (training U validation) forced to 0.0, then
// the first topn of the descending sort (sort_by_value, matrix.hh:253-257).
//
// This is the one dense contraction of the path, so it runs on the 5th-gen
// tensor cores.  fp32 fidelity of the ranking is kept with a split-bf16 product:
//   x = hi + lo (two bf16),  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi
// (three tcgen05.mma kind::f16 passes into the same fp32 TMEM accumulator; the
// dropped a_lo.b_lo term is ~2^-16 relative).  The bias terms ride in two extra
// K columns ([.., bias_u, 1] . [.., 1, bias_i]).
//
// One CTA owns 128 users (the UMMA M) and streams all items in tiles of 256 (the
// UMMA N).  The kernel is bound by its epilogue (selection), not by the tensor
// pipe (3 x 1.7 TFLOP = 3.6 ms of MMA at C5), so a CTA is kept SMALL -- one 96 KB
// operand stage, one 256-column accumulator -- and TWO CTAs share an SM: while one
// drains its accumulator the other one's MMAs run, and eight epilogue warps
// instead of four hide each other's latencies:
//   warp 0      TMA producer: cp.async.bulk.tensor 2D, 128B-swizzled K-major
//               boxes of the hi / lo operand matrices into the operand stage
//   warp 1      TMEM allocator + the single thread that issues tcgen05.mma
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time, one THREAD per user
//               row; a score is kept only if it beats the row's running
//               threshold (the topn-th best so far), so the 8.5e9 scores of the
//               Netflix-scale problem are never materialised.  Survivors go to
//               a per-row candidate buffer in global memory (L2 resident); when
//               a buffer runs full the warp selects the topn-th largest key with
//               an 8-bit MSB-first radix select (histogram in shared memory)
//               and compacts.
// Keys are 64 bit, (score bits << 32) | ~item: scores are >= 0 so the float bits
// order like the values, all keys are distinct, and "larger key" is exactly
// "higher score, ties by lower item" (the reference's qsort is unstable on ties;
// this repo fixes them by ascending item, as its parity tests do).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hpf {
namespace topk {

constexpr int kTileM = 128;        // users per CTA == UMMA M
constexpr int kTileN = 256;        // items per accumulator buffer == UMMA N
constexpr int kBlockK = 64;        // bf16 elements per 128-byte swizzle row
constexpr int kStages = 1;
constexpr uint32_t kTmemCols = 256; // one accumulator tile: two CTAs per SM share the 512 TMEM columns
constexpr int kCap = 512;          // candidate slots per user row
constexpr int kMaxTopN = 256;      // kCap - kTileN
constexpr int kThreads = 192;      // 6 warps
constexpr uint32_t kABytes = kTileM * kBlockK * 2;  // 16 KB per (hi | lo) box
constexpr uint32_t kBBytes = kTileN * kBlockK * 2;  // 32 KB
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes; // 96 KB
constexpr uint32_t kSortBytes = 4 * kMaxTopN * 8;   // final sort buffers, one per epilogue warp
constexpr uint32_t kHistBytes = 4 * 256 * 4;        // radix-select histograms, one per epilogue warp
constexpr uint32_t kSmemBytes = 1024 + kStages * kStageBytes + kSortBytes + kHistBytes + 256; // 111,872: two per SM
constexpr unsigned long long kSpinLimit = 1ull << 28;

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps (launch failure reported through the ABI) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done = 0;
  for (unsigned long long spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > kSpinLimit) __trap();
  }
}
// the same for the two single-thread roles (TMA producer, MMA issuer): they wait for most of the kernel's life and
// would otherwise spend a fifth of the SM's issue slots on the polling loop (ncu, round 1)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
  uint32_t done = 0;
  for (unsigned long long spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) __nanosleep(64);
    if (spin > (kSpinLimit >> 4)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32])
{
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// smem matrix descriptor, K-major operand tile with 128-byte swizzle (what the TMA
// box writes): rows are 128 B apart, 8-row groups 1024 B apart (SBO), version 1
// (Blackwell), layout type 2 (SWIZZLE_128B).  The tile base must be 1024-aligned;
// a K step of 16 bf16 (32 B) inside the swizzle row advances the start address.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1u << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1u << 46;                 // descriptor version
  d |= (uint64_t)2u << 61;                 // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// ---- operand preparation ------------------------------------------------------
// rows [0, rows_pad) x cols [0, Kpad) of hi / lo (bf16, row stride Kpad) from the
// fp32 expectation matrix Ev (row stride ld).  gather != nullptr picks source rows
// (the listed users).  With bias two extra columns carry {bias, 1} (users) or
// {1, bias} (items).  Rows >= rows and the K padding are zero.
__global__ void __launch_bounds__(256) split_kernel(const float *Ev, uint32_t ld, uint32_t K, const float *bias_ev,
                                                    int bias_first, const uint32_t *gather, uint32_t rows,
                                                    uint32_t rows_pad, uint32_t Kpad, __nv_bfloat16 *hi,
                                                    __nv_bfloat16 *lo)
{
  const uint64_t total = (uint64_t)rows_pad * Kpad;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(e / Kpad), k = (uint32_t)(e % Kpad);
    float v = 0.f;
    if (r < rows) {
      const uint32_t src = gather ? gather[r] : r;
      if (k < K) v = Ev[(size_t)src * ld + k];
      else if (bias_ev != nullptr && k == K) v = bias_first ? bias_ev[src] : 1.f;
      else if (bias_ev != nullptr && k == K + 1) v = bias_first ? 1.f : bias_ev[src];
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[e] = h;
    lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// key[j] = (row_of[j] << 32) | idx[j]   (exclusion entries, sorted by cub afterwards)
__global__ void excl_key_kernel(const uint32_t *row_of, const uint32_t *idx, uint64_t cnt, uint64_t *key)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cnt) key[j] = ((uint64_t)row_of[j] << 32) | idx[j];
}
__global__ void excl_unkey_kernel(const uint64_t *key, uint64_t cnt, uint32_t *idx)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cnt) idx[j] = (uint32_t)key[j];
}

// ---- the scoring + selection kernel -------------------------------------------
struct TopnArgs {
  uint32_t nu;          // users of this launch (rows >= nu are padding)
  uint32_t m;           // items
  uint32_t nkb;         // K blocks of 64
  uint32_t ntiles_n;    // item tiles of 256
  uint32_t topn;
  const uint64_t *excl_ptr; // [nu + 1] into excl_sorted (already offset to this launch's first user)
  const uint32_t *excl_sorted; // per user ascending item ids
  unsigned long long *cand;    // [gridDim.x * 128 * kCap]
  uint32_t *items_out;  // [nu x topn]
  float *scores_out;    // [nu x topn]
};

__device__ __forceinline__ bool is_excluded(const uint32_t *lst, uint32_t len, uint32_t item)
{
  uint32_t lo = 0, hi = len;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t v = __ldg(lst + mid);
    if (v == item) return true;
    if (v < item) lo = mid + 1; else hi = mid;
  }
  return false;
}

// Warp-cooperative: among the cnt (<= kCap) keys of one row keep the `keep`
// largest (keep <= cnt), compacted to the front of buf in arbitrary order;
// returns the smallest kept key (the row's new threshold).  Keys are distinct
// and nonzero.  The keep-th largest key is found by an MSB-first radix select, 8
// bits per pass, over the 64-bit keys held in registers (kCap / 32 per lane):
// a 256-bin histogram of the keys still in play (shared memory atomics), a
// descending scan of the bins (8 per lane + a warp scan), the bin that holds the
// keep-th key narrows the prefix.  Leading bytes on which all keys agree are
// skipped, and the search stops as soon as the chosen bin holds exactly the keys
// still needed, so a typical call takes 2-3 passes (scores differ within their
// upper three bytes) instead of the ~40 bit steps of a bitwise count-select.
__device__ __forceinline__ unsigned long long warp_select(unsigned long long *buf, uint32_t cnt, uint32_t keep, int lane, uint32_t *hist)
{
  constexpr int PER = kCap / 32;
  unsigned long long key[PER];
  unsigned long long kor = 0ull, kand = ~0ull;
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const uint32_t j = lane + 32 * t;
    key[t] = j < cnt ? buf[j] : 0ull; // 0 is not a real key (that would be item 2^32-1)
    if (j < cnt) { kor |= key[t]; kand &= key[t]; }
  }
  kor = ((unsigned long long)__reduce_or_sync(0xffffffffu, (uint32_t)(kor >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (uint32_t)kor);
  kand = ((unsigned long long)__reduce_and_sync(0xffffffffu, (uint32_t)(kand >> 32)) << 32) | __reduce_and_sync(0xffffffffu, (uint32_t)kand);
  unsigned long long thr = 0ull;  // the keep-th largest key
  if (keep == cnt) {               // everything stays: the threshold is the smallest key
    unsigned long long mn = ~0ull;
#pragma unroll
    for (int t = 0; t < PER; ++t)
      if (key[t] != 0ull && key[t] < mn) mn = key[t];
    const uint32_t mh = __reduce_min_sync(0xffffffffu, (uint32_t)(mn >> 32));
    const uint32_t ml = __reduce_min_sync(0xffffffffu, (uint32_t)(mn >> 32) == mh ? (uint32_t)mn : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
  }
  const unsigned long long vary = kor & ~kand;
  int shift = vary == 0ull ? 0 : (63 - __clzll((long long)vary)) & ~7; // first byte on which the keys differ
  unsigned long long prefix = shift >= 56 ? 0ull : (kand >> (shift + 8)) << (shift + 8);
  unsigned long long pmask = shift >= 56 ? 0ull : ~0ull << (shift + 8);
  uint32_t need = keep;            // the answer is the need-th largest of the keys matching the prefix
  for (;; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane + 32 * i] = 0u;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < PER; ++t)
      if (key[t] != 0ull && (key[t] & pmask) == prefix) atomicAdd(hist + (uint32_t)((key[t] >> shift) & 255ull), 1u);
    __syncwarp();
    // lane L owns the digits 255 - 8L ... 248 - 8L, i.e. bins in DESCENDING digit order across the warp
    uint32_t h[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = hist[255 - (8 * lane + i)]; sum += h[i]; }
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const uint32_t excl = incl - sum;
    const bool mine = excl < need && need <= incl; // exactly one lane
    uint32_t digit = 0, left = 0, inbin = 0;
    if (mine) {
      uint32_t run = excl;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (inbin == 0u && run + h[i] >= need) { digit = 255u - (uint32_t)(8 * lane + i); left = need - run; inbin = h[i]; }
        run += h[i];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    digit = __shfl_sync(0xffffffffu, digit, src);
    need = __shfl_sync(0xffffffffu, left, src);
    inbin = __shfl_sync(0xffffffffu, inbin, src);
    prefix |= (unsigned long long)digit << shift;
    pmask |= 255ull << shift;
    if (inbin == need || shift == 0) break; // every key of the bin is kept (or the key is fully determined)
  }
  // threshold: the smallest key matching the prefix (all of them are kept when the loop stopped on inbin == need;
  // with shift == 0 the prefix IS the key)
  {
    unsigned long long mn = ~0ull;
#pragma unroll
    for (int t = 0; t < PER; ++t)
      if (key[t] != 0ull && (key[t] & pmask) == prefix && key[t] < mn) mn = key[t];
    const uint32_t mh = __reduce_min_sync(0xffffffffu, (uint32_t)(mn >> 32));
    const uint32_t ml = __reduce_min_sync(0xffffffffu, (uint32_t)(mn >> 32) == mh ? (uint32_t)mn : 0xffffffffu);
    thr = ((unsigned long long)mh << 32) | ml;
  }
  // compact: kept keys to the front (each lane writes its own, offsets by warp scan)
  uint32_t mine_cnt = 0;
#pragma unroll
  for (int t = 0; t < PER; ++t) mine_cnt += (key[t] >= thr && key[t] != 0ull) ? 1u : 0u;
  uint32_t incl = mine_cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  uint32_t pos = incl - mine_cnt;
  __syncwarp();
#pragma unroll
  for (int t = 0; t < PER; ++t)
    if (key[t] >= thr && key[t] != 0ull) buf[pos++] = key[t];
  __syncwarp();
  return thr;
}

// bitonic sort (descending) of P = 2^p <= kMaxTopN keys in shared memory by one warp
__device__ __forceinline__ void warp_sort_desc(unsigned long long *s, uint32_t P, int lane)
{
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = lane; i < P; i += 32) {
        const uint32_t l = i ^ j;
        if (l > i) {
          const unsigned long long a = s[i], b = s[l];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { s[i] = b; s[l] = a; }
        }
      }
      __syncwarp();
    }
}

__global__ void __launch_bounds__(kThreads, 2)
topn_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const TopnArgs a)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u; // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sort_off = kStages * kStageBytes;
  const uint32_t hist_off = sort_off + kSortBytes;
  const uint32_t bar_off = hist_off + kHistBytes;
  const uint32_t bar_full = base + bar_off;            // [kStages]
  const uint32_t bar_empty = bar_full + 8 * kStages;   // [kStages]
  const uint32_t bar_tfull = bar_empty + 8 * kStages;  // accumulator complete
  const uint32_t bar_tempty = bar_tfull + 8;           // accumulator drained
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen_base + bar_off + 8 * (2 * kStages + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile_m = blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) { // TMEM: one 256-column accumulator (the SM's other CTA takes the other half); this warp also frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t j = 0; j < a.ntiles_n; ++j)
        for (uint32_t kb = 0; kb < a.nkb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait_relaxed(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t st = base + s * kStageBytes;
          mbar_expect_tx(bar_full + 8 * s, kStageBytes);
          tma_load_2d(st, &map_a_hi, bar_full + 8 * s, (int)(kb * kBlockK), (int)(tile_m * kTileM));
          tma_load_2d(st + kABytes, &map_a_lo, bar_full + 8 * s, (int)(kb * kBlockK), (int)(tile_m * kTileM));
          tma_load_2d(st + 2 * kABytes, &map_b_hi, bar_full + 8 * s, (int)(kb * kBlockK), (int)(j * kTileN));
          tma_load_2d(st + 2 * kABytes + kBBytes, &map_b_lo, bar_full + 8 * s, (int)(kb * kBlockK), (int)(j * kTileN));
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t j = 0; j < a.ntiles_n; ++j) {
        mbar_wait_relaxed(bar_tempty, (j & 1u) ^ 1u); // epilogue has drained the accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base;
        for (uint32_t kb = 0; kb < a.nkb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait_relaxed(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = base + s * kStageBytes;
          const uint64_t da_hi = make_desc_sw128(st), da_lo = make_desc_sw128(st + kABytes);
          const uint64_t db_hi = make_desc_sw128(st + 2 * kABytes), db_lo = make_desc_sw128(st + 2 * kABytes + kBBytes);
#pragma unroll
          for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 16 * 2) >> 4); // 32 bytes per K step of 16 bf16
            tc_mma_bf16(d_tmem, da_hi + adv, db_hi + adv, kIdesc, (kb | kk) != 0u);
            tc_mma_bf16(d_tmem, da_hi + adv, db_lo + adv, kIdesc, 1u);
            tc_mma_bf16(d_tmem, da_lo + adv, db_hi + adv, kIdesc, 1u);
          }
          tc_commit(bar_empty + 8 * s);  // smem stage reusable once these MMAs have read it
        }
        tc_commit(bar_tfull);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: one thread per user row =====
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const uint32_t row_in_tile = (uint32_t)(q * 32 + lane);
    const uint32_t row = tile_m * kTileM + row_in_tile;
    const bool live = row < a.nu;
    unsigned long long *my = a.cand + (size_t)row * kCap;
    const uint32_t *ex = nullptr;
    uint32_t exlen = 0;
    if (live) {
      const uint64_t e0 = a.excl_ptr[row], e1 = a.excl_ptr[row + 1];
      ex = a.excl_sorted + e0;
      exlen = (uint32_t)(e1 - e0);
    }
    uint32_t tau_hi = live ? 0u : 0xffffffffu;     // score word of the row's threshold: smaller scores cannot make the top-n
    uint32_t cnt = 0;
    uint32_t excur = 0;                            // cursor into the (ascending) exclusion list
    uint32_t *hist = reinterpret_cast<uint32_t *>(gen_base + hist_off) + (size_t)q * 256;
    for (uint32_t j = 0; j < a.ntiles_n; ++j) {
      const uint32_t col0 = j * kTileN;
      const uint32_t ncols = a.m - col0 < (uint32_t)kTileN ? a.m - col0 : (uint32_t)kTileN;
      // this tile's excluded columns as a 256-bit mask: the list is sorted, so a cursor walks it once per row
      uint32_t mask[kTileN / 32];
#pragma unroll
      for (int w = 0; w < kTileN / 32; ++w) mask[w] = 0u;
      while (excur < exlen) {
        const uint32_t v = __ldg(ex + excur);
        if (v >= col0 + (uint32_t)kTileN) break;
        const uint32_t o = v - col0, bit = 1u << (o & 31u);
#pragma unroll
        for (int w = 0; w < kTileN / 32; ++w) mask[w] |= ((o >> 5) == (uint32_t)w) ? bit : 0u;
        ++excur;
      }
      mbar_wait(bar_tfull, j & 1u);
      tc_fence_after();
#pragma unroll
      for (uint32_t c = 0; c < kTileN / 32; ++c) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
        tc_wait_ld();
        const uint32_t valid = ncols > c * 32 ? ncols - c * 32 : 0u;
        const uint32_t mw = mask[c];                                   // excluded columns of this chunk
        const uint32_t keepw = valid >= 32u ? 0xffffffffu : ((1u << valid) - 1u); // columns that exist
        const uint32_t inv0 = ~(col0 + c * 32);                        // ~item of column 0; ~(x + i) == ~x - i
        // branch-free filter: a score is appended when its float bits reach the row's threshold (scores
        // are >= 0, so the bits order like the values).  An excluded item scores 0 (hgaprec.cc:1729-1735)
        // and therefore only survives while the row has not seen topn candidates yet (tau_hi == 0).
        // Appending on ">= the high word" is a superset of "key > tau"; the prune selects exactly.
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          uint32_t bits = ((mw >> i) & 1u) ? 0u : r[i];
          if ((int)bits < 0) bits = 0u;
          const bool push = ((keepw >> i) & 1u) && bits >= tau_hi;
          if (push) {
            my[cnt] = ((unsigned long long)bits << 32) | (unsigned long long)(inv0 - (uint32_t)i);
            ++cnt;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty);
      // rows that could overflow during the next tile are pruned now, warp-cooperatively
      __syncwarp();
      uint32_t need = __ballot_sync(0xffffffffu, cnt > (uint32_t)(kCap - kTileN));
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const uint32_t rcnt = __shfl_sync(0xffffffffu, cnt, src);
        unsigned long long *rb = a.cand + (size_t)(tile_m * kTileM + q * 32 + src) * kCap;
        const uint32_t keep = rcnt < a.topn ? rcnt : a.topn;
        const unsigned long long thr = warp_select(rb, rcnt, keep, lane, hist);
        if (lane == src) {
          cnt = keep;
          if (rcnt >= a.topn) tau_hi = (uint32_t)(thr >> 32);
        }
      }
    }
    // final: per row select the topn, sort them, emit
    __syncwarp();
    unsigned long long *sbuf = reinterpret_cast<unsigned long long *>(gen_base + sort_off) + (size_t)q * kMaxTopN;
    uint32_t P = 1;
    while (P < a.topn) P <<= 1;
    for (int src = 0; src < 32; ++src) {
      const uint32_t r_row = tile_m * kTileM + q * 32 + src;
      if (r_row >= a.nu) break;
      const uint32_t rcnt = __shfl_sync(0xffffffffu, cnt, src);
      unsigned long long *rb = a.cand + (size_t)r_row * kCap;
      const uint32_t keep = rcnt < a.topn ? rcnt : a.topn;
      if (rcnt > keep) warp_select(rb, rcnt, keep, lane, hist);
      for (uint32_t i = lane; i < P; i += 32) sbuf[i] = i < keep ? rb[i] : 0ull;
      __syncwarp();
      warp_sort_desc(sbuf, P, lane);
      for (uint32_t i = lane; i < a.topn; i += 32) {
        const unsigned long long key = sbuf[i];
        const bool have = i < keep;
        a.items_out[(size_t)r_row * a.topn + i] = have ? ~(uint32_t)key : 0xffffffffu;
        a.scores_out[(size_t)r_row * a.topn + i] = have ? __uint_as_float((uint32_t)(key >> 32)) : 0.f;
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Item ranks (compute_itemrank, hgaprec.cc:1607-1701): for listed (user, query
// item) pairs, the position of the item in the user's FULL descending list (same
// scores, same exclusion rule and same tie order as the top-N: key = score bits,
// ~item).  position = number of OTHER items whose key is larger.
//
// Two kernels.  rank_prep_kernel (one warp per user): the queries' own keys from
// fp32 dot products -- the score the caller gets back.  rank_mma_kernel: the same
// TMA / tcgen05 scoring pipeline as topn_kernel (all items of 128 users per CTA,
// 3 x bf16 split), with a COUNTING epilogue instead of a selecting one.  A row's
// queries are handled 32 at a time: their keys are sorted (ascending) into shared
// memory, and every score of the row is located among them by a 5-step branch-
// free binary search, b = #{queries with a smaller key}, and counted in hist[b];
// afterwards rank_j = sum_{b > j} hist[b].  A row with more than 32 queries takes
// more passes over the items (the contraction is cheap: the kernel is bound by
// this epilogue).  The query item meets ITSELF in the stream with the tensor-core
// score, which need not equal its fp32 key bit for bit; it is recognised by its
// item id next to the insertion point and taken out of its own count.
// Shared-memory tables are laid out [slot][row] so that the 32 rows of a warp hit
// 32 different banks whatever slot each one reads.
// ---------------------------------------------------------------------------
constexpr int kRankWarps = 8;          // rank_prep_kernel: users per CTA
constexpr int kQBatch = 32;            // queries of a row per pass
constexpr int kRankEpiWarps = 16;      // rank_mma_kernel: four epilogue threads per row, a quarter of a tile's columns each
constexpr int kRankColsPerThread = kTileN / (kRankEpiWarps / 4);
constexpr int kRankThreads = 64 + 32 * kRankEpiWarps;
constexpr uint32_t kRankKeyBytes = kQBatch * kTileM * 8;         // sorted query keys: score words [slot][row], then item words [slot][row]
constexpr uint32_t kRankHistBytes = (kQBatch + 1) * kTileM * 4;  // counts per bucket      [bucket][row]
constexpr uint32_t kRankOrigBytes = kQBatch * kTileM;            // query index in the batch, self flag: [slot][row] bytes
constexpr uint32_t kRankSmemBytes = 1024 + kStageBytes + kRankKeyBytes + kRankHistBytes + 2 * kRankOrigBytes + 256;

struct RankArgs {
  uint32_t nu, m, K, ld;
  uint32_t nkb, ntiles_n;           // rank_mma_kernel: K blocks of 64, item tiles of 256
  const uint32_t *users;
  const float *Et, *Eb, *Etb, *Ebb; // E[theta], E[beta] (row stride ld), bias expectations or nullptr
  const uint64_t *excl_ptr; const uint32_t *excl_sorted;
  const uint64_t *q_ptr; const uint32_t *q_idx; // query items per listed user
  unsigned long long *q_key; // scratch, one per query
  uint32_t *rank_out; float *score_out;
};

__device__ __forceinline__ float rank_dot(const float *th, const float *be, uint32_t K)
{
  float s = 0.f;
  for (uint32_t k = 0; k < K; ++k) s = fmaf(th[k], be[k], s);
  return s;
}

__global__ void __launch_bounds__(kRankWarps * 32) rank_prep_kernel(const RankArgs a)
{
  extern __shared__ float rsm[]; // [kRankWarps][K]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ua = blockIdx.x * kRankWarps + w; // position in the user list
  if (ua >= a.nu) return;
  const uint32_t u = a.users[ua];
  float *myth = rsm + (size_t)w * a.K;
  for (uint32_t k = lane; k < a.K; k += 32) myth[k] = a.Et[(size_t)u * a.ld + k];
  const float ub = a.Etb ? a.Etb[u] : 0.f;
  const uint32_t *ex = a.excl_sorted + a.excl_ptr[ua];
  const uint32_t exlen = (uint32_t)(a.excl_ptr[ua + 1] - a.excl_ptr[ua]);
  __syncwarp();
  for (uint64_t q = a.q_ptr[ua] + lane; q < a.q_ptr[ua + 1]; q += 32) {
    const uint32_t it = a.q_idx[q];
    float s = rank_dot(myth, a.Eb + (size_t)it * a.ld, a.K);
    if (a.Etb) s += ub + a.Ebb[it];
    uint32_t bits = __float_as_uint(s);
    if ((int)bits < 0 || is_excluded(ex, exlen, it)) bits = 0u;
    a.q_key[q] = ((unsigned long long)bits << 32) | (unsigned long long)(~it);
    a.score_out[q] = __uint_as_float(bits);
    a.rank_out[q] = 0u;
  }
}

__global__ void __launch_bounds__(kRankThreads, 1)
rank_mma_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const RankArgs a)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
  uint32_t *qhi = reinterpret_cast<uint32_t *>(gen_base + kStageBytes);                               // [kQBatch][128] score words
  uint32_t *qlo = qhi + kQBatch * kTileM;                                                             // [kQBatch][128] ~item
  uint32_t *hist = reinterpret_cast<uint32_t *>(gen_base + kStageBytes + kRankKeyBytes);             // [kQBatch + 1][128]
  uint8_t *qorig = gen_base + kStageBytes + kRankKeyBytes + kRankHistBytes;                           // [kQBatch][128]
  uint8_t *selfgt = qorig + kRankOrigBytes;                                                           // [kQBatch][128]
  const uint32_t bar_off = kStageBytes + kRankKeyBytes + kRankHistBytes + 2 * kRankOrigBytes;
  const uint32_t bar_full = base + bar_off, bar_empty = bar_full + 8, bar_tfull = bar_full + 16, bar_tempty = bar_full + 24;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen_base + bar_off + 32);
  uint32_t *maxq_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile_m = blockIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1); mbar_init(bar_empty, 1); mbar_init(bar_tfull, 1); mbar_init(bar_tempty, 32 * kRankEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *maxq_slot = 0u;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // the passes this CTA needs: the largest query count among its rows, 32 at a time
  if (threadIdx.x < kTileM) {
    const uint32_t row = tile_m * kTileM + threadIdx.x;
    if (row < a.nu) atomicMax(maxq_slot, (uint32_t)(a.q_ptr[row + 1] - a.q_ptr[row]));
  }
  __syncthreads();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t npass = (*maxq_slot + kQBatch - 1) / kQBatch;
  const uint32_t total_tiles = npass * a.ntiles_n;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t jt = 0; jt < total_tiles; ++jt) {
        const uint32_t j = jt % a.ntiles_n;
        for (uint32_t kb = 0; kb < a.nkb; ++kb, ++it) {
          mbar_wait_relaxed(bar_empty, (it & 1u) ^ 1u);
          mbar_expect_tx(bar_full, kStageBytes);
          tma_load_2d(base, &map_a_hi, bar_full, (int)(kb * kBlockK), (int)(tile_m * kTileM));
          tma_load_2d(base + kABytes, &map_a_lo, bar_full, (int)(kb * kBlockK), (int)(tile_m * kTileM));
          tma_load_2d(base + 2 * kABytes, &map_b_hi, bar_full, (int)(kb * kBlockK), (int)(j * kTileN));
          tma_load_2d(base + 2 * kABytes + kBBytes, &map_b_lo, bar_full, (int)(kb * kBlockK), (int)(j * kTileN));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t jt = 0; jt < total_tiles; ++jt) {
        mbar_wait_relaxed(bar_tempty, (jt & 1u) ^ 1u);
        tc_fence_after();
        for (uint32_t kb = 0; kb < a.nkb; ++kb, ++it) {
          mbar_wait_relaxed(bar_full, it & 1u);
          tc_fence_after();
          const uint64_t da_hi = make_desc_sw128(base), da_lo = make_desc_sw128(base + kABytes);
          const uint64_t db_hi = make_desc_sw128(base + 2 * kABytes), db_lo = make_desc_sw128(base + 2 * kABytes + kBBytes);
#pragma unroll
          for (uint32_t kk = 0; kk < kBlockK / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 16 * 2) >> 4);
            tc_mma_bf16(tmem_base, da_hi + adv, db_hi + adv, kIdesc, (kb | kk) != 0u);
            tc_mma_bf16(tmem_base, da_hi + adv, db_lo + adv, kIdesc, 1u);
            tc_mma_bf16(tmem_base, da_lo + adv, db_hi + adv, kIdesc, 1u);
          }
          tc_commit(bar_empty);
        }
        tc_commit(bar_tfull);
      }
    }
  } else {
    // ===== epilogue: four threads per user row, each a quarter of every tile's columns =====
    const int ew = warp - 2;                       // 0..15
    const int q4 = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = ew >> 2;                      // which part of the columns (the quarter order is 2,3,0,1 in every group of four warps)
    const uint32_t r_in = (uint32_t)(q4 * 32 + lane);
    const uint32_t row = tile_m * kTileM + r_in;
    const bool live = row < a.nu;
    const uint32_t *ex = nullptr;
    uint32_t exlen = 0;
    uint64_t q0 = 0, q1 = 0;
    if (live) {
      const uint64_t e0 = a.excl_ptr[row], e1 = a.excl_ptr[row + 1];
      ex = a.excl_sorted + e0;
      exlen = (uint32_t)(e1 - e0);
      q0 = a.q_ptr[row]; q1 = a.q_ptr[row + 1];
    }
    uint32_t jt = 0;
    for (uint32_t pass = 0; pass < npass; ++pass) {
      const uint64_t b0 = q0 + (uint64_t)pass * kQBatch;
      const uint32_t nq = b0 < q1 ? (uint32_t)((q1 - b0) < (uint64_t)kQBatch ? (q1 - b0) : (uint64_t)kQBatch) : 0u;
      if (half == 0) {
        // this row's batch: keys ascending by insertion (<= 32 of them), unused slots = +inf; counts cleared
        for (uint32_t j = 0; j < (uint32_t)kQBatch; ++j) {
          const unsigned long long k = j < nq ? a.q_key[b0 + j] : ~0ull;
          // insert into the sorted prefix [0, j)
          uint32_t p = j;
          while (p > 0) {
            const unsigned long long prev = ((unsigned long long)qhi[(p - 1) * kTileM + r_in] << 32) | qlo[(p - 1) * kTileM + r_in];
            if (prev <= k) break;
            qhi[p * kTileM + r_in] = qhi[(p - 1) * kTileM + r_in];
            qlo[p * kTileM + r_in] = qlo[(p - 1) * kTileM + r_in];
            qorig[p * kTileM + r_in] = qorig[(p - 1) * kTileM + r_in];
            --p;
          }
          qhi[p * kTileM + r_in] = (uint32_t)(k >> 32);
          qlo[p * kTileM + r_in] = (uint32_t)k;
          qorig[p * kTileM + r_in] = (uint8_t)j;
          selfgt[j * kTileM + r_in] = 0;
        }
        for (uint32_t b = 0; b <= (uint32_t)kQBatch; ++b) hist[b * kTileM + r_in] = 0u;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kRankEpiWarps) : "memory");
      uint32_t excur = 0;
      for (uint32_t j = 0; j < a.ntiles_n; ++j, ++jt) {
        constexpr uint32_t kCols = (uint32_t)kRankColsPerThread;
        const uint32_t col0 = j * kTileN + (uint32_t)half * kCols;
        const uint32_t lim = a.m > col0 ? a.m - col0 : 0u;
        const uint32_t ncols = lim < kCols ? lim : kCols;
        uint32_t mask[kCols / 32];
#pragma unroll
        for (uint32_t w = 0; w < kCols / 32; ++w) mask[w] = 0u;
        while (excur < exlen) { // the sorted exclusion list, restricted to this thread's columns of the tile
          const uint32_t v = __ldg(ex + excur);
          if (v >= col0 + kCols) break;
          if (v >= col0) {
            const uint32_t o = v - col0, bit = 1u << (o & 31u);
#pragma unroll
            for (uint32_t w = 0; w < kCols / 32; ++w) mask[w] |= ((o >> 5) == w) ? bit : 0u;
          }
          ++excur;
        }
        mbar_wait(bar_tfull, jt & 1u);
        tc_fence_after();
#pragma unroll
        for (uint32_t c = 0; c < kCols / 32; ++c) {
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)half * kCols + c * 32, r);
          tc_wait_ld();
          const uint32_t valid = ncols > c * 32 ? ncols - c * 32 : 0u;
          const uint32_t mw = mask[c];
          const uint32_t inv0 = ~(col0 + c * 32);
          if (live && nq > 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if ((uint32_t)i >= valid) continue; // only the last tile has columns past the last item
              uint32_t xhi = ((mw >> i) & 1u) ? 0u : r[i];
              if ((int)xhi < 0) xhi = 0u;
              const uint32_t xlo = inv0 - (uint32_t)i;
              // b = number of batch keys smaller than x = (xhi, xlo).  First by score word alone (32-bit compares): a
              // 5-step search over the 32 sorted slots plus the 6th comparison 32 slots need; then the (rare) run
              // of keys with the SAME score word is walked comparing item words.  Pads are +inf.
              uint32_t b = 0;
#pragma unroll
              for (uint32_t step = kQBatch / 2; step >= 1; step >>= 1)
                if (qhi[(b + step - 1) * kTileM + r_in] < xhi) b += step;
              if (qhi[b * kTileM + r_in] < xhi) ++b;
              while (b < (uint32_t)kQBatch && qhi[b * kTileM + r_in] == xhi && qlo[b * kTileM + r_in] < xlo) ++b;
              atomicAdd(hist + b * kTileM + r_in, 1u); // the row's other threads count into the same table
              // the query item itself, with the tensor-core score: next to the insertion point.  It must not count
              // towards its own rank, whichever of its two scores is larger.
              if (b > 0 && qlo[(b - 1) * kTileM + r_in] == xlo) selfgt[(b - 1) * kTileM + r_in] = 1; // x > key_q: counted, take it out
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_tempty);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kRankEpiWarps) : "memory");
      if (half == 0 && live) {
        // rank of sorted slot j = items counted in buckets above j, minus the query item itself if it landed there
        uint32_t above = 0;
        for (int j = kQBatch - 1; j >= 0; --j) {
          above += hist[(j + 1) * kTileM + r_in];
          if ((uint32_t)j < nq) a.rank_out[b0 + qorig[j * kTileM + r_in]] = above - (uint32_t)selfgt[j * kTileM + r_in];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kRankEpiWarps) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

} // namespace topk
} // namespace hpf

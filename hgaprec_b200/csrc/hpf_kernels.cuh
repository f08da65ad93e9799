// hpf_kernels.cuh -- sm_100a device code of the HPF CAVI engine.
//
// Formulation.  The reference forms, per nonzero (u,i,y),
//   phi_k = exp(Elog theta_uk + Elog beta_ik - logsumexp_k')      hgaprec.cc:206-239,
//                                                                 matrix.hh:367-389
// and adds y*phi to both shape rows (gpbase.hh:175-180).  Softmax of a SUM of
// logs is a normalised PRODUCT, so with A_uk = exp(Elog theta_uk - M_u) and
// B_ik = exp(Elog beta_ik - M_i) (row-max shifted, computed once per iteration
// in the dense row update):
//   Z_ui      = sum_k A_uk B_ik            (one dot product, no transcendental)
//   S^theta_uk = prior + A_uk * sum_i (y_ui / Z_ui) B_ik
//   S^beta_ik  = prior + B_ik * sum_u (y_ui / Z_ui) A_uk
// The two sums are the SAME kernel run over the CSR (rows = users, gathering
// item rows) and over the CSC (rows = items, gathering user rows): every
// accumulator lives in registers, there are no global atomics and the result
// is deterministic.  A nonzero whose Z under/overflows in fp32 takes an exact
// log-domain fallback (sweep_slow_path) that adds y*phi straight into a side
// buffer.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace hpf {

constexpr int kSweepThreads = 256;
// tuning knobs of the sweep kernels (tools/variants.sh builds alternatives side by side)
#ifndef HPF_UNROLL_T
#define HPF_UNROLL_T 1   // nonzeros of a chunk unrolled together
#endif
#ifndef HPF_SWEEP_MINBLOCKS
#define HPF_SWEEP_MINBLOCKS 3
#endif
#define HPF_PRAGMA(x) _Pragma(#x)
#define HPF_UNROLL(n) HPF_PRAGMA(unroll n)
constexpr int kUpdateWarps = 8;
constexpr float kZMin = 1e-30f;
constexpr float kZMax = 1e30f;

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// streaming (read-once) loads: keep them out of L1 so the gathered factor rows stay
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u8(const uint8_t *p)
{
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// 1/x, one MUFU (max error ~1 ulp); callers keep x inside [kZMin, kZMax]
__device__ __forceinline__ float frcp(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Blackwell packed fp32x2 arithmetic (FFMA2): two FMAs per issued instruction
__device__ __forceinline__ float2 lo2(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 &v) { return make_float2(v.z, v.w); }
template <int V> __device__ __forceinline__ float dot_rows(const float4 (&ar)[V], const float4 (&b)[V])
{
  float2 d = make_float2(0.f, 0.f), e = make_float2(0.f, 0.f); // two independent chains
#pragma unroll
  for (int v = 0; v < V; ++v) {
    d = __ffma2_rn(lo2(ar[v]), lo2(b[v]), d);
    e = __ffma2_rn(hi2(ar[v]), hi2(b[v]), e);
  }
  d = __fadd2_rn(d, e);
  return d.x + d.y;
}
template <int V> __device__ __forceinline__ void axpy_rows(float sc, const float4 (&b)[V], float4 (&acc)[V])
{
  const float2 s2 = make_float2(sc, sc);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float2 l = __ffma2_rn(s2, lo2(b[v]), lo2(acc[v]));
    const float2 h = __ffma2_rn(s2, hi2(b[v]), hi2(acc[v]));
    acc[v] = make_float4(l.x, l.y, h.x, h.y);
  }
}

// digamma for x > 0 in fp32: upward recurrence to x >= 6, then the Stirling
// series.  Replaces gsl_sf_psi at gpbase.hh:260,593,923.
__device__ __forceinline__ float digammaf(float x)
{
  // psi(x) = psi(x + 6) - sum_{i<6} 1/(x + i); the sum is formed as ONE quotient p/q
  // (q = prod (x+i), p = sum of the products leaving one factor out): one division instead of six
  float r = 0.f;
  if (x < 6.f) {
    float q = x, p = 1.f;
#pragma unroll
    for (int i = 1; i < 6; ++i) {
      const float t = x + (float)i;
      p = fmaf(p, t, q);
      q *= t;
    }
    r = -__fdividef(p, q);
    x += 6.f;
  }
  const float inv = __fdividef(1.f, x);
  const float f = inv * inv;
  const float t = f * (-1.f / 12.f + f * (1.f / 120.f + f * (-1.f / 252.f + f * (1.f / 240.f))));
  return r + __logf(x) - 0.5f * inv + t;
}

// ---------------------------------------------------------------------------
// K1: phi sweep.  One GROUP of G lanes owns one work segment (<= seg_len
// consecutive nonzeros of ONE row); 32/G segments advance in lock-step per warp.
// Lane g of a group holds float4 #(g + v*G), v < V, of the row-side factor row
// and of the accumulator; per nonzero the group gathers the column-side row with
// 128-bit loads, forms Z with a G-lane xor-shuffle reduction and accumulates
// (y/Z) * B.  Replaces hgaprec.cc:1340-1366 / 928-942 / 1227-1248.
// ---------------------------------------------------------------------------
struct SweepArgs {
  const uint4 *seg;        // {begin_lo, begin_hi, row, len}
  const uint32_t *seg_out; // < R: row of T;  >= R: slot (out - R) of Tpart
  uint32_t nsegs;
  uint32_t R;              // rows on the row side
  const uint32_t *idx;     // per nonzero: row number on the column side
  const uint8_t *y;        // per nonzero rating, or nullptr (all ones)
  const float *Arow;       // [R x ld]  exp(Elog - shift), row side
  const float *Acol;       // [C x ld]  same, column side
  float *T;                // [R x ld]  out: sum (y/Z) * Acol
  float *Tpart;            // [P x ld]  out: partial sums of multi-segment rows
  // -bias (phi has K+2 slots, hgaprec.cc:222-239)
  const float2 *row_aux;   // {exp(Elogbias_r - shift_r), exp(-shift_r)}
  const float2 *col_aux;
  float *Tb;               // [R] out: sum (y/Z) * col_aux.y
  float *Tbpart;           // [P]
  // exact fallback
  const float *ElogRow, *ElogCol;   // [. x ld], padding = -inf
  const float *ElogbRow, *ElogbCol; // bias logs (or nullptr)
  float *Tdirect;          // [R x ld] += y*phi
  float *Tbdirect;         // [R]
  uint32_t *direct_flag;   // set to 1 when Tdirect was touched
  unsigned long long *slow_count;
  uint32_t K, K4;  // factors; float4 per row that hold data (K rounded up to 4)
  uint32_t ld, ld4; // row stride of every matrix in floats / float4 (rows are 128-byte aligned)
};

template <int G>
__device__ __forceinline__ uint32_t group_mask(int lane)
{
  if (G == 32) return 0xffffffffu;
  return ((1u << (G & 31)) - 1u) << (lane & ~(G - 1));
}

// Exact log-domain phi for one nonzero (called when Z left the fp32 range):
// phi_k = exp(x_k - max) / sum, x_k = ElogRow_k + ElogCol_k (+ the two bias
// logs), i.e. what get_phi + lognormalize compute; y*phi is added to Tdirect.
struct SlowArgs { // by value: taking the address of kernel parameters would force a local copy
  const float *ElogRow, *ElogCol, *ElogbRow, *ElogbCol;
  float *Tdirect, *Tbdirect;
  uint32_t *direct_flag;
  unsigned long long *slow_count;
  uint32_t K, K4, ld, ld4;
};

template <int G, int V, bool BIAS>
__device__ __noinline__ void sweep_slow_path(const SlowArgs a, uint32_t row, uint32_t c, float yv, int lane)
{
  const int gl = lane & (G - 1);
  const uint32_t mask = group_mask<G>(lane);
  const float4 *er = reinterpret_cast<const float4 *>(a.ElogRow) + (size_t)row * a.ld4;
  const float4 *ec = reinterpret_cast<const float4 *>(a.ElogCol) + (size_t)c * a.ld4;
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint32_t q = gl + v * G;
    if (q < a.K4) {
      const float4 r4 = er[q], c4 = ec[q];
      mx = fmaxf(mx, fmaxf(fmaxf(r4.x + c4.x, r4.y + c4.y), fmaxf(r4.z + c4.z, r4.w + c4.w)));
    }
  }
  float xbr = -CUDART_INF_F, xbc = -CUDART_INF_F;
  if (BIAS) {
    xbr = a.ElogbRow[row];
    xbc = a.ElogbCol[c];
    mx = fmaxf(mx, fmaxf(xbr, xbc));
  }
#pragma unroll
  for (int off = G / 2; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(mask, mx, off));
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint32_t q = gl + v * G;
    if (q < a.K4) {
      const float4 r4 = er[q], c4 = ec[q];
      sum += expf(r4.x + c4.x - mx) + expf(r4.y + c4.y - mx) + expf(r4.z + c4.z - mx) + expf(r4.w + c4.w - mx);
    }
  }
  if (BIAS && gl == 0) sum += expf(xbr - mx) + expf(xbc - mx);
#pragma unroll
  for (int off = G / 2; off >= 1; off >>= 1) sum += __shfl_xor_sync(mask, sum, off);
  const float sc = yv / sum;
  float *td = a.Tdirect + (size_t)row * a.ld;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint32_t q = gl + v * G;
    if (q < a.K4) {
      const float4 r4 = er[q], c4 = ec[q];
      const uint32_t k0 = q * 4;
      if (k0 + 0 < a.K) atomicAdd(td + k0 + 0, sc * expf(r4.x + c4.x - mx));
      if (k0 + 1 < a.K) atomicAdd(td + k0 + 1, sc * expf(r4.y + c4.y - mx));
      if (k0 + 2 < a.K) atomicAdd(td + k0 + 2, sc * expf(r4.z + c4.z - mx));
      if (k0 + 3 < a.K) atomicAdd(td + k0 + 3, sc * expf(r4.w + c4.w - mx));
    }
  }
  if (gl == 0) {
    if (BIAS) atomicAdd(a.Tbdirect + row, sc * expf(xbr - mx));
    atomicAdd(a.slow_count, 1ull);
    *a.direct_flag = 1u;
  }
}

template <int G, int V, bool BIAS>
__global__ void __launch_bounds__(kSweepThreads, HPF_SWEEP_MINBLOCKS) sweep_kernel(const SweepArgs a)
{
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const uint32_t group = (blockIdx.x * (uint32_t)kSweepThreads + threadIdx.x) / G;
  const bool have = group < a.nsegs;

  uint64_t begin = 0;
  uint32_t row = 0, len = 0, out = 0;
  if (have) {
    const uint4 s = __ldg(a.seg + group);
    begin = (uint64_t)s.x | ((uint64_t)s.y << 32);
    row = s.z;
    len = s.w;
    out = __ldg(a.seg_out + group);
  }

  float4 ar[V], acc[V];
  {
    const float4 *rp = reinterpret_cast<const float4 *>(a.Arow) + (size_t)row * a.ld4;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t q = gl + v * G;
      ar[v] = (have && q < a.K4) ? ldg4(rp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float2 raux = make_float2(0.f, 0.f);
  float accb = 0.f;
  if (BIAS && have) raux = __ldg(a.row_aux + row);

  const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
  const uint32_t *ip = a.idx + begin;
  const uint8_t *yp = a.y ? a.y + begin : nullptr;
  const float4 *acol = reinterpret_cast<const float4 *>(a.Acol);
  // Lanes past the last float4 of a row re-read that float4 (their row-side value is 0 and their sums are
  // never stored), so the loop carries no predicates.  Only the last of the V slots can be out of range.
  const uint32_t q_last = min((uint32_t)(gl + (V - 1) * G), a.K4 - 1u);

  // chunks of G nonzeros: every lane of the group fetches one (index, rating), then the chunk is unrolled
  // so that the gathers of one nonzero overlap the arithmetic of the previous one
  for (uint32_t j0 = 0; j0 < maxlen; j0 += G) {
    const uint32_t jj = j0 + gl;
    const uint32_t cbuf = (jj < len) ? ld_stream_u32(ip + jj) : 0u; // 0 past the end: a valid row
    const float ybuf = (jj < len) ? (yp != nullptr ? (float)ld_stream_u8(yp + jj) : 1.f) : 0.f; // 0: no contribution
    HPF_UNROLL(HPF_UNROLL_T)
    for (int t = 0; t < G; ++t) {
      const uint32_t c = __shfl_sync(0xffffffffu, cbuf, t, G);
      const float yv = __shfl_sync(0xffffffffu, ybuf, t, G);
      float4 b[V];
      const float4 *cp = acol + (size_t)c * a.ld4;
#pragma unroll
      for (int v = 0; v < V - 1; ++v) b[v] = ldg4(cp + gl + v * G);
      b[V - 1] = ldg4(cp + q_last);
      float2 caux = make_float2(0.f, 0.f);
      if (BIAS) caux = __ldg(a.col_aux + c);

      float dot = dot_rows<V>(ar, b);
#pragma unroll
      for (int off = G / 2; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
      float z = dot;
      if (BIAS) z += raux.x * caux.y + caux.x * raux.y;

      const bool ok = z > kZMin && z < kZMax;
      const float sc = ok ? yv * frcp(z) : 0.f;
      axpy_rows<V>(sc, b, acc);
      if (BIAS) accb = fmaf(sc, caux.y, accb);
      if (!ok && yv != 0.f) { // Z left the fp32 range: exact log-domain path (rare)
        SlowArgs sa;
        sa.ElogRow = a.ElogRow; sa.ElogCol = a.ElogCol; sa.ElogbRow = a.ElogbRow; sa.ElogbCol = a.ElogbCol;
        sa.Tdirect = a.Tdirect; sa.Tbdirect = a.Tbdirect; sa.direct_flag = a.direct_flag;
        sa.slow_count = a.slow_count; sa.K = a.K; sa.K4 = a.K4; sa.ld = a.ld; sa.ld4 = a.ld4;
        sweep_slow_path<G, V, BIAS>(sa, row, c, yv, lane);
      }
    }
  }

  if (have) {
    float4 *dst = reinterpret_cast<float4 *>(out < a.R ? a.T + (size_t)out * a.ld
                                                        : a.Tpart + (size_t)(out - a.R) * a.ld);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t q = gl + v * G;
      if (q < a.K4) dst[q] = acc[v];
    }
    if (BIAS && gl == 0) {
      if (out < a.R) a.Tb[out] = accb;
      else a.Tbpart[out - a.R] = accb;
    }
  }
}

// ---------------------------------------------------------------------------
// K1b: tile sweep.  Same arithmetic as sweep_kernel, but the GATHERED factor rows
// come from shared memory: a tile of up to tile_rows rows of the column side is
// staged once per CTA (packed K floats per row) and every nonzero whose column
// falls into that tile reads it from there -- the L2->SM gather, which bounds
// sweep_kernel, disappears.  Two uses:
//   item pass   rows = items, tiles = consecutive blocks of users; an item's
//               nonzeros inside one user block form one segment;
//   user pass   rows = users, ONE tile holding the most popular items (a Zipf
//               head carries most of the nonzeros); the tail stays in sweep_kernel.
// A row's sum is spread over many CTAs, so partial sums are added to T with
// vector reductions (red.global.add.v4.f32) -- T is cleared (item pass) or
// written by sweep_kernel (user pass) beforehand on the same stream.
// Work: chunk c = (tile c / cpt, part c % cpt) covers an even share of the
// tile's segments (tile_seg_ptr); CTAs take chunks round-robin.
// ---------------------------------------------------------------------------
#ifndef HPF_TILE_THREADS
#define HPF_TILE_THREADS 640
#endif
constexpr int kTileThreads = HPF_TILE_THREADS;

struct TileArgs {
  const uint4 *seg;             // {begin_lo, begin_hi, row, len}, sorted by (tile, len desc)
  const uint32_t *tile_seg_ptr; // [ntiles + 1]
  uint32_t ntiles, cpt;         // chunks per tile
  uint32_t tile_rows;           // rows per tile (last tile may hold fewer)
  uint32_t C;                   // rows on the column side
  const uint32_t *tile_row_ids; // explicit row list of tile 0 (single-tile use) or nullptr: tile t = rows [t*tile_rows, ...)
  uint32_t tile0_count;         // rows in the explicit list
  const uint32_t *idx;          // per nonzero: SLOT inside its tile
  const uint8_t *y;
  const float *Arow, *Acol;     // [. x ld]
  float *T;                     // [R x ld]  +=
  const float2 *row_aux, *col_aux;
  float *Tb;                    // [R] +=
  const float *ElogRow, *ElogCol, *ElogbRow, *ElogbCol;
  float *Tdirect, *Tbdirect;
  uint32_t *direct_flag;
  unsigned long long *slow_count;
  uint32_t K, K4, ld, ld4;
};

__device__ __forceinline__ void red_add_v4(float *addr, float4 v)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int G, int V, bool BIAS>
__global__ void __launch_bounds__(kTileThreads, 1) tile_sweep_kernel(const TileArgs a)
{
  extern __shared__ float4 tile_sm[]; // [tile_rows x K4] (+ float2 aux[tile_rows] with BIAS)
  float2 *aux_sm = reinterpret_cast<float2 *>(tile_sm + (size_t)a.tile_rows * a.K4);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const uint32_t gid = threadIdx.x / G;              // group inside the CTA
  constexpr uint32_t kGroups = kTileThreads / G;
  const uint32_t nchunks = a.ntiles * a.cpt;
  uint32_t cur_tile = 0xffffffffu;
  bool pq[V];
#pragma unroll
  for (int v = 0; v < V; ++v) pq[v] = (uint32_t)(gl + v * G) < a.K4;
  // lanes past the last float4 of a row re-read it (row-side value 0, sums never stored): only the last slot can be
  const uint32_t q_last = min((uint32_t)(gl + (V - 1) * G), a.K4 - 1u);

  for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const uint32_t tile = ch / a.cpt, part = ch % a.cpt;
    // the tile's segments are sorted by length: part p takes every cpt-th one, so the parts carry equal work
    const uint32_t t0 = __ldg(a.tile_seg_ptr + tile), t1 = __ldg(a.tile_seg_ptr + tile + 1);
    if (t0 + part >= t1) continue;
    const uint32_t nmine = (t1 - t0 - part + a.cpt - 1) / a.cpt; // segments t0 + part + cpt * i, i < nmine
    const uint32_t tile_base = tile * a.tile_rows;
    if (tile != cur_tile) { // stage the tile's factor rows (and bias terms)
      __syncthreads();
      const uint32_t nrows = a.tile_row_ids ? a.tile0_count : min(a.tile_rows, a.C - tile_base);
      const float4 *acol = reinterpret_cast<const float4 *>(a.Acol);
      for (uint32_t e = threadIdx.x; e < nrows * a.K4; e += kTileThreads) {
        const uint32_t r = e / a.K4, q = e - r * a.K4;
        const uint32_t src = a.tile_row_ids ? __ldg(a.tile_row_ids + r) : tile_base + r;
        tile_sm[e] = __ldg(acol + (size_t)src * a.ld4 + q);
      }
      if (BIAS)
        for (uint32_t r = threadIdx.x; r < nrows; r += kTileThreads)
          aux_sm[r] = __ldg(a.col_aux + (a.tile_row_ids ? __ldg(a.tile_row_ids + r) : tile_base + r));
      cur_tile = tile;
      __syncthreads();
    }
    // the 32/G groups of a warp advance in lock-step over consecutive (equally long) segments
    for (uint32_t ib = gid & ~(uint32_t)(32 / G - 1); ib < nmine; ib += kGroups) {
      const uint32_t i = ib + (gid & (32 / G - 1));
      const bool have = i < nmine;
      const uint32_t sidx = t0 + part + a.cpt * i;
      uint64_t begin = 0;
      uint32_t row = 0, len = 0;
      if (have) {
        const uint4 s = __ldg(a.seg + sidx);
        begin = (uint64_t)s.x | ((uint64_t)s.y << 32);
        row = s.z;
        len = s.w;
      }
      float4 ar[V], acc[V], b[V];
      {
        const float4 *rp = reinterpret_cast<const float4 *>(a.Arow) + (size_t)row * a.ld4;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          ar[v] = (have && pq[v]) ? ldg4(rp + gl + v * G) : make_float4(0.f, 0.f, 0.f, 0.f);
          acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          b[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float2 raux = make_float2(0.f, 0.f);
      float accb = 0.f;
      if (BIAS && have) raux = __ldg(a.row_aux + row);
      const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
      const uint32_t *ip = a.idx + begin;
      const uint8_t *yp = a.y ? a.y + begin : nullptr;
      for (uint32_t j0 = 0; j0 < maxlen; j0 += G) {
        const uint32_t jj = j0 + gl;
        const uint32_t cbuf = (jj < len) ? ld_stream_u32(ip + jj) : 0u; // slot 0 past the end: valid
        const float ybuf = (jj < len) ? (yp != nullptr ? (float)ld_stream_u8(yp + jj) : 1.f) : 0.f;
        HPF_UNROLL(HPF_UNROLL_T)
        for (int t = 0; t < G; ++t) {
          const uint32_t c = __shfl_sync(0xffffffffu, cbuf, t, G);
          const float yv = __shfl_sync(0xffffffffu, ybuf, t, G);
          const float4 *cp = tile_sm + (size_t)c * a.K4;
#pragma unroll
          for (int v = 0; v < V - 1; ++v) b[v] = cp[gl + v * G];
          b[V - 1] = cp[q_last];
          float dot = dot_rows<V>(ar, b);
#pragma unroll
          for (int off = G / 2; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
          float z = dot;
          float2 caux = make_float2(0.f, 0.f);
          if (BIAS) {
            caux = aux_sm[c];
            z += raux.x * caux.y + caux.x * raux.y;
          }
          const bool ok = z > kZMin && z < kZMax;
          const float sc = ok ? yv * frcp(z) : 0.f;
          axpy_rows<V>(sc, b, acc);
          if (BIAS) accb = fmaf(sc, caux.y, accb);
          if (!ok && yv != 0.f) { // Z left the fp32 range: exact log-domain path (rare)
            SlowArgs sa;
            sa.ElogRow = a.ElogRow; sa.ElogCol = a.ElogCol; sa.ElogbRow = a.ElogbRow; sa.ElogbCol = a.ElogbCol;
            sa.Tdirect = a.Tdirect; sa.Tbdirect = a.Tbdirect; sa.direct_flag = a.direct_flag;
            sa.slow_count = a.slow_count; sa.K = a.K; sa.K4 = a.K4; sa.ld = a.ld; sa.ld4 = a.ld4;
            const uint32_t cg = a.tile_row_ids ? __ldg(a.tile_row_ids + c) : tile_base + c;
            sweep_slow_path<G, V, BIAS>(sa, row, cg, yv, lane);
          }
        }
      }
      if (have && len > 0) {
        float *dst = a.T + (size_t)row * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (pq[v]) red_add_v4(dst + (size_t)(gl + v * G) * 4, acc[v]);
        if (BIAS && gl == 0) atomicAdd(a.Tb + row, accb);
      }
    }
  }
}

// ---- device-side work list of the tile sweep ----------------------------------
// run q = tile * R + row covers nonzeros [run_ptr[q], run_ptr[q+1]); it is cut into
// segments of <= L nonzeros.  cnt[q] = segments of run q.
__global__ void seg_count_kernel(const uint64_t *run_ptr, uint64_t nruns, uint32_t L, uint32_t *cnt)
{
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nruns) {
    const uint64_t len = run_ptr[q + 1] - run_ptr[q];
    cnt[q] = (uint32_t)((len + L - 1) / L);
  }
}

// segments of run q go to slots [off[q], off[q] + cnt); key orders them by (tile, descending length)
__global__ void seg_emit_kernel(const uint64_t *run_ptr, const uint32_t *off, uint64_t nruns, uint32_t R, uint32_t L,
                                uint4 *seg, uint32_t *key)
{
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nruns) return;
  const uint64_t b0 = run_ptr[q], len = run_ptr[q + 1] - b0;
  if (len == 0) return;
  const uint32_t tile = (uint32_t)(q / R), row = (uint32_t)(q % R);
  const uint32_t cnt = (uint32_t)((len + L - 1) / L);
  uint32_t o = off[q];
  for (uint32_t s = 0; s < cnt; ++s, ++o) {
    const uint64_t sb = b0 + (uint64_t)s * L;
    const uint32_t sl = (uint32_t)min((uint64_t)L, len - (uint64_t)s * L);
    seg[o] = make_uint4((uint32_t)sb, (uint32_t)(sb >> 32), row, sl);
    key[o] = tile * (L + 1) + (L - sl);
  }
}

// tile_ptr[t] = first sorted segment whose tile is >= t, t in [0, ntiles]
__global__ void tile_ptr_kernel(const uint32_t *sorted_key, uint32_t nsegs, uint32_t Lp1, uint32_t ntiles, uint32_t *tile_ptr)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > nsegs) return;
  const uint32_t prev = j == 0 ? 0u : sorted_key[j - 1] / Lp1 + 1u;
  const uint32_t cur = j == nsegs ? ntiles + 1u : sorted_key[j] / Lp1 + 1u;
  for (uint32_t t = prev; t < cur && t <= ntiles; ++t) tile_ptr[t] = j;
}

// column-side index -> slot inside its tile of tile_rows consecutive rows
__global__ void to_slot_kernel(uint32_t *idx, uint64_t nnz, uint32_t tile_rows)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) idx[j] %= tile_rows;
}

// ---- head / tail split of the user pass ----------------------------------------
// item degrees from the item-pass runs (run q = tile * R + item): no atomics on the hot items
__global__ void degree_kernel(const uint64_t *run_ptr, uint32_t R, uint32_t ntiles, uint32_t *deg)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  uint64_t d = 0;
  for (uint32_t t = 0; t < ntiles; ++t) {
    const size_t q = (size_t)t * R + i;
    d += run_ptr[q + 1] - run_ptr[q];
  }
  deg[i] = (uint32_t)d;
}
// keys that sort items by descending degree (ties by ascending id through the stable sort)
__global__ void neg_key_kernel(const uint32_t *deg, uint32_t m, uint32_t *key, uint32_t *id)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) { key[i] = 0xffffffffu - deg[i]; id[i] = i; }
}
// slot_of[item] = position among the H most popular items, or 0xffffffff
__global__ void head_slot_kernel(const uint32_t *sorted_id, uint32_t H, uint32_t *slot_of)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < H) slot_of[sorted_id[r]] = r;
}
__global__ void head_flag_kernel(const uint32_t *col, const uint32_t *slot_of, uint64_t nnz, uint32_t *is_tail)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) is_tail[j] = slot_of[col[j]] == 0xffffffffu ? 1u : 0u;
  else if (j == nnz) is_tail[j] = 0u; // sentinel: the exclusive scan's last entry is the tail total
}
// stable partition of the CSR nonzeros into a tail CSR (column ids) and a head list (slots),
// both still ordered by user; tail_pos is the exclusive scan of is_tail
__global__ void head_split_kernel(const uint32_t *col, const uint8_t *y, const uint32_t *slot_of, const uint32_t *tail_pos,
                                  uint64_t nnz, uint32_t *tail_idx, uint8_t *tail_y, uint32_t *head_idx, uint8_t *head_y)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const uint32_t cc = col[j], sl = slot_of[cc];
  const uint32_t tp = tail_pos[j];
  if (sl == 0xffffffffu) {
    tail_idx[tp] = cc;
    if (y) tail_y[tp] = y[j];
  } else if (head_idx != nullptr) { // the dense head keeps no list
    const uint64_t hp = j - tp;
    head_idx[hp] = sl;
    if (y) head_y[hp] = y[j];
  }
}
// row pointers of both parts from the scan: tail_ptr[u] = tail_pos[row_ptr[u]], head_ptr[u] = row_ptr[u] - tail_ptr[u]
// (tail_pos has nnz + 1 entries: the last one is the tail total)
__global__ void split_ptr_kernel(const uint64_t *row_ptr, const uint32_t *tail_pos, uint32_t n, uint64_t *tail_ptr,
                                 uint64_t *head_ptr)
{
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u > n) return;
  const uint64_t p = row_ptr[u];
  const uint64_t tp = tail_pos[p];
  tail_ptr[u] = tp;
  head_ptr[u] = p - tp;
}

// ---------------------------------------------------------------------------
// combine the partial sums of rows that were split over several segments
// (fixed order => deterministic).  One block per multi-segment row.
// ---------------------------------------------------------------------------
struct CombineArgs {
  const uint32_t *multi_row, *multi_first, *multi_cnt;
  uint32_t nmulti, Kp, ld; // Kp: K rounded up to 4 (loop bound); ld: row stride in floats
  const float *Tpart;
  float *T;
  const float *Tbpart; // or nullptr
  float *Tb;
};

__global__ void __launch_bounds__(kUpdateWarps * 32) combine_kernel(const CombineArgs a)
{
  extern __shared__ float sm[]; // [kUpdateWarps][Kp] (+ kUpdateWarps for bias)
  const uint32_t r = blockIdx.x;
  if (r >= a.nmulti) return;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row = a.multi_row[r], first = a.multi_first[r], cnt = a.multi_cnt[r];
  float *mine = sm + (size_t)w * a.Kp;
  for (uint32_t k = lane; k < a.Kp; k += 32) {
    float s = 0.f;
    for (uint32_t p = w; p < cnt; p += kUpdateWarps) s += a.Tpart[(size_t)(first + p) * a.ld + k];
    mine[k] = s;
  }
  float *smb = sm + (size_t)kUpdateWarps * a.Kp;
  if (a.Tbpart != nullptr && lane == 0) {
    float s = 0.f;
    for (uint32_t p = w; p < cnt; p += kUpdateWarps) s += a.Tbpart[first + p];
    smb[w] = s;
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < a.Kp; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kUpdateWarps; ++q) s += sm[(size_t)q * a.Kp + k];
    a.T[(size_t)row * a.ld + k] = s;
  }
  if (a.Tbpart != nullptr && threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kUpdateWarps; ++q) s += smb[q];
    a.Tb[row] = s;
  }
}

// ---------------------------------------------------------------------------
// K2/K3/K4/K5: dense row update, one warp per row.  Fuses, for one parameter
// matrix: shape = prior + A*T (the add_slice sums), rate (set_prior_rate +
// update_rate_next, gpbase.hh:163-173,218-223 / GR 560-564), swap (240-246),
// compute_expectations (248-262), sum_rows / sum_cols (264-280), the GPArray
// xi/eta step (hgaprec.cc:1398-1414, gpbase.hh:877-925), the bias step
// (hgaprec.cc:1388-1396) and the next iteration's shifted exponentials.
// ---------------------------------------------------------------------------
struct UpdateArgs {
  uint32_t R, K, Kp, ld; // Kp: K rounded up to 4; ld: row stride in floats
  const float *T;
  float *Tdirect;
  const uint32_t *direct_flag;
  float *A, *Elog, *Ev, *shape, *rate; // rate: [R x ld] (hier) or [Kp] (global rate)
  float *shift;                        // [R]
  int hier;
  const float *colsum_other; // [Kp]  sum over the OTHER side's rows of Ev
  float prior_shape, prior_rate;
  float *pr_shape, *pr_rate, *pr_Ev;   // GPArray xi / eta (hier)
  float pr_prior_shape, pr_prior_rate;
  int bias;
  const float *Tb;
  float *Tbdirect;
  float *b_shape, *b_rate, *b_Ev, *b_Elog;
  float2 *aux;
  float bias_prior_shape, bias_rate_total; // rate prior + (m or n_global)
  float *colsum_partial;               // [gridDim.x x Kp]
  // dense-head plan: the split-bf16 copy of A (row stride split_ld) that head_kernel reads by TMA
  __nv_bfloat16 *split_hi, *split_lo;
  uint32_t split_ld;
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

__device__ __forceinline__ float floor30(float v) { return v > 0.f ? v : 1e-30f; } // make_nonzero, gpbase.hh:27-44

__global__ void __launch_bounds__(kUpdateWarps * 32) update_kernel(const UpdateArgs a)
{
  extern __shared__ float cs[]; // [kUpdateWarps][Kp] column partial sums of Ev
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *mycs = cs + (size_t)w * a.Kp;
  for (uint32_t k = lane; k < a.Kp; k += 32) mycs[k] = 0.f;
  const bool use_direct = *a.direct_flag != 0u;
  const uint32_t warps_total = gridDim.x * kUpdateWarps;

  // the global-rate (GPMatrixGR) vector is written once, by block 0
  if (!a.hier && blockIdx.x == 0)
    for (uint32_t k = threadIdx.x; k < a.Kp; k += blockDim.x)
      a.rate[k] = k < a.K ? a.prior_rate + a.colsum_other[k] : 1.f;

  for (uint32_t r = blockIdx.x * kUpdateWarps + w; r < a.R; r += warps_total) {
    const size_t base = (size_t)r * a.ld;
    const float rprior = a.hier ? a.pr_Ev[r] : a.prior_rate;
    float mx = -CUDART_INF_F, rowsum = 0.f;
    for (uint32_t k = lane; k < a.Kp; k += 32) {
      if (k < a.K) {
        float s = a.prior_shape + a.A[base + k] * a.T[base + k];
        if (use_direct) {
          s += a.Tdirect[base + k];
          a.Tdirect[base + k] = 0.f;
        }
        const float rt = rprior + a.colsum_other[k];
        const float sa = floor30(s), rb = floor30(rt);
        const float ev = sa / rb;
        const float el = digammaf(sa) - logf(rb);
        a.shape[base + k] = s;
        if (a.hier) a.rate[base + k] = rt;
        a.Ev[base + k] = ev;
        a.Elog[base + k] = el;
        mx = fmaxf(mx, el);
        rowsum += ev;
        mycs[k] += ev;
      } else {
        a.shape[base + k] = 0.f;
        if (a.hier) a.rate[base + k] = 1.f;
        a.Ev[base + k] = 0.f;
        a.Elog[base + k] = -CUDART_INF_F;
      }
    }
    mx = warp_max(mx);
    rowsum = warp_sum(rowsum);
    for (uint32_t k = lane; k < a.Kp; k += 32) {
      const float av = k < a.K ? expf(a.Elog[base + k] - mx) : 0.f;
      a.A[base + k] = av;
      if (a.split_hi != nullptr) { // x = hi + lo in bf16 (representation error 2^-18)
        const __nv_bfloat16 h = __float2bfloat16_rn(av);
        a.split_hi[(size_t)r * a.split_ld + k] = h;
        a.split_lo[(size_t)r * a.split_ld + k] = __float2bfloat16_rn(av - __bfloat162float(h));
      }
    }
    if (lane == 0) {
      a.shift[r] = mx;
      if (a.hier) {
        // hgaprec.cc:1399-1405: shape = a' + K a', rate = b' + sum_k E[theta_uk]
        const float ps = a.pr_prior_shape + (float)a.K * a.pr_prior_shape;
        const float pr = a.pr_prior_rate + rowsum;
        a.pr_shape[r] = ps;
        a.pr_rate[r] = pr;
        a.pr_Ev[r] = floor30(ps) / floor30(pr);
      }
      if (a.bias) {
        float bs = a.bias_prior_shape + a.aux[r].x * a.Tb[r];
        if (use_direct) {
          bs += a.Tbdirect[r];
          a.Tbdirect[r] = 0.f;
        }
        const float br = a.bias_rate_total;
        const float sa = floor30(bs), rb = floor30(br);
        const float bel = digammaf(sa) - logf(rb);
        a.b_shape[r] = bs;
        a.b_rate[r] = br;
        a.b_Ev[r] = sa / rb;
        a.b_Elog[r] = bel;
        const float2 ax = make_float2(expf(bel - mx), expf(-mx));
        a.aux[r] = ax;
        if (a.split_hi != nullptr) { // the two bias columns of the dense head's operand copy
          const size_t o = (size_t)r * a.split_ld + a.Kp;
          const __nv_bfloat16 hx = __float2bfloat16_rn(ax.x), hy = __float2bfloat16_rn(ax.y);
          a.split_hi[o] = hx; a.split_hi[o + 1] = hy;
          a.split_lo[o] = __float2bfloat16_rn(ax.x - __bfloat162float(hx));
          a.split_lo[o + 1] = __float2bfloat16_rn(ax.y - __bfloat162float(hy));
        }
      }
    }
  }
  __syncthreads();
  float *dst = a.colsum_partial + (size_t)blockIdx.x * a.Kp;
  for (uint32_t k = threadIdx.x; k < a.Kp; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kUpdateWarps; ++q) s += cs[(size_t)q * a.Kp + k];
    dst[k] = s;
  }
}

// colsum[k] = sum over blocks of partial[b][k], accumulated in double in a fixed
// order (one block per 32 columns, 8 row groups per block); also clears the side's
// direct_flag for the next iteration.
__global__ void __launch_bounds__(256) colsum_finalize_kernel(const float *partial, uint32_t nblocks, uint32_t Kp,
                                                             float *colsum, uint32_t *direct_flag)
{
  __shared__ double part[8][32];
  const uint32_t kx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const uint32_t k = blockIdx.x * 32 + kx;
  double s = 0.0;
  if (k < Kp)
    for (uint32_t b = g; b < nblocks; b += 8) s += (double)partial[(size_t)b * Kp + k];
  part[g][kx] = s;
  __syncthreads();
  if (g == 0 && k < Kp) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][kx];
    colsum[k] = (float)t;
  }
  if (direct_flag != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *direct_flag = 0u;
}

// column sums of an existing Ev matrix (after hpf_set_state)
__global__ void __launch_bounds__(kUpdateWarps * 32) colsum_partial_kernel(const float *Ev, uint32_t R, uint32_t Kp,
                                                                          uint32_t ld, float *partial)
{
  extern __shared__ float cs[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *mycs = cs + (size_t)w * Kp;
  for (uint32_t k = lane; k < Kp; k += 32) mycs[k] = 0.f;
  const uint32_t warps_total = gridDim.x * kUpdateWarps;
  for (uint32_t r = blockIdx.x * kUpdateWarps + w; r < R; r += warps_total)
    for (uint32_t k = lane; k < Kp; k += 32) mycs[k] += Ev[(size_t)r * ld + k];
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < Kp; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kUpdateWarps; ++q) s += cs[(size_t)q * Kp + k];
    partial[(size_t)blockIdx.x * Kp + k] = s;
  }
}

// ---------------------------------------------------------------------------
// state import / export (host fp64 row-major, stride K  <->  device fp32, stride ld)
// ---------------------------------------------------------------------------
// matrix import; when Elog != nullptr also derives shift and A in double.
__global__ void __launch_bounds__(kUpdateWarps * 32) import_matrix_kernel(const double *src, uint32_t R, uint32_t K, uint32_t Kp,
                                                                         uint32_t ld, float *dst, float pad)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warps_total = gridDim.x * kUpdateWarps;
  for (uint32_t r = blockIdx.x * kUpdateWarps + w; r < R; r += warps_total)
    for (uint32_t k = lane; k < Kp; k += 32)
      dst[(size_t)r * ld + k] = k < K ? (float)src[(size_t)r * K + k] : pad;
}

__global__ void __launch_bounds__(kUpdateWarps * 32) import_elog_kernel(const double *src, uint32_t R, uint32_t K, uint32_t Kp,
                                                                       uint32_t ld, float *Elog, float *A, float *shift)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warps_total = gridDim.x * kUpdateWarps;
  for (uint32_t r = blockIdx.x * kUpdateWarps + w; r < R; r += warps_total) {
    double mx = -CUDART_INF;
    for (uint32_t k = lane; k < K; k += 32) mx = fmax(mx, src[(size_t)r * K + k]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    // the stored shift is the fp32 value the sweep's aux terms are built from
    const float mxf = (float)mx;
    for (uint32_t k = lane; k < Kp; k += 32) {
      if (k < K) {
        const double e = src[(size_t)r * K + k];
        Elog[(size_t)r * ld + k] = (float)e;
        A[(size_t)r * ld + k] = (float)exp(e - (double)mxf);
      } else {
        Elog[(size_t)r * ld + k] = -CUDART_INF_F;
        A[(size_t)r * ld + k] = 0.f;
      }
    }
    if (lane == 0) shift[r] = mxf;
  }
}

__global__ void import_vector_kernel(const double *src, uint32_t n, float *dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

// aux[r] = {exp(Elogbias_r - shift_r), exp(-shift_r)} from uploaded state
__global__ void build_aux_kernel(const float *b_Elog, const float *shift, uint32_t R, float2 *aux)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) aux[r] = make_float2(expf(b_Elog[r] - shift[r]), expf(-shift[r]));
}

__global__ void __launch_bounds__(kUpdateWarps * 32) export_matrix_kernel(const float *src, uint32_t R, uint32_t K, uint32_t ld,
                                                                         double *dst)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warps_total = gridDim.x * kUpdateWarps;
  for (uint32_t r = blockIdx.x * kUpdateWarps + w; r < R; r += warps_total)
    for (uint32_t k = lane; k < K; k += 32) dst[(size_t)r * K + k] = (double)src[(size_t)r * ld + k];
}

__global__ void export_vector_kernel(const float *src, uint32_t n, double *dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}

// Elog of a GPArray (xi / eta) is only materialised on export: psi(shape) - log(rate)
__global__ void export_gparray_elog_kernel(const float *shape, const float *rate, uint32_t n, double *dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)(digammaf(floor30(shape[i])) - logf(floor30(rate[i])));
}

// ---------------------------------------------------------------------------
// K6: held-out log-likelihood (hgaprec.cc:1439-1470, 1503-1570).  One group of
// G lanes per (u, i, y) triple; fp32 dot product, fp64 log-likelihood and sum.
// ---------------------------------------------------------------------------
struct HeldoutArgs {
  const uint32_t *u, *i;
  const uint8_t *y;
  uint64_t npairs;
  const float *Et, *Eb;   // Ev matrices [. x ld]
  const float *Etb, *Ebb; // bias Ev (or nullptr)
  uint32_t K4, ld4;
  int binary;
  const double *logfact;  // [256]  log(y!) as the reference sums it (hgaprec.cc:1563-1570)
  double *block_sums;     // [gridDim.x]
};

template <int G, int V>
__global__ void __launch_bounds__(kSweepThreads) heldout_kernel(const HeldoutArgs a)
{
  __shared__ double wsum[kSweepThreads / 32];
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), w = threadIdx.x >> 5;
  const uint64_t groups_total = (uint64_t)gridDim.x * kSweepThreads / G;
  double local = 0.0;
  const uint64_t g0 = ((uint64_t)blockIdx.x * kSweepThreads + threadIdx.x) / G;
  const uint64_t rounds = (a.npairs + groups_total - 1) / groups_total;
  for (uint64_t rd = 0; rd < rounds; ++rd) {
    const uint64_t p = g0 + rd * groups_total;
    const bool active = p < a.npairs;
    const uint32_t uu = active ? a.u[p] : 0u, ii = active ? a.i[p] : 0u;
    const float4 *tp = reinterpret_cast<const float4 *>(a.Et) + (size_t)uu * a.ld4;
    const float4 *bp = reinterpret_cast<const float4 *>(a.Eb) + (size_t)ii * a.ld4;
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t q = gl + v * G;
      if (active && q < a.K4) {
        const float4 t4 = ldg4(tp + q), b4 = ldg4(bp + q);
        dot = fmaf(t4.x, b4.x, dot);
        dot = fmaf(t4.y, b4.y, dot);
        dot = fmaf(t4.z, b4.z, dot);
        dot = fmaf(t4.w, b4.w, dot);
      }
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if (active && gl == 0) {
      double s = (double)dot;
      if (a.Etb != nullptr) s += (double)a.Etb[uu] + (double)a.Ebb[ii];
      if (s < 1e-30) s = 1e-30;
      const uint32_t yy = a.y[p];
      if (a.binary)
        local += yy == 0 ? -s : log(1.0 - exp(-s));
      else
        local += (double)yy * log(s) - s - a.logfact[yy];
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if (lane == 0) wsum[w] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int q = 0; q < kSweepThreads / 32; ++q) s += wsum[q];
    a.block_sums[blockIdx.x] = s;
  }
}

__global__ void sum_blocks_kernel(const double *block_sums, uint32_t nblocks, double *out)
{
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (uint32_t b = 0; b < nblocks; ++b) s += block_sums[b];
    *out = s;
  }
}

// ---------------------------------------------------------------------------
// ratings set-up.  Both sweeps walk the nonzeros grouped by (tile of the
// GATHERED side, row): a tile is a contiguous range of gathered rows whose
// factor rows fit the L2 budget, so every gather of a tile's pass is an L2 hit.
// The orderings are built with cub stable radix sorts of a permutation; these
// kernels produce the keys and apply the permutation.
// ---------------------------------------------------------------------------
__global__ void expand_rows_kernel(const uint64_t *row_ptr, uint32_t nrows, uint64_t nnz, uint32_t *row_of)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  uint32_t lo = 0, hi = nrows; // largest r with row_ptr[r] <= j
  while (hi - lo > 1) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (row_ptr[mid] <= j) lo = mid; else hi = mid;
  }
  row_of[j] = lo;
}

__global__ void iota_kernel(uint32_t *p, uint64_t n)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) p[j] = (uint32_t)j;
}

// key[j] = src[perm[j]] / div   (div == 1: plain gather)
__global__ void gather_key_kernel(const uint32_t *perm, const uint32_t *src, uint32_t div, uint64_t nnz, uint32_t *key)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) key[j] = src[perm[j]] / div;
}

// apply the final permutation: gathered-side index, rating, and the composite
// key tile * R + row that the run pointers are derived from
__global__ void apply_perm_kernel(const uint32_t *perm, const uint32_t *row, const uint32_t *col, const uint8_t *y,
                                  uint32_t tile_cols, uint32_t R, uint64_t nnz, uint32_t *out_idx, uint8_t *out_y,
                                  uint32_t *out_key)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const uint32_t p = perm[j];
  const uint32_t cc = col[p];
  out_idx[j] = cc;
  if (y != nullptr) out_y[j] = y[p];
  out_key[j] = (cc / tile_cols) * R + row[p];
}

// run_ptr[c] = first position whose sorted key is >= c, for c in [0, nkeys]
__global__ void run_ptr_kernel(const uint32_t *sorted_key, uint64_t nnz, uint32_t nkeys, uint64_t *run_ptr)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > nnz) return;
  const uint32_t prev = j == 0 ? 0u : sorted_key[j - 1] + 1u;
  const uint32_t cur = j == nnz ? nkeys + 1u : sorted_key[j] + 1u;
  for (uint32_t c = prev; c < cur && c <= nkeys; ++c) run_ptr[c] = j;
}

// any index >= limit?  (argument check of hpf_set_ratings_csr, done on the device)
__global__ void check_range_kernel(const uint32_t *idx, uint64_t nnz, uint32_t limit, uint32_t *bad)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz && idx[j] >= limit) atomicMax(bad, idx[j]);
}

} // namespace hpf

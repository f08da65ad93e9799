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

__device__ __forceinline__ float floor30(float v) { return v > 0.f ? v : 1e-30f; } // make_nonzero, gpbase.hh:27-44

// digamma for x > 0 in fp32: upward recurrence to x >= 6, then the Stirling
// series.  Replaces gsl_sf_psi at gpbase.hh:260,593,923.
// The special-function unit's approximations, without the denormal rescaling nvcc wraps around them (every argument
// here is a normal number: shapes and rates are floored at 1e-30): rcp 1 ulp, lg2 absolute error 2^-22, ex2 2 ulp.
__device__ __forceinline__ float rcp_fast(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float log_fast(float x) // absolute error <= 2e-7 + 1 ulp: what E[log v] = psi(shape) - log(rate) can carry
{
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.693147182f;
}
__device__ __forceinline__ float exp_fast(float x) // x <= 0; relative error 2 ulp + |x| * 2^-24
{
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.44269502f));
  return r;
}

// CONVERGED: all 32 lanes of the warp are here, so the recurrence is skipped by a warp vote when no lane needs it
// (rows of large shapes) and costs no divergence when some do; same arithmetic, same bits either way.
template <bool CONVERGED = false> __device__ __forceinline__ float digammaf(float x)
{
  // psi(x) = psi(x + 6) - sum_{i<6} 1/(x + i); the sum is formed as ONE quotient p/q
  // (q = prod (x+i), p = sum of the products leaving one factor out): one division instead of six.
  // No per-lane branch: lanes of a warp fall on both sides of 6, and the four values a lane works on overlap.
  const bool small = x < 6.f;
  float r = 0.f;
  if (!CONVERGED || __any_sync(0xffffffffu, small)) {
    float q = x, p = 1.f;
#pragma unroll
    for (int i = 1; i < 6; ++i) {
      const float t = x + (float)i;
      p = fmaf(p, t, q);
      q *= t;
    }
    r = small ? -p * rcp_fast(q) : 0.f; // (x >= 6: p / q may be inf / inf -- never selected)
  }
  x = small ? x + 6.f : x;
  const float inv = rcp_fast(x);
  const float f = inv * inv;
  const float t = f * (-1.f / 12.f + f * (1.f / 120.f + f * (-1.f / 252.f + f * (1.f / 240.f))));
  return r + log_fast(x) - 0.5f * inv + t;
}

// E[log v] = psi(shape) - log(rate) as update_kernel, derive_kernel and the exact fallback's load_elog4 all form it
// (the same operations, so the three agree bit for bit)
template <bool CONVERGED = false> __device__ __forceinline__ float elog_of(float shape_floored, float rate_floored)
{
  return digammaf<CONVERGED>(shape_floored) - log_fast(rate_floored);
}

// ---------------------------------------------------------------------------
// K1: phi sweep.  One GROUP of G lanes owns one work segment (<= seg_len
// consecutive nonzeros of ONE row); 32/G segments advance in lock-step per warp.
// Lane g of a group holds float4 #(g + v*G), v < V, of the row-side factor row
// and of the accumulator; per nonzero the group gathers the column-side row with
// 128-bit loads, forms Z with a G-lane xor-shuffle reduction and accumulates
// (y/Z) * B.  Replaces hgaprec.cc:1340-1366 / 928-942 / 1227-1248.
// ---------------------------------------------------------------------------
// E[log v] of a parameter matrix as the exact fallback needs it.  The dense update does not store it per iteration
// (nothing on the fast path reads it): while `valid` is 0 it is recomputed from the stored shape and the two rate terms
// with the operations update_kernel / derive_kernel use, so the value is bit for bit what they would have stored.
// After hpf_set_state (expectations given by the caller, not functions of shape and rate) and after derive_kernel
// the array itself is valid.
struct ElogSrc {
  const float *Elog, *shape;  // [. x ld]
  const float *rate_row;      // [R]  (hier)
  const float *rate_col;      // [Kp] (hier: column term; GR: the rate vector)
  int valid, hier;
};
__device__ __forceinline__ float4 load_elog4(const ElogSrc &e, uint32_t row, uint32_t q, uint32_t ld4, uint32_t K)
{
  if (e.valid) return reinterpret_cast<const float4 *>(e.Elog)[(size_t)row * ld4 + q];
  const float4 sh = reinterpret_cast<const float4 *>(e.shape)[(size_t)row * ld4 + q];
  const float4 ct = reinterpret_cast<const float4 *>(e.rate_col)[q];
  const float rr = e.hier ? e.rate_row[row] : 0.f;
  float4 out;
  const float *s = reinterpret_cast<const float *>(&sh), *c = reinterpret_cast<const float *>(&ct);
  float *o = reinterpret_cast<float *>(&out);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = q * 4 + j < K ? elog_of(floor30(s[j]), floor30(e.hier ? rr + c[j] : c[j])) : -CUDART_INF_F;
  return out;
}

struct SweepArgs {
  const uint4 *seg;        // {begin_lo, begin_hi, row, len}
  const uint32_t *seg_out; // < R: row of T;  >= R: slot (out - R) of Tpart
  uint32_t nsegs;
  uint32_t R;              // rows on the row side
  const uint32_t *idx;     // per nonzero: row number on the column side; with `packed` the rating sits in its top byte
  const uint8_t *y;        // per nonzero rating, or nullptr (all ones; unused with `packed`)
  uint32_t packed;         // the column side has < 2^24 rows: one word (index | rating << 24) per nonzero, and one
                           // shuffle fewer per nonzero in a loop that is bound by the L1 data pipe
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
  ElogSrc ElogRow, ElogCol;         // E[log v] of both sides (padding = -inf)
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
// Takes the kernel's own parameter block by reference (a __grid_constant__ parameter may have its address taken
// without a local copy), so the rare call costs the hot loop no registers.
template <int G, int V, bool BIAS>
__device__ __noinline__ void sweep_slow_path(const SweepArgs &a, uint32_t row, uint32_t c, float yv, int lane)
{
  const int gl = lane & (G - 1);
  const uint32_t mask = group_mask<G>(lane);
  float4 xs[V]; // x_k = Elog_row_k + Elog_col_k of this lane's float4 slots
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint32_t q = gl + v * G;
    xs[v] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    if (q < a.K4) {
      const float4 r4 = load_elog4(a.ElogRow, row, q, a.ld4, a.K), c4 = load_elog4(a.ElogCol, c, q, a.ld4, a.K);
      xs[v] = make_float4(r4.x + c4.x, r4.y + c4.y, r4.z + c4.z, r4.w + c4.w);
      mx = fmaxf(mx, fmaxf(fmaxf(xs[v].x, xs[v].y), fmaxf(xs[v].z, xs[v].w)));
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
    if (q < a.K4) sum += expf(xs[v].x - mx) + expf(xs[v].y - mx) + expf(xs[v].z - mx) + expf(xs[v].w - mx);
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
      const uint32_t k0 = q * 4;
      if (k0 + 0 < a.K) atomicAdd(td + k0 + 0, sc * expf(xs[v].x - mx));
      if (k0 + 1 < a.K) atomicAdd(td + k0 + 1, sc * expf(xs[v].y - mx));
      if (k0 + 2 < a.K) atomicAdd(td + k0 + 2, sc * expf(xs[v].z - mx));
      if (k0 + 3 < a.K) atomicAdd(td + k0 + 3, sc * expf(xs[v].w - mx));
    }
  }
  if (gl == 0) {
    if (BIAS) atomicAdd(a.Tbdirect + row, sc * expf(xbr - mx));
    atomicAdd(a.slow_count, 1ull);
    *a.direct_flag = 1u;
  }
}

// Second walk over the warp's segments for the nonzeros whose Z left the fp32 range.  The hot loop only COUNTS them
// (a call inside it costs registers and local-memory traffic in a kernel that is bound by the L1 data pipe); when a
// warp has seen any, it comes here once: same chunks, same arithmetic for Z (so the same nonzeros are found), and
// the exact log-domain path for each of them.  Called by all 32 lanes.
template <int G, int V, bool BIAS>
__device__ __noinline__ void sweep_redo_slow(const SweepArgs &a, uint32_t row, uint64_t begin, uint32_t len, uint32_t maxlen, bool have, int lane)
{
  const int gl = lane & (G - 1);
  float4 ar[V];
  const float4 *rp = reinterpret_cast<const float4 *>(a.Arow) + (size_t)row * a.ld4;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint32_t q = gl + v * G;
    ar[v] = (have && len > 0u && q < a.K4) ? ldg4(rp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float2 raux = make_float2(0.f, 0.f);
  if (BIAS && have) raux = __ldg(a.row_aux + row);
  const uint32_t *ip = a.idx + begin;
  const uint8_t *yp = a.y ? a.y + begin : nullptr;
  const float4 *acol = reinterpret_cast<const float4 *>(a.Acol);
  const uint32_t q_last = min((uint32_t)(gl + (V - 1) * G), a.K4 - 1u);
  for (uint32_t j0 = 0; j0 < maxlen; j0 += G) {
    const uint32_t jj = j0 + gl;
    const uint32_t cword = (jj < len) ? ip[jj] : 0u;
    const uint32_t cbuf = a.packed ? (cword & 0x00ffffffu) : cword;
    const float ybuf = (jj < len) ? (a.packed ? (float)(cword >> 24) : (yp != nullptr ? (float)yp[jj] : 1.f)) : 0.f;
    const int tmax = (int)min((uint32_t)G, maxlen - j0);
    for (int t = 0; t < tmax; ++t) {
      const uint32_t c = __shfl_sync(0xffffffffu, cbuf, t, G);
      const float yv = __shfl_sync(0xffffffffu, ybuf, t, G);
      float4 b[V];
      const float4 *cp = acol + (size_t)c * a.ld4;
#pragma unroll
      for (int v = 0; v < V - 1; ++v) b[v] = ldg4(cp + gl + v * G);
      b[V - 1] = ldg4(cp + q_last);
      float dot = dot_rows<V>(ar, b);
#pragma unroll
      for (int off = G / 2; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
      float z = dot;
      if (BIAS) {
        const float2 caux = __ldg(a.col_aux + c);
        z += raux.x * caux.y + caux.x * raux.y;
      }
      const bool ok = z > kZMin && z < kZMax;
      if (!ok && yv != 0.f) sweep_slow_path<G, V, BIAS>(a, row, c, yv, lane);
    }
  }
}

template <int G, int V, bool BIAS>
__global__ void __launch_bounds__(kSweepThreads, HPF_SWEEP_MINBLOCKS) sweep_kernel(const __grid_constant__ SweepArgs a)
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
      ar[v] = (have && len > 0u && q < a.K4) ? ldg4(rp + q) : make_float4(0.f, 0.f, 0.f, 0.f); // an empty row only clears its T row
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
  const bool packed = a.packed != 0u;
  uint32_t nslow = 0;
  // Lanes past the last float4 of a row re-read that float4 (their row-side value is 0 and their sums are
  // never stored), so the loop carries no predicates.  Only the last of the V slots can be out of range.
  const uint32_t q_last = min((uint32_t)(gl + (V - 1) * G), a.K4 - 1u);

  // chunks of G nonzeros: every lane of the group fetches one (index, rating), then walks the chunk.  The loop is
  // deliberately NOT software-pipelined: the kernel sits at the practical ceiling of the L1 data pipe (4 wavefronts per
  // gathered 512-byte row + 1.25 for the shuffles; ncu lsu write-back ~60 % busy, the same as the pure-gather
  // micro-benchmark at its 22 TB/s peak), and prefetching the next row only costs registers and occupancy
  // (measured 10-65 % slower: profiles/r02d_sweep_pipeline_experiment.txt)
  for (uint32_t j0 = 0; j0 < maxlen; j0 += G) {
    const uint32_t jj = j0 + gl;
    const uint32_t cbuf = (jj < len) ? ld_stream_u32(ip + jj) : 0u; // 0 past the end: a valid row (and rating 0 when packed)
    float ybuf = 0.f;                                                // 0: no contribution
    if (!packed && jj < len) ybuf = yp != nullptr ? (float)ld_stream_u8(yp + jj) : 1.f;
    // the last chunk of the warp's longest segment may be short: walk only what exists (warp-uniform bound).  With
    // users sharded over GPUs most item rows hold a handful of local nonzeros, and a full chunk would gather row 0
    // for every missing one.
    const int tmax = (int)min((uint32_t)G, maxlen - j0);
    HPF_UNROLL(HPF_UNROLL_T)
    for (int t = 0; t < tmax; ++t) {
      const uint32_t cw = __shfl_sync(0xffffffffu, cbuf, t, G);
      uint32_t c = cw;
      float yv;
      if (packed) { // warp-uniform
        c = cw & 0x00ffffffu;
        yv = (float)(cw >> 24);
      } else {
        yv = __shfl_sync(0xffffffffu, ybuf, t, G);
      }
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
      nslow += (!ok && yv != 0.f) ? 1u : 0u; // Z left the fp32 range: contributes nothing here, redone exactly below
    }
  }

  if (__any_sync(0xffffffffu, nslow != 0u)) sweep_redo_slow<G, V, BIAS>(a, row, begin, len, maxlen, have, lane);

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
// K2/K3/K4/K5: dense row update, one warp per row, 128-bit accesses (lane l owns
// float4 #(l + 32 v), v < V, of the row).  Fuses, for one parameter matrix:
// shape = prior + A*T (the add_slice sums), rate (set_prior_rate +
// update_rate_next, gpbase.hh:163-173,218-223 / GR 560-564), swap (240-246),
// compute_expectations (248-262), sum_rows / sum_cols (264-280), the GPArray
// xi/eta step (hgaprec.cc:1398-1414, gpbase.hh:877-925), the bias step
// (hgaprec.cc:1388-1396) and the next iteration's shifted exponentials.
//
// Per iteration only what the NEXT iteration reads is stored: A (the sweep
// operand) and the shape.  The rate is rank-2 structured -- rate_uk = E[xi_u] +
// sum_i E[beta_ik] (hier) or a K-vector (GR) -- so the kernel keeps its two terms
// (rate_row[r], rate_col[k]) instead of an R x K matrix; E[v] =
// shape / rate and E[log v] = psi(shape) - log(rate) are only needed by the
// report-window consumers (held-out ll, top-N, ELBO, hpf_get_state) and by the
// exact fallback: derive_kernel materialises rate, E[v] and E[log v] from the
// shape and the two terms on demand, and the fallback recomputes the E[log v] it
// needs (load_elog4) -- all with the same fp32 operations, so the values are the
// ones this kernel used for its row / column sums and its shifted exponentials.
// ---------------------------------------------------------------------------
struct UpdateArgs {
  uint32_t R, K, Kp, K4, ld4; // Kp: K rounded up to 4; K4 = Kp / 4; ld4: row stride in float4
  const float4 *T;
  float4 *Tdirect;
  const uint32_t *direct_flag;
  const float *direct_flag_all;        // multi-GPU exact mode: the all-reduced item-side flag (or nullptr)
  float4 *A, *shape;
  float *shift;                        // [R]
  int hier;
  const float *colsum_other;           // [Kp]  sum over the OTHER side's rows of Ev
  float *rate_vec;                     // [Kp]  written by block 0 -- GR: the rate vector; hier: a copy of colsum_other
  float *rate_row;                     // [R]   hier: the E[xi] / E[eta] this update used
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

__device__ __forceinline__ float &f4at(float4 &v, int j) { return reinterpret_cast<float *>(&v)[j]; }
__device__ __forceinline__ float f4get(const float4 &v, int j) { return reinterpret_cast<const float *>(&v)[j]; }

// resident blocks the register budget is set for: V float4 per lane of row operands, twice (two rows in flight)
constexpr int update_min_blocks(int V) { return V <= 2 ? 3 : (V <= 4 ? 2 : 1); }

// FULL: K is a multiple of 4, so every float4 a lane works on holds four live factors and the per-value bounds
// checks (one reconvergence region each, which kept the four chains from overlapping) disappear
template <int V, bool FULL>
__global__ void __launch_bounds__(kUpdateWarps * 32, update_min_blocks(V)) update_kernel(const UpdateArgs a)
{
  extern __shared__ float cs[]; // [kUpdateWarps][Kp] column partial sums of Ev
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool use_direct = *a.direct_flag != 0u || (a.direct_flag_all != nullptr && *a.direct_flag_all != 0.f);
  const uint32_t warps_total = gridDim.x * kUpdateWarps;

  // block 0 keeps the column term of the rate for derive_kernel: the global-rate (GPMatrixGR) vector itself, or
  // (hier) the column sums this update used
  if (blockIdx.x == 0)
    for (uint32_t k = threadIdx.x; k < a.Kp; k += blockDim.x)
      a.rate_vec[k] = a.hier ? a.colsum_other[k] : (k < a.K ? a.prior_rate + a.colsum_other[k] : 1.f);

  // this lane's columns: the rate's column term and the running column sums stay in registers
  float4 cterm[V], csum[V];
  bool act[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint32_t q = lane + 32 * v;
    act[v] = q < a.K4;
    csum[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    cterm[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act[v]) cterm[v] = __ldg(reinterpret_cast<const float4 *>(a.colsum_other) + q);
  }

  uint32_t r = blockIdx.x * kUpdateWarps + w;
  float4 av[V], tv[V];
  if (r < a.R) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      av[v] = tv[v] = make_float4(0.f, 0.f, 0.f, 0.f); // lanes past the row's end work on zeros (all 32 stay converged)
      if (act[v]) {
        av[v] = a.A[(size_t)r * a.ld4 + lane + 32 * v];
        tv[v] = __ldg(a.T + (size_t)r * a.ld4 + lane + 32 * v);
      }
    }
  }
  for (; r < a.R; r += warps_total) {
    const size_t base = (size_t)r * a.ld4;
    // the next row's operands are requested before this row's arithmetic (two rows in flight per warp)
    const uint32_t rn = r + warps_total;
    float4 an[V], tn[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      an[v] = tn[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rn < a.R && act[v]) {
        an[v] = a.A[(size_t)rn * a.ld4 + lane + 32 * v];
        tn[v] = __ldg(a.T + (size_t)rn * a.ld4 + lane + 32 * v);
      }
    }
    const float rprior = a.hier ? a.pr_Ev[r] : a.prior_rate;
    float mx = -CUDART_INF_F, rowsum = 0.f;
    float4 el[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t q = lane + 32 * v;
      float4 sh;
      float4 td = make_float4(0.f, 0.f, 0.f, 0.f);
      if (use_direct && act[v]) {
        td = a.Tdirect[base + q];
        a.Tdirect[base + q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { // every lane, live value or not: the warp stays converged for digammaf's vote
        const bool live = act[v] && (FULL || q * 4 + j < a.K);
        float s = a.prior_shape + f4get(av[v], j) * f4get(tv[v], j);
        if (use_direct) s += f4get(td, j);
        const float rt = rprior + f4get(cterm[v], j);
        const float sa = floor30(s), rb = floor30(rt);
        const float ev = live ? sa * rcp_fast(rb) : 0.f; // 1 ulp: feeds the row / column sums only (derive_kernel repeats it)
        const float e0 = elog_of<true>(sa, rb); // outside the select: the vote inside needs every lane
        const float e = live ? e0 : -CUDART_INF_F;
        f4at(sh, j) = live ? s : 0.f;
        f4at(el[v], j) = e;
        mx = fmaxf(mx, e);
        rowsum += ev;
        f4at(csum[v], j) += ev;
      }
      if (act[v]) a.shape[base + q] = sh;
    }
    mx = warp_max(mx);
    rowsum = warp_sum(rowsum);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (!act[v]) continue;
      const uint32_t q = lane + 32 * v;
      float4 an4;
#pragma unroll
      for (int j = 0; j < 4; ++j) f4at(an4, j) = (FULL || q * 4 + j < a.K) ? exp_fast(f4get(el[v], j) - mx) : 0.f;
      a.A[base + q] = an4;
      if (a.split_hi != nullptr) { // x = hi + lo in bf16 (representation error 2^-18)
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h[j] = __float2bfloat16_rn(f4get(an4, j));
          l[j] = __float2bfloat16_rn(f4get(an4, j) - __bfloat162float(h[j]));
        }
        const size_t o = (size_t)r * a.split_ld + (size_t)q * 4;
        *reinterpret_cast<uint2 *>(a.split_hi + o) = *reinterpret_cast<const uint2 *>(h);
        *reinterpret_cast<uint2 *>(a.split_lo + o) = *reinterpret_cast<const uint2 *>(l);
      }
    }
    if (lane == 0) {
      a.shift[r] = mx;
      if (a.hier) {
        a.rate_row[r] = rprior;
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
#pragma unroll
    for (int v = 0; v < V; ++v) { av[v] = an[v]; tv[v] = tn[v]; }
  }
  // block partial of the column sums: warps in a fixed order
#pragma unroll
  for (int v = 0; v < V; ++v)
    if (act[v]) reinterpret_cast<float4 *>(cs + (size_t)w * a.Kp)[lane + 32 * v] = csum[v];
  __syncthreads();
  float *dst = a.colsum_partial + (size_t)blockIdx.x * a.Kp;
  for (uint32_t k = threadIdx.x; k < a.Kp; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kUpdateWarps; ++q) s += cs[(size_t)q * a.Kp + k];
    dst[k] = s;
  }
}

// rate and E[v] of a parameter matrix from its shape and the two rate terms the last update used
// (report-window consumers only; see update_kernel).  One thread per float4.
struct DeriveArgs {
  uint32_t R, K, K4, ld4;
  const float4 *shape;
  float4 *Ev, *Elog, *rate; // rate: [R x ld] (hier) or nullptr
  int hier;
  const float *rate_row;    // [R]   hier
  const float *rate_col;    // [Kp]  hier: the column sums the update used; GR: the rate vector
};
__global__ void __launch_bounds__(256) derive_kernel(const DeriveArgs a)
{
  const uint64_t total = (uint64_t)a.R * a.K4;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(e / a.K4), q = (uint32_t)(e - (uint64_t)r * a.K4);
    const size_t o = (size_t)r * a.ld4 + q;
    const float4 sh = a.shape[o];
    const float4 ct = __ldg(reinterpret_cast<const float4 *>(a.rate_col) + q);
    const float rr = a.hier ? __ldg(a.rate_row + r) : 0.f;
    float4 ev, rt, el;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (q * 4 + j < a.K) {
        const float t = a.hier ? rr + f4get(ct, j) : f4get(ct, j);
        const float sa = floor30(f4get(sh, j)), rb = floor30(t);
        f4at(rt, j) = t;
        f4at(ev, j) = sa * rcp_fast(rb); // the operation update_kernel used for its sums
        f4at(el, j) = elog_of(sa, rb);
      } else {
        f4at(rt, j) = 1.f;
        f4at(ev, j) = 0.f;
        f4at(el, j) = -CUDART_INF_F;
      }
    }
    a.Ev[o] = ev;
    a.Elog[o] = el;
    if (a.rate != nullptr) a.rate[o] = rt;
  }
}

// colsum[k] = sum over blocks of partial[b][k], accumulated in double in a fixed
// order (one block per 32 columns, 32 row groups per block); also clears the side's
// direct_flag for the next iteration.  Multi-GPU bookkeeping of the exact fallback
// (one thread): flag_out[0] = 1.0 if *flag_in (or *flag_in2, when given) is set (the
// theta update publishes the ITEM side's flag -- with a sharded beta update also its
// own -- into the tail of the reduce block), sticky[0] += reduced[0] (the
// beta update accumulates the all-reduced flag; read by hpf_iterate at its end).
constexpr int kFinalizeGroups = 32;
__global__ void __launch_bounds__(32 * kFinalizeGroups) colsum_finalize_kernel(const float *partial, uint32_t nblocks, uint32_t Kp,
                                                             float *colsum, uint32_t *direct_flag, const uint32_t *flag_in,
                                                             const uint32_t *flag_in2, float *flag_out, const float *reduced,
                                                             float *sticky)
{
  __shared__ double part[kFinalizeGroups][32];
  const uint32_t kx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const uint32_t k = blockIdx.x * 32 + kx;
  double s = 0.0;
  if (k < Kp)
    for (uint32_t b = g; b < nblocks; b += kFinalizeGroups) s += (double)partial[(size_t)b * Kp + k];
  part[g][kx] = s;
  __syncthreads();
  if (g == 0 && k < Kp) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < kFinalizeGroups; ++q) t += part[q][kx];
    colsum[k] = (float)t;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (flag_out != nullptr) flag_out[0] = (*flag_in != 0u || (flag_in2 != nullptr && *flag_in2 != 0u)) ? 1.f : 0.f;
    if (sticky != nullptr) sticky[0] += reduced[0];
    if (direct_flag != nullptr) *direct_flag = 0u;
  }
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
// row_of[j] = the row whose range [row_ptr[r], row_ptr[r + 1]) holds position j.  One warp per row (grid-stride),
// coalesced stores -- a binary search per nonzero costs 19 dependent loads each at Netflix scale (1.1 ms vs 0.2 ms)
__global__ void __launch_bounds__(256) expand_rows_kernel(const uint64_t *row_ptr, uint32_t nrows, uint64_t nnz, uint32_t *row_of)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
    const uint64_t b = row_ptr[r], e = min(row_ptr[r + 1], nnz);
    for (uint64_t j = b + lane; j < e; j += 32) row_of[j] = (uint32_t)r;
  }
}

__global__ void iota_kernel(uint32_t *p, uint64_t n)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) p[j] = (uint32_t)j;
}

// sort key and payload of one nonzero for an orientation: key = tile(col) * R + row, payload = its position.  One
// stable radix sort of (key, position) orders the nonzeros by (tile, row); orient_gather_kernel then fetches the
// gathered-side index and the rating through the sorted positions.  The ratings are not needed before that second
// step, so their host->device copy runs under the sort (hpf_set_ratings_csr uploads them last, on a second stream).
__global__ void orient_key_kernel(const uint32_t *row, const uint32_t *col, uint32_t tile_cols, uint32_t R, uint64_t nnz,
                                  uint32_t *key, uint32_t *pos)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  key[j] = (col[j] / tile_cols) * R + row[j];
  pos[j] = (uint32_t)j;
}
// pack != 0: one word per nonzero for the sweep, index | rating << 24 (the gathered side has < 2^24 rows)
__global__ void orient_gather_kernel(const uint32_t *pos, const uint32_t *col, const uint8_t *y, uint64_t nnz, uint32_t *out_idx,
                                     uint8_t *out_y, int pack)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const uint32_t p = pos[j];
  const uint32_t cc = col[p];
  if (pack) {
    out_idx[j] = cc | ((y != nullptr ? (uint32_t)y[p] : 1u) << 24);
  } else {
    out_idx[j] = cc;
    if (out_y != nullptr) out_y[j] = y[p];
  }
}
// the same packing for an orientation that needs no sort (the CSR as given)
__global__ void pack_kernel(const uint32_t *idx, const uint8_t *y, uint64_t nnz, uint32_t *out)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) out[j] = idx[j] | ((y != nullptr ? (uint32_t)y[j] : 1u) << 24);
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

// ---------------------------------------------------------------------------
// work lists, built on the device.  The nonzeros of one orientation arrive as
// RUNS: run (t, r) holds the nonzeros of row r whose gathered-side index lies in
// L2 tile t; run_ptr[t * R + r] is where it starts (ntiles == 1: the plain row
// pointer).  Every run is cut into segments of <= L nonzeros -- the unit one
// group of sweep_kernel lanes owns.  A row with one segment writes its T row
// itself; a row with several writes partial slots that combine_kernel adds in a
// fixed order.  A row without nonzeros gets one empty segment (it clears its T
// row); a skipped row (head item of the dense-head plan) gets none.  Segments are
// ordered by (chunk of rows, tile, descending length): blocks are scheduled in
// index order, so a launch walks one tile's L2-resident factor rows at a time,
// the 32/G segments a warp advances in lock-step have equal trip counts, and long
// work starts first.  CHUNKS are equal ranges of rows that are launched one after
// the other, so that the all-reduce of a finished chunk's T rows can run under the
// next chunk's sweep (multi-GPU; one chunk otherwise).
// ---------------------------------------------------------------------------
struct WlArgs {
  const uint64_t *run_ptr;   // [ntiles * R + 1]
  uint32_t R, ntiles, L;
  uint32_t chunk_rows;       // rows per chunk
  const uint32_t *skip_slot; // [R] != 0xffffffff: the row is skipped; or nullptr
  uint32_t *segcnt, *multicnt, *ismulti; // [R] counts (phase 1), then their exclusive scans in *_off
  uint32_t *seg_off, *first_off, *multi_off;
  uint4 *seg_u;              // unsorted segments {begin_lo, begin_hi, row, len}
  uint32_t *out_u, *key_u;   // unsorted output slot, sort key
  uint32_t *multi_row, *multi_first, *multi_cnt;
};

__global__ void wl_count_kernel(const WlArgs a)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  uint32_t cnt = 0;
  if (a.skip_slot == nullptr || a.skip_slot[r] == 0xffffffffu) {
    for (uint32_t t = 0; t < a.ntiles; ++t) {
      const size_t q = (size_t)t * a.R + r;
      const uint64_t len = a.run_ptr[q + 1] - a.run_ptr[q];
      cnt += (uint32_t)((len + a.L - 1) / a.L);
    }
    if (cnt == 0) cnt = 1; // the empty segment that clears the row's T
  }
  a.segcnt[r] = cnt;
  a.multicnt[r] = cnt > 1 ? cnt : 0u;
  a.ismulti[r] = cnt > 1 ? 1u : 0u;
}

__global__ void wl_emit_kernel(const WlArgs a)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  const uint32_t cnt = a.segcnt[r];
  if (cnt == 0) return;
  const uint32_t chunk = r / a.chunk_rows;
  uint32_t o = a.seg_off[r], j = 0;
  const uint32_t first = a.first_off[r];
  if (cnt > 1) {
    const uint32_t mp = a.multi_off[r];
    a.multi_row[mp] = r; a.multi_first[mp] = first; a.multi_cnt[mp] = cnt;
  }
  for (uint32_t t = 0; t < a.ntiles; ++t) {
    const size_t q = (size_t)t * a.R + r;
    const uint64_t b0 = a.run_ptr[q], len = a.run_ptr[q + 1] - b0;
    const uint32_t n = (uint32_t)((len + a.L - 1) / a.L);
    for (uint32_t s = 0; s < n; ++s, ++o, ++j) {
      const uint64_t sb = b0 + (uint64_t)s * a.L;
      const uint32_t sl = (uint32_t)min((uint64_t)a.L, len - (uint64_t)s * a.L);
      a.seg_u[o] = make_uint4((uint32_t)sb, (uint32_t)(sb >> 32), r, sl);
      a.out_u[o] = cnt == 1 ? r : a.R + first + j;
      a.key_u[o] = (chunk * a.ntiles + t) * (a.L + 1) + (a.L - sl);
    }
  }
  if (j == 0) { // no nonzeros at all: one empty segment in tile 0
    const uint64_t b0 = a.run_ptr[r];
    a.seg_u[o] = make_uint4((uint32_t)b0, (uint32_t)(b0 >> 32), r, 0u);
    a.out_u[o] = r;
    a.key_u[o] = (chunk * a.ntiles) * (a.L + 1) + a.L;
  }
}

// apply the sort: seg[j] = seg_u[perm[j]], seg_out[j] = out_u[perm[j]]
__global__ void wl_gather_kernel(const uint32_t *perm, const uint4 *seg_u, const uint32_t *out_u, const uint32_t *total,
                                 uint4 *seg, uint32_t *seg_out)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *total) return;
  const uint32_t p = perm[j];
  seg[j] = seg_u[p];
  seg_out[j] = out_u[p];
}

// totals and chunk boundaries -> info: {nsegs, nslots, nmulti, 0, seg bound[0..nchunks], multi bound[0..nchunks]}
__global__ void wl_info_kernel(const WlArgs a, const uint32_t *sorted_key, uint32_t nchunks, uint32_t *info)
{
  if (blockIdx.x != 0) return;
  __shared__ uint32_t tot[3];
  if (threadIdx.x == 0) {
    const uint32_t l = a.R - 1;
    tot[0] = a.seg_off[l] + a.segcnt[l];
    tot[1] = a.first_off[l] + a.multicnt[l];
    tot[2] = a.multi_off[l] + a.ismulti[l];
    info[0] = tot[0]; info[1] = tot[1]; info[2] = tot[2]; info[3] = 0;
  }
  __syncthreads();
  const uint32_t nsegs = tot[0], nmulti = tot[2];
  for (uint32_t c = threadIdx.x; c <= nchunks; c += blockDim.x) {
    // first sorted segment whose key belongs to chunk >= c
    const uint64_t want = (uint64_t)c * a.ntiles * (a.L + 1);
    uint32_t lo = 0, hi = nsegs;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if ((uint64_t)sorted_key[mid] < want) lo = mid + 1; else hi = mid;
    }
    info[4 + c] = c == nchunks ? nsegs : lo;
    // first multi-segment row >= c * chunk_rows (multi_row is ascending)
    const uint64_t wrow = (uint64_t)c * a.chunk_rows;
    lo = 0; hi = nmulti;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if ((uint64_t)a.multi_row[mid] < wrow) lo = mid + 1; else hi = mid;
    }
    info[4 + nchunks + 1 + c] = c == nchunks ? nmulti : lo;
  }
}

// any index >= limit?  (argument check of hpf_set_ratings_csr, done on the device)
__global__ void check_range_kernel(const uint32_t *idx, uint64_t nnz, uint32_t limit, uint32_t *bad)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz && idx[j] >= limit) atomicMax(bad, idx[j]);
}

} // namespace hpf

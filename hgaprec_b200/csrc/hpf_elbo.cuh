// hpf_elbo.cuh -- HGAPRec::logl() (src/hgaprec.cc:2160-2255) on the device: the variational lower bound the
// reference appends to logl.txt in every report window when run with -logl.  A diagnostic, not the hot path:
// plain kernels, fp64 accumulation, every block writes one partial sum and the host adds them in a fixed order.
//
// Per training nonzero (u, i, y) the reference forms phi = softmax_k(Elog theta_uk + Elog beta_ik [, the two bias
// logs]), scales it by y when y > 1 (2214-2215) and adds  sum_k y * phi_k * (x_k - log phi_k)  with the SCALED
// phi (2217-2224).  Since x_k - log(y softmax_k) = logsumexp(x) - log y for every slot and softmax sums to one,
// that is  y^2 * (logsumexp(x) - log y);  then it subtracts  E[theta_u] . E[beta_i]  (+ the two bias
// expectations, 2226-2231).  The Gamma terms are compute_elbo_term_helper of src/gpbase.hh:360-387 (GPMatrix),
// 717-741 (GPMatrixGR) and 951-969 (GPArray).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace hpf {
namespace elbo {

constexpr int kThreads = 256;

// psi(x), x > 0: recurrence up to x >= 6, then the asymptotic series (abs error < 1e-12)
__device__ __forceinline__ double digamma_d(double x)
{
  double r = 0.0;
  while (x < 6.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}

__device__ __forceinline__ double floor30_d(double v) { return v > 0.0 ? v : 1e-30; } // make_nonzero, gpbase.hh:27-44

// sum of `local` over the block, written to out[blockIdx.x] (fixed order: lanes by shuffle tree, warps in order)
__device__ __forceinline__ void block_store(double local, double *out)
{
  __shared__ double wsum[kThreads / 32];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int q = 0; q < kThreads / 32; ++q) s += wsum[q];
    out[blockIdx.x] = s;
  }
}

struct NnzArgs {
  const uint64_t *row_ptr;  // [n + 1] CSR row pointer (device copy kept for HPF_LOGL)
  const uint32_t *idx;      // item of every nonzero, CSR order
  const uint8_t *y;         // rating, or nullptr (all ones)
  uint32_t n, K, ld;
  const float *ElogT, *ElogB, *EvT, *EvB;          // [rows x ld]
  const float *ElogbT, *ElogbB, *EvbT, *EvbB;      // bias sets [rows], or nullptr
  double *block_sums;       // [gridDim.x]
};

// one warp per user row (warps stride over the users), lanes over the factors
__global__ void __launch_bounds__(kThreads) nnz_kernel(const NnzArgs a)
{
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * (uint32_t)kThreads + threadIdx.x) >> 5;
  const uint32_t nwarps = gridDim.x * (uint32_t)(kThreads / 32);
  const bool bias = a.ElogbT != nullptr;
  double local = 0.0; // lane 0 carries the sum
  for (uint32_t u = warp; u < a.n; u += nwarps) {
    const uint64_t beg = a.row_ptr[u], end = a.row_ptr[u + 1];
    const float *et = a.ElogT + (size_t)u * a.ld, *vt = a.EvT + (size_t)u * a.ld;
    const float xbu = bias ? a.ElogbT[u] : -CUDART_INF_F;
    for (uint64_t j = beg; j < end; ++j) {
      const uint32_t i = a.idx[j];
      const float *eb = a.ElogB + (size_t)i * a.ld, *vb = a.EvB + (size_t)i * a.ld;
      const float xbi = bias ? a.ElogbB[i] : -CUDART_INF_F;
      float mx = fmaxf(xbu, xbi);
      for (uint32_t k = lane; k < a.K; k += 32) mx = fmaxf(mx, et[k] + eb[k]);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float se = 0.f, dot = 0.f;
      for (uint32_t k = lane; k < a.K; k += 32) {
        se += expf(et[k] + eb[k] - mx);
        dot = fmaf(vt[k], vb[k], dot);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, off);
        dot += __shfl_xor_sync(0xffffffffu, dot, off);
      }
      if (lane == 0) {
        double sed = (double)se, sub = (double)dot;
        if (bias) {
          sed += exp((double)(xbu - mx)) + exp((double)(xbi - mx));
          sub += (double)a.EvbT[u] + (double)a.EvbB[i];
        }
        const double lse = (double)mx + log(sed);
        const double yd = a.y ? (double)a.y[j] : 1.0;
        local += (yd > 0.0 ? yd * yd * (lse - log(yd)) : 0.0) - sub; // y = 0: every y * phi_k factor is 0
      }
    }
  }
  block_store(local, a.block_sums);
}

struct GammaArgs {
  uint32_t R, K, ld;            // rows, columns, row stride of the four arrays
  const float *shape, *rate, *Ev, *Elog;
  int rate_is_vector;           // GPMatrixGR: rate is a K-vector shared by all rows
  double sprior, rprior, lg_sprior; // (shape, rate) prior and lgamma(shape prior)
  // hier theta / beta: the rate prior of row r is the xi / eta expectation the last set_prior_rate stored
  // (gpbase.hh:163-173): a / b and psi(a) - log(b) of the PREVIOUS xi / eta (shape, rate); nullptr: constants
  const float *pri_shape, *pri_rate;
  double *block_sums;           // [gridDim.x]
};

// compute_elbo_term_helper of a GPMatrix / GPMatrixGR (gpbase.hh:360-387, 717-741): one thread per element
__global__ void __launch_bounds__(kThreads) gamma_matrix_kernel(const GammaArgs g)
{
  const uint64_t total = (uint64_t)g.R * g.K;
  double local = 0.0;
  for (uint64_t e = (uint64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (uint64_t)gridDim.x * kThreads) {
    const uint32_t r = (uint32_t)(e / g.K), k = (uint32_t)(e % g.K);
    const size_t o = (size_t)r * g.ld + k;
    const double ev = (double)g.Ev[o], el = (double)g.Elog[o];
    const double a = floor30_d((double)g.shape[o]);
    const double b = floor30_d((double)(g.rate_is_vector ? g.rate[k] : g.rate[o]));
    double rp = g.rprior, lrp = log(g.rprior);
    if (g.pri_shape != nullptr) {
      const double pa = floor30_d((double)g.pri_shape[r]), pb = floor30_d((double)g.pri_rate[r]);
      rp = pa / pb;
      lrp = digamma_d(pa) - log(pb);
    }
    double s = g.sprior * lrp + (g.sprior - 1.0) * el;
    s -= rp * ev + g.lg_sprior;
    s -= a * log(b) + (a - 1.0) * el;
    s += b * ev + lgamma(a);
    local += s;
  }
  block_store(local, g.block_sums);
}

// GPArray::compute_elbo_term_helper (gpbase.hh:951-969) for xi / eta: expectations from (shape, rate) as
// compute_expectations left them (gpbase.hh:912-925)
__global__ void __launch_bounds__(kThreads) gamma_array_kernel(const float *shape, const float *rate, uint32_t R, double sprior,
                                                               double rprior, double lg_sprior, double *block_sums)
{
  double local = 0.0;
  for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < R; r += gridDim.x * kThreads) {
    const double a = floor30_d((double)shape[r]), b = floor30_d((double)rate[r]);
    const double ev = a / b, el = digamma_d(a) - log(b);
    double s = sprior * log(rprior) + (sprior - 1.0) * el;
    s -= rprior * ev + lg_sprior;
    s -= a * log(b) + (a - 1.0) * el;
    s += b * ev + lgamma(a);
    local += s;
  }
  block_store(local, block_sums);
}

} // namespace elbo
} // namespace hpf

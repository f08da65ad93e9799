// Probe of the tcgen05 MN-major shared-memory descriptor (not part of the product): one CTA computes
// D[128 x 128] = sum_k A[k][m] * B[k][n] with BOTH operands stored "MN-major" (k rows of 64 contiguous
// mn-elements, 128-byte swizzled), for candidate (LBO, SBO) encodings, and prints the max error.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61; // SWIZZLE_128B
  return d;
}

// element (k, mn) of an MN-major operand tile [K=128][MN=128]: 64-wide mn blocks of [128 k-rows x 128 B]
__device__ __forceinline__ uint32_t mn_offset(uint32_t k, uint32_t mn)
{
  const uint32_t blk = mn >> 6, c = (mn & 63u) >> 3, e = mn & 7u;
  return blk * 16384u + k * 128u + ((c ^ (k & 7u)) << 4) + e * 2u;
}

__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *B, float *D, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes)
{
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t *g = raw + (base - smem_u32(raw));
  uint8_t *sa = g, *sb = g + 32768;
  uint64_t *bar = reinterpret_cast<uint64_t *>(g + 65536);
  uint32_t *slot = reinterpret_cast<uint32_t *>(g + 65536 + 16);
  for (uint32_t e = threadIdx.x; e < 128 * 128; e += 128) {
    const uint32_t k = e / 128, mn = e % 128;
    *reinterpret_cast<__nv_bfloat16 *>(sa + mn_offset(k, mn)) = __float2bfloat16_rn(A[e]);
    *reinterpret_cast<__nv_bfloat16 *>(sb + mn_offset(k, mn)) = __float2bfloat16_rn(B[e]);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    // D fp32, A/B bf16, A and B MN-major (bits 15, 16), M = 128, N = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (uint32_t ks = 0; ks < 8; ++ks) {
      const uint64_t da = make_desc(base + ks * kstep_bytes, lbo, sbo), db = make_desc(base + 32768 + ks * kstep_bytes, lbo, sbo);
      const uint32_t acc = ks != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  {
    uint32_t done = 0;
    for (unsigned long long spin = 0; !done; ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
      if (spin > (1ull << 26)) __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t warp = threadIdx.x >> 5, row = threadIdx.x;
  for (uint32_t c = 0; c < 4; ++c) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((warp * 32u) << 16) + c * 32u) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[row * 128 + c * 32 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main()
{
  const int N = 128 * 128;
  float *hA = (float *)malloc(N * 4), *hB = (float *)malloc(N * 4), *hD = (float *)malloc(N * 4), *ref = (float *)malloc(N * 4);
  srand(1);
  for (int i = 0; i < N; ++i) { hA[i] = (float)(rand() % 17 - 8) / 8.f; hB[i] = (float)(rand() % 13 - 6) / 4.f; } // exact in bf16
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < 128; ++k) s += (double)hA[k * 128 + m] * hB[k * 128 + n];
      ref[m * 128 + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, N * 4); cudaMalloc(&dB, N * 4); cudaMalloc(&dD, N * 4);
  cudaMemcpy(dA, hA, N * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const uint32_t cand[][3] = { { 16384, 1024, 2048 }, { 1024, 16384, 2048 }, { 16384, 1024, 32 }, { 1024, 16384, 32 },
                               { 16384, 2048, 2048 }, { 2048, 16384, 2048 }, { 8192, 1024, 2048 }, { 1024, 8192, 2048 } };
  for (auto &c : cand) {
    cudaMemset(dD, 0, N * 4);
    probe<<<1, 128, 70000>>>(dA, dB, dD, c[0], c[1], c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo=%u sbo=%u kstep=%u: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, N * 4, cudaMemcpyDeviceToHost);
    double mx = 0; int bad = 0;
    for (int i = 0; i < N; ++i) { double d = fabs((double)hD[i] - ref[i]); if (d > mx) mx = d; if (d > 1e-3) ++bad; }
    printf("lbo=%5u sbo=%5u kstep=%4u: max err %.4g, wrong %d / %d  (D[0]=%g ref %g, D[1]=%g ref %g, D[128]=%g ref %g)\n", c[0], c[1], c[2], mx, bad, N,
           hD[0], ref[0], hD[1], ref[1], hD[128], ref[128]);
  }
  return 0;
}

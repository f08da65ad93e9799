// gather_bench.cu -- what this box's L2 -> SM path delivers for the access pattern of hpf::sweep_kernel:
// random gathers of 512-byte rows (128 fp32: one K=100 factor row at the engine's row stride) from a table that
// lives in L2.  The result is the denominator of bench.py's roofline at Netflix scale (bound "l2"),
// profiles/l2_gather_peak.json.  Measurement aid, not part of the library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o gather_bench tools/gather_bench.cu -lcuda
//   ./gather_bench MODE ROWS [NGATHERS_LOG2]      one JSON line per kernel configuration on stdout
// MODE
//   ldg8     8 lanes per row, 4 x LDG.128 (ld.global.nc) per lane -- sweep_kernel's own pattern; U rows in flight per group
//   ldg32    32 lanes per row, one LDG.128 per lane; U rows in flight per warp
//   bulk     one elected lane per warp issues cp.async.bulk (TMA, 512 B per row) into a shared-memory ring; the warp
//            reads the rows back from shared memory
//   gather4  the same ring filled by cp.async.bulk.tensor.2d ... tile::gather4 (four rows per instruction)
// Every kernel folds what it loaded into a checksum that must equal the host's, so a wrong copy cannot pass as a fast one.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kRowF4 = 32;            // float4 per row (512 bytes)
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float fold(float4 v) { return (v.x + v.y) + (v.z + v.w); }

// ---- ldg8: G = 8 lanes per row, V = 4 float4 per lane, U rows in flight ------------------------------------------
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) ldg8_kernel(const float4 *table, const uint32_t *idx, uint64_t n_per_group, double *out)
{
  const int lane = threadIdx.x & 31, gl = lane & 7;
  const uint64_t group = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) / 8;
  const uint32_t *ip = idx + group * n_per_group;
  float acc = 0.f;
  for (uint64_t j0 = 0; j0 < n_per_group; j0 += 8) {
    const uint32_t mine = ld_stream_u32(ip + j0 + gl);
#pragma unroll
    for (int t0 = 0; t0 < 8; t0 += U) {
      float4 b[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t r = __shfl_sync(0xffffffffu, mine, t0 + u, 8);
        const float4 *rp = table + (size_t)r * kRowF4;
#pragma unroll
        for (int v = 0; v < 4; ++v) b[u][v] = __ldg(rp + gl + v * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc += fold(b[u][v]);
    }
  }
  atomicAdd(out + (group & 1023), (double)acc);
}

// ---- ldg32: a warp per row -------------------------------------------------------------------------------------------
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) ldg32_kernel(const float4 *table, const uint32_t *idx, uint64_t n_per_warp, double *out)
{
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) / 32;
  const uint32_t *ip = idx + warp * n_per_warp;
  float acc = 0.f;
  for (uint64_t j0 = 0; j0 < n_per_warp; j0 += 32) {
    const uint32_t mine = ld_stream_u32(ip + j0 + lane);
#pragma unroll
    for (int t0 = 0; t0 < 32; t0 += U) {
      float4 b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) b[u] = __ldg(table + (size_t)__shfl_sync(0xffffffffu, mine, t0 + u) * kRowF4 + lane);
#pragma unroll
      for (int u = 0; u < U; ++u) acc += fold(b[u]);
    }
  }
  atomicAdd(out + (warp & 1023), (double)acc);
}

// ---- TMA paths: per-warp ring of D batches of 4 rows (2 KB), one mbarrier per batch ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap(); // a copy that never lands must not hang the box
  }
}
__device__ __forceinline__ void bulk_copy_row(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void gather4_rows(uint32_t dst, const CUtensorMap *map, int r0, int r1, int r2, int r3, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

template <int D, bool GATHER4, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) tma_kernel(const float4 *table, const __grid_constant__ CUtensorMap map, const uint32_t *idx,
                                                             uint64_t n_per_warp, double *out)
{
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kWarps = kThreads / 32;
  constexpr uint32_t kBatch = 4 * 512;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t warp = (uint64_t)blockIdx.x * kWarps + w;
  uint8_t *ring = smem + (size_t)w * D * kBatch;
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t bars = smem_u32(smem + (size_t)kWarps * D * kBatch) + (uint32_t)w * D * 8;
  if (lane == 0)
    for (int d = 0; d < D; ++d) mbar_init(bars + d * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint32_t *ip = idx + warp * n_per_warp;
  const uint64_t nb = n_per_warp / 4; // batches of four rows
  auto issue = [&](uint64_t b) {     // lane 0 only
    const uint32_t slot = (uint32_t)(b % D);
    const uint32_t bar = bars + slot * 8, dst = ring_u + slot * kBatch;
    const uint32_t r0 = ld_stream_u32(ip + b * 4), r1 = ld_stream_u32(ip + b * 4 + 1), r2 = ld_stream_u32(ip + b * 4 + 2),
                   r3 = ld_stream_u32(ip + b * 4 + 3);
    mbar_expect_tx(bar, kBatch);
    if (GATHER4) gather4_rows(dst, &map, (int)r0, (int)r1, (int)r2, (int)r3, bar);
    else {
      bulk_copy_row(dst, table + (size_t)r0 * kRowF4, 512, bar);
      bulk_copy_row(dst + 512, table + (size_t)r1 * kRowF4, 512, bar);
      bulk_copy_row(dst + 1024, table + (size_t)r2 * kRowF4, 512, bar);
      bulk_copy_row(dst + 1536, table + (size_t)r3 * kRowF4, 512, bar);
    }
  };
  if (lane == 0)
    for (uint64_t b = 0; b < (uint64_t)D && b < nb; ++b) issue(b);
  float acc = 0.f;
  for (uint64_t b = 0; b < nb; ++b) {
    const uint32_t slot = (uint32_t)(b % D);
    mbar_wait(bars + slot * 8, (uint32_t)((b / D) & 1));
    const float4 *rows = reinterpret_cast<const float4 *>(ring + slot * kBatch);
#pragma unroll
    for (int r = 0; r < 4; ++r) acc += fold(rows[r * kRowF4 + lane]);
    __syncwarp(); // every lane has read the slot before it is refilled
    if (lane == 0 && b + D < nb) issue(b + D);
  }
  atomicAdd(out + (warp & 1023), (double)acc);
}

// ---- host ------------------------------------------------------------------------------------------------------------
static double checksum_host(const std::vector<float> &rowsum, const std::vector<uint32_t> &idx, uint64_t n)
{
  double s = 0;
  for (uint64_t j = 0; j < n; ++j) s += rowsum[idx[j]];
  return s;
}

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: %s ldg8|ldg32|bulk|gather4 ROWS [log2 gathers]\n", argv[0]); return 2; }
  const char *mode = argv[1];
  const uint32_t R = (uint32_t)atoll(argv[2]);
  const int lg = argc > 3 ? atoi(argv[3]) : 26;
  const uint64_t N = 1ull << lg;
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // table: small integers, so that fp32 sums of a group's slice are exact whatever the order
  std::vector<float> h_table((size_t)R * 128), rowsum(R);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  for (uint32_t r = 0; r < R; ++r) {
    float s = 0;
    for (int k = 0; k < 128; ++k) { const float v = (float)(rnd() % 3); h_table[(size_t)r * 128 + k] = v; s += v; }
    rowsum[r] = s;
  }
  std::vector<uint32_t> h_idx(N);
  for (uint64_t j = 0; j < N; ++j) h_idx[j] = (uint32_t)(rnd() % R);
  float4 *d_table; uint32_t *d_idx; double *d_out;
  CK(cudaMalloc(&d_table, (size_t)R * 512)); CK(cudaMalloc(&d_idx, N * 4)); CK(cudaMalloc(&d_out, 1024 * 8));
  CK(cudaMemcpy(d_table, h_table.data(), (size_t)R * 512, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, h_idx.data(), N * 4, cudaMemcpyHostToDevice));
  CUtensorMap map;
  memset(&map, 0, sizeof map);
  if (!strcmp(mode, "gather4")) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    cuuint64_t dims[2] = { 128, R };
    cuuint64_t strides[1] = { 512 };
    cuuint32_t box[2] = { 128, 1 };   // tile::gather4: the box is ONE row; the instruction names four of them
    cuuint32_t estr[2] = { 1, 1 };
    CUresult rc = ((PFN_cuTensorMapEncodeTiled_v12000)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_table, dims, strides, box, estr,
                                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("{\"mode\": \"gather4\", \"error\": \"cuTensorMapEncodeTiled %d\"}\n", (int)rc); return 0; }
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto report = [&](const char *name, int param, int minb, int grid, float ms, int reps, uint64_t n_eff) {
    std::vector<double> h_out(1024);
    CK(cudaMemcpy(h_out.data(), d_out, 8192, cudaMemcpyDeviceToHost));
    double got = 0;
    for (double v : h_out) got += v;
    const double want = checksum_host(rowsum, h_idx, n_eff);
    const uint64_t N = n_eff;
    const double gbs = (double)N * 512 / (ms / reps * 1e-3) / 1e9;
    printf("{\"mode\": \"%s\", \"param\": %d, \"min_blocks\": %d, \"grid\": %d, \"rows\": %u, \"table_MB\": %.1f, \"gathers\": %llu, "
           "\"ms\": %.4f, \"GBps\": %.1f, \"checksum_ok\": %s}\n",
           name, param, minb, grid, R, R * 512.0 / 1e6, (unsigned long long)N, ms / reps, gbs,
           fabs(got - want * reps) <= 1e-9 * want * reps ? "true" : "false");
    fflush(stdout);
  };
  const int reps = 10;
#define RUN(NAME, PARAM, MINB, NEFF, LAUNCH)                             \
  do {                                                                    \
    CK(cudaMemset(d_out, 0, 8192));                                       \
    LAUNCH; LAUNCH;                                                       \
    CK(cudaDeviceSynchronize());                                          \
    CK(cudaMemset(d_out, 0, 8192));                                       \
    CK(cudaEventRecord(e0));                                              \
    for (int i_ = 0; i_ < reps; ++i_) { LAUNCH; }                         \
    CK(cudaEventRecord(e1));                                              \
    CK(cudaEventSynchronize(e1));                                         \
    CK(cudaGetLastError());                                               \
    float ms_ = 0; CK(cudaEventElapsedTime(&ms_, e0, e1));                \
    report(NAME, PARAM, MINB, grid, ms_, reps, NEFF);                     \
  } while (0)

  if (!strcmp(mode, "ldg8")) {
#define LDG8P(U, MINB)                                                                                        \
  {                                                                                                           \
    const int grid = sms * MINB; /* one resident wave */                                                      \
    const uint64_t groups = (uint64_t)grid * kThreads / 8, per = N / groups / 8 * 8;                          \
    RUN("ldg8", U, MINB, per * groups, (ldg8_kernel<U, MINB><<<grid, kThreads>>>(d_table, d_idx, per, d_out)));             \
  }
    LDG8P(1, 3) LDG8P(1, 4) LDG8P(1, 6) LDG8P(1, 8) LDG8P(2, 2) LDG8P(2, 3) LDG8P(2, 4) LDG8P(4, 2) LDG8P(4, 3) LDG8P(8, 1) LDG8P(8, 2)
  } else if (!strcmp(mode, "ldg32")) {
#define LDG32P(U, MINB)                                                                                       \
  {                                                                                                           \
    const int grid = sms * MINB;                                                                              \
    const uint64_t warps = (uint64_t)grid * kThreads / 32, per = N / warps / 32 * 32;                         \
    RUN("ldg32", U, MINB, per * warps, (ldg32_kernel<U, MINB><<<grid, kThreads>>>(d_table, d_idx, per, d_out)));           \
  }
    LDG32P(2, 4) LDG32P(4, 4) LDG32P(4, 8) LDG32P(8, 4) LDG32P(8, 8) LDG32P(16, 4) LDG32P(16, 6) LDG32P(32, 2) LDG32P(32, 4)
  } else {
    const bool g4 = !strcmp(mode, "gather4");
#define TMAP(D, MINB)                                                                                         \
  {                                                                                                           \
    const int grid = sms * MINB;                                                                              \
    const uint64_t warps = (uint64_t)grid * kThreads / 32, per = N / warps / 4 * 4;                           \
    const size_t sm = (size_t)(kThreads / 32) * D * 2048 + (kThreads / 32) * D * 8 + 128;                     \
    if (g4) {                                                                                                 \
      CK(cudaFuncSetAttribute(tma_kernel<D, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      RUN("gather4", D, MINB, per * warps, (tma_kernel<D, true, MINB><<<grid, kThreads, sm>>>(d_table, map, d_idx, per, d_out))); \
    } else {                                                                                                  \
      CK(cudaFuncSetAttribute(tma_kernel<D, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      RUN("bulk", D, MINB, per * warps, (tma_kernel<D, false, MINB><<<grid, kThreads, sm>>>(d_table, map, d_idx, per, d_out))); \
    }                                                                                                         \
  }
    TMAP(2, 4) TMAP(4, 2) TMAP(4, 3) TMAP(6, 2) TMAP(8, 1) TMAP(3, 4)
  }
  return 0;
}

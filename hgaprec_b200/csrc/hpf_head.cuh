// hpf_head.cuh -- K1c: the DENSE HEAD of the phi sweep on tcgen05.
//
// Item popularity is heavy-tailed: the kHead most popular items carry about half
// of all nonzeros (Zipf(1), 17.8K items: 52 %), and the (users x head items)
// block of the ratings matrix is ~25 % dense.  Gathering a 400-byte factor row
// per nonzero is the wrong tool there; the block is GEMM-shaped:
//
//   Z   = A_tile . B_head^T                 [128 users x 128 items]   (MMA 1)
//   P   = Y (/) Z   elementwise, 0 where there is no rating            (epilogue 1)
//   O   = P . B_head      -> T_theta[user tile] +=                     (MMA 2)
//   dB += P^T . A_tile    -> T_beta[head items] += (once per CTA)      (MMA 3)
//
// i.e. the per-nonzero work of sweep_kernel (hgaprec.cc:1340-1366) for every
// (user, head item) pair at once, with A = exp(Elog theta - shift), B likewise
// (DESIGN.md section 3).  The same shape as an attention forward + key-gradient
// pass, and built the same way: operands in 128B-swizzled shared memory (A and
// B_head by TMA, P written by the epilogue threads), accumulators in TMEM
// (Z | O | dB = 3 x 128 columns), one thread issuing tcgen05.mma.  fp32 fidelity
// comes from the split x = hi + lo (two bf16) with three MMAs per product
// (hi.hi + hi.lo + lo.hi; representation and dropped-term error 2^-18 each).
// MMA 2 and MMA 3 read operands "MN-major" (contraction index along the rows of the
// stored tile): descriptor LBO = 16 KB (next 64-column block), SBO = 1 KB (next 8
// rows), 2 KB per 16-row K step -- tools/umma_probe/probe.cu pins that encoding.
//
// The tail (all other items) stays in sweep_kernel; head items have no rows in
// the item pass and head nonzeros are absent from the user pass, so this kernel
// removes about half of all gathers of an iteration.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hpf_topn.cuh" // PTX wrappers (mbarrier, TMA, tcgen05)

namespace hpf {
namespace head {

using namespace topk; // smem_u32, mbar_*, tma_load_2d, tc_*

constexpr int kHead = 128;     // head items  (N of MMA 1, K of MMA 2, M of MMA 3)
constexpr int kUsers = 128;    // users per tile (M of MMA 1/2, K of MMA 3)
constexpr int kFact = 128;     // padded factor dimension (K of MMA 1, N of MMA 2/3)
constexpr int kEpiWarps = 8;    // two warps per TMEM lane quarter, each taking half of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps; // warp 0: TMA, warp 1: MMA + TMEM, warps 2-9: epilogue
constexpr uint32_t kBlk = 16384;            // one [128 rows x 64 bf16] swizzled block
constexpr uint32_t kOperand = 2 * kBlk;     // one 128 x 128 bf16 operand (hi or lo)
constexpr uint32_t kSmemA = 0, kSmemB = 2 * kOperand, kSmemP = 4 * kOperand, kSmemBar = 6 * kOperand;
constexpr uint32_t kSmemBytes = 1024 + 6 * kOperand + 256;

// K-major descriptor (rows = M/N index, 64-element K blocks): as topk::make_desc_sw128
// MN-major descriptor (rows = K index, 64-element M/N blocks 16 KB apart)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)(kBlk >> 4) << 16;   // LBO: next block of 64 M/N elements
  d |= (uint64_t)(1024u >> 4) << 32;  // SBO: next group of 8 K rows
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;            // SWIZZLE_128B
  return d;
}
// instruction descriptors: D fp32, A/B bf16, M = 128, N = 128; bit 15 / 16 = A / B is MN-major
constexpr uint32_t kIdescBase = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescKK = kIdescBase;                          // MMA 1: A K-major, B K-major
constexpr uint32_t kIdescKM = kIdescBase | (1u << 16);             // MMA 2: A K-major (P), B MN-major (B_head)
constexpr uint32_t kIdescMM = kIdescBase | (1u << 15) | (1u << 16);// MMA 3: A MN-major (P^T), B MN-major (A_tile)

struct HeadArgs {
  uint32_t n;            // users (rows >= n of the padded operand arrays are zero)
  uint32_t ntiles;       // user tiles of 128
  uint32_t K, ld;        // factors rounded up to 4, row stride of T in floats
  uint32_t Ktrue;        // factors
  const uint8_t *Y;      // [ntiles * 128 x 128] ratings of the head block (0: none)
  const uint32_t *head_ids; // [128] item row of each head slot (0xffffffff: unused slot)
  float *T_theta;        // [n x ld]  += O        (rows owned by this tile: plain read-modify-write)
  float *dB_part;        // [gridDim.x x 128 x 128]  this CTA's dB, summed in a fixed order by head_reduce_kernel
  // -bias: operand columns K, K+1 carry {aux.x, aux.y} (users) / {aux.y, aux.x} (items), so that Z gains the two
  // bias slots of phi, O[:, K] is the user-bias sum and dB[:, K+1] the item-bias sum (hgaprec.cc:222-239, 1361-1364)
  float *Tb_theta;       // [n] += O[:, K]   (nullptr without -bias)
  const float *ElogbT, *ElogbB;
  float *TbdirectT, *TbdirectB;
  // exact fallback for a pair whose Z left the fp32 range
  ElogSrc ElogT, ElogB;
  float *TdirectT, *TdirectB;
  uint32_t *flagT, *flagB;
  unsigned long long *slow_count;
};

// exact log-domain phi for one (user, item) pair, added to BOTH sides' fallback buffers (one thread)
__device__ __noinline__ void slow_pair(const HeadArgs &a, uint32_t u, uint32_t it, float yv)
{
  const bool bias = a.Tb_theta != nullptr;
  const float xbu = bias ? a.ElogbT[u] : -CUDART_INF_F, xbi = bias ? a.ElogbB[it] : -CUDART_INF_F;
  // x_k = Elog theta_uk + Elog beta_ik, four at a time (a.K is a multiple of 4; pad columns give -inf); E[log v] is
  // recomputed from shape and rate when the arrays are not current (load_elog4)
  auto x4 = [&](uint32_t q) {
    const float4 t4 = load_elog4(a.ElogT, u, q, a.ld / 4, a.Ktrue), b4 = load_elog4(a.ElogB, it, q, a.ld / 4, a.Ktrue);
    return make_float4(t4.x + b4.x, t4.y + b4.y, t4.z + b4.z, t4.w + b4.w);
  };
  float mx = fmaxf(xbu, xbi);
  for (uint32_t q = 0; q < a.K / 4; ++q) {
    const float4 x = x4(q);
    mx = fmaxf(mx, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
  }
  float sum = bias ? expf(xbu - mx) + expf(xbi - mx) : 0.f;
  for (uint32_t q = 0; q < a.K / 4; ++q) {
    const float4 x = x4(q);
    sum += expf(x.x - mx) + expf(x.y - mx) + expf(x.z - mx) + expf(x.w - mx);
  }
  const float sc = yv / sum;
  for (uint32_t q = 0; q < a.K / 4; ++q) {
    const float4 x = x4(q);
    const float v[4] = { sc * expf(x.x - mx), sc * expf(x.y - mx), sc * expf(x.z - mx), sc * expf(x.w - mx) };
    for (uint32_t j = 0; j < 4; ++j) {
      atomicAdd(a.TdirectT + (size_t)u * a.ld + q * 4 + j, v[j]);
      atomicAdd(a.TdirectB + (size_t)it * a.ld + q * 4 + j, v[j]);
    }
  }
  if (bias) {
    atomicAdd(a.TbdirectT + u, sc * expf(xbu - mx));
    atomicAdd(a.TbdirectB + it, sc * expf(xbi - mx));
  }
  atomicAdd(a.slow_count, 2ull); // the gather path counts a nonzero once per pass
  *a.flagT = 1u;
  *a.flagB = 1u;
}

// VARIANT (HPF_HEAD_VARIANT, experiments; 0 is the measured default and compiles to the code it always was):
//   bit 0  epilogue 2 adds O to T_theta with red.global.add.v4.f32 instead of load + add + store (a row has one
//          adder, so the result is the same and stays deterministic; no load round trip, half the traffic)
//   bit 1  epilogue 1 fetches its 64 bytes of Y before waiting for Z (Y does not depend on the MMA)
//   bit 2  the two epilogues get their own warps (2-9: epilogue 1, 10-17: epilogue 2; 576 threads): epilogue 2 of tile i
//          runs under epilogue 1 and MMA 2/3 of tile i + 1.  One more barrier: "P free", committed with o_full, because
//          epilogue 1 no longer follows epilogue 2 in program order.  The MMA thread's waits already order the rest
//          (Z is rewritten only after p_full, O only after o_empty).
constexpr int head_threads(int variant) { return (variant & 4) != 0 ? kThreads + 32 * kEpiWarps : kThreads; }

template <int VARIANT>
__global__ void __launch_bounds__(head_threads(VARIANT), 1)
head_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const HeadArgs a)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_a_full = base + kSmemBar, bar_a_empty = bar_a_full + 8, bar_b_full = bar_a_full + 16,
                 bar_z_full = bar_a_full + 24, bar_p_full = bar_a_full + 32, bar_o_full = bar_a_full + 40,
                 bar_o_empty = bar_a_full + 48, bar_p_free = bar_a_full + 56;
  constexpr bool kSplit = (VARIANT & 4) != 0;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + kSmemBar + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(bar_a_full, 1); mbar_init(bar_a_empty, 1); mbar_init(bar_b_full, 1);
    mbar_init(bar_z_full, 1); mbar_init(bar_p_full, 32 * kEpiWarps); mbar_init(bar_o_full, 1); mbar_init(bar_o_empty, 32 * kEpiWarps);
    if constexpr (kSplit) mbar_init(bar_p_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_z = tmem, tm_o = tmem + 128u, tm_db = tmem + 256u;
  const uint32_t sA_hi = base + kSmemA, sA_lo = sA_hi + kOperand, sB_hi = base + kSmemB, sB_lo = sB_hi + kOperand,
                 sP_hi = base + kSmemP, sP_lo = sP_hi + kOperand;
  const uint32_t my_tiles = a.ntiles > blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

  if (warp == 0) {
    // ===== TMA producer: B_head once, then one A tile per user tile =====
    if (lane == 0) {
      mbar_expect_tx(bar_b_full, 2 * kOperand);
      for (int blk = 0; blk < 2; ++blk) {
        tma_load_2d(sB_hi + blk * kBlk, &map_b_hi, bar_b_full, blk * 64, 0);
        tma_load_2d(sB_lo + blk * kBlk, &map_b_lo, bar_b_full, blk * 64, 0);
      }
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t tile = blockIdx.x + i * gridDim.x;
        mbar_wait(bar_a_empty, (i & 1u) ^ 1u); // MMA 3 of the previous tile has read A
        mbar_expect_tx(bar_a_full, 2 * kOperand);
        for (int blk = 0; blk < 2; ++blk) {
          tma_load_2d(sA_hi + blk * kBlk, &map_a_hi, bar_a_full, blk * 64, (int)(tile * kUsers));
          tma_load_2d(sA_lo + blk * kBlk, &map_a_lo, bar_a_full, blk * 64, (int)(tile * kUsers));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      mbar_wait(bar_b_full, 0u);
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t ph = i & 1u;
        mbar_wait(bar_a_full, ph);
        tc_fence_after();
        // MMA 1: Z = A . B^T  (K = factors: two 64-wide blocks x 4 steps of 16)
#pragma unroll
        for (uint32_t kb = 0; kb < 2; ++kb)
#pragma unroll
          for (uint32_t kk = 0; kk < 4; ++kk) {
            const uint32_t off = kb * kBlk + kk * 32u;
            const uint64_t ah = make_desc_sw128(sA_hi + off), al = make_desc_sw128(sA_lo + off);
            const uint64_t bh = make_desc_sw128(sB_hi + off), bl = make_desc_sw128(sB_lo + off);
            tc_mma_bf16(tm_z, ah, bh, kIdescKK, (kb | kk) != 0u);
            tc_mma_bf16(tm_z, ah, bl, kIdescKK, 1u);
            tc_mma_bf16(tm_z, al, bh, kIdescKK, 1u);
          }
        tc_commit(bar_z_full);
        mbar_wait(bar_p_full, ph);         // epilogue 1 has written P (and drained Z)
        mbar_wait(bar_o_empty, ph ^ 1u);   // epilogue 2 of the previous tile has drained O
        tc_fence_after();
        // MMA 2: O = P . B_head   (K = head items, 8 steps of 16; P K-major, B_head MN-major)
        // MMA 3: dB += P^T . A    (K = users,      8 steps of 16; both MN-major)
#pragma unroll
        for (uint32_t ks = 0; ks < 8; ++ks) {
          const uint32_t kmaj = (ks >> 2) * kBlk + (ks & 3u) * 32u; // K-major: block of 64, then 32 B per step
          const uint32_t mmaj = ks * 2048u;                          // MN-major: 16 rows of 128 B per step
          const uint64_t ph_k = make_desc_sw128(sP_hi + kmaj), pl_k = make_desc_sw128(sP_lo + kmaj);
          const uint64_t bh_m = make_desc_mn(sB_hi + mmaj), bl_m = make_desc_mn(sB_lo + mmaj);
          tc_mma_bf16(tm_o, ph_k, bh_m, kIdescKM, ks != 0u);
          tc_mma_bf16(tm_o, ph_k, bl_m, kIdescKM, 1u);
          tc_mma_bf16(tm_o, pl_k, bh_m, kIdescKM, 1u);
          const uint64_t ph_m = make_desc_mn(sP_hi + mmaj), pl_m = make_desc_mn(sP_lo + mmaj);
          const uint64_t ah_m = make_desc_mn(sA_hi + mmaj), al_m = make_desc_mn(sA_lo + mmaj);
          tc_mma_bf16(tm_db, ph_m, ah_m, kIdescMM, (i | ks) != 0u);
          tc_mma_bf16(tm_db, ph_m, al_m, kIdescMM, 1u);
          tc_mma_bf16(tm_db, pl_m, ah_m, kIdescMM, 1u);
        }
        tc_commit(bar_o_full);
        tc_commit(bar_a_empty);
        if constexpr (kSplit) tc_commit(bar_p_free);
      }
    }
  } else {
    // ===== epilogue: two threads per user row (one per half of the columns) =====
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const bool do_ep1 = !kSplit || warp < 2 + kEpiWarps, do_ep2 = !kSplit || warp >= 2 + kEpiWarps;
    const uint32_t half = (uint32_t)(kSplit && warp >= 2 + kEpiWarps ? warp - 2 - kEpiWarps : warp - 2) >> 2; // 0: columns 0-63, 1: 64-127
    const uint32_t r = (uint32_t)(q * 32 + lane);          // row inside the tile == TMEM lane
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    for (uint32_t i = 0; i < my_tiles; ++i) {
      const uint32_t tile = blockIdx.x + i * gridDim.x, ph = i & 1u;
      const uint32_t u = tile * kUsers + r;
      const uint8_t *yrow = a.Y + (size_t)u * kHead;
      // ---- epilogue 1: P = y / Z, split into bf16 hi / lo, written K-major (items along the row) ----
      if (do_ep1) {
      uint4 yp0 = make_uint4(0u, 0u, 0u, 0u), yp1 = yp0, yp2 = yp0, yp3 = yp0;
      if constexpr ((VARIANT & 2) != 0) {
        const uint4 *yq = reinterpret_cast<const uint4 *>(yrow + half * 64u);
        yp0 = __ldg(yq); yp1 = __ldg(yq + 1); yp2 = __ldg(yq + 2); yp3 = __ldg(yq + 3);
      }
      if constexpr (kSplit) mbar_wait(bar_p_free, ph ^ 1u); // MMA 2/3 of the previous tile has read P
      mbar_wait(bar_z_full, ph);
      tc_fence_after();
#pragma unroll 1
      for (uint32_t c = half * 2u; c < half * 2u + 2u; ++c) {
        uint32_t z[32];
        tc_ld32(tm_z + lane_addr + c * 32u, z);
        uint4 y0, y1;
        if constexpr ((VARIANT & 2) != 0) {
          const bool first = c == half * 2u;
          y0 = first ? yp0 : yp2;
          y1 = first ? yp1 : yp3;
        } else {
          y0 = __ldg(reinterpret_cast<const uint4 *>(yrow + c * 32u));
          y1 = __ldg(reinterpret_cast<const uint4 *>(yrow + c * 32u + 16u));
        }
        const uint32_t yw[8] = { y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w };
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) { // 8 items -> one 16-byte chunk of the swizzled row, for hi and for lo
          uint32_t hi_pk[4], lo_pk[4];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float w[2];
#pragma unroll
            for (int d = 0; d < 2; ++d) {
              const int j = g * 8 + e + d;
              const float yv = (float)((yw[j >> 2] >> ((j & 3) * 8)) & 0xffu);
              const float zz = __uint_as_float(z[j]);
              const bool ok = zz > kZMin && zz < kZMax;
              w[d] = (yv != 0.f && ok) ? yv * frcp(zz) : 0.f;
              if (yv != 0.f && !ok) {
                const uint32_t it = __ldg(a.head_ids + c * 32u + j);
                if (u < a.n && it != 0xffffffffu) slow_pair(a, u, it, yv);
              }
            }
            const __nv_bfloat16 h0 = __float2bfloat16_rn(w[0]), h1 = __float2bfloat16_rn(w[1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(w[0] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w[1] - __bfloat162float(h1));
            hi_pk[e >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo_pk[e >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const uint32_t item0 = c * 32u + g * 8u;                 // first item of this chunk
          const uint32_t off = (item0 >> 6) * kBlk + r * 128u + ((((item0 & 63u) >> 3) ^ (r & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_hi + off), "r"(hi_pk[0]), "r"(hi_pk[1]), "r"(hi_pk[2]), "r"(hi_pk[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_lo + off), "r"(lo_pk[0]), "r"(lo_pk[1]), "r"(lo_pk[2]), "r"(lo_pk[3]) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes of P -> visible to the MMA
      tc_fence_before();
      mbar_arrive(bar_p_full);
      }
      // ---- epilogue 2: T_theta[u] += O[u] ----
      if (do_ep2) {
      mbar_wait(bar_o_full, ph);
      tc_fence_after();
      float *trow = a.T_theta + (size_t)u * a.ld;
#pragma unroll 1
      for (uint32_t c = half * 2u; c < half * 2u + 2u; ++c) {
        uint32_t o[32];
        tc_ld32(tm_o + lane_addr + c * 32u, o);
        tc_wait_ld();
        if (u < a.n && c * 32u <= a.K) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint32_t k = c * 32u + j;
            if (k == a.K && a.Tb_theta != nullptr) a.Tb_theta[u] += __uint_as_float(o[j]); // the user-bias slot of phi
            if constexpr ((VARIANT & 1) != 0) {
              if (k < a.K)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(trow + k), "f"(__uint_as_float(o[j])),
                             "f"(__uint_as_float(o[j + 1])), "f"(__uint_as_float(o[j + 2])), "f"(__uint_as_float(o[j + 3]))
                             : "memory");
            } else if (k < a.K) { // K is a multiple of 4 in storage (Kp): whole float4s, pad lanes are zero on both sides
              float4 t = *reinterpret_cast<float4 *>(trow + k);
              t.x += __uint_as_float(o[j]); t.y += __uint_as_float(o[j + 1]);
              t.z += __uint_as_float(o[j + 2]); t.w += __uint_as_float(o[j + 3]);
              *reinterpret_cast<float4 *>(trow + k) = t;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_o_empty);
      }
    }
    // ---- after the last tile: this CTA's dB (head items x factors), one thread per head item ----
    if (do_ep2) {
      float *brow = a.dB_part + ((size_t)blockIdx.x * kHead + r) * kFact;
#pragma unroll 1
      for (uint32_t c = half * 2u; c < half * 2u + 2u; ++c) {
        uint32_t d[32];
        if (my_tiles > 0) { // the last bar_o_full wait above already covers the final MMA 3
          tc_ld32(tm_db + lane_addr + c * 32u, d);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) d[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4 *>(brow + c * 32u + j) =
              make_float4(__uint_as_float(d[j]), __uint_as_float(d[j + 1]), __uint_as_float(d[j + 2]), __uint_as_float(d[j + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- set-up / per-iteration helpers ---------------------------------------------
// dense ratings of the head block: Y[u * 128 + slot] += y for every nonzero of a head item
// (a repeated (user, item) line adds up: same Z, so the contributions add as in the reference's walk)
// slot s = block * 128 + position; every block has its own [n_pad x 128] byte matrix (block_stride bytes apart)
// A cell is one byte: repeated lines whose ratings add up past 255 do not fit.  The add that crosses 255 sees it in
// the value atomicAdd returns and raises *overflow; hpf_set_ratings_csr then plans without the dense head.
__global__ void dense_y_kernel(const uint32_t *row_of, const uint32_t *col, const uint8_t *y, const uint32_t *slot_of, uint64_t nnz,
                               size_t block_stride, uint32_t *Yw, uint32_t *overflow)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const uint32_t s = slot_of[col[j]];
  if (s == 0xffffffffu) return;
  const uint32_t yv = y ? y[j] : 1u;
  const size_t e = (size_t)(s / kHead) * block_stride + (size_t)row_of[j] * kHead + (s % kHead);
  const uint32_t sh = (uint32_t)(e & 3u) * 8u;
  const uint32_t old = atomicAdd(Yw + (e >> 2), yv << sh);
  if (((old >> sh) & 0xffu) + yv > 255u) *overflow = 1u;
}

// split-bf16 operand rows with the two -bias columns: [A_r (K values) | 0.. | at Kp: aux.x, aux.y (users) or aux.y, aux.x (items)]
__global__ void __launch_bounds__(256) split_aux_kernel(const float *A, uint32_t ld, uint32_t Kp, const float2 *aux, int swap,
                                                        const uint32_t *gather, uint32_t rows, uint32_t rows_pad,
                                                        __nv_bfloat16 *hi, __nv_bfloat16 *lo)
{
  const uint64_t total = (uint64_t)rows_pad * kFact;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(e / kFact), k = (uint32_t)(e % kFact);
    float v = 0.f;
    if (r < rows) {
      const uint32_t src = gather ? gather[r] : r;
      if (k < Kp) v = A[(size_t)src * ld + k];
      else if (k == Kp) v = swap ? aux[src].y : aux[src].x;
      else if (k == Kp + 1) v = swap ? aux[src].x : aux[src].y;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[e] = h;
    lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// T_beta[head item] = sum over CTAs of their dB partials, in a fixed order (deterministic); pad columns 0.  One launch
// for all head blocks: CTA x serves slot x % 128 of block x / 128 (that block's partials start block_stride floats
// further on); thread (k, g) adds the partials p = g, g + 4, ... of column k, the four sub-sums are added in order.
constexpr int kReduceSplit = 4;
__global__ void __launch_bounds__(kFact * kReduceSplit)
head_reduce_kernel(const float *dB_part, size_t block_stride, uint32_t nparts, const uint32_t *head_ids, uint32_t Kp, uint32_t ld,
                   float *T, float *Tb)
{
  __shared__ float sub[kReduceSplit][kFact];
  const uint32_t slot = blockIdx.x % kHead;
  const uint32_t it = head_ids[blockIdx.x];
  if (it == 0xffffffffu) return;
  const float *part = dB_part + (size_t)(blockIdx.x / kHead) * block_stride;
  const uint32_t k = threadIdx.x % kFact, g = threadIdx.x / kFact;
  float s = 0.f;
  // columns < Kp hold the factor sums; with -bias column Kp + 1 is the item-bias sum
  if (k < Kp || (Tb != nullptr && k == Kp + 1))
    for (uint32_t p = g; p < nparts; p += kReduceSplit) s += part[((size_t)p * kHead + slot) * kFact + k];
  sub[g][k] = s;
  __syncthreads();
  if (g == 0) {
    float t = sub[0][k];
#pragma unroll
    for (int q = 1; q < kReduceSplit; ++q) t += sub[q][k];
    if (k < ld) T[(size_t)it * ld + k] = k < Kp ? t : 0.f;
    if (Tb != nullptr && k == Kp + 1) Tb[it] = t;
  }
}

} // namespace head
} // namespace hpf

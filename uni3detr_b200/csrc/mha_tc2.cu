// K5 v2 — multi-head self-attention core softmax(Q K^T / sqrt(32)) V on tcgen05, second design.
//
// Reference: the attention core of nn.MultiheadAttention used through mmcv MultiheadAttention in every decoder
// layer (config uni3detr_sunrgbd.py:79-83): per (sequence, head), seq_len in {300, 900} keys,
// head_dim 32.
//
// What the first kernel (mha_tc.cu) left on the table (profiles/: 230 us per call at 1024 (sequence, head) units,
// torch SDPA: cuDNN 78 us, flash 109 us): one CTA of 128 threads per unit with every step serialised, V transposed
// through 2-byte shared stores, P bounced through shared memory, keys split 256 + 44.
// Here: ONE persistent CTA per SM, 13 warps:
//   warps 0-11  three "warpgroups" of 128 threads; warpgroup g owns the 128-query blocks g, g+3, ... of the current
//               unit and runs them through rounds of <= 160 keys: S = Q K^T (tcgen05.mma, M = 128, N <= 160, fp32 in
//               ITS 160 TMEM columns), softmax straight from the TMEM lanes (thread = query row), P written BACK
//               INTO TMEM as packed bf16 over the S columns it just consumed (tcgen05.st) and fed to the tensor core
//               as the A operand of O = P V (tcgen05.mma with A in TMEM - no shared-memory round trip), V taken as it
//               lies in memory ([key][d], an MN-major B operand: no transpose), online-softmax merge of the 32 O
//               columns in registers across key rounds. The three warpgroups overlap each other's MMA / softmax /
//               barrier latencies; each issues its own MMAs (one elected thread).
//   warp 12     producer: Q / K / V tiles of the next unit by TMA tensor loads (64-byte rows, SWIZZLE_64B) into a
//               two-stage ring, so the loads of unit u+1 hide under the math of unit u.
// TMEM: 3 x 160 = 480 of 512 columns: S / P at [0,160) of a warpgroup's region, O at [128,160) (written only after
// the softmax has consumed the S columns it overlays).
#include <stdlib.h>
#include "tma.cuh"

namespace u3d {
namespace mha2 {

using namespace tc;
using namespace tma;

constexpr int kHd = 32;
constexpr int kQB = 128;            // queries per block = UMMA M
constexpr int kKR = 160;            // keys per round = UMMA N max used (multiple of 16)
constexpr int kWG = 3;
constexpr int kThreads = kWG * 128 + 32;
constexpr int kMaxKeys = 1024;
constexpr int kRegion = 160;        // TMEM columns per warpgroup
constexpr int kOCol = 128;          // O accumulator columns inside the region

using SwQK = Swz<32>;               // 64-byte rows (32 dims), SWIZZLE_64B

struct Smem {
  uint64_t full[2];
  uint64_t empty[2];
  uint64_t mma_bar[kWG];
  uint32_t tmem_base;
};

__device__ __forceinline__ void wg_bar(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}
// D[tmem] (+)= A[tmem] x B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 2^x on the SFU without exp2f()'s range-scaling wrapper (arguments here are <= 0; flush-to-zero is what we want)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1)
k_mha_tc2(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
          const __grid_constant__ CUtensorMap tmap_v, int seq_len, int heads, int n_units, int n_qb, int n_kr,
          int stages, uint32_t stage_bytes, uint32_t kv_bytes, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const uint32_t data_s = smem_u32(smem_raw) + 1024u;   // stage s: K | V | Q blocks, each 1024-aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], kWG * 128);
    }
    for (int g = 0; g < kWG; ++g) mbar_init(&S.mma_bar[g], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t q_off = 2u * kv_bytes;                 // Q blocks follow K and V inside a stage

  if (warp == kWG * 4) {
    // ======================= producer: TMA loads of Q / K / V, one unit ahead =======================
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
        const int st = it % stages;
        mbar_wait(&S.empty[st], (((uint32_t)(it / stages)) & 1u) ^ 1u);
        const int seq = u / heads, head = u - seq * heads;
        const int row0 = seq * seq_len, col = head * kHd;
        const uint32_t base = data_s + (uint32_t)st * stage_bytes;
        mbar_expect_tx(&S.full[st], (uint32_t)(2 * n_kr * kKR + n_qb * kQB) * 64u);
        for (int r = 0; r < n_kr; ++r) {
          tma_load_2d(base + (uint32_t)r * (kKR * 64), &tmap_k, col, row0 + r * kKR, &S.full[st]);
          tma_load_2d(base + kv_bytes + (uint32_t)r * (kKR * 64), &tmap_v, col, row0 + r * kKR, &S.full[st]);
        }
        for (int b = 0; b < n_qb; ++b)
          tma_load_2d(base + q_off + (uint32_t)b * (kQB * 64), &tmap_q, col, row0 + b * kQB, &S.full[st]);
      }
    }
  } else {
    // ======================= three warpgroups: MMA issue + softmax + output =======================
    const int wg = warp >> 2;
    const int row = tid & 127;                           // query row inside the block = TMEM lane
    const bool issuer = row == 0;
    const uint32_t region = tmem + (uint32_t)(wg * kRegion);
    const uint32_t lane_base = region + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;   // 1/sqrt(32) * log2(e)
    const uint32_t idesc_m = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kQB >> 4) << 24);
    uint32_t mph = 0u;                                   // parity of this warpgroup's MMA barrier
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int st = it % stages;
      const int seq = u / heads, head = u - seq * heads;
      const uint32_t base = data_s + (uint32_t)st * stage_bytes;
      mbar_wait(&S.full[st], (uint32_t)(it / stages) & 1u);
      for (int qb = wg; qb < n_qb; qb += kWG) {
        float o[kHd];
#pragma unroll
        for (int d = 0; d < kHd; ++d) o[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int r = 0; r < n_kr; ++r) {
          const int k0 = r * kKR;
          const int nk = seq_len - k0 < kKR ? seq_len - k0 : kKR;      // valid keys of the round
          const int nk16 = (nk + 15) & ~15;
          // ---- S = Q K^T  (M = 128, N = nk16, K = 2 x 16)
          if (issuer) {
            tc_fence_after();
            const uint32_t idesc = idesc_m | ((uint32_t)(nk16 >> 3) << 17);
            const uint64_t a_desc = SwQK::desc(base + q_off + (uint32_t)qb * (kQB * 64));
            const uint64_t b_desc = SwQK::desc(base + (uint32_t)k0 * 64);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              umma_bf16(region, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, kk);
            umma_commit(&S.mma_bar[wg]);
          }
          mbar_wait(&S.mma_bar[wg], mph);
          mph ^= 1u;
          tc_fence_after();
          // ---- softmax of this thread's row: sweep 1 = max (only the last 16-column chunk of a round can
          // hold padding keys: full chunks run without per-element masks)
          const int n_full = nk >> 4;                                   // chunks whose 16 keys are all valid
          float bmax = -INFINITY;
#pragma unroll 2
          for (int ch = 0; ch < n_full; ++ch) {
            uint32_t sv[16];
            tmem_ld16(lane_base + (uint32_t)(ch * 16), sv);
            tmem_ld_wait();
            float m4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              m4[j] = fmaxf(fmaxf(__uint_as_float(sv[j]), __uint_as_float(sv[j + 4])),
                            fmaxf(__uint_as_float(sv[j + 8]), __uint_as_float(sv[j + 12])));
            bmax = fmaxf(bmax, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          }
          if (n_full * 16 < nk16) {
            uint32_t sv[16];
            tmem_ld16(lane_base + (uint32_t)(n_full * 16), sv);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n_full * 16 + j < nk) bmax = fmaxf(bmax, __uint_as_float(sv[j]));
          }
          const float m_new = fmaxf(m_run, bmax);
          const float corr = ex2((m_run - m_new) * sl2);               // first round: 2^(-inf) = 0
          const float mb = m_new * sl2;
          // sweep 2 = exp, row sum, P packed bf16 back into the S columns already consumed
          float ps0 = 0.f, ps1 = 0.f;
#pragma unroll 2
          for (int ch = 0; ch < n_full; ++ch) {
            uint32_t sv[16];
            tmem_ld16(lane_base + (uint32_t)(ch * 16), sv);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = ex2(fmaf(__uint_as_float(sv[j]), sl2, -mb));
              const float p1 = ex2(fmaf(__uint_as_float(sv[j + 1]), sl2, -mb));
              ps0 += p0;
              ps1 += p1;
              __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            tmem_st8(lane_base + (uint32_t)(ch * 8), pk);
          }
          if (n_full * 16 < nk16) {
            const int c0 = n_full * 16;
            uint32_t sv[16];
            tmem_ld16(lane_base + (uint32_t)c0, sv);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = c0 + j < nk ? ex2(fmaf(__uint_as_float(sv[j]), sl2, -mb)) : 0.f;
              const float p1 = c0 + j + 1 < nk ? ex2(fmaf(__uint_as_float(sv[j + 1]), sl2, -mb)) : 0.f;
              ps0 += p0;
              ps1 += p1;
              __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            tmem_st8(lane_base + (uint32_t)(c0 >> 1), pk);
          }
          const float psum = ps0 + ps1;
          l_run = l_run * corr + psum;
          m_run = m_new;
          tmem_st_wait();
          tc_fence_before();
          wg_bar(wg);                                                   // every row's P is in TMEM
          // ---- O_round = P V  (A = P from TMEM, B = V[key][d] MN-major, N = 32, K = nk16)
          if (issuer) {
            tc_fence_after();
            const uint32_t idesc = idesc_m | ((uint32_t)(kHd >> 3) << 17) | (1u << 16);   // bit 16: B MN-major
            for (int ks = 0; ks < nk16 / 16; ++ks) {
              const uint64_t b_desc = SwQK::desc(base + kv_bytes + (uint32_t)(k0 + ks * 16) * 64);
              umma_bf16_ts(region + kOCol, region + (uint32_t)(ks * 8), b_desc, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(&S.mma_bar[wg]);
          }
          mbar_wait(&S.mma_bar[wg], mph);
          mph ^= 1u;
          tc_fence_after();
          {
            uint32_t ov[32];
            tmem_ld32(lane_base + kOCol, ov);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < kHd; ++d) o[d] = o[d] * corr + __uint_as_float(ov[d]);
          }
          tc_fence_before();
          wg_bar(wg);                       // every row has read O before the next S overwrites the region
        }
        const int q = qb * kQB + row;
        if (q < seq_len) {
          const float inv = 1.f / l_run;
          uint32_t pk[16];
#pragma unroll
          for (int d = 0; d < kHd; d += 2) {
            __nv_bfloat162 h = __floats2bfloat162_rn(o[d] * inv, o[d + 1] * inv);
            pk[d >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* op = reinterpret_cast<uint4*>(out + ((size_t)seq * seq_len + q) * (size_t)(heads * kHd) + head * kHd);
#pragma unroll
          for (int c = 0; c < 4; ++c) op[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
      mbar_arrive(&S.empty[st]);            // this thread is done with the unit's Q / K / V
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace mha2

bool mha_tc2_supported(int seq_len, int ldq, int ldk, int ldv, const void* q, const void* k, const void* v,
                       const void* out) {
  if (seq_len < 1 || seq_len > mha2::kMaxKeys) return false;
  if ((ldq | ldk | ldv) % 8 != 0) return false;
  if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) != 0) return false;
  return true;
}

int mha_core_tc2(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int n_seq, int seq_len,
                 int heads, void* out, cudaStream_t st) {
  using namespace mha2;
  const int n_kr = cdiv(seq_len, kKR), n_qb = cdiv(seq_len, kQB);
  const uint32_t kv_bytes = (uint32_t)n_kr * kKR * 64;                  // multiple of 10240 -> 1024-aligned
  const uint32_t stage_bytes = 2 * kv_bytes + (uint32_t)n_qb * kQB * 64;
  const int stages = 2 * (size_t)stage_bytes + 2048 <= 200 * 1024 ? 2 : 1;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 1024;
  U3D_CHECK_ARG(smem <= 227 * 1024, "mha: sequence of %d keys does not fit shared memory", seq_len);
  const int rows = n_seq * seq_len, cols = heads * kHd;
  CUtensorMap tq, tk, tv;
  if (tma::encode_2d_bf16(&tq, q, cols, rows, ldq, kHd, kQB) != U3D_OK) return U3D_EINVAL;
  if (tma::encode_2d_bf16(&tk, k, cols, rows, ldk, kHd, kKR) != U3D_OK) return U3D_EINVAL;
  if (tma::encode_2d_bf16(&tv, v, cols, rows, ldv, kHd, kKR) != U3D_OK) return U3D_EINVAL;
  static int cur_smem = 0;
  U3D_CUDA(ensure_dynamic_smem(k_mha_tc2, smem, &cur_smem));
  const int n_units = n_seq * heads;
  const int grid = n_units < kNumSMs ? n_units : kNumSMs;
  k_mha_tc2<<<grid, kThreads, smem, st>>>(tq, tk, tv, seq_len, heads, n_units, n_qb, n_kr, stages, stage_bytes,
                                          kv_bytes, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // namespace u3d

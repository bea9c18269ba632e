// K5 (tensor-core flavour) — multi-head self-attention core softmax(Q K^T / sqrt(32)) V on tcgen05.
//
// Reference: the attention core of nn.MultiheadAttention used through mmcv MultiheadAttention in
// every decoder layer (config uni3detr_sunrgbd.py:79-83): per (sequence, head),
// seq_len in {300, 900} keys, head_dim 32.
//
// One CTA (128 threads) per (head, sequence); it stages all K rows and V^T ONCE and then loops over
// the 128-query blocks of the sequence (U3D_MHA_QSPLIT=1: one CTA per query block as in the first
// version, K / V^T staged once per block):
//   Q block, all K rows and V^T live in shared memory in the UMMA K-major swizzled layouts
//   (Q, K: 64-byte rows, SWIZZLE_64B; V^T and P: 64-key blocks of 128-byte rows, SWIZZLE_128B);
//   per key block of <= 256 keys:
//     S = Q K^T          2 x tcgen05.mma (M=128, N=keys, K=2x16) -> 256 fp32 TMEM columns
//     softmax            thread r owns row r: two sweeps over its TMEM lane with tcgen05.ld
//                        (row max, then exp2 / row sum), P written to shared memory as bf16
//     O_blk = P V        keys/16 x tcgen05.mma (M=128, N=32) -> 32 TMEM columns (reusing S's)
//     online-softmax merge of O_blk into 32 fp32 registers per row
//   out = O / l, one 64-byte bf16 row per thread.
// TMEM: 256 columns per CTA, so two CTAs share an SM.
#include <stdlib.h>
#include "tc_common.cuh"

namespace u3d {
namespace mha {

using namespace tc;

constexpr int kHd = 32;
constexpr int kQB = 128;       // queries per CTA = UMMA M
constexpr int kKB = 256;       // keys per S tile = UMMA N max
constexpr int kThreads = 128;
constexpr int kMaxKeys = 1024;

using SwQK = Swz<32>;          // 64-byte rows (32 dims)
using SwP = Swz<64>;           // 128-byte rows (64 keys)

struct Layout {
  uint32_t q, k, vt, p, total;
};
__host__ __device__ inline Layout make_layout(int keys_pad64) {
  Layout L;
  L.q = 0;                                         // 128 x 64 B
  L.k = L.q + kQB * 64;                            // keys_pad64 x 64 B (keys_pad64 % 64 == 0 -> 1024-aligned)
  L.vt = L.k + (uint32_t)keys_pad64 * 64;          // (keys_pad64/64) blocks of 32 x 128 B
  L.p = L.vt + (uint32_t)(keys_pad64 / 64) * 4096; // 4 blocks of 128 x 128 B
  L.total = L.p + 4 * 16384;
  return L;
}

// kVT = true (default): V is staged TRANSPOSED (V^T[d][key], K-major B operand of O = P V).
// kVT = false (U3D_MHA_VMN=1; validated in round 2, test_mha_core_v_mn_major): V is staged as loaded ([key][d], 64-byte
// rows, the layout of the K tile) and handed to the tensor core as an MN-major B operand (instruction
// descriptor bit 16), which removes the 2-byte transposing shared stores.
template <bool kVT>
__global__ void __launch_bounds__(kThreads)
k_mha_tc(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
         const __nv_bfloat16* __restrict__ v, int ldq, int ldk, int ldv, int seq_len, int heads,
         int qblocks_per_cta, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int head = blockIdx.y, seq = blockIdx.z;
  const int qb_first = blockIdx.x * qblocks_per_cta;
  const size_t row0 = (size_t)seq * seq_len;
  const int keys_pad64 = (seq_len + 63) & ~63;
  const Layout L = make_layout(keys_pad64);
  uint8_t* sm = smem_raw;
  const uint32_t sm_u = smem_u32(smem_raw);

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- zero the padded K / V^T regions, then fill Q, K (row-major 64-byte rows) and V transposed
  {
    uint4* z = reinterpret_cast<uint4*>(sm + L.k);
    const int n16 = (int)((L.p - L.k) / 16);
    for (int i = tid; i < n16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int e = tid; e < seq_len * 4; e += kThreads) {      // K rows and V^T columns
    const int r = e >> 2, c = e & 3;
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(k + (row0 + r) * ldk + head * kHd + c * 8));
    *reinterpret_cast<uint4*>(sm + L.k + SwQK::offset(r, c)) = kv;
    const uint4 vv = __ldg(reinterpret_cast<const uint4*>(v + (row0 + r) * ldv + head * kHd + c * 8));
    if (kVT) {
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
      const int blk = r >> 6, col = r & 63;
#pragma unroll
      for (int j = 0; j < 8; ++j) {                          // V^T[d][key]: row d = c*8+j, column = key
        const int d = c * 8 + j;
        *reinterpret_cast<__nv_bfloat16*>(sm + L.vt + blk * 4096 + SwP::offset(d, col >> 3) + (col & 7) * 2) = ve[j];
      }
    } else {
      *reinterpret_cast<uint4*>(sm + L.vt + SwQK::offset(r, c)) = vv;   // V[key][d], same layout as the K tile
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kQB >> 4) << 24);
  const float sl2 = 0.17677669529663687f * 1.4426950408889634f;   // 1/sqrt(32) * log2(e)
  uint32_t phase = 0;

  for (int qb = qb_first; qb < qb_first + qblocks_per_cta && qb * kQB < seq_len; ++qb) {
  const int q0 = qb * kQB;
  // Q block: 128 rows x 4 chunks (the MMAs that read the previous block's Q have completed: their
  // commit was waited for before the last __syncthreads of the previous iteration)
  for (int e = tid; e < kQB * 4; e += kThreads) {
    const int r = e >> 2, c = e & 3;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < seq_len)
      val = __ldg(reinterpret_cast<const uint4*>(q + (row0 + q0 + r) * ldq + head * kHd + c * 8));
    *reinterpret_cast<uint4*>(sm + L.q + SwQK::offset(r, c)) = val;
  }
  fence_proxy_async();            // Q (and, first time, K / V^T) stores -> visible to the tensor core
  __syncthreads();

  float o[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) o[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int kb0 = 0; kb0 < seq_len; kb0 += kKB) {
    const int nk = seq_len - kb0 < kKB ? seq_len - kb0 : kKB;
    const int nk16 = (nk + 15) & ~15;
    // ---- S = Q K^T
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = idesc_base | ((uint32_t)(nk16 >> 3) << 17);
      const uint64_t a_desc = SwQK::desc(sm_u + L.q);
      const uint64_t b_desc = SwQK::desc(sm_u + L.k + (uint32_t)kb0 * 64);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) umma_bf16(tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, kk);
      umma_commit(&s_bar);
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- softmax of this thread's row: sweep 1 = max
    float bmax = -INFINITY;
    for (int c0 = 0; c0 < nk16; c0 += 32) {
      uint32_t sv[32];
      tmem_ld32(lane_base + (uint32_t)c0, sv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < nk) bmax = fmaxf(bmax, __uint_as_float(sv[j]));
    }
    const float m_new = fmaxf(m_run, bmax);
    const float corr = exp2f((m_run - m_new) * sl2);     // first block: exp2(-inf) = 0
    const float mb = m_new * sl2;
    // sweep 2 = exp, row sum, P -> shared memory (bf16, K-major 64-key blocks)
    float psum = 0.f;
    for (int c0 = 0; c0 < nk16; c0 += 32) {
      uint32_t sv[32];
      tmem_ld32(lane_base + (uint32_t)c0, sv);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = c0 + j < nk ? exp2f(__uint_as_float(sv[j]) * sl2 - mb) : 0.f;
        const float p1 = c0 + j + 1 < nk ? exp2f(__uint_as_float(sv[j + 1]) * sl2 - mb) : 0.f;
        psum += p0 + p1;
        __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
      const int blk = c0 >> 6, ch0 = (c0 & 63) >> 3;     // 32 columns = 4 chunks of a 64-key block
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        *reinterpret_cast<uint4*>(sm + L.p + blk * 16384 + SwP::offset(tid, ch0 + cc)) =
            make_uint4(pk[4 * cc], pk[4 * cc + 1], pk[4 * cc + 2], pk[4 * cc + 3]);
    }
    l_run = l_run * corr + psum;
    m_run = m_new;
    fence_proxy_async();            // P (generic-proxy stores) -> visible to the tensor core
    tc_fence_before();              // and this thread's TMEM reads of S are done
    __syncthreads();
    // ---- O_blk = P V  (accumulator reuses TMEM columns 0..31 of S)
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = idesc_base | ((uint32_t)(kHd >> 3) << 17) | (kVT ? 0u : (1u << 16));   // bit 16: B MN-major
      for (int ks = 0; ks < nk16 / 16; ++ks) {
        const int blk = ks >> 2, kk = ks & 3;
        const uint64_t a_desc = SwP::desc(sm_u + L.p + (uint32_t)blk * 16384) + (uint64_t)(kk * 2);
        const uint64_t b_desc =
            kVT ? SwP::desc(sm_u + L.vt + (uint32_t)((kb0 >> 6) + blk) * 4096) + (uint64_t)(kk * 2)
                : SwQK::desc(sm_u + L.vt + (uint32_t)(kb0 + ks * 16) * 64);   // 16 keys x 64-byte rows per K step
        umma_bf16(tmem, a_desc, b_desc, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(&s_bar);
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1u;
    tc_fence_after();
    {
      uint32_t ov[32];
      tmem_ld32(lane_base, ov);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < kHd; ++d) o[d] = o[d] * corr + __uint_as_float(ov[d]);
    }
    tc_fence_before();
    __syncthreads();                // every row has read O_blk before the next S overwrites TMEM / P
  }

  if (q0 + tid < seq_len) {
    const float inv = 1.f / l_run;
    uint32_t pk[16];
#pragma unroll
    for (int d = 0; d < kHd; d += 2) {
      __nv_bfloat162 h = __floats2bfloat162_rn(o[d] * inv, o[d + 1] * inv);
      pk[d >> 1] = *reinterpret_cast<uint32_t*>(&h);
    }
    uint4* op = reinterpret_cast<uint4*>(out + (row0 + q0 + tid) * (size_t)(heads * kHd) + head * kHd);
#pragma unroll
    for (int c = 0; c < 4; ++c) op[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
  }
  }   // query-block loop
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

}  // namespace mha

bool mha_tc_supported(int seq_len, int ldq, int ldk, int ldv, const void* q, const void* k, const void* v,
                      const void* out) {
  if (seq_len < 1 || seq_len > mha::kMaxKeys) return false;
  if ((ldq | ldk | ldv) % 8 != 0) return false;                       // 16-byte row chunks
  if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) != 0) return false;
  return mha::make_layout((seq_len + 63) & ~63).total + 1024 <= 227 * 1024;
}

int mha_core_tc(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int n_seq,
                int seq_len, int heads, void* out, cudaStream_t st) {
  using namespace mha;
  const size_t smem = make_layout((seq_len + 63) & ~63).total + 1024;
  const bool vmn = getenv("U3D_MHA_VMN") != nullptr && atoi(getenv("U3D_MHA_VMN")) == 1;
  static int cur_smem = 0, cur_smem_vmn = 0;
  if (vmn) U3D_CUDA(ensure_dynamic_smem(k_mha_tc<false>, smem, &cur_smem_vmn));
  else U3D_CUDA(ensure_dynamic_smem(k_mha_tc<true>, smem, &cur_smem));
  // one CTA per (head, sequence) looping over the query blocks (K / V^T staged once), unless the
  // launch would leave SMs idle (few sequences) or U3D_MHA_QSPLIT=1 asks for one CTA per query block
  const int n_qb = cdiv(seq_len, kQB);
  bool split = heads * n_seq < 2 * kNumSMs;
  if (const char* e = getenv("U3D_MHA_QSPLIT")) split = atoi(e) != 0;
  const int per_cta = split ? 1 : n_qb;
  dim3 grid(cdiv(n_qb, per_cta), heads, n_seq);
  if (vmn)
    k_mha_tc<false><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                   (const __nv_bfloat16*)v, ldq, ldk, ldv, seq_len, heads,
                                                   per_cta, (__nv_bfloat16*)out);
  else
    k_mha_tc<true><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                  (const __nv_bfloat16*)v, ldq, ldk, ldv, seq_len, heads,
                                                  per_cta, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // namespace u3d

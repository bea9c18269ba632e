// K7 — dense row-wise linear layer with a fused epilogue on tcgen05: the decoder / head GEMMs.
//
//   out = act2( LN( act1(A @ W^T + bias) * mul + res1 + res2 ) ),   out2 = out + add2
//
// Reference: every nn.Linear of the DETR decoder and the heads, together with the elementwise
// work the reference runs as separate kernels around them -
//   projects/mmdet3d_plugin/models/utils/uni3detr_transformer.py:18-30 (MLP), :179-186 (query_pos =
//   query_scale(output) * ref_point_head(sine)), :329-360 (UniCrossAtten output_proj + residual +
//   position encoder), mmcv BaseTransformerLayer (self-attn out-proj + identity + LN,
//   FFN + identity + LN), projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:365-411
//   (cls / reg / iou branches: Linear-LN-ReLU / Linear-ReLU stacks).
//
// One persistent CTA per SM, 128-row tiles, the full output row (N <= 256 per pass) in one fp32
// TMEM accumulator so that LayerNorm needs no cross-thread reduction:
//   warp 0   producer (one lane): per 64-wide K chunk one TMA tensor load of the A box
//            (128 rows x 64 bf16, hardware SWIZZLE_128B, rows beyond the matrix zero-filled) and one
//            cp.async.bulk of the pre-swizzled (N x 64) weight image; 4-stage mbarrier ring.
//   warp 1   MMA issue (one lane): K/16 x tcgen05.mma M=128, N<=256 per tile; accumulators
//            double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps tile i+1.
//   warps 2-5 epilogue: thread = output row (TMEM lane), tcgen05.ld 32 columns at a time; bias, ReLU,
//            elementwise multiplier, up to two residual rows; LayerNorm statistics in fp32 over the
//            row (the pre-norm row is parked back in TMEM with tcgen05.st), second sweep normalises;
//            bf16 (or fp32 for the narrow final heads) stores, optional second output out + add2
//            (the next GEMM's A operand, e.g. x + query_pos).
// N > 256 (in_proj 512, FFN 512) runs as independent 256-column halves.
#include <stdlib.h>
#include "tma.cuh"

namespace u3d {
namespace lin {

using namespace tc;
using namespace tma;

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStages = 3;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kInWarp = 2 + kEpiWarps;            // warp 10: loader of the TMA-staged epilogue operand
constexpr int kThreads = (kInWarp + 1) * 32;      // 352
constexpr int kABytes = kBM * kBK * 2;            // 16 KB per stage
constexpr int kSlab = 64;                         // output columns per staging slab (128-byte rows)
constexpr int kSlabBytes = kBM * kSlab * 2;       // 16 KB

enum : int {
  F_RELU1 = 1,      // ReLU right after the bias
  F_MUL = 2,        // * mul[r][c]
  F_RES1 = 4,
  F_RES2 = 8,
  F_LN = 16,
  F_RELU2 = 32,     // ReLU on the final value
  F_OUT2 = 64,      // out2 = out + add2
  F_OUT_F32 = 128,  // fp32 output (narrow heads)
  F_REF = 256,      // ref_out[r][0..2] = ref_in[r][0..2] + (v0, v1, v4): the decoder's reference refinement
  F_DBG_NOSTORE = 1 << 20,   // profiling switches (scripts/bench_linear.py): no global epilogue traffic,
  F_DBG_NOEPI = 1 << 21,     // epilogue releases the accumulator untouched,
  F_DBG_LDTM = 1 << 22,      // epilogue only reads TMEM
};

struct Params {
  int rows, K, n_pass, n_cols;   // n_pass = columns per pass (<= 256, multiple of 16); n_cols = real output columns
  int n_passes;                  // 1 or 2 (N = n_passes * n_pass)
  int ldo, ldr;                  // output / residual row strides in elements
  int flags;
  int in0_kind;                  // 0 none, 1 = multiplier, 2 = addend: the epilogue operand staged through TMA
  float eps;
  const uint8_t* wpk;            // [pass][K/64] images of (n_pass x 64) bf16, K-major, SWIZZLE_128B
  const float* bias;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* mul;      // operands NOT staged through TMA (thread = row loads); null if staged / absent
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  const __nv_bfloat16* add2;
  void* out;                     // fp32 narrow path only (bf16 outputs leave through the tensor maps)
  const float* ref_in;
  float* ref_out;
};

struct Smem {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t in_full[2];
  uint64_t in_empty[2];
  uint32_t tmem_base;
  alignas(16) float bias[512];
  alignas(16) float gamma[256];
  alignas(16) float beta[256];
  float2 stat[2][kBM];
};

__device__ __forceinline__ void epi_bar(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kEpiThreads) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// 32 consecutive bf16 of a global row (64 bytes, 16-byte aligned) combined into f[32]
template <bool kMul> __device__ __forceinline__ void apply_row32(float* f, const __nv_bfloat16* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    unpack8(__ldg(q + i), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[8 * i + j] = kMul ? f[8 * i + j] * t[j] : f[8 * i + j] + t[j];
  }
}

__global__ void __launch_bounds__(kThreads, 1)
k_linear_tc(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_in0,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
            const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  constexpr uint32_t kHeader = (uint32_t)((sizeof(Smem) + 1023) & ~(size_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (P.rows + kBM - 1) / kBM;
  const int n_items = n_tiles * P.n_passes;
  const int KC = P.K / kBK;
  const uint32_t w_bytes = (uint32_t)P.n_pass * 128u;
  const uint32_t stage_bytes = (uint32_t)kABytes + w_bytes;
  const uint32_t tiles_s = smem_u32(smem_raw) + kHeader;
  const uint32_t out_s = tiles_s + (uint32_t)kStages * stage_bytes;   // 2 output staging slabs
  const uint32_t in_s = out_s + 2u * kSlabBytes;                      // 2 input staging slabs
  const int n_slabs = P.n_pass / kSlab;                               // 0 on the narrow fp32 path

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S.acc_full[b], 1);
      mbar_init(&S.acc_empty[b], kEpiThreads);
      mbar_init(&S.in_full[b], 1);
      mbar_init(&S.in_empty[b], kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // per-column constants of the epilogue -> shared memory (broadcast reads)
  {
    const int ntot = P.n_pass * P.n_passes;
    for (int i = tid; i < 512; i += kThreads) S.bias[i] = (P.bias && i < P.n_cols && i < ntot) ? __ldg(&P.bias[i]) : 0.f;
    if (P.flags & F_LN)
      for (int i = tid; i < P.n_pass; i += kThreads) {
        S.gamma[i] = __ldg(&P.gamma[i]);
        S.beta[i] = __ldg(&P.beta[i]);
      }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // ======================= producer: A boxes (TMA) + weight images (bulk copy) =======================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      int slot = 0;
      uint32_t eph = 1u;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / P.n_passes, pass = item - tile * P.n_passes;
        const uint8_t* wsrc = P.wpk + (size_t)pass * KC * w_bytes;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&S.empty[slot], eph);
          const uint32_t st_s = tiles_s + (uint32_t)slot * stage_bytes;
          mbar_expect_tx(&S.full[slot], stage_bytes);
          tma_load_2d(st_s, &tmap_a, kc * kBK, tile * kBM, &S.full[slot]);
          bulk_g2s(st_s + kABytes, wsrc + (size_t)kc * w_bytes, w_bytes, &S.full[slot]);
          if (++slot == kStages) { slot = 0; eph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      using SW = Swz<64>;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.n_pass >> 3) << 17) |
                             ((uint32_t)(kBM >> 4) << 24);
      int slot = 0, t = 0;
      uint32_t fph = 0u;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++t) {
        const int ab = t & 1;
        mbar_wait(&S.acc_empty[ab], ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(ab * 256);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&S.full[slot], fph);
          tc_fence_after();
          const uint32_t st_s = tiles_s + (uint32_t)slot * stage_bytes;
          const uint64_t a_desc = SW::desc(st_s);
          const uint64_t b_desc = SW::desc(st_s + kABytes);
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk)
            umma_bf16(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc,
                      (kc > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&S.empty[slot]);
          if (++slot == kStages) { slot = 0; fph ^= 1u; }
        }
        umma_commit(&S.acc_full[ab]);
      }
      tc_fence_before();
    }
  } else if (warp == kInWarp) {
    // ======================= loader of the TMA-staged epilogue operand (multiplier or addend) ============
    if (lane == 0 && P.in0_kind != 0 && n_slabs > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_in0)) : "memory");
      int ii = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / P.n_passes, pass = item - tile * P.n_passes;
        for (int slab = 0; slab < n_slabs; ++slab, ++ii) {
          const int ib = ii & 1;
          mbar_wait(&S.in_empty[ib], (((uint32_t)(ii >> 1)) & 1u) ^ 1u);
          mbar_expect_tx(&S.in_full[ib], (uint32_t)kSlabBytes);
          tma_load_2d(in_s + (uint32_t)ib * kSlabBytes, &tmap_in0, pass * P.n_pass + slab * kSlab, tile * kBM,
                      &S.in_full[ib]);
        }
      }
    }
  } else {
    // ======================= epilogue: 8 warps, thread = (row, 32-column half of a 64-column slab) ========
    using SW = Swz<64>;
    const int e = warp - 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int h = e >> 2;                         // column half inside a slab
    const int row = q * 32 + lane;                // row of the tile
    const bool elected = (e == 0 && lane == 0);   // issues the TMA stores
    const int flags = P.flags;
    const int N = P.n_pass;
    const bool dbg_nostore = (flags & F_DBG_NOSTORE) != 0;
    int t = 0, jj = 0, ii = 0;                    // tile, output staging job, input staging slab counters
    if (elected) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_out)) : "memory");
    }
    // one output box: every warp stages its own 32 rows x 32 columns (2 KB, SWIZZLE_64B) in a private double buffer
    // and its lane 0 issues the TMA store - no CTA-wide barrier in the epilogue's inner loop (the first form staged
    // 64-column slabs of the whole tile behind two named barriers per slab: 12.7 -> 11.9 us plain, 32.9 -> 27.0 us
    // for the mul + second-output launch, LayerNorm launches 19.9 -> 19.5 us).
    using SW32 = Swz<32>;
    uint8_t* wstage = smem_raw + (out_s - smem_u32(smem_raw)) + (uint32_t)e * 4096u;   // 2 x 2 KB per warp
    auto emit = [&](const float* f, const CUtensorMap* mapw, int col, int tile) {
      uint8_t* dst = wstage + (uint32_t)(jj & 1) * 2048u;
      if (lane == 0) bulk_wait_read<1>();         // this lane's store that last read the buffer has left shared memory
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(dst + SW32::offset(lane, i)) = pack8(f + 8 * i);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !dbg_nostore) {
        tma_store_2d(mapw, col + h * 32, tile * kBM + q * 32, smem_u32(dst));
        bulk_commit();
      }
      ++jj;
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++t) {
      const int tile = item / P.n_passes, pass = item - tile * P.n_passes;
      const int ab = t & 1;
      const int r = tile * kBM + row;
      const bool row_ok = r < P.rows && !dbg_nostore;
      const int cbase = pass * N;                 // first output column of this pass
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256);
      mbar_wait_relaxed(&S.acc_full[ab], (uint32_t)(t >> 1) & 1u, 1000u);
      tc_fence_after();
      if (flags & F_DBG_NOEPI) {
        tc_fence_before();
        mbar_arrive(&S.acc_empty[ab]);
        continue;
      }
      if (n_slabs == 0) {
        // ---- narrow fp32 heads (N <= 48 columns, no staging): the h == 0 warps store directly
        if (h == 0) {
          for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(lane_base + (uint32_t)c0, v);
            tmem_ld_wait();
            if (row_ok) {
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                f[j] = __uint_as_float(v[j]) + S.bias[c0 + j];
                if (flags & F_RELU1) f[j] = fmaxf(f[j], 0.f);
              }
              float* o = reinterpret_cast<float*>(P.out) + (size_t)r * P.ldo;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < P.n_cols) o[c0 + j] = f[j];
              if ((flags & F_REF) && c0 == 0) {
                P.ref_out[(size_t)r * 3 + 0] = __ldg(&P.ref_in[(size_t)r * 3 + 0]) + f[0];
                P.ref_out[(size_t)r * 3 + 1] = __ldg(&P.ref_in[(size_t)r * 3 + 1]) + f[1];
                P.ref_out[(size_t)r * 3 + 2] = __ldg(&P.ref_in[(size_t)r * 3 + 2]) + f[4];
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&S.acc_empty[ab]);
        continue;
      }
      float sum = 0.f, sumsq = 0.f;
      for (int slab = 0; slab < n_slabs; ++slab) {
        const int c0 = slab * kSlab + h * 32;     // this thread's 32 columns of the pass
        uint32_t v[32];
        tmem_ld32(lane_base + (uint32_t)c0, v);
        tmem_ld_wait();
        if (flags & F_DBG_LDTM) continue;
        float f[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(&S.bias[cbase + c0 + 4 * i]);
          f[4 * i] = __uint_as_float(v[4 * i]) + b4.x;
          f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
          f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
          f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
        }
        if (flags & F_RELU1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (P.in0_kind != 0) {                    // the TMA-staged operand slab (same swizzled layout as the output)
          const int ib = ii & 1;
          mbar_wait(&S.in_full[ib], (uint32_t)(ii >> 1) & 1u);
          const uint8_t* src = smem_raw + (in_s - smem_u32(smem_raw)) + (uint32_t)ib * kSlabBytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float tv[8];
            unpack8(*reinterpret_cast<const uint4*>(src + SW::offset(row, h * 4 + i)), tv);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[8 * i + j] = P.in0_kind == 1 ? f[8 * i + j] * tv[j] : f[8 * i + j] + tv[j];
          }
          mbar_arrive(&S.in_empty[ib]);
          ++ii;
        }
        if (row_ok) {
          const size_t off = (size_t)r * P.ldr + cbase + c0;
          if (P.mul) apply_row32<true>(f, P.mul + off);
          if (P.res1) apply_row32<false>(f, P.res1 + off);
          if (P.res2) apply_row32<false>(f, P.res2 + off);
        }
        if (flags & F_LN) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sum += f[j];
            sumsq += f[j] * f[j];
            v[j] = __float_as_uint(f[j]);
          }
          tmem_st32(lane_base + (uint32_t)c0, v);      // park the pre-norm row in the accumulator
        } else {
          if (flags & F_RELU2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          emit(f, &tmap_out, cbase + slab * kSlab, tile);
          if (flags & F_OUT2) {
            if (row_ok) apply_row32<false>(f, P.add2 + (size_t)r * P.ldr + cbase + c0);
            emit(f, &tmap_out2, cbase + slab * kSlab, tile);
          }
        }
      }
      if ((flags & F_LN) && !(flags & F_DBG_LDTM)) {
        tmem_st_wait();
        S.stat[h][row] = make_float2(sum, sumsq);
        epi_bar(3);
        const float2 s0 = S.stat[0][row], s1 = S.stat[1][row];
        const float mean = (s0.x + s1.x) / (float)N;
        const float var = fmaxf((s0.y + s1.y) / (float)N - mean * mean, 0.f);
        const float rstd = rsqrtf(var + P.eps);
        for (int slab = 0; slab < n_slabs; ++slab) {
          const int c0 = slab * kSlab + h * 32;
          uint32_t v[32];
          tmem_ld32(lane_base + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 g4 = *reinterpret_cast<const float4*>(&S.gamma[c0 + 4 * i]);
            const float4 b4 = *reinterpret_cast<const float4*>(&S.beta[c0 + 4 * i]);
            f[4 * i] = (__uint_as_float(v[4 * i]) - mean) * rstd * g4.x + b4.x;
            f[4 * i + 1] = (__uint_as_float(v[4 * i + 1]) - mean) * rstd * g4.y + b4.y;
            f[4 * i + 2] = (__uint_as_float(v[4 * i + 2]) - mean) * rstd * g4.z + b4.z;
            f[4 * i + 3] = (__uint_as_float(v[4 * i + 3]) - mean) * rstd * g4.w + b4.w;
          }
          if (flags & F_RELU2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          emit(f, &tmap_out, slab * kSlab, tile);
          if (flags & F_OUT2) {
            if (row_ok) apply_row32<false>(f, P.add2 + (size_t)r * P.ldr + c0);
            emit(f, &tmap_out2, slab * kSlab, tile);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&S.acc_empty[ab]);
    }
    if (lane == 0) bulk_wait_all();               // every output box of this warp has reached global memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// Linear(3 -> C) + LayerNorm + ReLU: the first half of UniCrossAtten.position_encoder
// (uni3detr_transformer.py:256-260). K = 3 is no GEMM: one warp per row, 8 channels per lane (C = 256).
template <typename T>
__global__ void __launch_bounds__(256)
k_pos3_ln_relu(const float* __restrict__ ref, const float* __restrict__ w,   // w (C,3)
               const float* __restrict__ b, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, int rows, int C, T* __restrict__ out) {
  // a warp keeps the parameters of its 8 channels per lane in registers and walks rows grid-stride:
  // per row 3 loads, 24 FMAs, two 5-step shuffle reductions and one 16-byte (bf16) store per lane
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float wx[8], wy[8], wz[8], bb[8], gg[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane * 8 + j;
    const bool ok = c < C;
    wx[j] = ok ? __ldg(&w[c * 3]) : 0.f;
    wy[j] = ok ? __ldg(&w[c * 3 + 1]) : 0.f;
    wz[j] = ok ? __ldg(&w[c * 3 + 2]) : 0.f;
    bb[j] = ok ? __ldg(&b[c]) : 0.f;
    gg[j] = ok ? __ldg(&gamma[c]) : 0.f;
    be[j] = ok ? __ldg(&beta[c]) : 0.f;
  }
  const float invC = 1.f / (float)C;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const float x = __ldg(&ref[(size_t)row * 3]), y = __ldg(&ref[(size_t)row * 3 + 1]), z = __ldg(&ref[(size_t)row * 3 + 2]);
    float v[8];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = lane * 8 + j < C ? fmaf(wx[j], x, fmaf(wy[j], y, fmaf(wz[j], z, bb[j]))) : 0.f;
      s += v[j];
      ss = fmaf(v[j], v[j], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const float mean = s * invC;
    const float rstd = rsqrtf(fmaxf(ss * invC - mean * mean, 0.f) + eps);
    T* o = out + (size_t)row * C + lane * 8;
    if (lane * 8 + 8 <= C && sizeof(T) == 2) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float a0 = fmaxf((v[j] - mean) * rstd * gg[j] + be[j], 0.f);
        const float a1 = fmaxf((v[j + 1] - mean) * rstd * gg[j + 1] + be[j + 1], 0.f);
        __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane * 8 + j < C) o[j] = from_f32<T>(fmaxf((v[j] - mean) * rstd * gg[j] + be[j], 0.f));
    }
  }
}

// Box assembly of Uni3DETRHead.forward (dense_heads/uni3detr_head.py:470-496): the reference point goes
// through sigmoid -> inverse_sigmoid(eps = 1e-5) (transformer :129 -> head :475), is added to the raw
// (x, y) = tmp[0:2] and z = tmp[4] offsets, squashed and scaled to pc_range; the other code entries pass.
__global__ void k_box_assemble(const float* __restrict__ tmp, const float* __restrict__ ref_logit, int rows, int code,
                               float x0, float y0, float z0, float sx, float sy, float sz, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* t = tmp + (size_t)r * code;
  float* o = out + (size_t)r * code;
  float inv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = 1.f / (1.f + expf(-__ldg(&ref_logit[(size_t)r * 3 + c])));
    s = fminf(fmaxf(s, 0.f), 1.f);
    inv[c] = logf(fmaxf(s, 1e-5f) / fmaxf(1.f - s, 1e-5f));
  }
  const float bx = 1.f / (1.f + expf(-(t[0] + inv[0])));
  const float by = 1.f / (1.f + expf(-(t[1] + inv[1])));
  const float bz = 1.f / (1.f + expf(-(t[4] + inv[2])));
  for (int c = 0; c < code; ++c) o[c] = t[c];
  o[0] = bx * sx + x0;
  o[1] = by * sy + y0;
  o[4] = bz * sz + z0;
}

}  // namespace lin
}  // namespace u3d

using namespace u3d;

extern "C" {

size_t u3d_linear_packed_bytes(int N, int K) {
  if (N < 1 || K < 64 || K % 64 != 0) return 0;
  const int n_pass = N > 256 ? 256 : (N + 15) / 16 * 16;
  if (N > 256 && N % 256 != 0) return 0;
  const int passes = N > 256 ? N / 256 : 1;
  if (passes > 2) return 0;
  return (size_t)passes * (K / 64) * n_pass * 128;
}

// packs W (N, K) bf16 row-major (nn.Linear.weight) into [pass][K/64] images of (n_pass x 64) bf16 in the
// UMMA K-major SWIZZLE_128B layout (rows >= N zero)
__global__ void k_linear_pack(const __nv_bfloat16* __restrict__ w, int N, int K, int n_pass, int passes,
                              __nv_bfloat16* __restrict__ packed) {
  const int KC = K / 64;
  const long long total = (long long)passes * KC * n_pass * 8;   // 16-byte chunks
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(i & 7);
    long long t = i >> 3;
    const int n = (int)(t % n_pass);
    t /= n_pass;
    const int kc = (int)(t % KC);
    const int pass = (int)(t / KC);
    const int row = pass * n_pass + n;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (row < N) val = *reinterpret_cast<const uint4*>(w + (size_t)row * K + kc * 64 + chunk * 8);
    uint8_t* img = reinterpret_cast<uint8_t*>(packed) + ((size_t)pass * KC + kc) * n_pass * 128;
    *reinterpret_cast<uint4*>(img + tc::Swz<64>::offset(n, chunk)) = val;
  }
}

int u3d_linear_pack_weights(const void* w, int N, int K, void* packed, void* stream) {
  const size_t bytes = u3d_linear_packed_bytes(N, K);
  U3D_CHECK_ARG(bytes != 0, "linear pack: N=%d K=%d unsupported (K %% 64 == 0, N <= 256 or N == 512)", N, K);
  U3D_CHECK_ARG((((uintptr_t)w | (uintptr_t)packed) & 15) == 0, "linear pack: buffers must be 16-byte aligned");
  const int n_pass = N > 256 ? 256 : (N + 15) / 16 * 16;
  const int passes = N > 256 ? N / 256 : 1;
  const long long total = (long long)bytes / 16;
  k_linear_pack<<<cdiv(total, 256) < 1024 ? cdiv(total, 256) : 1024, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)w, N, K, n_pass, passes, (__nv_bfloat16*)packed);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

int u3d_linear_tc(const void* a, int lda, int rows, int K, const void* w_packed, int N, const float* bias,
                  int flags, const void* mul, const void* res1, const void* res2, int ldr,
                  const float* gamma, const float* beta, float eps, const void* add2, void* out2,
                  void* out, int ldo, const float* ref_in, float* ref_out, void* stream) {
  using namespace lin;
  U3D_CHECK_ARG(u3d_linear_packed_bytes(N, K) != 0, "linear: N=%d K=%d unsupported", N, K);
  U3D_CHECK_ARG(rows >= 0 && lda >= K && lda % 8 == 0, "linear: lda=%d must be >= K=%d and a multiple of 8", lda, K);
  if (rows == 0) return U3D_OK;
  Params P;
  P.rows = rows;
  P.K = K;
  P.n_passes = N > 256 ? N / 256 : 1;
  P.n_pass = N > 256 ? 256 : (N + 15) / 16 * 16;
  P.n_cols = N;
  P.ldo = ldo;
  P.ldr = ldr;
  P.flags = flags;
  P.eps = eps;
  P.wpk = (const uint8_t*)w_packed;
  P.bias = bias;
  P.gamma = gamma;
  P.beta = beta;
  P.mul = (const __nv_bfloat16*)mul;
  P.res1 = (const __nv_bfloat16*)res1;
  P.res2 = (const __nv_bfloat16*)res2;
  P.add2 = (const __nv_bfloat16*)add2;
  P.out = out;
  P.ref_in = ref_in;
  P.ref_out = ref_out;
  const bool f32 = (flags & F_OUT_F32) != 0;
  U3D_CHECK_ARG(!(flags & F_LN) || (P.n_passes == 1 && P.n_pass == N && gamma && beta),
                "linear: LayerNorm needs the whole row in one pass (N=%d <= 256)", N);
  U3D_CHECK_ARG(f32 ? N <= 48 : (N % 64 == 0 && ldo % 8 == 0),
                "linear: bf16 outputs need N %% 64 == 0 and ldo %% 8 == 0; fp32 outputs N <= 48 (N=%d)", N);
  U3D_CHECK_ARG(!f32 || !(flags & (F_MUL | F_RES1 | F_RES2 | F_OUT2 | F_LN | F_RELU2)),
                "linear: the narrow fp32 path takes bias / ReLU / reference refinement only");
  U3D_CHECK_ARG(!(flags & (F_MUL | F_RES1 | F_RES2 | F_OUT2)) || ldr % 8 == 0, "linear: epilogue operands need ldr %% 8 == 0");
  U3D_CHECK_ARG(!(flags & F_MUL) || mul, "linear: F_MUL without a multiplier");
  U3D_CHECK_ARG(!(flags & F_RES1) || res1, "linear: F_RES1 without res1");
  U3D_CHECK_ARG(!(flags & F_RES2) || res2, "linear: F_RES2 without res2");
  U3D_CHECK_ARG(!(flags & F_OUT2) || (add2 && out2 && !f32), "linear: F_OUT2 needs add2, out2 and a bf16 output");
  U3D_CHECK_ARG(!(flags & F_REF) || (f32 && ref_in && ref_out && N >= 5), "linear: F_REF needs an fp32 output with >= 5 columns");
  U3D_CHECK_ARG((((uintptr_t)a | (uintptr_t)w_packed | (uintptr_t)out | (uintptr_t)mul | (uintptr_t)res1 |
                  (uintptr_t)res2 | (uintptr_t)add2 | (uintptr_t)out2) & 15) == 0,
                "linear: buffers must be 16-byte aligned");
  if (!(flags & F_MUL)) P.mul = nullptr;
  if (!(flags & F_RES1)) P.res1 = nullptr;
  if (!(flags & F_RES2)) P.res2 = nullptr;

  CUtensorMap tmap_a, tmap_in0, tmap_out, tmap_out2;
  if (tma::encode_2d_bf16(&tmap_a, a, K, rows, lda, kBK, kBM) != U3D_OK) return U3D_EINVAL;
  tmap_in0 = tmap_a;
  tmap_out = tmap_a;
  tmap_out2 = tmap_a;
  // the first epilogue operand (multiplier, else the first residual) is staged through TMA; the others
  // (rare: output_proj's second residual, query_scale's `+ x`) are read by the row threads
  P.in0_kind = 0;
  if (flags & (F_DBG_NOEPI | F_DBG_LDTM)) P.mul = P.res1 = P.res2 = nullptr;   // profiling: epilogue operands unused
  if (!f32) {
    const void* in0 = nullptr;
    if (P.mul) { in0 = P.mul; P.in0_kind = 1; P.mul = nullptr; }
    else if (P.res1) { in0 = P.res1; P.in0_kind = 2; P.res1 = nullptr; }
    else if (P.res2) { in0 = P.res2; P.in0_kind = 2; P.res2 = nullptr; }
    if (in0 && tma::encode_2d_bf16(&tmap_in0, in0, N, rows, ldr, kSlab, kBM) != U3D_OK) return U3D_EINVAL;
    // outputs leave as per-warp boxes of 32 columns x 32 rows
    if (tma::encode_2d_bf16(&tmap_out, out, N, rows, ldo, 32, 32) != U3D_OK) return U3D_EINVAL;
    if ((flags & F_OUT2) && tma::encode_2d_bf16(&tmap_out2, out2, N, rows, ldr, 32, 32) != U3D_OK) return U3D_EINVAL;
  }

  const size_t header = (sizeof(Smem) + 1023) & ~(size_t)1023;
  const size_t smem = header + (size_t)kStages * (kABytes + (size_t)P.n_pass * 128) + 4 * (size_t)kSlabBytes;
  static int cur_smem = 0;
  U3D_CUDA(ensure_dynamic_smem(k_linear_tc, smem, &cur_smem));
  const int items = cdiv(rows, kBM) * P.n_passes;
  const int grid = items < kNumSMs ? items : kNumSMs;
  k_linear_tc<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmap_a, tmap_in0, tmap_out, tmap_out2, P);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

int u3d_box_assemble(const float* tmp, const float* ref_logit, int rows, int code, const float* pc_range,
                     float* out, void* stream) {
  U3D_CHECK_ARG(tmp && ref_logit && out && pc_range && code >= 6, "box_assemble: bad argument (code=%d)", code);
  if (rows <= 0) return U3D_OK;
  lin::k_box_assemble<<<cdiv(rows, 128), 128, 0, (cudaStream_t)stream>>>(
      tmp, ref_logit, rows, code, pc_range[0], pc_range[1], pc_range[2], pc_range[3] - pc_range[0],
      pc_range[4] - pc_range[1], pc_range[5] - pc_range[2], out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

int u3d_pos3_ln_relu(const float* ref, const float* w, const float* b, const float* gamma, const float* beta,
                     float eps, int rows, int C, void* out, int dtype, void* stream) {
  U3D_CHECK_ARG(C >= 1 && C <= 256, "pos3_ln_relu: C=%d must be <= 256", C);
  if (rows <= 0) return U3D_OK;
  const int wpb = 8;
  int grid = cdiv(rows, wpb * 4);                 // >= 4 rows per warp: the per-warp parameter load amortises
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (grid < 1) grid = 1;
  U3D_CHECK_ARG(C % 8 == 0 || dtype == U3D_F32 || true, "pos3_ln_relu");
  if (dtype == U3D_BF16)
    lin::k_pos3_ln_relu<__nv_bfloat16><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(
        ref, w, b, gamma, beta, eps, rows, C, (__nv_bfloat16*)out);
  else
    lin::k_pos3_ln_relu<float><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(ref, w, b, gamma, beta, eps, rows, C,
                                                                             (float*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // extern "C"

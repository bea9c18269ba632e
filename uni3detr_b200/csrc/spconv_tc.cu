// K3 (tensor-core flavour) — sparse conv forward as an output-stationary gather-GEMM on the
// 5th-generation tensor cores: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with the accumulator
// tile in TMEM, operands staged in shared memory through an mbarrier ring.
//
// Reference semantics: SURVEY.md A.3/A.4 (spconv indice_conv: out[o] += in[i] @ W[k] over the
// rulebook pairs, then BatchNorm1d(eval) + ReLU, SparseBasicBlock identity add) for the layers of
// projects/mmdet3d_plugin/models/pts_encoder/sparse_encoder_hd.py:106-132.
//
// Persistent, warp-specialised kernel: one CTA per SM loops over 128-row output tiles (UMMA M =
// 128 = the 128 TMEM lanes; N = Cout <= 256 per instruction, two instructions for Cout = 512).
// The reduction of a tile runs over its ACTIVE kernel offsets (per-tile bit masks written by the
// rulebook build: offsets whose 128 entries are all empty are skipped) and over Cin in blocks of
// CIN_BLK in {16,32,64} elements; one pipeline stage always carries K = 64 reduction columns
// (one 64-column block of one offset, or 2 / 4 whole offsets for Cin = 32 / 16).
//   warp 5      slice loader: bulk-copies (cp.async.bulk) the active 512-byte rows of the next
//               tiles' rulebook slices into a 3-deep shared-memory ring.
//   warps 6-13  producers, one STAGE per warp at a time (warp w owns ring slots w, w+8): the
//               warp gathers the stage's 128 input rows x CIN_BLK bf16 with 16-byte cp.async
//               (zero-fill for missing neighbours) into the K-major stage with the hardware
//               swizzle of 2*CIN_BLK bytes - 32 copies per lane in an unrolled loop, so barrier and
//               loop overhead is paid once per stage per warp - and publishes it with
//               cp.async.mbarrier.arrive.noinc (the hardware arrives when the copies land; no
//               producer-side wait or fence). Lane 0 adds the stage's weight tile W[k][:, block] -
//               a pre-swizzled (Cout x CIN_BLK) image from u3d_spconv_pack_weights - with ONE
//               cp.async.bulk (TMA) completing on the same mbarrier.
//               (A cp.async.bulk.tensor tile::gather4 producer was measured 2-3x slower than this
//               LSU gather for 32..128-byte rows: 0.47 vs 0.21 ms on the 64->64 layers.)
//   warp 4      MMA: one thread issues CIN_BLK/16 tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) per
//               stage, releases the stage with tcgen05.commit, and commits the tile's accumulator.
//   warps 0-3   epilogue: tcgen05.ld (32x32b.x16) of their TMEM lane quarter, scale/shift
//               (+residual) (+ReLU), bf16 rows out with 16-byte stores. The accumulator is
//               double-buffered in TMEM (2 x Cout columns), so the epilogue of tile i overlaps the
//               gathers and MMAs of tile i+1.
#include <stdlib.h>
#include "tc_common.cuh"

namespace u3d {

namespace tc {

constexpr int kRows = 128;             // UMMA M
constexpr int kEpiThreads = 128;       // warps 0-3: epilogue (TMEM lane quarter = warp id)
constexpr int kMmaWarp = 4;            // warp 4: MMA issue + TMEM allocation
constexpr int kSliceWarp = 5;          // warp 5: rulebook-slice loader
constexpr int kProdWarp0 = 6;          // warps 6-13: producers
constexpr int kNumProd = 8;            // producer warps with one CTA per SM (4 with two)
constexpr int kThreads = (kProdWarp0 + kNumProd) * 32;   // 448
constexpr int kSliceBufs = 2;
constexpr int kMaxK = 27;
constexpr int kMaxStages = 16;

struct Smem {
  uint64_t full[kMaxStages];    // 32 cp.async arrives (owning warp) + 1 arrive.expect_tx (weight TMA)
  uint64_t empty[kMaxStages];   // tcgen05.commit: the MMAs that read the stage have retired
  uint64_t acc_full[2];         // tcgen05.commit: accumulator of a tile is complete
  uint64_t acc_empty[2];        // 128 epilogue threads: accumulator drained
  uint64_t slice_full[kSliceBufs];   // tx bytes: rulebook slice of a tile has landed
  uint64_t slice_empty[kSliceBufs];  // the active producer warps are done with the slice
  uint32_t tmem_base;
  float scale[512];
  float shift[512];
  alignas(128) int nbr[kSliceBufs][kMaxK][kRows];
};

template <int CIN_BLK, int CTAS>
__global__ void __launch_bounds__(CTAS == 1 ? kThreads : (kProdWarp0 + kNumProd / 2) * 32, CTAS)
k_spconv_tc(const __nv_bfloat16* __restrict__ in, const int32_t* __restrict__ nbr, int nbr_stride,
            const uint32_t* __restrict__ tile_mask, const int32_t* __restrict__ n_out_p, int K,
            const __nv_bfloat16* __restrict__ wpk, const float* __restrict__ scale,
            const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual, int relu,
            __nv_bfloat16* __restrict__ out, int Cin, int Cout, int stages, int acc_bufs,
            uint32_t tmem_cols, int dbg) {
  using SW = Swz<CIN_BLK>;
  constexpr int kNP = kNumProd / CTAS;            // producer warps of this CTA
  constexpr int kNT = (kProdWarp0 + kNP) * 32;    // threads of this CTA
  constexpr int kChunks = CIN_BLK / 8;            // 16-byte chunks per A row
  constexpr int kRowsPerPass = 32 / kChunks;      // rows one warp-wide cp.async covers
  constexpr int kPasses = kRows / kRowsPerPass;   // 32 / 16 / 8 copies per lane per unit
  constexpr int kG = 64 / CIN_BLK;                // (offset, Cin-block) units per stage: K = 64 per stage
  constexpr int kUnitA = kRows * SW::P;           // one 128-row A sub-tile
  constexpr int kABytes = kG * kUnitA;            // 16 KB for every CIN_BLK
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  constexpr uint32_t kHeader = (uint32_t)((sizeof(Smem) + 1023) & ~(size_t)1023);

  const int n_out = *n_out_p;
  const int n_tiles = (n_out + kRows - 1) / kRows;
  if ((int)blockIdx.x >= n_tiles) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = Cin / CIN_BLK;
  const uint32_t b_bytes = (uint32_t)Cout * SW::P;                       // one weight sub-tile
  const uint32_t b_unit = (b_bytes + 1023u) & ~1023u;                     // 1024-aligned sub-tiles
  const uint32_t stage_bytes = kABytes + kG * b_unit;
  const uint32_t tiles_s = smem_u32(smem_raw) + kHeader;  // 1024-aligned (dynamic smem base is)
  const uint32_t all_mask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);

  // ---- prologue: barriers, TMEM, epilogue constants
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&S.full[s], 32 + 1);
      mbar_init(&S.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S.acc_full[b], 1);
      mbar_init(&S.acc_empty[b], kEpiThreads);
    }
    for (int b = 0; b < kSliceBufs; ++b) {
      mbar_init(&S.slice_full[b], 1);
      mbar_init(&S.slice_empty[b], stages < kNP ? stages : kNP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&S.tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = tid; c < Cout; c += kNT) {
    S.scale[c] = scale ? __ldg(&scale[c]) : 1.f;
    S.shift[c] = shift ? __ldg(&shift[c]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp >= kProdWarp0) {
    // ======================= producers: one stage per warp at a time =======================
    const int w = warp - kProdWarp0;
    const int chunk = lane % kChunks;
    const int rsub = lane / kChunks;
    // per-lane constants: swizzled destination offset of passes 0 and 1; the XOR term of the swizzle
    // has a period of 8 rows (128-byte rows) or less, i.e. at most 2 passes
    const uint32_t off_even = SW::offset(rsub, chunk);
    const uint32_t off_odd = SW::offset(rsub + kRowsPerPass, chunk) - (uint32_t)(kRowsPerPass * SW::P);
    const uint64_t row_bytes = (uint64_t)Cin * 2;
    // Each active producer warp OWNS ring slots w and w+8 (ring <= 16): it fills the stages
    // g = w, w+8 (mod ring) in increasing order, so it meets the generations of each of its slots
    // in order and the 1-bit mbarrier parity is never ambiguous.
    const bool active = w < stages;
    int ng_a = w, ng_b = (w + kNP < stages) ? w + kNP : 0x7fffffff;   // next stage per slot
    uint32_t eph_a = 1u, eph_b = 1u;   // parity to wait for on empty[slot]; flips on every visit
    int g0 = 0;          // global index of the first stage of the current tile
    int t = 0;
    for (int tile = blockIdx.x; active && tile < n_tiles; tile += gridDim.x, ++t) {
      const int buf = t % kSliceBufs;
      const int m0 = tile * kRows;
      const uint32_t mask = tile_mask ? __ldg(&tile_mask[tile]) : all_mask;
      const int n_units = __popc(mask) * nkb;
      const int n_st = (n_units + kG - 1) / kG;   // stages of this tile (kG units each, last one partial)
      const int rows_live = n_out - m0;   // rulebook entries of rows >= n_out are uninitialised memory
      // every active warp passes through every slice (even one it owns no stage of): a warp can
      // then never arrive twice on slice_empty[buf] within one phase
      if (nbr) mbar_wait(&S.slice_full[buf], (uint32_t)(t / kSliceBufs) & 1u);
      // my stages of this tile
      for (;;) {
        const bool second = ng_b < ng_a;
        const int ng = second ? ng_b : ng_a;
        if (ng >= g0 + n_st) break;
        const int slot = second ? w + kNP : w;
        const uint32_t eph = second ? eph_b : eph_a;
        if (second) { ng_b += stages; eph_b ^= 1u; } else { ng_a += stages; eph_a ^= 1u; }
        const int u0 = (ng - g0) * kG;
        const int cnt = n_units - u0 < kG ? n_units - u0 : kG;
        mbar_wait(&S.empty[slot], eph);
        const uint32_t a_s = tiles_s + (uint32_t)slot * stage_bytes;
        if (lane == 0) mbar_expect_tx(&S.full[slot], (uint32_t)cnt * b_bytes);
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          if (j < cnt) {
            const int u = u0 + j;
            const int ki = u / nkb, kb = u - ki * nkb;
            const int k = __fns(mask, 0, ki + 1);            // position of the ki-th set bit
            if (lane == 0)
              bulk_g2s(a_s + kABytes + (uint32_t)j * b_unit,
                       (const uint8_t*)wpk + ((size_t)k * nkb + kb) * b_bytes, b_bytes, &S.full[slot]);
            const uint8_t* src_base =
                reinterpret_cast<const uint8_t*>(in) + (size_t)(kb * CIN_BLK + chunk * 8) * 2;
            const int* nb = nbr ? &S.nbr[buf][k][rsub] : nullptr;
            const uint32_t a_u = a_s + (uint32_t)(j * kUnitA);
            if (!(dbg & 1))
#pragma unroll
            for (int i = 0; i < kPasses; ++i) {
              const int src_row = nb ? nb[i * kRowsPerPass] : m0 + i * kRowsPerPass + rsub;
              const bool ok = src_row >= 0 && i * kRowsPerPass + rsub < rows_live;
              const uint8_t* src = src_base + (ok ? (uint64_t)(uint32_t)src_row * row_bytes : 0ull);
              cp_async16(a_u + ((i & 1) ? off_odd : off_even) + (uint32_t)(i * kRowsPerPass * SW::P), src,
                         ok ? 16u : 0u);
            }
          }
        }
        // asynchronous publish: the hardware arrives on full[slot] when this lane's copies land
        cp_async_arrive(&S.full[slot]);
      }
      g0 += n_st;
      if (nbr) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.slice_empty[buf]);
      }
    }
  } else if (warp == kSliceWarp) {
    // ======================= rulebook-slice loader (one thread) =======================
    if (lane == 0 && nbr != nullptr) {
      int t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int buf = t % kSliceBufs;
        mbar_wait(&S.slice_empty[buf], (((uint32_t)(t / kSliceBufs)) & 1u) ^ 1u);
        uint32_t m = tile_mask ? __ldg(&tile_mask[tile]) : all_mask;
        mbar_expect_tx(&S.slice_full[buf], (uint32_t)__popc(m) * (uint32_t)(kRows * 4));
        while (m) {
          const int k = __ffs(m) - 1;
          m &= m - 1;
          bulk_g2s(smem_u32(&S.nbr[buf][k][0]), nbr + (size_t)k * nbr_stride + (size_t)tile * kRows,
                   kRows * 4, &S.slice_full[buf]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ======================= MMA issuer (one thread) =======================
    if (lane == 0) {
      const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kRows >> 4) << 24);
      int slot = 0, t = 0;
      uint32_t fph = 0u;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int ab = t % acc_bufs;
        const uint32_t mask = tile_mask ? __ldg(&tile_mask[tile]) : all_mask;
        const int n_units = __popc(mask) * nkb;
        const int n_st = (n_units + kG - 1) / kG;
        mbar_wait(&S.acc_empty[ab], ((uint32_t)(t / acc_bufs) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(ab * Cout);
        for (int st = 0; st < n_st; ++st) {
          const int cnt = n_units - st * kG < kG ? n_units - st * kG : kG;
          mbar_wait(&S.full[slot], fph);
          if (!(dbg & 8)) tc_fence_after();
          const uint32_t a_s = tiles_s + (uint32_t)slot * stage_bytes;
          const uint32_t b_s = a_s + kABytes;
          if (dbg & 2) {   // timing experiment: no MMA, release the stage directly
            mbar_arrive(&S.empty[slot]);
            if (++slot == stages) { slot = 0; fph ^= 1u; }
            continue;
          }
          for (int rep = 0; rep < ((dbg & 4) ? 2 : 1); ++rep)
          for (int n0 = 0; n0 < Cout; n0 += 256) {
            const int n = Cout - n0 < 256 ? Cout - n0 : 256;
            const uint32_t idesc = idesc_base | ((uint32_t)(n >> 3) << 17);
#pragma unroll
            for (int j = 0; j < kG; ++j) {
              if (j < cnt) {
                const uint64_t a_desc = SW::desc(a_s + (uint32_t)(j * kUnitA));
                const uint64_t b_desc = SW::desc(b_s + (uint32_t)j * b_unit + (uint32_t)n0 * SW::P);
#pragma unroll
                for (int kk = 0; kk < CIN_BLK / 16; ++kk)
                  umma_bf16(d_tmem + (uint32_t)n0, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2),
                            idesc, (st > 0 || j > 0 || kk > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&S.empty[slot]);   // frees the stage once these MMAs have read it
          if (++slot == stages) { slot = 0; fph ^= 1u; }
        }
        umma_commit(&S.acc_full[ab]);
      }
      tc_fence_before();
    }
  } else {
    // ======================= epilogue: TMEM -> registers -> global =======================
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int ab = t % acc_bufs;
      mbar_wait(&S.acc_full[ab], (uint32_t)(t / acc_bufs) & 1u);
      tc_fence_after();
      const int o = tile * kRows + tid;
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ab * Cout);
      for (int c0 = 0; c0 < Cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)c0, v);   // warp-collective
        tmem_ld_wait();
        if (o < n_out) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * S.scale[c0 + j] + S.shift[c0 + j];
          if (residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(residual + (size_t)o * Cout + c0);
            uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 a = __bfloat1622float2(h0[j]), b = __bfloat1622float2(h1[j]);
              f[2 * j] += a.x; f[2 * j + 1] += a.y;
              f[8 + 2 * j] += b.x; f[8 + 2 * j + 1] += b.y;
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 w0, w1;
          __nv_bfloat162* p0 = reinterpret_cast<__nv_bfloat162*>(&w0);
          __nv_bfloat162* p1 = reinterpret_cast<__nv_bfloat162*>(&w1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            p0[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            p1[j] = __floats2bfloat162_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
          }
          uint4* op = reinterpret_cast<uint4*>(out + (size_t)o * Cout + c0);
          op[0] = w0;
          op[1] = w1;
        }
      }
      tc_fence_before();
      mbar_arrive(&S.acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols)
                 : "memory");
  }
}

// (K,Cin,Cout) row-major bf16 -> [K][Cin/CIN_BLK] images of (Cout x CIN_BLK), K-major, swizzled;
// for Cout < 128 a second set of 128-row images follows in which the tile is replicated every
// 32 (Cout <= 32) / 64 rows (the A operand of the rows-on-N kernel, spconv_tn.cu; rows between a
// tile and the next replica stay zero).
template <int CIN_BLK>
__global__ void __launch_bounds__(256)
k_pack_w(const __nv_bfloat16* __restrict__ w, int K, int Cin, int Cout, __nv_bfloat16* __restrict__ out) {
  using SW = Swz<CIN_BLK>;
  const int nkb = Cin / CIN_BLK;
  const size_t total = (size_t)K * Cin * Cout;
  const int rep_span = Cout <= 32 ? 32 : 64;
  uint8_t* rep = reinterpret_cast<uint8_t*>(out) + total * 2;    // second image set (Cout < 128 only)
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(e % Cout);
    const int c = (int)((e / Cout) % Cin);
    const int k = (int)(e / ((size_t)Cout * Cin));
    const int kb = c / CIN_BLK, cc = c % CIN_BLK;
    const size_t img = ((size_t)k * nkb + kb) * ((size_t)Cout * SW::P);
    const uint32_t off = SW::offset(n, cc / 8) + (uint32_t)(cc % 8) * 2;
    const __nv_bfloat16 val = w[e];
    *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(out) + img + off) = val;
    if (Cout < 128) {
      const size_t img2 = ((size_t)k * nkb + kb) * ((size_t)128 * SW::P);
      for (int row = n; row < 128; row += rep_span)
        *reinterpret_cast<__nv_bfloat16*>(rep + img2 + SW::offset(row, cc / 8) + (uint32_t)(cc % 8) * 2) = val;
    }
  }
}

static inline int cin_blk_for(int Cin) { return Cin % 64 == 0 ? 64 : (Cin == 32 ? 32 : (Cin == 16 ? 16 : 0)); }

}  // namespace tc

bool spconv_tc_supported(int Cin, int Cout, int dtype) {
  if (dtype != U3D_BF16) return false;
  if (tc::cin_blk_for(Cin) == 0 || Cin > 512) return false;
  if (Cout < 16 || Cout > 512 || (Cout & (Cout - 1)) != 0) return false;  // power of two: TMEM columns
  return true;
}

int spconv_fwd_tc(const void* in, const int32_t* nbr, int nbr_stride,
                  const uint32_t* tile_mask, const int32_t* slot_row, const int32_t* n_out, int out_cap,
                  int K, const void* wpk,
                  const float* scale, const float* shift, const void* residual, int relu, void* out,
                  int Cin, int Cout, cudaStream_t st) {
  using namespace tc;
  // Cout <= 128 with a rulebook: the rows-on-N kernel (spconv_tn.cu) - full-size MMAs, half the
  // weight traffic. U3D_TC_KERNEL=1 forces this rows-on-M kernel (A/B timing, tests).
  {
    const char* e = getenv("U3D_TC_KERNEL");
    if (!(e && atoi(e) == 1) && spconv_tn_supported(Cin, Cout, nbr))
      return spconv_fwd_tn(in, nbr, nbr_stride, tile_mask, slot_row, n_out, out_cap, K, wpk, scale, shift,
                           residual, relu, out, Cin, Cout, st);
  }
  U3D_CHECK_ARG(slot_row == nullptr, "spconv tc: sorted tiles (slot_row) need the rows-on-N kernel "
                "(Cout <= 128, U3D_TC_KERNEL != 1); pass the natural-order rulebook (Cin=%d Cout=%d)", Cin, Cout);
  U3D_CHECK_ARG(K >= 1 && K <= kMaxK, "spconv tc: K=%d unsupported", K);
  U3D_CHECK_ARG((((uintptr_t)in | (uintptr_t)out | (uintptr_t)wpk | (uintptr_t)residual) & 15) == 0,
                "spconv tc: buffers must be 16-byte aligned");
  int tiles = cdiv(out_cap, kRows);
  if (tiles < 1) return U3D_OK;
  if (nbr) {
    U3D_CHECK_ARG((((uintptr_t)nbr) & 15) == 0 && nbr_stride % 4 == 0 && nbr_stride >= tiles * kRows,
                  "spconv tc: the rulebook must be 16-byte aligned with a row stride that is a multiple of 4 "
                  "and >= 128*ceil(out_cap/128) (stride=%d, out_cap=%d)", nbr_stride, out_cap);
  }
  const int blk = cin_blk_for(Cin);
  const uint32_t P = 2 * blk;
  const uint32_t b_bytes = (uint32_t)Cout * P;
  const uint32_t kg = 64 / blk;   // units per stage (see kG in the kernel)
  const uint32_t stage_bytes = kg * kRows * P + kg * ((b_bytes + 1023u) & ~1023u);
  // persistent CTAs: one per SM with as deep an operand ring as ~200 KB allows, or (U3D_TC_CTAS=2)
  // two per SM with half the ring, half the producer warps and half of TMEM each
  int ctas = 1;
  if (const char* e = getenv("U3D_TC_CTAS")) ctas = atoi(e) == 2 ? 2 : 1;
  if (Cout > 256) ctas = 1;
  const size_t header = (sizeof(Smem) + 1023) & ~(size_t)1023;
  const size_t budget = ctas == 1 ? 200u * 1024u : 110u * 1024u;
  int stages = (int)((budget - header) / stage_bytes);
  if (const char* e = getenv("U3D_TC_STAGES")) stages = atoi(e);
  const int max_stages = 2 * (kNumProd / ctas);   // at most two ring slots per producer warp
  if (stages > max_stages) stages = max_stages;
  if (stages < 2) stages = 2;
  const size_t smem = header + (size_t)stages * stage_bytes;
  U3D_CHECK_ARG(smem <= 227 * 1024, "spconv tc: tile does not fit shared memory (Cin=%d Cout=%d)", Cin, Cout);
  const int acc_bufs = 2 * Cout * ctas <= 512 ? 2 : 1;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < acc_bufs * Cout) tmem_cols <<= 1;
  const int grid = tiles < kNumSMs * ctas ? tiles : kNumSMs * ctas;
  int dbg = 0;   // U3D_TC_DEBUG: timing experiments only (1 = skip the gathers, 2 = skip the MMAs)
  if (const char* e = getenv("U3D_TC_DEBUG")) dbg = atoi(e);

#define U3D_TC_LAUNCH(BLK) \
  do {                     \
    if (ctas == 1) U3D_TC_LAUNCH2(BLK, 1); else U3D_TC_LAUNCH2(BLK, 2); \
  } while (0)
#define U3D_TC_LAUNCH2(BLK, CT)                                                                     \
  do {                                                                                              \
    static int cur_smem = 0;                                                                        \
    U3D_CUDA(ensure_dynamic_smem(k_spconv_tc<BLK, CT>, smem, &cur_smem));                           \
    k_spconv_tc<BLK, CT><<<grid, (kProdWarp0 + kNumProd / CT) * 32, smem, st>>>(                    \
        (const __nv_bfloat16*)in, nbr, nbr_stride, tile_mask, n_out, K, (const __nv_bfloat16*)wpk,  \
        scale, shift,                                                                               \
        (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)out, Cin, Cout, stages, acc_bufs,     \
        tmem_cols, dbg);                                                                            \
  } while (0)
  if (blk == 64) U3D_TC_LAUNCH(64);
  else if (blk == 32) U3D_TC_LAUNCH(32);
  else U3D_TC_LAUNCH(16);
#undef U3D_TC_LAUNCH
#undef U3D_TC_LAUNCH2
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // namespace u3d

using namespace u3d;

extern "C" size_t u3d_spconv_packed_bytes(int K, int Cin, int Cout) {
  if (!spconv_tc_supported(Cin, Cout, U3D_BF16) || K < 1) return 0;
  // rows-on-M images (Cout rows each) + for Cout < 128 the replicated 128-row images of the rows-on-N kernel
  return (size_t)K * Cin * ((size_t)Cout + (Cout < 128 ? 128 : 0)) * sizeof(__nv_bfloat16);
}

extern "C" int u3d_spconv_pack_weights(const void* w, int K, int Cin, int Cout, void* packed,
                                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(w && packed, "u3d_spconv_pack_weights: null buffer");
  U3D_CHECK_ARG(spconv_tc_supported(Cin, Cout, U3D_BF16) && K >= 1 && K <= tc::kMaxK,
                "u3d_spconv_pack_weights: unsupported shape K=%d Cin=%d Cout=%d", K, Cin, Cout);
  const size_t total = (size_t)K * Cin * Cout;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  const int blk = tc::cin_blk_for(Cin);
  if (Cout < 128)   // rows of the replicated images that no weight lands on (Cout = 16) must read as zero
    U3D_CUDA(cudaMemsetAsync((uint8_t*)packed + total * 2, 0, (size_t)K * Cin * 128 * 2, st));
  if (blk == 64) tc::k_pack_w<64><<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, K, Cin, Cout, (__nv_bfloat16*)packed);
  else if (blk == 32) tc::k_pack_w<32><<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, K, Cin, Cout, (__nv_bfloat16*)packed);
  else tc::k_pack_w<16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, K, Cin, Cout, (__nv_bfloat16*)packed);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_spconv_fwd_packed(const void* in, const int32_t* nbr, int nbr_stride,
                                     const uint32_t* tile_mask, const int32_t* slot_row,
                                     const int32_t* n_out, int out_cap, int K, const void* w_packed,
                                     const float* scale, const float* shift, const void* residual,
                                     int relu, void* out, int Cin, int Cout, int flags, void* stream) {
  U3D_CHECK_ARG(in && n_out && w_packed && out, "u3d_spconv_fwd_packed: null buffer");
  U3D_CHECK_ARG(nbr != nullptr || K == 1, "u3d_spconv_fwd_packed: nbr==NULL requires K==1");
  U3D_CHECK_ARG(spconv_tc_supported(Cin, Cout, U3D_BF16),
                "u3d_spconv_fwd_packed: needs Cin in {16,32,64k<=512}, Cout a power of two in [16,512] "
                "(Cin=%d Cout=%d)", Cin, Cout);
  // flags bit 1 (U3D_CONV_REVERSE_TILES): rows-on-N kernel walks its tiles from the last row back
  {
    const char* e = getenv("U3D_TC_KERNEL");
    if ((flags & 2) && !(e && atoi(e) == 1) && spconv_tn_supported(Cin, Cout, nbr))
      return spconv_fwd_tn_ex(in, nbr, nbr_stride, tile_mask, slot_row, n_out, out_cap, K, w_packed, scale, shift,
                              residual, relu, out, Cin, Cout, 2, Cin, Cout, 0, Cout, (cudaStream_t)stream);
  }
  return spconv_fwd_tc(in, nbr, nbr_stride, tile_mask, slot_row, n_out, out_cap, K, w_packed, scale, shift,
                       residual, relu, out, Cin, Cout, (cudaStream_t)stream);
}

// 3xBF16 evaluation of an fp32 sparse conv on the rows-on-N tcgen05 kernel (spconv_tn.cu, kX3): see there.
// in (rows, 2*Cin) bf16 [hi | lo]; w_packed: the first (and, for Cout < 128, second) image sets of
// u3d_spconv_pack_weights for the three K-block groups [w_hi ; w_lo ; w_hi] concatenated per offset;
// out / residual (rows, 2*cout_total) bf16 [hi | lo]; this launch computes channels [cout_off, cout_off + Cout),
// Cout <= 128; scale / shift point at channel cout_off.
extern "C" int u3d_spconv_fwd_packed_x3(const void* in, const int32_t* nbr, int nbr_stride, const uint32_t* tile_mask,
                                        const int32_t* slot_row, const int32_t* n_out, int out_cap, int K,
                                        const void* w_packed, const float* scale, const float* shift,
                                        const void* residual, int relu, void* out, int Cin, int Cout, int cout_off,
                                        int cout_total, int flags, void* stream) {
  U3D_CHECK_ARG(in && nbr && n_out && w_packed && out, "u3d_spconv_fwd_packed_x3: null buffer (a rulebook is required; "
                "pointwise convs pass an identity table)");
  U3D_CHECK_ARG(spconv_tn_supported(Cin, Cout, nbr) && cout_off >= 0 && cout_off + Cout <= cout_total && (cout_off & 7) == 0 &&
                    (cout_total & 7) == 0,
                "u3d_spconv_fwd_packed_x3: needs Cin in {16,32,64k<=512}, even Cout <= 128 (Cin=%d Cout=%d off=%d total=%d)",
                Cin, Cout, cout_off, cout_total);
  return spconv_fwd_tn_ex(in, nbr, nbr_stride, tile_mask, slot_row, n_out, out_cap, K, w_packed, scale, shift, residual,
                          relu, out, Cin, Cout, 1 | (flags & 2), 2 * Cin, 2 * cout_total, cout_off, cout_total,
                          (cudaStream_t)stream);
}


// K3 (tensor-core flavour, rows-on-N variant) — the sparse-conv gather-GEMM with the operand roles
// swapped: D^T[Cout, rows] = W_k^T[Cout, Cin] x X_k^T[Cin, rows].
//
// Why: a tcgen05.mma with both operands in shared memory costs about as long as reading its A
// operand (128 rows x 32 B at ~32 B/clk = ~128 cycles) no matter how small N is, so the
// rows-on-M kernel of spconv_tc.cu (M = 128 rows, N = Cout <= 128) pays a full-size instruction
// for a quarter (Cout = 64) or less of its work. Here the WEIGHT tile is the A operand (M = 128
// lanes, of which the first Cout hold channels) and 256 gathered rows are the B operand
// (N = 256): every instruction carries 128 x 256 x 16 MACs, and a weight tile is fetched once
// per 256 rows instead of once per 128.
//
// Reference semantics as spconv_tc.cu: SURVEY.md A.3/A.4 (spconv indice_conv + BatchNorm1d(eval)
// + ReLU, SparseBasicBlock identity add), projects/mmdet3d_plugin/models/pts_encoder/
// sparse_encoder_hd.py:106-132.
//
// Persistent, warp-specialised, one CTA per SM, 256-row output tiles:
//   warps 0-7    epilogue: TMEM lane = output channel, TMEM column = row of the tile. A warp may
//                only read the lane quarter warp%4 (hardware rule). Cout = 32 / 64 (round 2): the
//                weight-stationary instruction form tcgen05.mma.ws with M = Cout, whose accumulator
//                spreads the 256 columns over the lane groups a plain M < 128 instruction leaves idle
//                (M = 64: lanes 64-127 hold columns 128-255; M = 32: quarter q holds columns 64q..64q+63):
//                every lane of every warp holds live data, the datapath is full (80 / 76 cycles per
//                dispatch instead of 128, scripts/ubench/mma_cost.cu) and the weight image is the
//                un-replicated Cout-row one. Other widths (and U3D_TN_WS=0): plain tcgen05.mma with the
//                weight image replicated every 32 / 64 A rows so that every quarter holds all channels
//                (or M = 64 on the un-replicated image for Cout = 64).
//                tcgen05.ld 32x32b.x32, scale/shift (per-lane constants), adjacent channels paired
//                by one shuffle so that a lane stores bf16x2, residual (prefetched one chunk ahead),
//                ReLU, 64 contiguous bytes per row per warp.
//   warp 8       MMA issue (one thread), accumulators double-buffered in TMEM (2 x 256 columns). The
//                instruction form is a TEMPLATE parameter: a run-time branch around the two asm blocks
//                slowed every layer shape by 10 - 30 %.
//   warp 9       rulebook loader (cp.async.bulk of 1 KB rulebook rows): whole-tile slices of the active
//                offsets, double-buffered (55 KB), or - Cout = 128 layers - the rows of each gather stage
//                into a small ring (kSliceBufs = 0), which frees shared memory for one more stage.
//   warps 10+    producers: two warps per stage (128 rows each), 3 - 4 stages of 36 - 48 KB (weight
//                images + 32 KB gathered rows). A lane owns a CONTIGUOUS run of tile rows, so its
//                rulebook entries arrive with a few 16-byte shared loads up front, followed by
//                back-to-back 16-byte cp.async gathers (zero fill for missing neighbours) into the
//                K-major swizzled B tile, published with cp.async.mbarrier.arrive.noinc; the first
//                warp of the pair also bulk-copies the stage's weight images, one cp.async.bulk per
//                (offset, Cin block).
// Sorted tiles (tilesort.cu; every SubM level since round 2): with a `slot_row` list the rulebook is in
// slot order, the MMA thread bulk-copies the tile's 256 slot -> row entries next to the accumulator
// buffer (they complete on acc_full together with the tcgen05.commit) and the epilogue reads the
// residual of / writes row slot_row[s]; inputs keep their natural row numbers.
// Measured (profiles/): first version (one rulebook load per pass, one weight copy) 0.47 ms on the
// 64->64 layers at batch 32 vs 0.64 ms rows-on-M; batched rulebook loads 0.33 ms; M = 64 on un-replicated
// images + a 4th stage 0.316; tcgen05.mma.ws 0.280; sorted tiles 0.222 ms = 570 TFLOP/s of real pairs, at
// which point the launch moves 11.6 - 11.9 TB/s over the L2 -> SM crossbar (its cap): DESIGN.md section 4.
#include <stdlib.h>
#include "tc_common.cuh"

namespace u3d {

namespace tn {

using namespace tc;

constexpr int kTile = 256;             // rows per tile = UMMA N
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kMmaWarp = 8;
constexpr int kSliceWarp = 9;
constexpr int kProdWarp0 = 10;
// producer warps: 6 = 3 pairs = 3 ring slots (a stage is always 48 KB; default), or 8 with the
// single-buffered rulebook slices of the 4-stage variant (U3D_TN_SLICE_BUFS=1: measured slower)
constexpr int threads_of(int n_prod) { return (kProdWarp0 + n_prod) * 32; }   // 512 / 576
constexpr int kMaxK = 27;

// kSliceBufs: rulebook-slice buffers, 2 (whole-tile slices, double-buffered: 55 KB) or 1 (single-buffered, measured
// slower), or 0 = STAGE ROWS: the 1 KB rulebook rows travel per gather stage - the rows of stage g (one per unit) sit
// in slot g % (2 x stages) of a small ring with a full / empty barrier pair per slot, filled by the loader thread a
// few stages ahead and released by the two producer warps that own the stage (10 - 20 KB instead of 55 KB, which
// frees shared memory for one more gather stage);
// kMaxStages: one ring slot per producer warp pair
template <int kSliceBufs, int kMaxStages, int kG>
struct Smem {
  uint64_t full[kMaxStages];    // 64 cp.async arrives (the stage's two warps) + 1 arrive.expect_tx
  uint64_t empty[kMaxStages];   // tcgen05.commit: the MMAs that read the stage have retired
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t slice_full[kSliceBufs];
  uint64_t slice_empty[kSliceBufs];
  uint32_t tmem_base;
  alignas(128) int srow[2][kTile];   // output row of every slot of the tile in accumulator buffer ab (sorted tiles)
  alignas(128) int nbr[kSliceBufs][kMaxK][kTile];
};
template <int kMaxStages, int kG>
struct Smem<0, kMaxStages, kG> {
  static constexpr int kSlots = 2 * kMaxStages;
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t slice_full[kSlots];     // rows of stage g: slot g % kSlots, generation g / kSlots
  uint64_t slice_empty[kSlots];    // 2 arrivals: the producer warps of the stage have read their entries
  uint32_t tmem_base;
  alignas(128) int srow[2][kTile];
  alignas(128) int nbr[kSlots][kG][kTile];
};

__device__ __forceinline__ uint32_t mask_of_tile(const uint32_t* __restrict__ tile_mask, int tile,
                                                 int n_tiles128, uint32_t all_mask) {
  if (!tile_mask) return all_mask;
  uint32_t m = __ldg(&tile_mask[2 * tile]);
  if (2 * tile + 1 < n_tiles128) m |= __ldg(&tile_mask[2 * tile + 1]);
  return m & all_mask;
}

// kX3: "3xBF16" evaluation of an fp32 layer. Activations travel as (rows, 2*C) bf16 matrices [hi | lo] with
// hi = bf16(v), lo = bf16(v - hi) (16 significant bits); the weights are packed as three Cin-block groups
// [w_hi ; w_lo ; w_hi] and the K loop multiplies x_hi*w_hi + x_hi*w_lo + x_lo*w_hi into the same fp32
// accumulator (the dropped lo*lo term is 2^-16 relative), so fp32 layers (BASELINE configs 3 and 5) run on
// the tensor cores within the 1e-3 parity bound instead of on FFMA. `Cin` is the REAL channel count; `in_ld` /
// `out_ld` are the row strides of the input / output (and residual) matrices, `cout_off` the first output
// channel of this launch (Cout > 128 runs as several 128-channel launches), `cout_total` the real width.
template <int CIN_BLK, int kSliceBufs, int kNumProd, bool kX3, bool kWS>
__global__ void __launch_bounds__(threads_of(kNumProd), 1)
k_spconv_tn(const __nv_bfloat16* __restrict__ in, const int32_t* __restrict__ nbr, int nbr_stride,
            const uint32_t* __restrict__ tile_mask, const int32_t* __restrict__ slot_row,
            const int32_t* __restrict__ n_out_p, int K,
            const __nv_bfloat16* __restrict__ wpk, const float* __restrict__ scale,
            const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual, int relu,
            __nv_bfloat16* __restrict__ out, int Cin, int Cout, int stages, int in_ld, int out_ld, int cout_off,
            int cout_total, int m64) {
  using SW = Swz<CIN_BLK>;
  constexpr int kChunks = CIN_BLK / 8;            // 16-byte chunks per gathered row
  constexpr int kRowsPerPass = 32 / kChunks;      // rows one warp-wide cp.async covers
  constexpr int kPasses = 128 / kRowsPerPass;     // copies per lane per unit (a warp fills 128 rows)
  constexpr int kG = 64 / CIN_BLK;                // (offset, Cin-block) units per stage: K = 64
  constexpr int kUnitX = kTile * SW::P;           // one 256-row feature sub-tile
  constexpr int kXBytes = kG * kUnitX;            // 32 KB for every CIN_BLK
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool kRing = kSliceBufs == 0;
  constexpr int kSlots = kNumProd;                // stage-row slots (kRing): 2 x the ring's stages
  using SmemT = Smem<kSliceBufs, kNumProd / 2, kG>;
  SmemT& S = *reinterpret_cast<SmemT*>(smem_raw);
  constexpr uint32_t kHeader = (uint32_t)((sizeof(SmemT) + 1023) & ~(size_t)1023);

  const int n_out = *n_out_p;
  const int n_tiles = (n_out + kTile - 1) / kTile;
  const int n_tiles128 = (n_out + 127) / 128;
  if ((int)blockIdx.x >= n_tiles) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb_r = Cin / CIN_BLK;                 // Cin blocks of one operand part
  const int nkb = (kX3 ? 3 : 1) * nkb_r;           // K-loop blocks per offset
  // The A tile always spans the 128 TMEM lanes. For Cout < 128 the packed weights carry a second
  // set of images in which the (Cout x CIN_BLK) tile is REPLICATED every rep_span rows (written by
  // u3d_spconv_pack_weights), so every TMEM lane quarter holds a copy of the channels and all eight
  // epilogue warps drain a tile (with one copy only the warps of the first quarter(s) can, and a
  // 16/32-channel layer becomes epilogue-bound). One bulk copy per (offset, Cin block) either way.
  const int rep_span = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128);           // rows between replicas
  const int n_rep = 128 / rep_span;                                         // 4 / 2 / 1
  // m64 (Cout == 64): an M = 64 instruction on the UN-replicated 64-row weight image. A dispatch costs the same
  // cycles as M = 128, but the image is half the bytes (8 instead of 16 KB per unit: 40 % less L2 -> SM traffic
  // for this kernel) and a stage shrinks to 40 KB, which buys a 4th ring slot. The M = 64 accumulator puts 16
  // rows in the first 16 lanes of every TMEM lane quarter, so all eight epilogue warps still drain a tile.
  const int flags_ = m64;
  m64 &= 1;
  // ws (bit 2, Cout == 32 / 64): the weight-stationary instruction form `tcgen05.mma.ws` with M = Cout. Its accumulator
  // spreads the N = 256 columns over the idle lane groups (M = 64: lanes 64-127 hold columns 128-255; M = 32: lane
  // quarter q holds columns 64q .. 64q+63), so the datapath is full: a dispatch costs N/4 = 64 cycles instead of 128
  // (scripts/ubench/mma_cost.cu, profiles/r02_ubench_mma.txt), the un-replicated image is Cout rows, and every lane
  // of all eight epilogue warps holds live data.
  constexpr bool ws = kWS;   // a template parameter: a run-time branch around the MMA issue cost 30 % (measured)
  const uint32_t a_rows = ws ? (uint32_t)Cout : (m64 ? 64u : 128u);         // rows of one weight image
  const uint32_t w_unit = a_rows * SW::P;                                   // 4 / 8 / 16 KB (m64: 8 KB)
  const uint8_t* wimg = reinterpret_cast<const uint8_t*>(wpk) +
                        ((Cout < 128 && !m64 && !ws) ? (size_t)K * nkb * CIN_BLK * Cout * 2 : (size_t)0);   // 128-row images
  const uint32_t w_region = kG * w_unit;                                    // 16 KB for every CIN_BLK (m64: 8 KB)
  const uint32_t stage_bytes = w_region + kXBytes;
  const uint32_t tiles_s = smem_u32(smem_raw) + kHeader;                   // 1024-aligned
  const uint32_t all_mask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&S.full[s], 64 + 1);
      mbar_init(&S.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S.acc_full[b], slot_row ? 2 : 1);   // tcgen05.commit (+ the tile's slot->row list)
      mbar_init(&S.acc_empty[b], kEpiThreads);
    }
    for (int b = 0; b < (kRing ? kNumProd : kSliceBufs); ++b) {
      mbar_init(&S.slice_full[b], 1);
      mbar_init(&S.slice_empty[b], kRing ? 2 : kNumProd);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&S.tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp >= kProdWarp0) {
    // ======================= producers: a warp pair per stage =======================
    // Lane (rsub, chunk) copies the 16-byte chunk `chunk` of the tile rows rsub*kPasses + i,
    // i = 0..kPasses-1, of its half: the kPasses rulebook entries it needs are CONTIGUOUS, so they
    // come in with kPasses/4 independent 16-byte shared loads issued before the first cp.async
    // (a load per pass, as in the first version of this loop, serialises LDS -> address -> LDGSTS
    // behind the asm memory clobbers: ~45 cycles per pass instead of ~8).
    const int w = warp - kProdWarp0;
    const int pair = w >> 1, half = w & 1;
    const int chunk = lane % kChunks;
    const int rsub = lane / kChunks;
    const int rbase = half * 128 + rsub * kPasses;      // first tile row of this lane
    const uint32_t chunk16 = (uint32_t)chunk * 16u;
    const uint32_t lane_off = (uint32_t)(rbase * SW::P);
    const uint64_t row_bytes = (uint64_t)in_ld * 2;
    // pair p OWNS ring slot p: it fills the global stages g = p, p + stages, ... so it meets the
    // generations of its slot in order and the 1-bit mbarrier parity is never ambiguous (a pair
    // without a slot - ring shortened by U3D_TN_STAGES - only passes through the slices)
    const int slot = pair;
    int g = pair < stages ? pair : 0x7fffffff;
    uint32_t eph = 1u;   // parity to wait for on empty[slot]; flips on every visit
    int g0 = 0;          // global index of the first stage of the current tile
    int t = 0;
    const uint32_t st_s = tiles_s + (uint32_t)slot * stage_bytes;
    for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++t) {
      const int tile = (flags_ & 2) ? n_tiles - 1 - tl : tl;   // bit 1: walk the tiles from the last row back (L2 reuse)
      const int buf = kRing ? 0 : t % (kRing ? 1 : kSliceBufs);
      const int m0 = tile * kTile;
      const uint32_t mask = mask_of_tile(tile_mask, tile, n_tiles128, all_mask);
      const int n_units = __popc(mask) * nkb;
      const int n_st = (n_units + kG - 1) / kG;
      const int rows_live = n_out - m0 - rbase;   // rulebook entries of rows >= n_out are uninitialised
      // rank of this lane's bit among the set bits of the mask: "position of the r-th set bit" is
      // then one ballot away
      const bool my_bit = (mask >> lane) & 1u;
      const int my_rank = __popc(mask & ((1u << lane) - 1u));
      if (!kRing) mbar_wait(&S.slice_full[buf], (uint32_t)(t / (kRing ? 1 : kSliceBufs)) & 1u);
      for (; g < g0 + n_st; g += stages) {
        const int u0 = (g - g0) * kG;
        const int cnt = n_units - u0 < kG ? n_units - u0 : kG;
        const int rs = kRing ? g % kSlots : 0;     // stage rows: slot of this stage's rulebook rows
        mbar_wait(&S.empty[slot], eph);
        eph ^= 1u;
        if (half == 0 && lane == 0) mbar_expect_tx(&S.full[slot], (uint32_t)cnt * w_unit);
        if (kRing) mbar_wait(&S.slice_full[rs], (uint32_t)(g / kSlots) & 1u);
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          if (j < cnt) {
            const int u = u0 + j;
            const int ki = u / nkb, kb = u - ki * nkb;
            const int k = __ffs(__ballot_sync(0xffffffffu, my_bit && my_rank == ki)) - 1;
            // source columns of this K block: plain = block kb; 3xBF16 = [x_hi | x_hi | x_lo] blocks
            const int part = kX3 ? kb / nkb_r : 0;
            const int src_col = (part == 2 ? Cin : 0) + (kb - part * nkb_r) * CIN_BLK;
            if (half == 0 && lane == 0)
              bulk_g2s(st_s + (uint32_t)j * w_unit, wimg + ((size_t)k * nkb + kb) * w_unit, w_unit,
                       &S.full[slot]);
            const uint8_t* src_base =
                reinterpret_cast<const uint8_t*>(in) + (size_t)(src_col + chunk * 8) * 2;
            int idx[kPasses];
            const int4* nb4 = reinterpret_cast<const int4*>(kRing ? &S.nbr[rs][j][rbase] : &S.nbr[buf][k][rbase]);
#pragma unroll
            for (int v = 0; v < kPasses / 4; ++v) {
              const int4 q4 = nb4[v];
              idx[4 * v] = q4.x; idx[4 * v + 1] = q4.y; idx[4 * v + 2] = q4.z; idx[4 * v + 3] = q4.w;
            }
            const uint32_t x_u = st_s + w_region + (uint32_t)(j * kUnitX) + lane_off;
#pragma unroll
            for (int i = 0; i < kPasses; ++i) {
              // swizzle: 16-byte chunk index XOR address bits [7, 7+log2(P/16)); the row base of a
              // lane is a multiple of 8 rows, so the XOR term depends on i only
              constexpr uint32_t kSwzMask = SW::kMask;
              const uint32_t sw = (((uint32_t)(i * SW::P)) >> 7) & kSwzMask;
              const bool ok = idx[i] >= 0 && i < rows_live;
              const uint8_t* src = src_base + (ok ? (uint64_t)(uint32_t)idx[i] * row_bytes : 0ull);
              cp_async16(x_u + (uint32_t)(i * SW::P) + (chunk16 ^ (sw << 4)), src, ok ? 16u : 0u);
            }
          }
        }
        cp_async_arrive(&S.full[slot]);
        if (kRing) {     // every lane's entries are in registers (the gathers above consumed them): free the rows
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.slice_empty[rs]);
        }
      }
      g0 += n_st;
      if (!kRing) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.slice_empty[buf]);
      }
    }
  } else if (warp == kSliceWarp) {
    // ======================= rulebook-slice loader (one thread) =======================
    if (lane == 0) {
      int t = 0, g = 0;
      for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++t) {
        const int tile = (flags_ & 2) ? n_tiles - 1 - tl : tl;   // bit 1: walk the tiles from the last row back (L2 reuse)
        uint32_t m = mask_of_tile(tile_mask, tile, n_tiles128, all_mask);
        int ents = nbr_stride - tile * kTile;     // the last tile of a row may hold 128 entries only
        if (ents > kTile) ents = kTile;
        const uint32_t bytes = (uint32_t)ents * 4u;
        if (kRing) {
          // stage by stage, in the producers' unit order: unit u of the tile reads the row of its (u / nkb)-th
          // active offset (a Cin = 128 layer has two units - K blocks - per offset: the row is fetched for both)
          const int n_units = __popc(m) * nkb;
          const int n_st = (n_units + kG - 1) / kG;
          for (int st = 0; st < n_st; ++st, ++g) {
            const int u0 = st * kG;
            const int cnt = n_units - u0 < kG ? n_units - u0 : kG;
            const int rs = g % kSlots;
            mbar_wait_relaxed(&S.slice_empty[rs], (((uint32_t)(g / kSlots)) & 1u) ^ 1u, 500u);
            mbar_expect_tx(&S.slice_full[rs], (uint32_t)cnt * bytes);
            for (int j = 0; j < cnt; ++j) {
              const int k = (int)__fns(m, 0, (u0 + j) / nkb + 1);
              bulk_g2s(smem_u32(&S.nbr[rs][j][0]), nbr + (size_t)k * nbr_stride + (size_t)tile * kTile, bytes,
                       &S.slice_full[rs]);
            }
          }
        } else {
          const int buf = t % (kRing ? 1 : kSliceBufs);
          mbar_wait_relaxed(&S.slice_empty[buf], (((uint32_t)(t / (kRing ? 1 : kSliceBufs))) & 1u) ^ 1u, 2000u);
          mbar_expect_tx(&S.slice_full[buf], (uint32_t)__popc(m) * bytes);
          while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            bulk_g2s(smem_u32(&S.nbr[buf][k][0]), nbr + (size_t)k * nbr_stride + (size_t)tile * kTile, bytes,
                     &S.slice_full[buf]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ======================= MMA issuer (one thread) =======================
    if (lane == 0) {
      // kind::f16: D = f32, A = B = bf16, both K-major, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTile >> 3) << 17) |
                             ((uint32_t)(a_rows >> 4) << 24);
      int slot = 0, t = 0;
      uint32_t fph = 0u;
      for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++t) {
      const int tile = (flags_ & 2) ? n_tiles - 1 - tl : tl;   // bit 1: walk the tiles from the last row back (L2 reuse)
        const int ab = t & 1;
        const uint32_t mask = mask_of_tile(tile_mask, tile, n_tiles128, all_mask);
        const int n_units = __popc(mask) * nkb;
        const int n_st = (n_units + kG - 1) / kG;
        mbar_wait(&S.acc_empty[ab], ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (slot_row) {   // sorted tiles: the epilogue of this tile needs slot -> output row (same lifetime as ab)
          mbar_expect_tx(&S.acc_full[ab], (uint32_t)(kTile * 4));
          bulk_g2s(smem_u32(&S.srow[ab][0]), slot_row + (size_t)tile * kTile, kTile * 4, &S.acc_full[ab]);
        }
        const uint32_t d_tmem = tmem + (uint32_t)(ab * kTile);
        for (int st = 0; st < n_st; ++st) {
          const int cnt = n_units - st * kG < kG ? n_units - st * kG : kG;
          mbar_wait(&S.full[slot], fph);
          tc_fence_after();
          const uint32_t st_s = tiles_s + (uint32_t)slot * stage_bytes;
#pragma unroll
          for (int j = 0; j < kG; ++j) {
            if (j < cnt) {
              const uint64_t w_desc = SW::desc(st_s + (uint32_t)j * w_unit);
              const uint64_t x_desc = SW::desc(st_s + w_region + (uint32_t)(j * kUnitX));
#pragma unroll
              for (int kk = 0; kk < CIN_BLK / 16; ++kk) {
                const uint32_t acc = (st > 0 || j > 0 || kk > 0) ? 1u : 0u;
                if (ws) umma_bf16_ws(d_tmem, w_desc + (uint64_t)(kk * 2), x_desc + (uint64_t)(kk * 2), idesc, acc);
                else umma_bf16(d_tmem, w_desc + (uint64_t)(kk * 2), x_desc + (uint64_t)(kk * 2), idesc, acc);
              }
            }
          }
          umma_commit(&S.empty[slot]);
          if (++slot == stages) { slot = 0; fph ^= 1u; }
        }
        umma_commit(&S.acc_full[ab]);
      }
      tc_fence_before();
    }
  } else {
    // ======================= epilogue: TMEM -> registers -> global =======================
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int h = warp >> 2;           // column (= row of the tile) half
    // m64: accumulator row r sits in lane 32 * (r / 16) + r % 16: quarter q holds channels 16q .. 16q + 15
    // ws: TMEM lane 32q + lane holds channel (32q + lane) % Cout of the tile rows (32q / Cout) * (256 Cout / 128) + column
    const int c = ws ? (q * 32 + lane) % Cout : (m64 ? q * 16 + lane : (q * 32 + lane) % rep_span);   // output channel of this lane
    const int rho = m64 ? 0 : (q * 32) / rep_span;                     // which replica this quarter holds
    const int ws_cols = 2 * Cout;               // ws: TMEM columns of one accumulator (128 / 64)
    const int ncol = ws ? ws_cols / 2 : (m64 ? 128 : 128 / n_rep);   // columns this warp drains: 128 / 64 / 32
    const int tcol_lo = ws ? h * ncol : h * 128 + rho * ncol;         // first TMEM column of this warp
    const int col_lo = ws ? ((q * 32) / Cout) * ws_cols + h * ncol : tcol_lo;   // first tile row of this warp
    const bool lane_live = ws ? true : (m64 ? lane < 16 : c < Cout);
    const bool odd = lane & 1;
    const int cb = c & ~1;             // channel pair this lane stores
    const float sc = (lane_live && scale) ? __ldg(&scale[c]) : 1.f;
    const float sh = (lane_live && shift) ? __ldg(&shift[c]) : 0.f;
    int t = 0;
    for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++t) {
      const int tile = (flags_ & 2) ? n_tiles - 1 - tl : tl;   // bit 1: walk the tiles from the last row back (L2 reuse)
      const int ab = t & 1;
      const int row0 = tile * kTile + col_lo + (odd ? 1 : 0);   // slot of this lane: + col + 2p
      const int* srow = slot_row ? &S.srow[ab][col_lo + (odd ? 1 : 0)] : nullptr;
      uint32_t res[16], res_lo[kX3 ? 16 : 1];
      auto load_res = [&](int col) {
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int slot = row0 + col + 2 * p;
          if (residual && lane_live && slot < n_out) {
            const int o = srow ? srow[col + 2 * p] : slot;
            const __nv_bfloat16* rp = residual + (size_t)o * out_ld + cout_off + cb;
            res[p] = __ldg(reinterpret_cast<const uint32_t*>(rp));
            if (kX3) res_lo[p] = __ldg(reinterpret_cast<const uint32_t*>(rp + cout_total));
          } else {
            res[p] = 0u;
            if (kX3) res_lo[p] = 0u;
          }
        }
      };
      if (!srow) load_res(0);          // natural order: residual addresses do not depend on the slot list
      mbar_wait_relaxed(&S.acc_full[ab], (uint32_t)(t >> 1) & 1u, 2000u);
      tc_fence_after();
      if (srow) load_res(0);
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * kTile + tcol_lo);
#pragma unroll 1
      for (int col = 0; col < ncol; col += 32) {
        uint32_t v[32];
        tmem_ld32(lane_base + (uint32_t)col, v);   // warp-collective
        tmem_ld_wait();
        uint32_t cur[16], cur_lo[kX3 ? 16 : 1];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          cur[p] = res[p];
          if (kX3) cur_lo[p] = res_lo[p];
        }
        if (col + 32 < ncol) load_res(col + 32);     // prefetch the next chunk's residuals
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const float fe = __uint_as_float(v[2 * p]) * sc + sh;        // row col+2p
          const float fo = __uint_as_float(v[2 * p + 1]) * sc + sh;    // row col+2p+1
          const float recv = __shfl_xor_sync(0xffffffffu, odd ? fe : fo, 1);
          float lo = odd ? recv : fe;     // channel cb
          float hi = odd ? fo : recv;     // channel cb+1
          const int slot = row0 + col + 2 * p;
          if (lane_live && slot < n_out) {
            const int o = srow ? srow[col + 2 * p] : slot;
            const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cur[p]));
            lo += r.x;
            hi += r.y;
            if (kX3) {
              const float2 r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cur_lo[p]));
              lo += r2.x;
              hi += r2.y;
            }
            if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            __nv_bfloat16* op = out + (size_t)o * out_ld + cout_off + cb;
            const __nv_bfloat162 top = __floats2bfloat162_rn(lo, hi);
            *reinterpret_cast<__nv_bfloat162*>(op) = top;
            if (kX3) {            // second half of the row: what bf16 dropped (v - bf16(v)), again as bf16
              const float2 t = __bfloat1622float2(top);
              *reinterpret_cast<__nv_bfloat162*>(op + cout_total) = __floats2bfloat162_rn(lo - t.x, hi - t.y);
            }
          }
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace tn

bool spconv_tn_supported(int Cin, int Cout, const int32_t* nbr) {
  if (nbr == nullptr) return false;                       // pointwise convs stay on the rows-on-M kernel
  if (Cout > 128 || (Cout & 1)) return false;             // one M = 128 accumulator
  return spconv_tc_supported(Cin, Cout, U3D_BF16);
}

int spconv_fwd_tn(const void* in, const int32_t* nbr, int nbr_stride, const uint32_t* tile_mask,
                  const int32_t* slot_row, const int32_t* n_out, int out_cap, int K, const void* wpk,
                  const float* scale,
                  const float* shift, const void* residual, int relu, void* out, int Cin, int Cout,
                  cudaStream_t st) {
  return spconv_fwd_tn_ex(in, nbr, nbr_stride, tile_mask, slot_row, n_out, out_cap, K, wpk, scale, shift, residual,
                          relu, out, Cin, Cout, /*x3=*/0, /*in_ld=*/Cin, /*out_ld=*/Cout, /*cout_off=*/0,
                          /*cout_total=*/Cout, st);
}

int spconv_fwd_tn_ex(const void* in, const int32_t* nbr, int nbr_stride, const uint32_t* tile_mask,
                     const int32_t* slot_row, const int32_t* n_out, int out_cap, int K, const void* wpk,
                     const float* scale, const float* shift, const void* residual, int relu, void* out, int Cin,
                     int Cout, int x3, int in_ld, int out_ld, int cout_off, int cout_total, cudaStream_t st) {
  using namespace tn;
  U3D_CHECK_ARG(K >= 1 && K <= kMaxK, "spconv tn: K=%d unsupported", K);
  U3D_CHECK_ARG((((uintptr_t)in | (uintptr_t)out | (uintptr_t)wpk | (uintptr_t)residual) & 15) == 0,
                "spconv tn: buffers must be 16-byte aligned");
  const int tiles = cdiv(out_cap, kTile);
  if (tiles < 1) return U3D_OK;
  U3D_CHECK_ARG(slot_row == nullptr || ((((uintptr_t)slot_row) & 15) == 0 && nbr_stride >= kTile * tiles),
                "spconv tn: slot_row must be 16-byte aligned and, like the sorted rulebook, padded to whole "
                "256-slot tiles (stride=%d, out_cap=%d)", nbr_stride, out_cap);
  U3D_CHECK_ARG((((uintptr_t)nbr) & 15) == 0 && nbr_stride % 4 == 0 && nbr_stride >= 128 * cdiv(out_cap, 128),
                "spconv tn: the rulebook must be 16-byte aligned with a row stride that is a multiple of 4 "
                "and >= 128*ceil(out_cap/128) (stride=%d, out_cap=%d)", nbr_stride, out_cap);
  const int blk = Cin % 64 == 0 ? 64 : Cin;
  const uint32_t P = 2 * blk;
  const uint32_t kg = 64 / blk;
  // Cout == 64 on 64-wide K blocks: M = 64 instruction, un-replicated 8 KB weight images, 4 stages of 40 KB filled by
  // 8 producer warps (U3D_TN_M64=0 keeps the replicated M = 128 form)
  // tile order: the previous layer left its LAST rows in L2 (feature matrices of the wide levels exceed it at batch
  // 32), so every other launch walks the tiles backwards and starts on rows that are still resident
  const bool reverse_tiles = (x3 & 2) != 0;
  x3 &= 1;
  bool m64 = Cout == 64 && blk == 64;
  if (const char* e = getenv("U3D_TN_M64")) m64 = m64 && atoi(e) != 0;
  // Cout == 32 / 64: weight-stationary instruction form with M = Cout on the un-replicated images (U3D_TN_WS=0 falls
  // back to the M = 64 / replicated M = 128 forms); 8 producer warps, 4 stages of 36 / 40 KB
  bool ws = Cout == 32 || Cout == 64;
  if (const char* e = getenv("U3D_TN_WS")) ws = ws && atoi(e) != 0;
  if (ws) m64 = false;
  const uint32_t a_rows = ws ? (uint32_t)Cout : (m64 ? 64u : 128u);
  const uint32_t stage_bytes = kg * a_rows * P + kg * kTile * P;   // weight images (16 KB replicated / <= 8 KB) + 32 KB rows
  // Default: double-buffered rulebook slices, 6 producer warps, 3 stages. U3D_TN_SLICE_BUFS=1 single-buffers the slices,
  // which frees 28 KB for a 4th stage filled by a 4th producer pair (576 threads; measured slower).
  bool deep = false;
  if (const char* e = getenv("U3D_TN_SLICE_BUFS")) deep = atoi(e) == 1;
  const bool wide = m64 || ws;                                 // 8 producer warps, double-buffered slices
  if (wide) deep = false;
  // Per-stage rulebook rows (kSliceBufs = 0: 2 x stages slots of 64 / blk KB instead of 2 x 27 KB slices) free shared
  // memory for one more gather stage. Measured per step (batch 32): 128->128 0.725 -> 0.672 ms (4 x 48 KB stages instead
  // of 3), 64->64 unchanged (5 x 40 KB instead of 4: that layer moves 10 TB/s L2 -> SM, the fabric is the bound, not
  // the ring depth), 32->32 0.635 -> 0.825 ms (two rows per stage: the single loader thread falls behind). Default:
  // on for the 64-wide K blocks of the M = 128 form (Cout = 128 layers); U3D_TN_RING=1 / 0 forces it everywhere / off.
  bool ring = blk == 64 && !wide;
  if (const char* e = getenv("U3D_TN_RING")) ring = atoi(e) != 0;
  if (deep) ring = false;
  const size_t ring_hdr5 = blk == 64 ? sizeof(Smem<0, 5, 1>) : (blk == 32 ? sizeof(Smem<0, 5, 2>) : sizeof(Smem<0, 5, 4>));
  const size_t ring_hdr4 = blk == 64 ? sizeof(Smem<0, 4, 1>) : (blk == 32 ? sizeof(Smem<0, 4, 2>) : sizeof(Smem<0, 4, 4>));
  int max_stages = (deep || wide) ? 4 : 3;                     // one ring slot per producer pair
  size_t header = ((deep ? sizeof(Smem<1, 4, 1>) : (wide ? sizeof(Smem<2, 4, 1>) : sizeof(Smem<2, 3, 1>))) + 1023) & ~(size_t)1023;
  if (ring) {   // 5 stages (10 producer warps) when they fit next to the 5-stage header, else 4 (8 warps)
    const size_t h5 = (ring_hdr5 + 1023) & ~(size_t)1023, h4 = (ring_hdr4 + 1023) & ~(size_t)1023;
    const bool five = blk != 16 && h5 + 5 * (size_t)stage_bytes <= 227u * 1024u;
    max_stages = five ? 5 : 4;
    header = five ? h5 : h4;
  }
  int stages = (int)((227u * 1024u - header) / stage_bytes);
  if (const char* e = getenv("U3D_TN_STAGES")) stages = atoi(e);
  if (stages > max_stages) stages = max_stages;
  U3D_CHECK_ARG(stages >= 2, "spconv tn: tile does not fit shared memory (Cin=%d Cout=%d)", Cin, Cout);
  const size_t smem = header + (size_t)stages * stage_bytes;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;

#define U3D_TN_LAUNCH3(BLK, SB, NP, X3, WS)                                                           \
  do {                                                                                              \
    static int cur_smem = 0;                                                                        \
    U3D_CUDA(ensure_dynamic_smem(k_spconv_tn<BLK, SB, NP, X3, WS>, smem, &cur_smem));               \
    k_spconv_tn<BLK, SB, NP, X3, WS><<<grid, threads_of(NP), smem, st>>>(                           \
        (const __nv_bfloat16*)in, nbr, nbr_stride, tile_mask, slot_row, n_out, K,                   \
        (const __nv_bfloat16*)wpk, scale, shift, (const __nv_bfloat16*)residual, relu,              \
        (__nv_bfloat16*)out, Cin, Cout, stages, in_ld, out_ld, cout_off, cout_total,                \
        (m64 ? 1 : 0) | (reverse_tiles ? 2 : 0));                                                   \
  } while (0)
#define U3D_TN_LAUNCH2(BLK, SB, NP, WS)                                                             \
  do {                                                                                              \
    if (x3) U3D_TN_LAUNCH3(BLK, SB, NP, true, WS); else U3D_TN_LAUNCH3(BLK, SB, NP, false, WS);     \
  } while (0)
#define U3D_TN_LAUNCH(BLK)                                                                          \
  do {                                                                                              \
    if (deep) U3D_TN_LAUNCH2(BLK, 1, 8, false); else U3D_TN_LAUNCH2(BLK, 2, 6, false);              \
  } while (0)
#define U3D_TN_RING_LAUNCH(BLK, NP)                                                                 \
  do {                                                                                              \
    if (ws) U3D_TN_LAUNCH2(BLK, 0, NP, true); else U3D_TN_LAUNCH2(BLK, 0, NP, false);               \
  } while (0)
  if (ring) {
    if (blk == 64) { if (max_stages == 5) U3D_TN_RING_LAUNCH(64, 10); else U3D_TN_RING_LAUNCH(64, 8); }
    else if (blk == 32) { if (max_stages == 5) U3D_TN_RING_LAUNCH(32, 10); else U3D_TN_RING_LAUNCH(32, 8); }
    else U3D_TN_RING_LAUNCH(16, 8);
  }
  else if (ws) {
    if (blk == 64) U3D_TN_LAUNCH2(64, 2, 8, true);
    else if (blk == 32) U3D_TN_LAUNCH2(32, 2, 8, true);
    else U3D_TN_LAUNCH2(16, 2, 8, true);
  }
  else if (m64) U3D_TN_LAUNCH2(64, 2, 8, false);
  else if (blk == 64) U3D_TN_LAUNCH(64);
  else if (blk == 32) U3D_TN_LAUNCH(32);
  else U3D_TN_LAUNCH(16);
#undef U3D_TN_RING_LAUNCH
#undef U3D_TN_LAUNCH
#undef U3D_TN_LAUNCH2
#undef U3D_TN_LAUNCH3
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // namespace u3d

// K4 — batched D-FPS (furthest point sampling) + gather + min-max normalisation.
//
// Reference: mmcv.ops.PointsSampler([nq]) / furthest_point_sample + gather_points +
// shift_scale_points, projects/mmdet3d_plugin/models/detectors/uni3detr.py:138,178-187
// (semantics SURVEY.md A.5/A.6). The reference runs one CTA per scene inside a python
// loop over the batch and re-reads all N points from HBM on each of the nq iterations.
//
// Here one thread-block CLUSTER owns one scene: every thread keeps PPT points and their
// running min-distances in registers for the whole kernel (points are read from HBM
// exactly once: N*12 B). Per iteration: thread-local arg-max, warp arg-max with two REDUX
// instructions on a packed (distance bits, ~index) key, one __syncthreads, CTA arg-max in
// warp 0, then each CTA pushes one 32-byte candidate into every peer's shared memory
// (st.shared::cluster) and arrives on the peer's mbarrier (release.cluster); waiting on the
// local mbarrier (acquire.cluster) replaces the much slower barrier.cluster round trip.
// Candidate slots and barriers are double-buffered by iteration parity.
// All scenes of the batch run concurrently (grid = B clusters).
//
// Defined arithmetic (mirrored by oracle/): d = ((dx*dx + dy*dy) + dz*dz) with no FMA
// contraction, running distance initialised to 1e10, first pick = index 0. Exact distance ties
// (the rule on the integer voxel lattice of FPS #2): `tie_block` = 0 -> the lowest index;
// `tie_block` = T (the reference kernel's block-size cap, 1024 in mmcv) -> the winner of mmcv's
// reduction: that kernel runs bs = min(T, 2^floor(log2 n)) threads, thread t scans k = t, t + bs, ...
// keeping its FIRST maximum (strict >), then a shared-memory tree (s = bs/2 .. 1: slot t takes slot
// t + s only if strictly greater) - so among tied points the one whose thread id has the smallest
// BIT-REVERSED value wins, and within a thread the lowest k. Both are a total order on indices, so
// they fold into the arg-max key: key(k) = bitrev_log2(bs)(k mod bs) << 12 | k / bs.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace u3d {

constexpr int kFpsMaxNq = 4096;

// One candidate travelling between CTAs: 32 bytes = two 16-byte DSMEM stores.
// (hi, lo) is the arg-max key: hi = bits of the (non-negative) distance, lo = ~index, so an
// unsigned 64-bit max picks the largest distance and, on exact ties, the LOWEST tie key (fps_key).
struct __align__(16) FpsMsg {
  uint32_t hi, lo;
  float x, y, z;
  uint32_t pad[3];
};

__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// tie order of point i (smaller wins): i itself, or mmcv's reduction order for a block of 2^lg threads (lg < 0: off)
__device__ __forceinline__ uint32_t fps_key(uint32_t i, int lg) {
  if (lg < 0) return i;
  if (lg == 0) return i;
  const uint32_t r = i & ((1u << lg) - 1u);
  return ((__brev(r) >> (32 - lg)) << 12) | (i >> lg);
}
__device__ __forceinline__ uint32_t fps_unkey(uint32_t key, int lg) {
  if (lg <= 0) return key;
  const uint32_t r = __brev(key >> 12) >> (32 - lg);
  return ((key & 0xfffu) << lg) | r;
}

// warp arg-max of a (hi, lo) key with two REDUX instructions; returns true on the winning lane
__device__ __forceinline__ bool warp_argmax(uint32_t hi, uint32_t lo, uint32_t& whi, uint32_t& wlo) {
  whi = __reduce_max_sync(0xffffffffu, hi);
  wlo = __reduce_max_sync(0xffffffffu, hi == whi ? lo : 0u);
  return hi == whi && lo == wlo;
}

template <int THREADS, int PPT, int CS, bool SMEM_XYZ>
__global__ void __launch_bounds__(THREADS, 1)
k_fps(const float* __restrict__ dist_src, int dist_stride, int dist_seg_stride,
      const float* __restrict__ gather_src, int gather_stride, const int32_t* __restrict__ seg,
      int nq, int reverse, int tie_block, int async_msg, int32_t* __restrict__ idx_out, float* __restrict__ out) {
  constexpr int NW = THREADS / 32;
  constexpr bool kSimpleTie = ((THREADS * CS) % 1024) == 0;
  extern __shared__ float s_dyn[];              // SMEM_XYZ: x[PPT*THREADS], y[..], z[..]
  __shared__ FpsMsg s_warp[NW];
  __shared__ FpsMsg s_slot[2][CS];              // written by every CTA of the cluster (DSMEM)
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int s_sel[kFpsMaxNq];
  __shared__ float s_mm[6];
  __shared__ float s_red[NW][6];

  const int scene = blockIdx.x / CS;
  const int crank = blockIdx.x % CS;  // == cluster block rank (1-D cluster)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = seg[scene];
  const int n = seg[scene + 1] - s0;
  const float* base = dist_src + (size_t)s0 * dist_seg_stride;
  // log2 of the reference kernel's block size for this scene (-1: lowest-index ties)
  int lg = -1;
  if (tie_block > 0 && n > 0) {
    lg = 31 - __clz(n);
    const int cap = 31 - __clz(tie_block);
    if (lg > cap) lg = cap;
  }

  float px[SMEM_XYZ ? 1 : PPT], py[SMEM_XYZ ? 1 : PPT], pz[SMEM_XYZ ? 1 : PPT], td[PPT];
  float* sx = s_dyn;
  float* sy = s_dyn + PPT * THREADS;
  float* sz = s_dyn + 2 * PPT * THREADS;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = (j * CS + crank) * THREADS + tid;
    const bool ok = i < n;
    const float x = ok ? __ldg(base + (size_t)i * dist_stride + 0) : 0.f;
    const float y = ok ? __ldg(base + (size_t)i * dist_stride + 1) : 0.f;
    const float z = ok ? __ldg(base + (size_t)i * dist_stride + 2) : 0.f;
    if (SMEM_XYZ) {
      sx[j * THREADS + tid] = x; sy[j * THREADS + tid] = y; sz[j * THREADS + tid] = z;
    } else {
      px[j] = x; py[j] = y; pz[j] = z;
    }
    // padding keeps distance 0 and an index >= n: it loses against every real point
    td[j] = ok ? 1e10f : 0.f;
  }

  float lx = 0.f, ly = 0.f, lz = 0.f;
  if (n > 0) {
    lx = __ldg(base + 0);
    ly = __ldg(base + 1);
    lz = __ldg(base + 2);
  }
  if (tid == 0) {
    s_sel[0] = 0;
    if (CS > 1) {
      // async_msg: one local arrive.expect_tx per phase, completed by the peers' st.async bytes; else CS remote arrives
      for (int b = 0; b < 2; ++b)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fps_smem_u32(&s_bar[b])), "r"(async_msg ? 1 : CS));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  if (CS > 1) cg::this_cluster().sync();  // barriers initialised + every CTA resident before DSMEM use
  else __syncthreads();

  for (int it = 1; it < nq; ++it) {
    // 1. update running distances, thread-local arg-max. The kernel is ISSUE bound (20 k points x 300 iterations x 32
    // scenes), so the loop carries the bare minimum: 8 FP ops + min + compare + two selects per point. A thread's
    // points ascend in j; when THREADS * CS is a multiple of the reference block size they all share (i mod bs), so the
    // tie key ascends with j as well and a strict '>' keeps the thread's winner (kSimpleTie; always true for plain
    // lowest-index ties). The key itself and the coordinates are looked up afterwards, for the winner only.
    float bt = -1.f;
    int bj = 0;
    uint32_t bkey = 0xffffffffu;
    if (kSimpleTie || lg < 0) {
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float x = SMEM_XYZ ? sx[j * THREADS + tid] : px[j];
        const float y = SMEM_XYZ ? sy[j * THREADS + tid] : py[j];
        const float z = SMEM_XYZ ? sz[j * THREADS + tid] : pz[j];
        const float dx = __fsub_rn(x, lx), dy = __fsub_rn(y, ly), dz = __fsub_rn(z, lz);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float t = fminf(td[j], d);
        td[j] = t;
        if (t > bt) { bt = t; bj = j; }
      }
      const uint32_t i = (uint32_t)((bj * CS + crank) * THREADS + tid);
      // padding (distance 0, i >= n) keeps a key above every real point's: it loses every tie
      bkey = (int)i < n ? fps_key(i, lg) : (0x7f000000u | i);
    } else {
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float x = SMEM_XYZ ? sx[j * THREADS + tid] : px[j];
        const float y = SMEM_XYZ ? sy[j * THREADS + tid] : py[j];
        const float z = SMEM_XYZ ? sz[j * THREADS + tid] : pz[j];
        const float dx = __fsub_rn(x, lx), dy = __fsub_rn(y, ly), dz = __fsub_rn(z, lz);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float t = fminf(td[j], d);
        td[j] = t;
        const uint32_t i = (uint32_t)((j * CS + crank) * THREADS + tid);
        const uint32_t key = (int)i < n ? fps_key(i, lg) : (0x7f000000u | i);
        if (t > bt || (t == bt && key < bkey)) { bt = t; bkey = key; bj = j; }
      }
    }
    const uint32_t hi = __float_as_uint(bt);
    const uint32_t lo = 0xffffffffu - bkey;
    // 2. warp arg-max (2 REDUX); the winning lane publishes key + coordinates
    uint32_t whi, wlo;
    if (warp_argmax(hi, lo, whi, wlo)) {
      float bx = 0.f, by = 0.f, bz = 0.f;
      if (SMEM_XYZ) {
        bx = sx[bj * THREADS + tid]; by = sy[bj * THREADS + tid]; bz = sz[bj * THREADS + tid];
      } else {
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (j == bj) { bx = px[j]; by = py[j]; bz = pz[j]; }
      }
      FpsMsg m;
      m.hi = hi; m.lo = lo; m.x = bx; m.y = by; m.z = bz;
      s_warp[warp] = m;
    }
    __syncthreads();
    // 3. CTA arg-max by warp 0, then one 32-byte message to every CTA of the cluster
    const int par = it & 1;
    if (CS > 1 && async_msg && tid == 0)   // this phase completes when the CS x 32 message bytes have landed
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fps_smem_u32(&s_bar[par])), "r"(CS * 32) : "memory");
    if (warp == 0) {
      const bool have = lane < NW;
      const uint32_t chi = have ? s_warp[lane].hi : 0u, clo = have ? s_warp[lane].lo : 0u;
      uint32_t bhi, blo;
      const bool win = warp_argmax(chi, clo, bhi, blo) && have;
      const int src = __ffs(__ballot_sync(0xffffffffu, win)) - 1;
      if (CS == 1) {
        if (lane == src) s_slot[par][0] = s_warp[lane];
      } else if (lane < CS) {
        const uint4* m = reinterpret_cast<const uint4*>(&s_warp[src]);
        const uint4 m0 = m[0], m1 = m[1];
        __syncwarp(__activemask());   // every sender has read the winner before anyone arrives
        uint32_t rdst, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(fps_smem_u32(&s_slot[par][crank])), "r"(lane));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(fps_smem_u32(&s_bar[par])), "r"(lane));
        if (async_msg) {
          // st.async: the stores themselves complete the peer's mbarrier transaction count - no release fence, no
          // separate remote arrive on the exchange's critical path
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rdst), "r"(m0.x), "r"(m0.y), "r"(m0.z), "r"(m0.w), "r"(rbar) : "memory");
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rdst + 16), "r"(m1.x), "r"(m1.y), "r"(m1.z), "r"(m1.w), "r"(rbar) : "memory");
        } else {
          asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rdst), "r"(m0.x), "r"(m0.y), "r"(m0.z), "r"(m0.w) : "memory");
          asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rdst + 16), "r"(m1.x), "r"(m1.y), "r"(m1.z), "r"(m1.w) : "memory");
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
        }
      }
    }
    // 4. wait for the CS candidates of this iteration, resolve the winner (identical in every CTA)
    if (CS == 1) {
      __syncthreads();
    } else {
      const uint32_t parity = (uint32_t)((it - 1) >> 1) & 1u;   // each barrier is used every 2nd iteration
      if (async_msg) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "FPS_WAITA:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra FPS_DONEA;\n\t"
            "bra FPS_WAITA;\n\t"
            "FPS_DONEA:\n\t"
            "}" ::"r"(fps_smem_u32(&s_bar[par])), "r"(parity) : "memory");
      } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "FPS_WAIT:\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra FPS_DONE;\n\t"
            "bra FPS_WAIT;\n\t"
            "FPS_DONE:\n\t"
            "}" ::"r"(fps_smem_u32(&s_bar[par])), "r"(parity) : "memory");
      }
    }
    FpsMsg w = s_slot[par][0];
#pragma unroll
    for (int r = 1; r < CS; ++r) {
      const FpsMsg c = s_slot[par][r];
      if (c.hi > w.hi || (c.hi == w.hi && c.lo > w.lo)) w = c;
    }
    lx = w.x; ly = w.y; lz = w.z;
    if (tid == 0) s_sel[it] = (int)fps_unkey(0xffffffffu - w.lo, lg);
  }
  __syncthreads();
  if (crank != 0) return;  // every CTA holds the same selection; rank 0 writes it

  // gather + per-scene min/max of the sampled set + normalise to [0,1]
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int q = tid; q < nq; q += THREADS) {
    int i = n > 0 ? s_sel[q] : 0;
    const float* g = gather_src + (size_t)(s0 + i) * gather_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = n > 0 ? __ldg(g + c) : 0.f;
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s_red[warp][c] = mn[c]; s_red[warp][3 + c] = mx[c]; }
  }
  __syncthreads();
  if (tid < 6) {
    int c = tid;
    float v = s_red[0][c];
    for (int w = 1; w < NW; ++w) v = c < 3 ? fminf(v, s_red[w][c]) : fmaxf(v, s_red[w][c]);
    s_mm[c] = v;
  }
  __syncthreads();
  for (int q = tid; q < nq; q += THREADS) {
    int i = n > 0 ? s_sel[q] : 0;
    idx_out[(size_t)scene * nq + q] = i;
    const float* g = gather_src + (size_t)(s0 + i) * gather_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int sc = reverse ? 2 - c : c;
      float v = n > 0 ? __ldg(g + sc) : 0.f;
      // shift_scale_points: ((p - min) * 1) / (max - min) + 0
      float r = __fdiv_rn(__fsub_rn(v, s_mm[sc]), __fsub_rn(s_mm[3 + sc], s_mm[sc]));
      out[((size_t)scene * nq + q) * 3 + c] = r;
    }
  }
}

template <int THREADS, int PPT, int CS, bool SMEM_XYZ>
static int launch_fps(const float* dist_src, int dist_stride, int dist_seg_stride,
                      const float* gather_src, int gather_stride, const int32_t* seg, int B, int nq,
                      int reverse, int tie_block, int32_t* idx, float* out, cudaStream_t st) {
  auto kern = k_fps<THREADS, PPT, CS, SMEM_XYZ>;
  const size_t dyn = SMEM_XYZ ? (size_t)3 * PPT * THREADS * sizeof(float) : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * CS);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int async_msg = 1;                // st.async message exchange (U3D_FPS_ASYNC=0: plain DSMEM stores + remote arrive)
  if (const char* e = getenv("U3D_FPS_ASYNC")) async_msg = atoi(e) != 0;
  static bool configured = false;   // per template instantiation
  if (!configured) {
    if (CS > 8) U3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (dyn > 48 * 1024)
      U3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    configured = true;
  }
  U3D_CUDA(cudaLaunchKernelEx(&cfg, kern, dist_src, dist_stride, dist_seg_stride, gather_src,
                              gather_stride, seg, nq, reverse, tie_block, async_msg, idx, out));
  count_launch();
  return U3D_OK;
}

__global__ void k_coors_to_float(const int32_t* __restrict__ coors, int rows, float* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int4 c = __ldg(reinterpret_cast<const int4*>(coors) + r);
  out[(size_t)r * 3 + 0] = (float)c.y;
  out[(size_t)r * 3 + 1] = (float)c.z;
  out[(size_t)r * 3 + 2] = (float)c.w;
}

}  // namespace u3d

using namespace u3d;

extern "C" int u3d_fps(const float* dist_src, int dist_stride, int dist_seg_stride,
                       const float* gather_src, int gather_stride, const int32_t* seg, int B,
                       int max_n, int nq, int reverse, int tie_block, int32_t* idx, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(dist_src && gather_src && seg && idx && out, "u3d_fps: null buffer");
  U3D_CHECK_ARG(B >= 1 && nq >= 1 && nq <= kFpsMaxNq && max_n >= 0, "u3d_fps: bad B/nq (nq<=%d)",
                kFpsMaxNq);
  U3D_CHECK_ARG(dist_stride >= 3 && gather_stride >= 3 && dist_seg_stride >= 3, "u3d_fps: strides");
  U3D_CHECK_ARG(tie_block >= 0 && tie_block <= 1024 && (tie_block & (tie_block - 1)) == 0,
                "u3d_fps: tie_block=%d must be 0 or a power of two <= 1024", tie_block);
#define U3D_FPS_CASE(THREADS, PPT, CS, SM)                                                         \
  if ((long long)max_n <= (long long)THREADS * PPT * CS)                                          \
    return launch_fps<THREADS, PPT, CS, SM>(dist_src, dist_stride, dist_seg_stride, gather_src,   \
                                           gather_stride, seg, B, nq, reverse, tie_block, idx, out, st);
  // smallest configuration that keeps every point of a scene on chip: registers up to 16
  // points/thread, shared memory for the coordinates beyond that (distances stay in registers)
  U3D_FPS_CASE(256, 8, 1, false)     //   2 048
  U3D_FPS_CASE(256, 8, 2, false)     //   4 096
  U3D_FPS_CASE(256, 8, 4, false)     //   8 192
  U3D_FPS_CASE(256, 8, 8, false)     //  16 384
  if (const char* e = getenv("U3D_FPS_CFG")) {   // experiments: other splits of a <= 20 480-point scene
    const int c = atoi(e);
    if (c == 1) { U3D_FPS_CASE(512, 10, 4, false) }
    if (c == 2) { U3D_FPS_CASE(1024, 10, 2, false) }
    if (c == 3) { U3D_FPS_CASE(512, 5, 8, false) }
    if (c == 4) { U3D_FPS_CASE(128, 10, 16, false) }
    if (c == 5) { U3D_FPS_CASE(1024, 5, 4, false) }
    if (c == 6) { U3D_FPS_CASE(128, 20, 8, false) }
  }
  if (getenv("U3D_FPS_FAT") != nullptr) {
    // optional: few fat CTAs (2 x 1024 threads, coordinates in shared memory): one DSMEM exchange
    // partner and only 2 SMs per scene (measured: 0.54 vs 0.34 ms per launch, same step time)
    U3D_FPS_CASE(1024, 10, 2, true)  //  20 480
  }
  U3D_FPS_CASE(256, 10, 8, false)    //  20 480
  U3D_FPS_CASE(256, 16, 8, false)    //  32 768
  U3D_FPS_CASE(512, 16, 8, false)    //  65 536
  U3D_FPS_CASE(512, 13, 16, false)   // 106 496 (non-portable cluster of 16)
  U3D_FPS_CASE(512, 16, 16, false)   // 131 072
  U3D_FPS_CASE(512, 26, 16, true)    // 212 992
#undef U3D_FPS_CASE
  set_error("u3d_fps: max_n=%d exceeds the 212992-point on-chip limit", max_n);
  return U3D_EINVAL;
}

extern "C" int u3d_coors_to_float(const int32_t* coors, int rows, float* out, void* stream) {
  U3D_CHECK_ARG(coors && out && rows >= 0, "u3d_coors_to_float: bad argument");
  if (rows == 0) return U3D_OK;
  k_coors_to_float<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(coors, rows, out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

// K4 — batched D-FPS (furthest point sampling) + gather + min-max normalisation.
//
// Reference: mmcv.ops.PointsSampler([nq]) / furthest_point_sample + gather_points +
// shift_scale_points, projects/mmdet3d_plugin/models/detectors/uni3detr.py:138,178-187
// (semantics SURVEY.md A.5/A.6). The reference runs one CTA per scene inside a python
// loop over the batch and re-reads all N points from HBM on each of the nq iterations.
//
// Here one thread-block CLUSTER owns one scene: every thread keeps PPT points and their
// running min-distances in registers for the whole kernel (points are read from HBM
// exactly once: N*12 B), the per-iteration arg-max is a warp-shuffle reduction, a
// shared-memory stage and - for clusters - one DSMEM exchange + cluster barrier.
// All scenes of the batch run concurrently (grid = B clusters).
//
// Defined arithmetic (mirrored by oracle/): d = ((dx*dx + dy*dy) + dz*dz) with no FMA
// contraction, running distance initialised to 1e10, first pick = index 0, ties -> the
// lowest index.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace u3d {

constexpr int kFpsThreads = 1024;
constexpr int kFpsMaxNq = 4096;

struct Cand {
  float d;
  int idx;
  float x, y, z;
};

__device__ __forceinline__ bool better(float d, int i, float d2, int i2) {
  return d > d2 || (d == d2 && i < i2);
}

template <int PPT, int CS>
__global__ void __launch_bounds__(kFpsThreads, 1)
k_fps(const float* __restrict__ dist_src, int dist_stride, int dist_seg_stride,
      const float* __restrict__ gather_src, int gather_stride, const int32_t* __restrict__ seg,
      int nq, int reverse, int32_t* __restrict__ idx_out, float* __restrict__ out) {
  __shared__ Cand s_warp[32];
  __shared__ Cand s_slot[2][CS];
  __shared__ int s_sel[kFpsMaxNq];
  __shared__ float s_mm[6];

  const int scene = blockIdx.x / CS;
  const int crank = blockIdx.x % CS;  // == cluster block rank (1-D cluster)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s0 = seg[scene];
  const int n = seg[scene + 1] - s0;
  const float* base = dist_src + (size_t)s0 * dist_seg_stride;

  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    int i = (j * CS + crank) * kFpsThreads + threadIdx.x;
    bool ok = i < n;
    px[j] = ok ? __ldg(base + (size_t)i * dist_stride + 0) : 0.f;
    py[j] = ok ? __ldg(base + (size_t)i * dist_stride + 1) : 0.f;
    pz[j] = ok ? __ldg(base + (size_t)i * dist_stride + 2) : 0.f;
    td[j] = ok ? 1e10f : -1.f;  // padding can never win (real distances are >= 0)
  }

  float lx = 0.f, ly = 0.f, lz = 0.f;
  if (n > 0) {
    lx = __ldg(base + 0);
    ly = __ldg(base + 1);
    lz = __ldg(base + 2);
  }
  if (threadIdx.x == 0) s_sel[0] = 0;
  if (CS > 1) cg::this_cluster().sync();  // all CTAs of the cluster are resident before DSMEM use

  for (int it = 1; it < nq; ++it) {
    // 1. update running distances, local arg-max (ascending index => strict '>' keeps lowest)
    Cand best;
    best.d = -2.f; best.idx = 0x7fffffff; best.x = 0.f; best.y = 0.f; best.z = 0.f;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      float dx = __fsub_rn(px[j], lx), dy = __fsub_rn(py[j], ly), dz = __fsub_rn(pz[j], lz);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      float t = fminf(td[j], d);
      td[j] = t;
      if (t > best.d) {
        best.d = t;
        best.idx = (j * CS + crank) * kFpsThreads + threadIdx.x;
        best.x = px[j]; best.y = py[j]; best.z = pz[j];
      }
    }
    // 2. warp arg-max on (d, idx); the winning lane publishes its coordinates
    float wd = best.d;
    int wi = best.idx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, wd, o);
      int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (better(od, oi, wd, wi)) { wd = od; wi = oi; }
    }
    if (best.idx == wi && best.d == wd) s_warp[warp] = best;  // unique: idx is unique
    __syncthreads();
    // 3. CTA arg-max by warp 0
    if (warp == 0) {
      Cand c = s_warp[lane];
      float cd = c.d;
      int ci = c.idx;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, cd, o);
        int oi = __shfl_xor_sync(0xffffffffu, ci, o);
        if (better(od, oi, cd, ci)) { cd = od; ci = oi; }
      }
      if (c.idx == ci && c.d == cd) {
        if (CS == 1) {
          s_slot[it & 1][0] = c;
        } else {
          cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
          for (int r = 0; r < CS; ++r) {
            Cand* remote = cluster.map_shared_rank(&s_slot[it & 1][crank], r);
            *remote = c;
          }
        }
      }
    }
    if (CS == 1) {
      __syncthreads();
    } else {
      cg::this_cluster().sync();
    }
    // 4. every thread resolves the winner among the CS CTA candidates
    Cand w = s_slot[it & 1][0];
#pragma unroll
    for (int r = 1; r < CS; ++r) {
      Cand c = s_slot[it & 1][r];
      if (better(c.d, c.idx, w.d, w.idx)) w = c;
    }
    lx = w.x; ly = w.y; lz = w.z;
    if (threadIdx.x == 0) s_sel[it] = w.idx;
  }
  __syncthreads();
  if (crank != 0) return;  // every CTA holds the same selection; rank 0 writes it

  // gather + per-scene min/max of the sampled set + normalise to [0,1]
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int q = threadIdx.x; q < nq; q += kFpsThreads) {
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
  __shared__ float s_red[32][6];
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s_red[warp][c] = mn[c]; s_red[warp][3 + c] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int c = threadIdx.x;
    float v = s_red[0][c];
    for (int w = 1; w < 32; ++w) v = c < 3 ? fminf(v, s_red[w][c]) : fmaxf(v, s_red[w][c]);
    s_mm[c] = v;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += kFpsThreads) {
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

template <int PPT, int CS>
static int launch_fps(const float* dist_src, int dist_stride, int dist_seg_stride,
                      const float* gather_src, int gather_stride, const int32_t* seg, int B, int nq,
                      int reverse, int32_t* idx, float* out, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * CS);
  cfg.blockDim = dim3(kFpsThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CS > 8)
    U3D_CUDA(cudaFuncSetAttribute(k_fps<PPT, CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  U3D_CUDA(cudaLaunchKernelEx(&cfg, k_fps<PPT, CS>, dist_src, dist_stride, dist_seg_stride,
                              gather_src, gather_stride, seg, nq, reverse, idx, out));
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
                       int max_n, int nq, int reverse, int32_t* idx, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(dist_src && gather_src && seg && idx && out, "u3d_fps: null buffer");
  U3D_CHECK_ARG(B >= 1 && nq >= 1 && nq <= kFpsMaxNq && max_n >= 0, "u3d_fps: bad B/nq (nq<=%d)",
                kFpsMaxNq);
  U3D_CHECK_ARG(dist_stride >= 3 && gather_stride >= 3 && dist_seg_stride >= 3, "u3d_fps: strides");
#define U3D_FPS_CASE(PPT, CS)                                                                    \
  if ((long long)max_n <= (long long)PPT * CS * kFpsThreads)                                     \
    return launch_fps<PPT, CS>(dist_src, dist_stride, dist_seg_stride, gather_src, gather_stride, \
                               seg, B, nq, reverse, idx, out, st);
  // smallest register footprint that covers max_n; clusters first (latency), then PPT
  // (a 1024-thread CTA has 64 registers/thread: at most ~13 points of 4 floats each)
  U3D_FPS_CASE(4, 1)    //   4 096
  U3D_FPS_CASE(4, 2)    //   8 192
  U3D_FPS_CASE(4, 4)    //  16 384
  U3D_FPS_CASE(4, 8)    //  32 768
  U3D_FPS_CASE(8, 8)    //  65 536
  U3D_FPS_CASE(12, 8)   //  98 304
  U3D_FPS_CASE(8, 16)   // 131 072 (non-portable cluster of 16)
  U3D_FPS_CASE(13, 16)  // 212 992
#undef U3D_FPS_CASE
  set_error("u3d_fps: max_n=%d exceeds the 212992-point register-resident limit", max_n);
  return U3D_EINVAL;
}

extern "C" int u3d_coors_to_float(const int32_t* coors, int rows, float* out, void* stream) {
  U3D_CHECK_ARG(coors && out && rows >= 0, "u3d_coors_to_float: bad argument");
  if (rows == 0) return U3D_OK;
  k_coors_to_float<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(coors, rows, out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

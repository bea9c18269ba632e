// K1 — point-to-voxel scatter (hard + dynamic voxelization) fused with the simple-VFE mean.
//
// Reference semantics (call sites
// projects/mmdet3d_plugin/models/detectors/uni3detr.py:148-149 hard,
// :156-167 dynamic). HBM-bound integer/atomic work: N*C*4 B read, M*(C*4+16) B written.
//
// Pipeline (hard):  mark cells (warp-aggregated atomicOr into the VoxelMap) -> scan ->
//   insert every point into its voxel's K-slot list with an atomicMin cascade (slot j
//   ends up holding the (j+1)-th smallest point index == arrival order, independent of
//   thread interleaving) -> per-scene block scan over "first point of its voxel" flags
//   gives the reference's first-appearance voxel order and the max_voxels cut ->
//   emit rows with 16-byte vector stores.
#include "common.cuh"

namespace u3d {

struct VoxGeom {
  float lo[3];   // x,y,z
  float vs[3];   // x,y,z
  int grid[3];   // x,y,z voxelization grid = round((hi-lo)/vs) like mmcv
  int dims[3];   // x,y,z index space of the VoxelMap (W,H,D) = sparse_shape, >= grid
};

__device__ __forceinline__ int scene_of(const int32_t* __restrict__ s_off, int B, int p) {
  int lo = 0, hi = B;  // find b with off[b] <= p < off[b+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (s_off[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// mmcv: c = floor((p - lo) / voxel) evaluated in fp32, point dropped if any c outside [0,grid)
__device__ __forceinline__ bool point_cell(const float* __restrict__ pt, const VoxGeom& g, int& cx,
                                           int& cy, int& cz) {
  float fx = floorf(__fdiv_rn(__fsub_rn(pt[0], g.lo[0]), g.vs[0]));
  float fy = floorf(__fdiv_rn(__fsub_rn(pt[1], g.lo[1]), g.vs[1]));
  float fz = floorf(__fdiv_rn(__fsub_rn(pt[2], g.lo[2]), g.vs[2]));
  // compare in float first so huge/NaN values cannot overflow the int conversion
  bool ok = fx >= 0.f && fx < (float)g.grid[0] && fy >= 0.f && fy < (float)g.grid[1] &&
            fz >= 0.f && fz < (float)g.grid[2];
  cx = ok ? (int)fx : -1;
  cy = ok ? (int)fy : -1;
  cz = ok ? (int)fz : -1;
  return ok;
}

constexpr int kMaxScenes = 1024;

__global__ void __launch_bounds__(256)
k_vox_mark(const float* __restrict__ pts, const int32_t* __restrict__ pt_off, int Ntot, int B, int C,
           VoxGeom g, uint2* __restrict__ map, uint32_t* __restrict__ pt_lin,
           int32_t* __restrict__ pt_coors /* (N,4) or null */) {
  extern __shared__ int32_t s_off[];
  for (int i = threadIdx.x; i <= B; i += blockDim.x) s_off[i] = pt_off[i];
  __syncthreads();
  const int D = g.dims[2], H = g.dims[1], W = g.dims[0];
  // whole warps stay converged for the aggregated atomic
  const int per_round = gridDim.x * blockDim.x;
  int nrounds = (Ntot + per_round - 1) / per_round;
  for (int r = 0; r < nrounds; ++r) {
    int p = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool inb = p < Ntot;
    bool valid = false;
    uint32_t lin = 0xffffffffu;
    if (inb) {
      int cx, cy, cz;
      valid = point_cell(pts + (size_t)p * C, g, cx, cy, cz);
      int b = scene_of(s_off, B, p);
      if (valid) lin = (uint32_t)((((size_t)b * D + cz) * H + cy) * W + cx);
      pt_lin[p] = lin;
      if (pt_coors) {
        int4 c = make_int4(b, cz, cy, cx);
        reinterpret_cast<int4*>(pt_coors)[p] = c;
      }
    }
    map_set_bit_aggregated(map, valid, lin);
  }
}

// slot j of voxel r converges to the (j+1)-th smallest point index of that voxel
__global__ void __launch_bounds__(256)
k_vox_insert(const uint32_t* __restrict__ pt_lin, int Ntot, const uint2* __restrict__ map,
             int32_t* __restrict__ slots, int K) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Ntot; p += gridDim.x * blockDim.x) {
    uint32_t lin = pt_lin[p];
    if (lin == 0xffffffffu) continue;
    int r = map_rank_at(map, lin);
    int32_t* s = slots + (size_t)r * K;
    int v = p;
    for (int j = 0; j < K; ++j) {
      int old = atomicMin(&s[j], v);
      if (old == kSlotEmpty) break;  // landed in an empty slot, nothing displaced
      v = max(old, v);               // carry the larger index to the next slot
    }
  }
}

// one CTA per scene: first-appearance (or linear) voxel order + max_voxels cut
__global__ void __launch_bounds__(1024)
k_vox_order(const uint32_t* __restrict__ pt_lin, const int32_t* __restrict__ pt_off,
            const uint2* __restrict__ map, const int32_t* __restrict__ slots, int K, int B,
            uint32_t cells_per_scene, int max_voxels, int order,
            int32_t* __restrict__ row_of_rank, int32_t* __restrict__ scene_rows) {
  __shared__ int smem[33];
  const int b = blockIdx.x;
  int base = 0;  // rows of the scenes before b
  for (int bb = 0; bb < b; ++bb) {
    int u = map_rank_at(map, (uint32_t)(bb + 1) * cells_per_scene) -
            map_rank_at(map, (uint32_t)bb * cells_per_scene);
    base += (max_voxels > 0 && u > max_voxels) ? max_voxels : u;
  }
  const int rs = map_rank_at(map, (uint32_t)b * cells_per_scene);
  const int re = map_rank_at(map, (uint32_t)(b + 1) * cells_per_scene);
  const int u_b = re - rs;
  const int m_b = (max_voxels > 0 && u_b > max_voxels) ? max_voxels : u_b;
  if (threadIdx.x == 0) {
    scene_rows[b] = base;
    if (b == B - 1) scene_rows[B] = base + m_b;
  }
  if (order == U3D_ORDER_LINEAR) {
    for (int r = rs + threadIdx.x; r < re; r += blockDim.x) {
      int vid = r - rs;
      row_of_rank[r] = vid < m_b ? base + vid : -1;
    }
    return;
  }
  const int p0 = pt_off[b], p1 = pt_off[b + 1];
  int running = 0;
  for (int start = p0; start < p1; start += blockDim.x) {
    int p = start + threadIdx.x;
    int r = -1, flag = 0;
    if (p < p1) {
      uint32_t lin = pt_lin[p];
      if (lin != 0xffffffffu) {
        r = map_rank_at(map, lin);
        flag = (slots[(size_t)r * K] == p);
      }
    }
    int total;
    int ex = block_exclusive_scan(flag, smem, total);
    if (flag) {
      int vid = running + ex;
      row_of_rank[r] = vid < m_b ? base + vid : -1;
    }
    running += total;
  }
}

// one thread per point; the first point of every kept voxel writes that voxel's row
template <int CVEC>  // CVEC==4: C==4 fast path with float4 traffic, 0: generic
__global__ void __launch_bounds__(256)
k_vox_emit(const float* __restrict__ pts, const uint32_t* __restrict__ pt_lin, int Ntot, int C,
           const uint2* __restrict__ map, const int32_t* __restrict__ slots, int K,
           const int32_t* __restrict__ row_of_rank, int D, int H, int W,
           int32_t* __restrict__ coors, int32_t* __restrict__ num_points,
           float* __restrict__ voxels, float* __restrict__ feats) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Ntot; p += gridDim.x * blockDim.x) {
    uint32_t lin = pt_lin[p];
    if (lin == 0xffffffffu) continue;
    int r = map_rank_at(map, lin);
    const int32_t* s = slots + (size_t)r * K;
    if (s[0] != p) continue;
    int row = row_of_rank[r];
    if (row < 0) continue;
    uint32_t t = lin;
    int x = t % W; t /= W;
    int y = t % H; t /= H;
    int z = t % D; t /= D;
    reinterpret_cast<int4*>(coors)[row] = make_int4((int)t, z, y, x);
    int n = 0;
    if (CVEC == 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < K; ++j) {
        int q = s[j];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q != kSlotEmpty) {
          v = __ldg(reinterpret_cast<const float4*>(pts) + q);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          ++n;
        }
        if (voxels) reinterpret_cast<float4*>(voxels)[(size_t)row * K + j] = v;
      }
      if (feats) {
        float fn = (float)n;
        reinterpret_cast<float4*>(feats)[row] =
            make_float4(__fdiv_rn(acc.x, fn), __fdiv_rn(acc.y, fn), __fdiv_rn(acc.z, fn),
                        __fdiv_rn(acc.w, fn));
      }
    } else {
      for (int j = 0; j < K; ++j) n += (s[j] != kSlotEmpty);
      for (int c = 0; c < C; ++c) {
        float acc = 0.f;
        for (int j = 0; j < K; ++j) {
          int q = s[j];
          float v = (q != kSlotEmpty) ? __ldg(pts + (size_t)q * C + c) : 0.f;
          acc += v;
          if (voxels) voxels[((size_t)row * K + j) * C + c] = v;
        }
        if (feats) feats[(size_t)row * C + c] = __fdiv_rn(acc, (float)n);
      }
    }
    num_points[row] = n;
  }
}

// ------------------------------------------------------------- dynamic -------
__global__ void __launch_bounds__(256)
k_dyn_accum(const float* __restrict__ pts, const uint32_t* __restrict__ pt_lin, int Ntot, int C,
            const uint2* __restrict__ map, int D, int H, int W, int32_t* __restrict__ coors,
            float* __restrict__ sums, int32_t* __restrict__ cnt) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Ntot; p += gridDim.x * blockDim.x) {
    uint32_t lin = pt_lin[p];
    if (lin == 0xffffffffu) continue;
    int r = map_rank_at(map, lin);
    int old = atomicAdd(&cnt[r], 1);
    if (old == 0) {  // exactly one point per voxel writes the coordinate row
      uint32_t t = lin;
      int x = t % W; t /= W;
      int y = t % H; t /= H;
      int z = t % D; t /= D;
      reinterpret_cast<int4*>(coors)[r] = make_int4((int)t, z, y, x);
    }
    for (int c = 0; c < C; ++c) atomicAdd(&sums[(size_t)r * C + c], __ldg(pts + (size_t)p * C + c));
  }
}

__global__ void __launch_bounds__(256)
k_dyn_finalize(float* __restrict__ feats, const int32_t* __restrict__ cnt,
               const int32_t* __restrict__ n_rows, int C) {
  const int n = *n_rows;
  const long long total = (long long)n * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / C);
    feats[i] = __fdiv_rn(feats[i], (float)cnt[r]);
  }
}

__global__ void k_scene_rows_from_map(const uint2* __restrict__ map, int B, uint32_t cells_per_scene,
                                      int32_t* __restrict__ scene_rows) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= B) scene_rows[b] = map_rank_at(map, (uint32_t)b * cells_per_scene);
}

// mmcv Voxelization.__init__: grid = round((range[3:] - range[:3]) / voxel_size) in fp32.
// (D,H,W) is the index space of the map (the encoder's sparse_shape): SECOND-style configs
// declare it one cell deeper than the voxel grid (KITTI/nuScenes: grid z = 40, sparse z = 41).
static int fill_geom(VoxGeom& g, const float* pc_range, const float* voxel_size, int D, int H, int W) {
  const int dims[3] = {W, H, D};
  for (int i = 0; i < 3; ++i) {
    g.lo[i] = pc_range[i];
    g.vs[i] = voxel_size[i];
    if (!(voxel_size[i] > 0.f)) {
      set_error("voxel_size[%d] must be positive", i);
      return U3D_EINVAL;
    }
    g.grid[i] = (int)nearbyintf((pc_range[3 + i] - pc_range[i]) / voxel_size[i]);
    g.dims[i] = dims[i];
    if (g.grid[i] < 1 || g.grid[i] > dims[i]) {
      set_error("voxel grid axis %d = %d does not fit the index space %d", i, g.grid[i], dims[i]);
      return U3D_EINVAL;
    }
  }
  return U3D_OK;
}

static inline int grid_for(long long n, int threads, int max_ctas = kNumSMs * 8) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > max_ctas) g = max_ctas;
  return (int)g;
}

}  // namespace u3d

using namespace u3d;

extern "C" int u3d_voxelize_hard(const float* points, const int32_t* pt_off, int Ntot, int B, int C,
                                 const float* pc_range, const float* voxel_size, int D, int H, int W,
                                 int max_pts, int max_voxels, int order, void* map_,
                                 int32_t* scan_scratch, uint32_t* pt_lin, int32_t* slots,
                                 int32_t* row_of_rank, int32_t* coors, int32_t* num_points,
                                 float* voxels, float* feats, int32_t* scene_rows, int cap,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(points && pt_off && map_ && pt_lin && slots && row_of_rank && coors && num_points &&
                    scene_rows && scan_scratch,
                "u3d_voxelize_hard: null buffer");
  U3D_CHECK_ARG(B >= 1 && B <= kMaxScenes && C >= 3 && max_pts >= 1 && Ntot >= 0,
                "u3d_voxelize_hard: bad B=%d C=%d max_pts=%d Ntot=%d", B, C, max_pts, Ntot);
  size_t words = u3d_voxmap_words(B, D, H, W);
  if (!words) {
    set_error("u3d_voxelize_hard: B*D*H*W exceeds the 32-bit cell index; split the batch");
    return U3D_ERANGE;
  }
  long long need = max_voxels > 0 ? (long long)B * max_voxels : Ntot;
  if (need > Ntot) need = Ntot;
  U3D_CHECK_ARG(cap >= need, "u3d_voxelize_hard: cap=%d < %lld", cap, need);
  uint2* map = (uint2*)map_;
  VoxGeom g;
  if (int grc = fill_geom(g, pc_range, voxel_size, D, H, W)) return grc;
  U3D_CUDA(cudaMemsetAsync(map, 0, words * sizeof(uint2), st));
  U3D_CUDA(cudaMemsetAsync(slots, 0x7f, (size_t)max(Ntot, 1) * max_pts * sizeof(int32_t), st));
  if (voxels == nullptr && feats == nullptr) {
    // nothing to emit but coordinates; still fine
  }
  int threads = 256;
  k_vox_mark<<<grid_for(Ntot, threads), threads, (B + 1) * sizeof(int32_t), st>>>(
      points, pt_off, Ntot, B, C, g, map, pt_lin, nullptr);
  U3D_LAUNCH_CHECK();
  int rc = voxmap_scan(map, words, scan_scratch, nullptr, st);
  if (rc) return rc;
  k_vox_insert<<<grid_for(Ntot, threads), threads, 0, st>>>(pt_lin, Ntot, map, slots, max_pts);
  U3D_LAUNCH_CHECK();
  k_vox_order<<<B, 1024, 0, st>>>(pt_lin, pt_off, map, slots, max_pts, B,
                                  (uint32_t)((size_t)D * H * W), max_voxels, order, row_of_rank,
                                  scene_rows);
  U3D_LAUNCH_CHECK();
  if (C == 4 && (((uintptr_t)points | (uintptr_t)voxels | (uintptr_t)feats) & 15) == 0) {
    k_vox_emit<4><<<grid_for(Ntot, threads), threads, 0, st>>>(
        points, pt_lin, Ntot, C, map, slots, max_pts, row_of_rank, D, H, W, coors, num_points,
        voxels, feats);
  } else {
    k_vox_emit<0><<<grid_for(Ntot, threads), threads, 0, st>>>(
        points, pt_lin, Ntot, C, map, slots, max_pts, row_of_rank, D, H, W, coors, num_points,
        voxels, feats);
  }
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_voxelize_dynamic(const float* points, const int32_t* pt_off, int Ntot, int B,
                                    int C, const float* pc_range, const float* voxel_size, int D,
                                    int H, int W, void* map_, int32_t* scan_scratch,
                                    uint32_t* pt_lin, int32_t* pt_coors, int32_t* coors,
                                    float* feats, int32_t* cnt, int32_t* scene_rows, int cap,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(points && pt_off && map_ && pt_lin && coors && feats && cnt && scene_rows &&
                    scan_scratch,
                "u3d_voxelize_dynamic: null buffer");
  U3D_CHECK_ARG(B >= 1 && B <= kMaxScenes && C >= 3 && Ntot >= 0, "u3d_voxelize_dynamic: bad shape");
  U3D_CHECK_ARG(cap >= Ntot, "u3d_voxelize_dynamic: cap=%d < Ntot=%d", cap, Ntot);
  size_t words = u3d_voxmap_words(B, D, H, W);
  if (!words) {
    set_error("u3d_voxelize_dynamic: B*D*H*W exceeds the 32-bit cell index; split the batch");
    return U3D_ERANGE;
  }
  uint2* map = (uint2*)map_;
  VoxGeom g;
  if (int grc = fill_geom(g, pc_range, voxel_size, D, H, W)) return grc;
  U3D_CUDA(cudaMemsetAsync(map, 0, words * sizeof(uint2), st));
  U3D_CUDA(cudaMemsetAsync(cnt, 0, (size_t)max(cap, 1) * sizeof(int32_t), st));
  U3D_CUDA(cudaMemsetAsync(feats, 0, (size_t)max(cap, 1) * C * sizeof(float), st));
  int threads = 256;
  k_vox_mark<<<grid_for(Ntot, threads), threads, (B + 1) * sizeof(int32_t), st>>>(
      points, pt_off, Ntot, B, C, g, map, pt_lin, pt_coors);
  U3D_LAUNCH_CHECK();
  int rc = voxmap_scan(map, words, scan_scratch, nullptr, st);
  if (rc) return rc;
  k_scene_rows_from_map<<<cdiv(B + 1, 128), 128, 0, st>>>(map, B, (uint32_t)((size_t)D * H * W),
                                                          scene_rows);
  U3D_LAUNCH_CHECK();
  k_dyn_accum<<<grid_for(Ntot, threads), threads, 0, st>>>(points, pt_lin, Ntot, C, map, D, H, W,
                                                           coors, feats, cnt);
  U3D_LAUNCH_CHECK();
  k_dyn_finalize<<<grid_for((long long)cap * C, threads), threads, 0, st>>>(feats, cnt,
                                                                            scene_rows + B, C);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

// Input pre-stage ("the step before the path"): the point-cloud part of the
// reference's test pipeline on the device, feeding u3d_voxelize_* directly:
//   LoadPointsFromFile(load_dim, use_dim, shift_height)  -> column select + height feature
//   PointsRangeFilter(point_cloud_range)                 -> order-preserving compaction
//   PointSample(num_points)                              -> gather through host-drawn indices
// Reference call sites: projects/configs/uni3detr/uni3detr_sunrgbd.py:175-191 (test_pipeline); the
// transforms themselves are mmdet3d v1.0.0rc5 (not vendored): semantics restated in oracle/pipeline.py.
//
// STATUS: validated on hardware in round 2 (tests/test_prestage.py, un-gated). Nothing on the benchmarked path
// calls it: the bench starts from points that are already in the model's format.
//
// HBM-bound byte work: one read of the raw floats, one write of the kept rows; the only non-trivial
// part is the 0.99th percentile of z behind shift_height (np.percentile(z, 0.99), linear
// interpolation between two order statistics): an exact radix select per scene, no sort.
#include "common.cuh"

namespace u3d {

constexpr int kPrepThreads = 1024;
constexpr int kMaxUse = 8;

struct PrepCfg {
  int load_dim, n_use, shift_height, filter;
  int use[kMaxUse];
  float lo[3], hi[3];
};

__device__ __forceinline__ uint32_t float_key(float f) {   // order-preserving float -> uint32
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// block-wide sum / min over one value per thread (all threads call)
__device__ __forceinline__ int block_sum(int v, int* smem) {
  int total;
  block_exclusive_scan(v, smem, total);
  return total;
}

// One CTA per scene: (1) floor_z[b] = the 0.99th percentile of the z column (exact order
// statistics by 4 x 8-bit radix select, interpolated in double like np.percentile's 'linear'),
// (2) kept[b] = number of points inside the open range box.
__global__ void __launch_bounds__(kPrepThreads)
k_points_stats(const float* __restrict__ raw, const int32_t* __restrict__ off, PrepCfg cfg,
               float* __restrict__ floor_z, int32_t* __restrict__ kept) {
  __shared__ int s_hist[256];
  __shared__ int s_scan[33];
  __shared__ uint32_t s_prefix;
  __shared__ int s_rank;
  __shared__ unsigned int s_min;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int p0 = off[b], n = off[b + 1] - p0;
  const float* base = raw + (size_t)p0 * cfg.load_dim;
  const int zc = cfg.use[2];

  // ---- in-range count
  int cnt = 0;
  for (int i = tid; i < n; i += kPrepThreads) {
    const float* p = base + (size_t)i * cfg.load_dim;
    const float x = p[cfg.use[0]], y = p[cfg.use[1]], z = p[zc];
    const bool in = !cfg.filter || (x > cfg.lo[0] && y > cfg.lo[1] && z > cfg.lo[2] && x < cfg.hi[0] &&
                                    y < cfg.hi[1] && z < cfg.hi[2]);
    cnt += in ? 1 : 0;
  }
  const int total = block_sum(cnt, s_scan);
  if (tid == 0) kept[b] = total;

  if (!cfg.shift_height) {
    if (tid == 0) floor_z[b] = 0.f;
    return;
  }
  if (n == 0) {
    if (tid == 0) floor_z[b] = 0.f;
    return;
  }
  // virtual index (n-1)*q, q = 0.99/100, as numpy computes it: n*q + (1 - q) - 1
  const double q = 0.99 / 100.0;
  const double vi = (double)n * q + (1.0 + q * (-1.0)) - 1.0;
  int k_lo = (int)floor(vi);
  if (k_lo < 0) k_lo = 0;
  if (k_lo > n - 1) k_lo = n - 1;
  const double gamma = vi - floor(vi);

  // ---- radix select of the k_lo-th smallest key
  if (tid == 0) { s_prefix = 0u; s_rank = k_lo; }
  __syncthreads();
  int less_total = 0;     // number of keys strictly below the selected key (thread 0 keeps it)
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = tid; i < 256; i += kPrepThreads) s_hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = pass == 3 ? 0u : (0xffffffffu << (8 * (pass + 1)));
    for (int i = tid; i < n; i += kPrepThreads) {
      const uint32_t key = float_key(base[(size_t)i * cfg.load_dim + zc]);
      if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> (8 * pass)) & 255u], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int r = s_rank, bin = 0, acc = 0;
      for (; bin < 256; ++bin) {
        if (acc + s_hist[bin] > r) break;
        acc += s_hist[bin];
      }
      if (bin > 255) bin = 255;
      s_rank = r - acc;
      less_total += acc;
      s_prefix = prefix | ((uint32_t)bin << (8 * pass));
    }
    __syncthreads();
  }
  const uint32_t key_lo = s_prefix;
  // ---- the next order statistic: the same value if it has spare copies, else the smallest larger key
  int eq = 0;
  unsigned int mn = 0xffffffffu;
  for (int i = tid; i < n; i += kPrepThreads) {
    const uint32_t key = float_key(base[(size_t)i * cfg.load_dim + zc]);
    eq += key == key_lo ? 1 : 0;
    if (key > key_lo && key < mn) mn = key;
  }
  if (tid == 0) s_min = 0xffffffffu;
  __syncthreads();
  atomicMin(&s_min, mn);
  const int eq_total = block_sum(eq, s_scan);   // contains the barrier that publishes s_min
  if (tid == 0) {
    const int k_hi = k_lo + 1 < n ? k_lo + 1 : n - 1;
    const float a = key_float(key_lo);
    const float bb = (k_hi < less_total + eq_total || s_min == 0xffffffffu) ? a : key_float(s_min);
    const double r = (double)a + ((double)bb - (double)a) * gamma;
    floor_z[b] = (float)r;
  }
}

__global__ void k_points_offsets(const int32_t* __restrict__ kept, int B, int32_t* __restrict__ out_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) { out_off[b] = acc; acc += kept[b]; }
    out_off[B] = acc;
  }
}

// One CTA per scene: select the columns, add the height feature, keep the in-range rows in order.
__global__ void __launch_bounds__(kPrepThreads)
k_points_emit(const float* __restrict__ raw, const int32_t* __restrict__ off, PrepCfg cfg,
              const float* __restrict__ floor_z, const int32_t* __restrict__ out_off,
              float* __restrict__ out) {
  __shared__ int s_scan[33];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int p0 = off[b], n = off[b + 1] - p0;
  const float* base = raw + (size_t)p0 * cfg.load_dim;
  const int C = cfg.n_use + (cfg.shift_height ? 1 : 0);
  const float fz = floor_z[b];
  int running = out_off[b];
  for (int start = 0; start < n; start += kPrepThreads) {
    const int i = start + tid;
    bool in = false;
    float v[kMaxUse];
    if (i < n) {
      const float* p = base + (size_t)i * cfg.load_dim;
#pragma unroll
      for (int c = 0; c < kMaxUse; ++c) v[c] = c < cfg.n_use ? p[cfg.use[c]] : 0.f;
      in = !cfg.filter || (v[0] > cfg.lo[0] && v[1] > cfg.lo[1] && v[2] > cfg.lo[2] && v[0] < cfg.hi[0] &&
                           v[1] < cfg.hi[1] && v[2] < cfg.hi[2]);
    }
    int total;
    const int ex = block_exclusive_scan(in ? 1 : 0, s_scan, total);
    if (in) {
      float* o = out + (size_t)(running + ex) * C;
      // mmdet3d LoadPointsFromFile: [x, y, z, height, rest...]
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
      int w = 3;
      if (cfg.shift_height) o[w++] = __fsub_rn(v[2], fz);
#pragma unroll
      for (int c = 3; c < kMaxUse; ++c)
        if (c < cfg.n_use) o[w++] = v[c];
    }
    running += total;
  }
}

__global__ void __launch_bounds__(256)
k_points_gather(const float* __restrict__ in, int C, const int32_t* __restrict__ choices, long long total,
                float* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    out[e] = __ldg(&in[(size_t)__ldg(&choices[r]) * C + c]);
  }
}

}  // namespace u3d

using namespace u3d;

extern "C" int u3d_points_prepare(const float* raw, const int32_t* raw_off, int B, int load_dim,
                                  const int32_t* use_dim, int n_use, int shift_height,
                                  const float* pc_range, float* floor_z, int32_t* kept, float* out,
                                  int32_t* out_off, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(raw && raw_off && use_dim && floor_z && kept && out && out_off, "u3d_points_prepare: null buffer");
  U3D_CHECK_ARG(B >= 1 && load_dim >= 3 && n_use >= 3 && n_use <= kMaxUse,
                "u3d_points_prepare: bad shape (B=%d load_dim=%d n_use=%d, n_use in [3,%d])", B, load_dim, n_use,
                kMaxUse);
  PrepCfg cfg = {};
  cfg.load_dim = load_dim;
  cfg.n_use = n_use;
  cfg.shift_height = shift_height ? 1 : 0;
  cfg.filter = pc_range ? 1 : 0;
  for (int i = 0; i < n_use; ++i) {
    U3D_CHECK_ARG(use_dim[i] >= 0 && use_dim[i] < load_dim, "u3d_points_prepare: use_dim[%d]=%d outside the row", i,
                  use_dim[i]);
    cfg.use[i] = use_dim[i];
  }
  for (int i = 0; i < 3; ++i) {
    cfg.lo[i] = pc_range ? pc_range[i] : 0.f;
    cfg.hi[i] = pc_range ? pc_range[3 + i] : 0.f;
  }
  k_points_stats<<<B, kPrepThreads, 0, st>>>(raw, raw_off, cfg, floor_z, kept);
  U3D_LAUNCH_CHECK();
  k_points_offsets<<<1, 32, 0, st>>>(kept, B, out_off);
  U3D_LAUNCH_CHECK();
  k_points_emit<<<B, kPrepThreads, 0, st>>>(raw, raw_off, cfg, floor_z, out_off, out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_points_gather(const float* in, int C, const int32_t* choices, int n, float* out,
                                 void* stream) {
  U3D_CHECK_ARG(in && choices && out && C >= 1 && n >= 0, "u3d_points_gather: bad argument");
  if (n == 0) return U3D_OK;
  const long long total = (long long)n * C;
  long long g = (total + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  k_points_gather<<<(int)g, 256, 0, (cudaStream_t)stream>>>(in, C, choices, total, out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

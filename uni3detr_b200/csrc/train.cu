// Training-side kernels: backward of the sparse convolution and of the
// UniCrossAtten sampling block, the Hungarian matcher and the aligned rotated 3-D IoU of the loss.
//
// Reference call sites (the arithmetic of the backward passes is torch.autograd through spconv /
// F.grid_sample in the reference; the matcher is scipy.optimize.linear_sum_assignment on the CPU):
//   * sparse conv autograd      projects/mmdet3d_plugin/models/pts_encoder/sparse_encoder_hd.py:106-138
//   * grid_sample autograd      projects/mmdet3d_plugin/models/utils/uni3detr_transformer.py:329-360
//   * Hungarian matching        projects/mmdet3d_plugin/core/bbox/assigners/hungarian_assigner_3d.py:124-139
//   * bbox_overlaps_3d (diag)   projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:690
// All kernels here are fp32 (the training step keeps the sparse encoder and the losses in fp32).
#include <float.h>
#include "bev_geom.cuh"

namespace u3d {

// ------------------------------------------------------------ rulebook transpose ---
// nbrT[k][i] = o  <=>  nbr[k][o] = i  (each (k, input row) feeds at most one output row). The data
// gradient of a sparse conv is the same gather-GEMM run over this table with W_k^T:
//   dX[i] = sum_k dY[nbrT[k][i]] @ W_k^T.
__global__ void k_rulebook_transpose(const int32_t* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ n_out_p,
                                     int K, int32_t* __restrict__ nbrT, int t_stride) {
  const int n_out = *n_out_p;
  const long long total = (long long)K * n_out;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / n_out), o = (int)(e - (long long)k * n_out);
    const int i = __ldg(&nbr[(size_t)k * nbr_stride + o]);
    if (i >= 0) nbrT[(size_t)k * t_stride + i] = o;
  }
}

// ------------------------------------------------------------ weight gradient ---
// dW[k][ci][co] = sum_o x[nbr[k][o]][ci] * dy[o][co]: per offset an (Cin x rows) @ (rows x Cout) product.
// Block = one (k, 64 x 64 tile of dW[k], chunk of output rows): 256 threads, 4 x 4 outputs each, rows
// staged 16 at a time in shared memory; partial tiles are accumulated into dW with fp32 atomics.
constexpr int kWgTile = 64;
constexpr int kWgRows = 16;

__global__ void __launch_bounds__(256)
k_spconv_wgrad(const float* __restrict__ x, const float* __restrict__ dy, const int32_t* __restrict__ nbr, int nbr_stride,
               const int32_t* __restrict__ n_out_p, int Cin, int Cout, int rows_per_block, float* __restrict__ dW) {
  __shared__ float xs[kWgRows][kWgTile + 1];
  __shared__ float ds[kWgRows][kWgTile + 1];
  const int n_out = *n_out_p;
  const int k = blockIdx.y;
  const int tiles_co = (Cout + kWgTile - 1) / kWgTile;
  const int ci0 = (blockIdx.x / tiles_co) * kWgTile, co0 = (blockIdx.x % tiles_co) * kWgTile;
  const int r_begin = blockIdx.z * rows_per_block;
  const int r_end = min(n_out, r_begin + rows_per_block);
  if (r_begin >= r_end) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  bool any = false;
  for (int r0 = r_begin; r0 < r_end; r0 += kWgRows) {
    // 16 rows x 64 channels of the gathered inputs and of dy: 1024 elements each, 4 per thread
#pragma unroll
    for (int e = threadIdx.x; e < kWgRows * kWgTile; e += 256) {
      const int rr = e / kWgTile, c = e - rr * kWgTile;
      const int o = r0 + rr;
      float xv = 0.f, dv = 0.f;
      if (o < r_end) {
        const int i = __ldg(&nbr[(size_t)k * nbr_stride + o]);
        if (i >= 0) {
          if (ci0 + c < Cin) xv = __ldg(&x[(size_t)i * Cin + ci0 + c]);
          if (co0 + c < Cout) dv = __ldg(&dy[(size_t)o * Cout + co0 + c]);
        }
      }
      xs[rr][c] = xv;
      ds[rr][c] = dv;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kWgRows; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = xs[rr][ty * 4 + i]; b[i] = ds[rr][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    any = true;
    __syncthreads();
  }
  if (!any) return;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + ty * 4 + i, co = co0 + tx * 4 + j;
      if (ci < Cin && co < Cout && acc[i][j] != 0.f) atomicAdd(&dW[((size_t)k * Cin + ci) * Cout + co], acc[i][j]);
    }
}

// ------------------------------------------------------------ UniCrossAtten sampling backward ---
// forward (decoder.cu k_cross_sample): out[c] = gate * S[c], S[c] = sum_corner w_corner * V_corner[c],
// gate = sigmoid((q + qp) . gw + gb), sample position f = sigmoid(ref) * size - 0.5 per axis.
// One warp per query row; channels across lanes.
__global__ void __launch_bounds__(256)
k_cross_sample_bwd(const float* __restrict__ value, int D, int H, int W, int C, const float* __restrict__ ref,
                   const float* __restrict__ query, const float* __restrict__ query_pos,
                   const float* __restrict__ gate_w, float gate_b, int Q, int rows, const float* __restrict__ d_out,
                   float* __restrict__ d_value, float* __restrict__ d_q, float* __restrict__ d_gate_w,
                   float* __restrict__ d_gate_b, float* __restrict__ d_ref) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int b = r / Q;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      float qv = __ldg(&query[(size_t)r * C + c]);
      if (query_pos) qv += __ldg(&query_pos[(size_t)r * C + c]);
      dot = fmaf(qv, __ldg(&gate_w[c]), dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float gate = 1.f / (1.f + expf(-(dot + gate_b)));
    float s3[3], f3[3];
    const int size3[3] = {W, H, D};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      s3[a] = 1.f / (1.f + expf(-__ldg(&ref[(size_t)r * 3 + a])));
      const float g = (s3[a] - 0.5f) * 2.f;
      f3[a] = ((g + 1.f) * (float)size3[a] - 1.f) * 0.5f;
    }
    const float x0f = floorf(f3[0]), y0f = floorf(f3[1]), z0f = floorf(f3[2]);
    const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
    const float tx = f3[0] - x0f, ty = f3[1] - y0f, tz = f3[2] - z0f;
    float wgt[8], gwx[8], gwy[8], gwz[8];
    long long off[8];
    bool inb[8];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const int dz = ci >> 2, dy = (ci >> 1) & 1, dx = ci & 1;
      const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
      inb[ci] = x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D;
      const float wx = dx ? tx : 1.f - tx, wy = dy ? ty : 1.f - ty, wz = dz ? tz : 1.f - tz;
      wgt[ci] = wx * wy * wz;
      gwx[ci] = (dx ? 1.f : -1.f) * wy * wz;
      gwy[ci] = (dy ? 1.f : -1.f) * wx * wz;
      gwz[ci] = (dz ? 1.f : -1.f) * wx * wy;
      off[ci] = inb[ci] ? ((((long long)b * D + z) * H + y) * W + x) * C : 0;
    }
    float dgate = 0.f, dfx = 0.f, dfy = 0.f, dfz = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float go = __ldg(&d_out[(size_t)r * C + c]);
      const float dS = gate * go;
      float S = 0.f;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        if (inb[ci]) {
          const float v = __ldg(&value[off[ci] + c]);
          S = fmaf(wgt[ci], v, S);
          const float t = dS * v;
          dfx = fmaf(gwx[ci], t, dfx);
          dfy = fmaf(gwy[ci], t, dfy);
          dfz = fmaf(gwz[ci], t, dfz);
          if (d_value && wgt[ci] != 0.f) atomicAdd(&d_value[off[ci] + c], wgt[ci] * dS);
        }
      }
      dgate = fmaf(go, S, dgate);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dgate += __shfl_xor_sync(0xffffffffu, dgate, o);
      dfx += __shfl_xor_sync(0xffffffffu, dfx, o);
      dfy += __shfl_xor_sync(0xffffffffu, dfy, o);
      dfz += __shfl_xor_sync(0xffffffffu, dfz, o);
    }
    const float dlogit = dgate * gate * (1.f - gate);
    for (int c = lane; c < C; c += 32) {
      float qv = __ldg(&query[(size_t)r * C + c]);
      if (query_pos) qv += __ldg(&query_pos[(size_t)r * C + c]);
      d_q[(size_t)r * C + c] = dlogit * __ldg(&gate_w[c]);
      atomicAdd(&d_gate_w[c], dlogit * qv);
    }
    if (lane == 0) {
      atomicAdd(d_gate_b, dlogit);
      // f = s * size - 0.5, s = sigmoid(ref)
      d_ref[(size_t)r * 3 + 0] = dfx * (float)W * s3[0] * (1.f - s3[0]);
      d_ref[(size_t)r * 3 + 1] = dfy * (float)H * s3[1] * (1.f - s3[1]);
      d_ref[(size_t)r * 3 + 2] = dfz * (float)D * s3[2] * (1.f - s3[2]);
    }
  }
}

// ------------------------------------------------------------ aligned rotated 3-D IoU ---
// mmdet3d BaseInstance3DBoxes.overlaps(mode='iou') for LiDAR-convention boxes [x, y, z(bottom), dx, dy, dz,
// yaw], pair i of a with pair i of b: rotated-BEV intersection area x height overlap over the union volume.
__global__ void k_iou3d_aligned(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pa[7], pb[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) { pa[j] = a[(size_t)i * 7 + j]; pb[j] = b[(size_t)i * 7 + j]; }
  const float h = fminf(pa[2] + pa[5], pb[2] + pb[5]) - fmaxf(pa[2], pb[2]);
  float iou = 0.f;
  if (h > 0.f) {
    const float dx = pa[0] - pb[0], dy = pa[1] - pb[1];
    const float ra = 0.5f * sqrtf(pa[3] * pa[3] + pa[4] * pa[4]), rb = 0.5f * sqrtf(pb[3] * pb[3] + pb[4] * pb[4]);
    if (dx * dx + dy * dy <= (ra + rb) * (ra + rb)) {
      const float inter = rect_intersection(pa, pb) * h;
      iou = inter / fmaxf(pa[3] * pa[4] * pa[5] + pb[3] * pb[4] * pb[5] - inter, 1e-8f);
    }
  }
  out[i] = iou;
}

// ------------------------------------------------------------ Hungarian matcher ---
// Minimum-cost assignment of every ROW (a ground-truth slot) to a distinct COLUMN (a query) for a
// rows x cols cost block with rows <= cols: the shortest-augmenting-path algorithm (Jonker-Volgenant /
// Kuhn-Munkres with potentials) that scipy.optimize.linear_sum_assignment implements, one problem per
// CTA. Columns are spread over the threads; every step of a path search is one block-wide arg-min.
// Potentials and slack in double precision (scipy works on float64 copies of the float32 costs).
// cost: (n_prob, rows, ld) f32 with element [p][i][j]; row_to_col: (n_prob, rows) int32 out.
// Dynamic shared memory: doubles u[rows+1], v[cols+1], minv[cols+1]; ints p[cols+1], way[cols+1], used[cols+1].
constexpr int kHungThreads = 256;

__global__ void __launch_bounds__(kHungThreads)
k_hungarian(const float* __restrict__ cost, long long prob_stride, int ld, int rows, int cols,
            int32_t* __restrict__ row_to_col) {
  extern __shared__ double hs[];
  double* u = hs;                       // rows + 1
  double* v = u + rows + 1;             // cols + 1
  double* minv = v + cols + 1;          // cols + 1
  int* p = reinterpret_cast<int*>(minv + cols + 1);   // cols + 1: row matched to column j (0 = none)
  int* way = p + cols + 1;
  int* used = way + cols + 1;
  __shared__ double s_best[kHungThreads / 32];
  __shared__ int s_bidx[kHungThreads / 32];
  __shared__ double s_delta;
  __shared__ int s_j1;
  const float* a = cost + (size_t)blockIdx.x * prob_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j <= cols; j += kHungThreads) { v[j] = 0.0; p[j] = 0; }
  for (int i = tid; i <= rows; i += kHungThreads) u[i] = 0.0;
  __syncthreads();
  for (int i = 1; i <= rows; ++i) {
    for (int j = tid; j <= cols; j += kHungThreads) { minv[j] = DBL_MAX; used[j] = 0; }
    if (tid == 0) p[0] = i;
    __syncthreads();
    int j0 = 0;
    while (true) {
      if (tid == 0) used[j0] = 1;
      const int i0 = p[j0];
      __syncthreads();
      const double ui0 = u[i0];
      double best = DBL_MAX;
      int bidx = 0x7fffffff;
      for (int j = 1 + tid; j <= cols; j += kHungThreads) {
        if (!used[j]) {
          const double cur = (double)__ldg(&a[(size_t)(i0 - 1) * ld + (j - 1)]) - ui0 - v[j];
          if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
          const double m = minv[j];
          if (m < best || (m == best && j < bidx)) { best = m; bidx = j; }
        }
      }
      // block arg-min (ties -> lowest column index, the order of a serial scan)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
      if (lane == 0) { s_best[warp] = best; s_bidx[warp] = bidx; }
      __syncthreads();
      if (tid == 0) {
        double bb = s_best[0];
        int bi = s_bidx[0];
        for (int w = 1; w < kHungThreads / 32; ++w)
          if (s_best[w] < bb || (s_best[w] == bb && s_bidx[w] < bi)) { bb = s_best[w]; bi = s_bidx[w]; }
        s_delta = bb;
        s_j1 = bi;
      }
      __syncthreads();
      const double delta = s_delta;
      const int j1 = s_j1;
      for (int j = tid; j <= cols; j += kHungThreads) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }     // distinct rows p[j] for distinct used columns
        else minv[j] -= delta;
      }
      __syncthreads();
      j0 = j1;
      if (p[j0] == 0) break;
    }
    // augment along the path (serial, short)
    if (tid == 0) {
      int j = j0;
      while (j) {
        const int jn = way[j];
        p[j] = p[jn];
        j = jn;
      }
    }
    __syncthreads();
  }
  for (int j = 1 + tid; j <= cols; j += kHungThreads)
    if (p[j] > 0) row_to_col[(size_t)blockIdx.x * rows + (p[j] - 1)] = j - 1;
}

}  // namespace u3d

using namespace u3d;

extern "C" {

int u3d_rulebook_transpose(const int32_t* nbr, int nbr_stride, const int32_t* n_out, int out_cap, int K,
                           int32_t* nbr_t, int t_stride, int in_cap, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(nbr && n_out && nbr_t && K >= 1 && K <= 27 && t_stride >= in_cap && nbr_stride >= out_cap,
                "u3d_rulebook_transpose: bad argument");
  U3D_CUDA(cudaMemsetAsync(nbr_t, 0xff, (size_t)K * t_stride * sizeof(int32_t), st));   // -1
  if (out_cap <= 0) return U3D_OK;
  long long g = ((long long)K * out_cap + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  k_rulebook_transpose<<<(int)g, 256, 0, st>>>(nbr, nbr_stride, n_out, K, nbr_t, t_stride);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

int u3d_spconv_wgrad(const float* x, const float* dy, const int32_t* nbr, int nbr_stride, const int32_t* n_out,
                     int out_cap, int K, int Cin, int Cout, float* dW, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(x && dy && n_out && dW && K >= 1 && K <= 27 && Cin >= 1 && Cout >= 1, "u3d_spconv_wgrad: bad argument");
  U3D_CHECK_ARG(nbr != nullptr, "u3d_spconv_wgrad: pointwise convs have no rulebook; use a dense GEMM");
  U3D_CUDA(cudaMemsetAsync(dW, 0, (size_t)K * Cin * Cout * sizeof(float), st));
  if (out_cap <= 0) return U3D_OK;
  const int tiles = cdiv(Cin, kWgTile) * cdiv(Cout, kWgTile);
  // enough row chunks to fill the machine, at least 256 rows each
  int chunks = cdiv(kNumSMs * 4, tiles * K);
  if (chunks < 1) chunks = 1;
  int rpb = cdiv(out_cap, chunks);
  rpb = (rpb + kWgRows - 1) / kWgRows * kWgRows;
  if (rpb < 256) rpb = 256;
  chunks = cdiv(out_cap, rpb);
  dim3 grid(tiles, K, chunks);
  k_spconv_wgrad<<<grid, 256, 0, st>>>(x, dy, nbr, nbr_stride, n_out, Cin, Cout, rpb, dW);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

int u3d_cross_sample_bwd(const float* value, int B, int D, int H, int W, int C, const float* ref, const float* query,
                         const float* query_pos, const float* gate_w, float gate_b, int Q, const float* d_out,
                         float* d_value, float* d_q, float* d_gate_w, float* d_gate_b, float* d_ref, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(value && ref && query && gate_w && d_out && d_q && d_gate_w && d_gate_b && d_ref,
                "u3d_cross_sample_bwd: null buffer");
  const int rows = B * Q;
  U3D_CUDA(cudaMemsetAsync(d_gate_w, 0, (size_t)C * sizeof(float), st));
  U3D_CUDA(cudaMemsetAsync(d_gate_b, 0, sizeof(float), st));
  if (d_value) U3D_CUDA(cudaMemsetAsync(d_value, 0, (size_t)B * D * H * W * C * sizeof(float), st));
  if (rows <= 0) return U3D_OK;
  int grid = cdiv(rows, 8);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  k_cross_sample_bwd<<<grid, 256, 0, st>>>(value, D, H, W, C, ref, query, query_pos, gate_w, gate_b, Q, rows, d_out,
                                           d_value, d_q, d_gate_w, d_gate_b, d_ref);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

int u3d_iou3d_aligned(const float* a, const float* b, int n, float* out, void* stream) {
  U3D_CHECK_ARG(a && b && out && n >= 0, "u3d_iou3d_aligned: bad argument");
  if (n == 0) return U3D_OK;
  k_iou3d_aligned<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(a, b, n, out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

size_t u3d_hungarian_smem_bytes(int rows, int cols) {
  return (size_t)(rows + 1 + 2 * (cols + 1)) * sizeof(double) + (size_t)3 * (cols + 1) * sizeof(int);
}

int u3d_hungarian(const float* cost, long long prob_stride, int ld, int n_prob, int rows, int cols,
                  int32_t* row_to_col, void* stream) {
  U3D_CHECK_ARG(cost && row_to_col && n_prob >= 0 && rows >= 0 && cols >= 1 && ld >= cols, "u3d_hungarian: bad argument");
  U3D_CHECK_ARG(rows <= cols, "u3d_hungarian: needs rows (%d) <= cols (%d); transpose the problem", rows, cols);
  if (n_prob == 0 || rows == 0) return U3D_OK;
  const size_t smem = u3d_hungarian_smem_bytes(rows, cols);
  U3D_CHECK_ARG(smem <= 200 * 1024, "u3d_hungarian: problem too large for shared memory (rows=%d cols=%d)", rows, cols);
  static int cur_smem = 0;
  U3D_CUDA(ensure_dynamic_smem(k_hungarian, smem, &cur_smem));
  k_hungarian<<<n_prob, kHungThreads, smem, (cudaStream_t)stream>>>(cost, prob_stride, ld, rows, cols, row_to_col);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // extern "C"

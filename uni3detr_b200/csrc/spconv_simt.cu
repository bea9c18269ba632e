// K3 (SIMT flavour) — sparse conv forward as an output-stationary gather-GEMM with fp32
// accumulation and a fused BN(eval)/residual/ReLU epilogue, plus SparseConvTensor.dense().
//
// Reference semantics: SURVEY.md A.3/A.4 (spconv indice_conv + BatchNorm1d(eval) + ReLU,
// SparseBasicBlock identity add), called for every layer of
// projects/mmdet3d_plugin/models/pts_encoder/sparse_encoder_hd.py:106-133.
//
// This kernel is the fp32 parity path (BASELINE config 3 tolerance 1e-3; also bf16 shapes the
// tensor-core kernel does not take). In bf16 every encoder layer - the stem included, its
// reduction dim zero-padded to 16 - runs on the tcgen05 kernel in spconv_tc.cu.
//
// Tile: 64 output rows x BN output channels per CTA, 256 threads, each thread owns a
// 4 x (BN/16) micro-tile. Offsets whose 64-row slice of the rulebook is empty are skipped.
#include "common.cuh"

namespace u3d {

constexpr int BM = 64;
constexpr int BK = 16;

template <typename T, int BN>
__global__ void __launch_bounds__(256)
k_spconv_simt(const T* __restrict__ in, const int32_t* __restrict__ nbr, int nbr_stride,
              const int32_t* __restrict__ n_out_p, int K, const T* __restrict__ w,
              const float* __restrict__ scale, const float* __restrict__ shift,
              const T* __restrict__ residual, int relu, T* __restrict__ out, int Cin, int Cout) {
  constexpr int NJ = BN / 16;
  __shared__ float As[BM][BK + 1];
  __shared__ float Ws[BK][BN];
  __shared__ int s_row[BM];
  __shared__ int s_any;

  const int n_out = *n_out_p;
  const int tiles_m = (n_out + BM - 1) / BM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
    const int m0 = tile * BM;
    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    for (int k = 0; k < K; ++k) {
      __syncthreads();  // previous users of s_row / As / Ws are done
      if (threadIdx.x == 0) s_any = 0;
      __syncthreads();
      if (threadIdx.x < BM) {
        int o = m0 + threadIdx.x;
        int r = -1;
        if (o < n_out) r = nbr ? __ldg(&nbr[(size_t)k * nbr_stride + o]) : o;
        s_row[threadIdx.x] = r;
        if (r >= 0) s_any = 1;
      }
      __syncthreads();
      if (!s_any) continue;
      const T* wk = w + (size_t)k * Cin * Cout;
      for (int c0 = 0; c0 < Cin; c0 += BK) {
        // gather A: BM x BK
#pragma unroll
        for (int i = 0; i < (BM * BK) / 256; ++i) {
          int e = threadIdx.x + i * 256;
          int r = e / BK, c = e % BK;
          int row = s_row[r];
          float v = 0.f;
          if (row >= 0 && c0 + c < Cin) v = to_f32<T>(__ldg(&in[(size_t)row * Cin + c0 + c]));
          As[r][c] = v;
        }
        // W tile: BK x BN
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
          int kk = e / BN, n = e % BN;
          float v = 0.f;
          if (c0 + kk < Cin && n0 + n < Cout) v = to_f32<T>(__ldg(&wk[(size_t)(c0 + kk) * Cout + n0 + n]));
          Ws[kk][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float a[4], b[NJ];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][kk];
#pragma unroll
          for (int j = 0; j < NJ; ++j) b[j] = Ws[kk][tx + 16 * j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int o = m0 + ty * 4 + i;
      if (o >= n_out) continue;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int n = n0 + tx + 16 * j;
        if (n >= Cout) continue;
        float v = acc[i][j];
        if (scale) v = v * __ldg(&scale[n]);
        if (shift) v = v + __ldg(&shift[n]);
        if (residual) v += to_f32<T>(residual[(size_t)o * Cout + n]);
        if (relu) v = fmaxf(v, 0.f);
        out[(size_t)o * Cout + n] = from_f32<T>(v);
      }
    }
  }
}

template <typename T>
static int launch_simt(const void* in, const int32_t* nbr, int nbr_stride, const int32_t* n_out,
                       int out_cap, int K, const void* w, const float* scale, const float* shift,
                       const void* residual, int relu, void* out, int Cin, int Cout,
                       cudaStream_t st) {
  int tiles = cdiv(out_cap, BM);
  if (tiles < 1) tiles = 1;
  int gx = tiles < kNumSMs * 4 ? tiles : kNumSMs * 4;
#define U3D_SIMT_LAUNCH(BN_)                                                                     \
  k_spconv_simt<T, BN_><<<dim3(gx, cdiv(Cout, BN_)), 256, 0, st>>>(                              \
      (const T*)in, nbr, nbr_stride, n_out, K, (const T*)w, scale, shift, (const T*)residual,   \
      relu, (T*)out, Cin, Cout)
  if (Cout <= 16) U3D_SIMT_LAUNCH(16);
  else if (Cout <= 32) U3D_SIMT_LAUNCH(32);
  else U3D_SIMT_LAUNCH(64);
#undef U3D_SIMT_LAUNCH
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

// ------------------------------------------------------------------ dense() ---
template <typename T>
__global__ void __launch_bounds__(256)
k_to_dense(const T* __restrict__ feats, const int32_t* __restrict__ coors,
           const int32_t* __restrict__ n_rows_p, int D, int H, int W, int C, int channels_last,
           T* __restrict__ out) {
  const int n = *n_rows_p;
  const long long total = (long long)n * C;
  const long long spatial = (long long)D * H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int r = (int)(e / C), c = (int)(e % C);
    int4 q = __ldg(reinterpret_cast<const int4*>(coors) + r);
    long long cell = ((long long)q.y * H + q.z) * W + q.w;
    long long dst = channels_last ? ((long long)q.x * spatial + cell) * C + c
                                  : ((long long)q.x * C + c) * spatial + cell;
    out[dst] = feats[e];
  }
}

// NDHWC with 16-byte aligned rows: one 16-byte vector per thread (a row is C*sizeof(T)/16 of them)
__global__ void __launch_bounds__(256)
k_to_dense_ndhwc_vec(const uint4* __restrict__ feats, const int32_t* __restrict__ coors,
                     const int32_t* __restrict__ n_rows_p, int D, int H, int W, int vec_per_row,
                     uint4* __restrict__ out) {
  const int n = *n_rows_p;
  const long long total = (long long)n * vec_per_row;
  const long long spatial = (long long)D * H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / vec_per_row), v = (int)(e - (long long)r * vec_per_row);
    const int4 q = __ldg(reinterpret_cast<const int4*>(coors) + r);
    const long long cell = ((long long)q.y * H + q.z) * W + q.w;
    out[((long long)q.x * spatial + cell) * vec_per_row + v] = __ldg(&feats[e]);
  }
}

}  // namespace u3d

using namespace u3d;

extern "C" int u3d_spconv_fwd(const void* in, const int32_t* nbr, int nbr_stride,
                              const int32_t* n_out, int out_cap, int K, const void* w,
                              const float* scale, const float* shift, const void* residual,
                              int relu, void* out, int Cin, int Cout, int dtype, int impl,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(in && n_out && w && out, "u3d_spconv_fwd: null buffer");
  U3D_CHECK_ARG(K >= 1 && Cin >= 1 && Cout >= 1 && out_cap >= 0, "u3d_spconv_fwd: bad shape");
  U3D_CHECK_ARG(nbr != nullptr || K == 1, "u3d_spconv_fwd: nbr==NULL requires K==1");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_spconv_fwd: bad dtype %d", dtype);
  U3D_CHECK_ARG(impl == 0 || impl == 1, "u3d_spconv_fwd: impl %d (the tensor-core kernel takes packed "
                "weights: u3d_spconv_fwd_packed)", impl);
  if (dtype == U3D_F32)
    return launch_simt<float>(in, nbr, nbr_stride, n_out, out_cap, K, w, scale, shift, residual,
                              relu, out, Cin, Cout, st);
  return launch_simt<__nv_bfloat16>(in, nbr, nbr_stride, n_out, out_cap, K, w, scale, shift,
                                    residual, relu, out, Cin, Cout, st);
}

extern "C" int u3d_sparse_to_dense(const void* feats, const int32_t* coors, const int32_t* n_rows,
                                   int cap, int B, int D, int H, int W, int C, int dtype,
                                   int channels_last, void* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(feats && coors && n_rows && out, "u3d_sparse_to_dense: null buffer");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_sparse_to_dense: bad dtype");
  size_t bytes = (size_t)B * D * H * W * C * dtype_size(dtype);
  U3D_CUDA(cudaMemsetAsync(out, 0, bytes, st));
  long long total = (long long)cap * C;
  int grid = (int)((total + 255) / 256);
  if (grid < 1) grid = 1;
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  const size_t row_bytes = (size_t)C * dtype_size(dtype);
  if (channels_last && row_bytes % 16 == 0 && (((uintptr_t)feats | (uintptr_t)out) & 15) == 0) {
    const int vec_per_row = (int)(row_bytes / 16);
    long long tv = (long long)cap * vec_per_row;
    int gv = (int)((tv + 255) / 256);
    if (gv < 1) gv = 1;
    if (gv > kNumSMs * 16) gv = kNumSMs * 16;
    k_to_dense_ndhwc_vec<<<gv, 256, 0, st>>>((const uint4*)feats, coors, n_rows, D, H, W, vec_per_row,
                                             (uint4*)out);
  } else if (dtype == U3D_F32)
    k_to_dense<float><<<grid, 256, 0, st>>>((const float*)feats, coors, n_rows, D, H, W, C,
                                            channels_last, (float*)out);
  else
    k_to_dense<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)feats, coors, n_rows, D,
                                                    H, W, C, channels_last, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

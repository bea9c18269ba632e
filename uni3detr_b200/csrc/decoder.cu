// K5/K6 — decoder-side kernels: sine position embedding, multi-head self-attention core
// with warp-shuffle online softmax, and the UniCrossAtten sampling block (sigmoid gate +
// trilinear gather from the NDHWC voxel volume).
//
// Reference: projects/mmdet3d_plugin/models/utils/uni3detr_transformer.py
//   get_sine_pos_embed :33-65 (called :180), UniCrossAtten.forward :271-360, and the
//   nn.MultiheadAttention self-attention of mmcv's BaseTransformerLayer.
#include <stdlib.h>
#include "common.cuh"

namespace u3d {

// ------------------------------------------------------------ sine embedding ---
// out[r, c*128 + i] = i even ? sin(v_i) : cos(v_i),  v_i = sigmoid(ref[r,c]) * 2pi / T^(2*floor(i/2)/128)
template <typename T>
__global__ void __launch_bounds__(128)
k_sine_embed(const float* __restrict__ ref, int rows, T* __restrict__ out) {
  const int i = threadIdx.x;  // 0..127
  // dim_t = 10000 ** (2*(i//2)/128), evaluated like torch (fp32 pow); hoisted out of the row loop with its
  // reciprocal-free form kept (v = s * 2pi / dim_t, the reference's operation order)
  const float dim_t = powf(10000.f, (float)(2 * (i / 2)) / 128.f);
  const bool odd = i & 1;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __ldg(&ref[(size_t)r * 3 + c]);
      const float s = 1.f / (1.f + expf(-x));
      const float v = s * 6.283185307179586f / dim_t;      // v in [0, 2*pi]
      float e;
      const float vr = v > 3.14159265358979f ? v - 6.283185307179586f : v;   // [-pi, pi]: the SFU's accurate range
      if (sizeof(T) == 2) e = odd ? __cosf(vr) : __sinf(vr);   // bf16 output: SFU sin / cos (abs. error 2^-21)
      else e = odd ? cosf(v) : sinf(v);
      out[(size_t)r * 384 + c * 128 + i] = from_f32<T>(e);
    }
  }
}

// ------------------------------------------------------ self-attention core ---
// One CTA per (sequence, head, query block). K^T and V of the whole sequence tile are staged in shared
// memory (fp32); each warp processes QPW queries at a time: lane <-> key for the score
// pass (K^T laid out [d][key], conflict-free), warp-shuffle max/sum (online softmax over
// key tiles), lane <-> head-dim for the P.V pass.
constexpr int kHd = 32;
constexpr int kTK = 64;   // keys per smem tile
constexpr int kQPW = 4;   // queries per warp pass
constexpr int kPasses = 4;
constexpr int kMhaWarps = 8;
constexpr int kQPerWarp = kQPW * kPasses;          // 16
constexpr int kQBlock = kMhaWarps * kQPerWarp;     // queries per CTA = 128

template <typename T>
__global__ void __launch_bounds__(kMhaWarps * 32)
k_mha_core(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ldq,
           int ldk, int ldv, int seq_len, int heads, T* __restrict__ out) {
  __shared__ float s_kt[kHd][kTK + 1];
  __shared__ float s_v[kTK][kHd];
  __shared__ float s_q[kMhaWarps][kQPerWarp][kHd];
  __shared__ float s_p[kMhaWarps][kQPW][kTK];

  const int head = blockIdx.y, seq = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qw = blockIdx.x * kQBlock + warp * kQPerWarp;  // first query of this warp
  const size_t row0 = (size_t)seq * seq_len;
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  const int out_ld = heads * kHd;

  // pre-scaled queries of this warp (lane <-> dim)
#pragma unroll
  for (int j = 0; j < kQPerWarp; ++j) {
    int qi = qw + j;
    float x = 0.f;
    if (qi < seq_len) x = to_f32<T>(__ldg(&q[(row0 + qi) * ldq + head * kHd + lane])) * scale;
    s_q[warp][j][lane] = x;
  }
  float m[kQPerWarp], l[kQPerWarp], acc[kQPerWarp];
#pragma unroll
  for (int j = 0; j < kQPerWarp; ++j) { m[j] = -INFINITY; l[j] = 0.f; acc[j] = 0.f; }

  for (int kt = 0; kt < seq_len; kt += kTK) {
    __syncthreads();  // previous tile fully consumed by every warp
    for (int e = threadIdx.x; e < kTK * kHd; e += blockDim.x) {
      int key = e / kHd, d = e % kHd;
      int ki = kt + key;
      float kv = 0.f, vv = 0.f;
      if (ki < seq_len) {
        kv = to_f32<T>(__ldg(&k[(row0 + ki) * ldk + head * kHd + d]));
        vv = to_f32<T>(__ldg(&v[(row0 + ki) * ldv + head * kHd + d]));
      }
      s_kt[d][key] = kv;
      s_v[key][d] = vv;
    }
    __syncthreads();
    const int nkeys = min(kTK, seq_len - kt);
#pragma unroll
    for (int pass = 0; pass < kPasses; ++pass) {
      if (qw + pass * kQPW >= seq_len) break;  // warp-uniform
      // scores: lane <-> key
      float sc[kQPW][kTK / 32];
#pragma unroll
      for (int g = 0; g < kTK / 32; ++g) {
        float a[kQPW];
#pragma unroll
        for (int j = 0; j < kQPW; ++j) a[j] = 0.f;
#pragma unroll
        for (int d = 0; d < kHd; ++d) {
          float kv = s_kt[d][g * 32 + lane];
#pragma unroll
          for (int j = 0; j < kQPW; ++j) a[j] = fmaf(s_q[warp][pass * kQPW + j][d], kv, a[j]);
        }
        bool valid = (g * 32 + lane) < nkeys;
#pragma unroll
        for (int j = 0; j < kQPW; ++j) sc[j][g] = valid ? a[j] : -INFINITY;
      }
      // online softmax update per query
#pragma unroll
      for (int j = 0; j < kQPW; ++j) {
        const int jj = pass * kQPW + j;
        float tmax = sc[j][0];
#pragma unroll
        for (int g = 1; g < kTK / 32; ++g) tmax = fmaxf(tmax, sc[j][g]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        float mnew = fmaxf(m[jj], tmax);
        float corr = __expf(m[jj] - mnew);  // m=-inf on the first tile -> 0
        float psum = 0.f;
#pragma unroll
        for (int g = 0; g < kTK / 32; ++g) {
          float p = __expf(sc[j][g] - mnew);
          s_p[warp][j][g * 32 + lane] = p;
          psum += p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        l[jj] = l[jj] * corr + psum;
        acc[jj] *= corr;
        m[jj] = mnew;
      }
      __syncwarp();
      // P.V: lane <-> dim
      for (int key = 0; key < nkeys; ++key) {
        float vv = s_v[key][lane];
#pragma unroll
        for (int j = 0; j < kQPW; ++j)
          acc[pass * kQPW + j] = fmaf(s_p[warp][j][key], vv, acc[pass * kQPW + j]);
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int j = 0; j < kQPerWarp; ++j) {
    int qi = qw + j;
    if (qi < seq_len) out[(row0 + qi) * out_ld + head * kHd + lane] = from_f32<T>(acc[j] / l[j]);
  }
}

// ----------------------------------------------------- cross-attention sample ---
// one warp per query; lanes stride the channel dimension in vectors of 8 (bf16) / 4 (f32)
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  typedef float4 type;
  static __device__ __forceinline__ void unpack(const float4& v, float* f) {
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ float4 pack(const float* f) {
    return make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  typedef uint4 type;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
k_cross_sample(const T* __restrict__ value, int D, int H, int W, int C,
               const float* __restrict__ ref, const T* __restrict__ query,
               const T* __restrict__ query_pos, const float* __restrict__ gate_w, float gate_b,
               int Q, int rows, T* __restrict__ out) {
  typedef Vec<T> V;
  typedef typename V::type vec_t;
  constexpr int VN = V::N;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int b = r / Q;
    // gate = sigmoid((query + query_pos) . w + bias)
    float dot = 0.f;
    for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
      float qf[VN], pf[VN];
      V::unpack(__ldg(reinterpret_cast<const vec_t*>(query + (size_t)r * C + c0)), qf);
      if (query_pos) {
        V::unpack(__ldg(reinterpret_cast<const vec_t*>(query_pos + (size_t)r * C + c0)), pf);
#pragma unroll
        for (int i = 0; i < VN; ++i) qf[i] = to_f32<T>(from_f32<T>(qf[i] + pf[i]));
      }
#pragma unroll
      for (int i = 0; i < VN; ++i) dot = fmaf(qf[i], __ldg(&gate_w[c0 + i]), dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float gate = 1.f / (1.f + __expf(-(dot + gate_b)));

    // grid_sample(align_corners=False): g = 2*sigmoid(ref)-1 ; i = ((g+1)*S-1)/2
    float fx, fy, fz;
    {
      float sx = 1.f / (1.f + expf(-__ldg(&ref[(size_t)r * 3 + 0])));
      float sy = 1.f / (1.f + expf(-__ldg(&ref[(size_t)r * 3 + 1])));
      float sz = 1.f / (1.f + expf(-__ldg(&ref[(size_t)r * 3 + 2])));
      float gx = (sx - 0.5f) * 2.f, gy = (sy - 0.5f) * 2.f, gz = (sz - 0.5f) * 2.f;
      fx = ((gx + 1.f) * (float)W - 1.f) * 0.5f;
      fy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
      fz = ((gz + 1.f) * (float)D - 1.f) * 0.5f;
    }
    const float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
    const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
    const float tx = fx - x0f, ty = fy - y0f, tz = fz - z0f;
    float wgt[8];
    long long off[8];
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx) {
      int dz = cidx >> 2, dy = (cidx >> 1) & 1, dx = cidx & 1;
      int x = x0 + dx, y = y0 + dy, z = z0 + dz;
      bool inb = x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D;
      float wv = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
      wgt[cidx] = inb ? wv : 0.f;
      off[cidx] = inb ? ((((long long)b * D + z) * H + y) * W + x) * C : 0;
    }
    for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
      float acc[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] = 0.f;
#pragma unroll
      for (int cidx = 0; cidx < 8; ++cidx) {
        if (wgt[cidx] != 0.f) {
          float f[VN];
          V::unpack(__ldg(reinterpret_cast<const vec_t*>(value + off[cidx] + c0)), f);
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[i] = fmaf(wgt[cidx], f[i], acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] *= gate;
      *reinterpret_cast<vec_t*>(out + (size_t)r * C + c0) = V::pack(acc);
    }
  }
}

// ------------------------------------------------- fused residual add + LayerNorm ---
// out[r,:] = act( LN(a[r,:] (+ b[r,:]) (+ c[r,:])) * gamma + beta ); one warp per row, the row stays
// in registers between the statistics and the normalisation (one HBM read of each input, one write).
template <typename T>
__global__ void __launch_bounds__(256)
k_add_layernorm(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                const T* __restrict__ gamma, const T* __restrict__ beta, float eps, int rows, int C,
                int relu, T* __restrict__ out) {
  typedef Vec<T> V;
  typedef typename V::type vec_t;
  constexpr int VN = V::N;
  constexpr int kMaxVec = 4;                       // C <= 32 * VN * 4
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int nvec = C / (32 * VN);
  float x[kMaxVec][VN];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    if (j < nvec) {
      const size_t off = (size_t)r * C + (size_t)(j * 32 + lane) * VN;
      V::unpack(__ldg(reinterpret_cast<const vec_t*>(a + off)), x[j]);
      if (b) {
        float t[VN];
        V::unpack(__ldg(reinterpret_cast<const vec_t*>(b + off)), t);
#pragma unroll
        for (int i = 0; i < VN; ++i) x[j][i] += t[i];
      }
      if (c) {
        float t[VN];
        V::unpack(__ldg(reinterpret_cast<const vec_t*>(c + off)), t);
#pragma unroll
        for (int i = 0; i < VN; ++i) x[j][i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < VN; ++i) sum += x[j][i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec)
#pragma unroll
      for (int i = 0; i < VN; ++i) { const float d = x[j][i] - mean; var += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)C + eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    if (j < nvec) {
      const int col = (j * 32 + lane) * VN;
      float g[VN], be[VN], y[VN];
      V::unpack(__ldg(reinterpret_cast<const vec_t*>(gamma + col)), g);
      V::unpack(__ldg(reinterpret_cast<const vec_t*>(beta + col)), be);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        y[i] = (x[j][i] - mean) * rstd * g[i] + be[i];
        if (relu) y[i] = fmaxf(y[i], 0.f);
      }
      *reinterpret_cast<vec_t*>(out + (size_t)r * C + col) = V::pack(y);
    }
  }
}

// out = sum_i act_i(x_i + bias_i) over up to three (rows, C) operands, one pass: the level merge of
// SECOND3DFPN (necks/second3d_fpn.py:125-126) with the folded-BN bias and ReLU of the two
// ConvTranspose3d branches applied on the fly (the transposed convs run without an epilogue).
template <typename T>
__global__ void __launch_bounds__(256)
k_bias_act_sum(const T* __restrict__ x0, const T* __restrict__ x1, const T* __restrict__ x2,
               const float* __restrict__ b0, const float* __restrict__ b1, const float* __restrict__ b2,
               int relu_mask, long long nvec, int C, T* __restrict__ out) {
  typedef Vec<T> V;
  typedef typename V::type vec_t;
  constexpr int VN = V::N;
  const T* xs[3] = {x0, x1, x2};
  const float* bs[3] = {b0, b1, b2};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)((i * VN) % C);
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (xs[j] == nullptr) continue;
      float v[VN];
      V::unpack(__ldg(reinterpret_cast<const vec_t*>(xs[j]) + i), v);
      if (bs[j]) {
#pragma unroll
        for (int e = 0; e < VN; e += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bs[j] + ch + e));
          v[e] += bb.x; v[e + 1] += bb.y; v[e + 2] += bb.z; v[e + 3] += bb.w;
        }
      }
      if ((relu_mask >> j) & 1) {
#pragma unroll
        for (int e = 0; e < VN; ++e) v[e] = fmaxf(v[e], 0.f);
      }
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] += v[e];
    }
    reinterpret_cast<vec_t*>(out)[i] = V::pack(acc);
  }
}

}  // namespace u3d

using namespace u3d;

extern "C" int u3d_add_layernorm(const void* a, const void* b, const void* c, const void* gamma,
                                 const void* beta, float eps, int rows, int C, int relu, void* out,
                                 int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(a && gamma && beta && out && rows >= 0, "u3d_add_layernorm: bad argument");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_add_layernorm: bad dtype");
  const int vn = dtype == U3D_BF16 ? 8 : 4;
  U3D_CHECK_ARG(C >= 32 * vn && C % (32 * vn) == 0 && C <= 32 * vn * 4,
                "u3d_add_layernorm: C=%d must be a multiple of %d and <= %d", C, 32 * vn, 32 * vn * 4);
  if (rows == 0) return U3D_OK;
  const int grid = cdiv(rows, 8);
  if (dtype == U3D_F32)
    k_add_layernorm<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, (const float*)c,
                                                 (const float*)gamma, (const float*)beta, eps, rows, C, relu,
                                                 (float*)out);
  else
    k_add_layernorm<__nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)c, (const __nv_bfloat16*)gamma,
        (const __nv_bfloat16*)beta, eps, rows, C, relu, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_bias_act_sum(const void* x0, const void* x1, const void* x2, const float* b0,
                                const float* b1, const float* b2, int relu_mask, long long rows, int C,
                                int dtype, void* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(x0 && out && rows >= 0 && C >= 1, "u3d_bias_act_sum: bad argument");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_bias_act_sum: bad dtype");
  const int vn = dtype == U3D_BF16 ? 8 : 4;
  U3D_CHECK_ARG(C % vn == 0, "u3d_bias_act_sum: C=%d must be a multiple of %d", C, vn);
  U3D_CHECK_ARG((((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out | (uintptr_t)b0 |
                  (uintptr_t)b1 | (uintptr_t)b2) & 15) == 0,
                "u3d_bias_act_sum: buffers must be 16-byte aligned");
  if (rows == 0) return U3D_OK;
  const long long nvec = rows * C / vn;
  long long g = (nvec + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  if (dtype == U3D_F32)
    k_bias_act_sum<float><<<(int)g, 256, 0, st>>>((const float*)x0, (const float*)x1, (const float*)x2, b0, b1,
                                                  b2, relu_mask, nvec, C, (float*)out);
  else
    k_bias_act_sum<__nv_bfloat16><<<(int)g, 256, 0, st>>>(
        (const __nv_bfloat16*)x0, (const __nv_bfloat16*)x1, (const __nv_bfloat16*)x2, b0, b1, b2, relu_mask,
        nvec, C, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

// x = hi + lo with hi the value rounded to TF32 (10 explicit mantissa bits, round to nearest on the 13
// dropped bits) and lo the remainder: operands of the 3-pass "3xTF32" dense convolutions (second_3d.py).
__global__ void k_split_tf32(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    uint32_t b = __float_as_uint(v);
    const uint32_t e = b & 0x7f800000u;
    float h = v;
    if (e != 0x7f800000u) h = __uint_as_float((b + 0x1000u) & 0xffffe000u);   // finite: round half up in magnitude
    hi[i] = h;
    lo[i] = v - h;
  }
}

extern "C" int u3d_split_tf32(const float* x, long long n, float* hi, float* lo, void* stream) {
  U3D_CHECK_ARG(x && hi && lo && n >= 0, "u3d_split_tf32: bad argument");
  if (n == 0) return U3D_OK;
  long long g = (n + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  k_split_tf32<<<(int)g, 256, 0, (cudaStream_t)stream>>>(x, n, hi, lo);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_sine_embed(const float* ref, int rows, void* out, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(ref && out && rows >= 0, "u3d_sine_embed: bad argument");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_sine_embed: bad dtype");
  if (rows == 0) return U3D_OK;
  int grid = rows < kNumSMs * 16 ? rows : kNumSMs * 16;
  if (dtype == U3D_F32) k_sine_embed<float><<<grid, 128, 0, st>>>(ref, rows, (float*)out);
  else k_sine_embed<__nv_bfloat16><<<grid, 128, 0, st>>>(ref, rows, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_mha_core(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv,
                            int n_seq, int seq_len, int heads, void* out, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(q && k && v && out, "u3d_mha_core: null buffer");
  U3D_CHECK_ARG(n_seq >= 1 && seq_len >= 1 && heads >= 1 && ldq >= heads * kHd && ldk >= heads * kHd &&
                    ldv >= heads * kHd,
                "u3d_mha_core: bad shape");
  U3D_CHECK_ARG(n_seq <= 65535 && heads <= 65535, "u3d_mha_core: grid too large");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_mha_core: bad dtype");
  // bf16: tcgen05 kernel (mha_tc.cu); fp32 (parity mode) and unsupported shapes: SIMT kernel below
  if (dtype == U3D_BF16 && getenv("U3D_MHA_SIMT") == nullptr) {
    // v2 (mha_tc2.cu: persistent, three warpgroups, P in TMEM, TMA loads) unless U3D_MHA_V1=1 asks for the first kernel
    if (getenv("U3D_MHA_V1") == nullptr && mha_tc2_supported(seq_len, ldq, ldk, ldv, q, k, v, out))
      return mha_core_tc2(q, k, v, ldq, ldk, ldv, n_seq, seq_len, heads, out, st);
    if (mha_tc_supported(seq_len, ldq, ldk, ldv, q, k, v, out))
      return mha_core_tc(q, k, v, ldq, ldk, ldv, n_seq, seq_len, heads, out, st);
  }
  dim3 grid(cdiv(seq_len, kQBlock), heads, n_seq);
  if (dtype == U3D_F32)
    k_mha_core<float><<<grid, kMhaWarps * 32, 0, st>>>((const float*)q, (const float*)k,
                                                       (const float*)v, ldq, ldk, ldv, seq_len, heads,
                                                       (float*)out);
  else
    k_mha_core<__nv_bfloat16><<<grid, kMhaWarps * 32, 0, st>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldq, ldk, ldv,
        seq_len, heads, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_cross_sample(const void* value, int B, int D, int H, int W, int C,
                                const float* ref, const void* query, const void* query_pos,
                                const float* gate_w, float gate_b, int Q, void* out, int dtype,
                                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(value && ref && query && gate_w && out, "u3d_cross_sample: null buffer");
  U3D_CHECK_ARG(dtype == U3D_F32 || dtype == U3D_BF16, "u3d_cross_sample: bad dtype");
  int vn = dtype == U3D_BF16 ? 8 : 4;
  U3D_CHECK_ARG(C % vn == 0, "u3d_cross_sample: C=%d must be a multiple of %d", C, vn);
  int rows = B * Q;
  if (rows == 0) return U3D_OK;
  int grid = cdiv(rows, 8);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (dtype == U3D_F32)
    k_cross_sample<float><<<grid, 256, 0, st>>>((const float*)value, D, H, W, C, ref,
                                                (const float*)query, (const float*)query_pos,
                                                gate_w, gate_b, Q, rows, (float*)out);
  else
    k_cross_sample<__nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)value, D, H, W, C, ref, (const __nv_bfloat16*)query,
        (const __nv_bfloat16*)query_pos, gate_w, gate_b, Q, rows, (__nv_bfloat16*)out);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

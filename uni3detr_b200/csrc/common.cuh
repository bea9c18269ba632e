// Shared device/host helpers for libu3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/u3d.h"

#ifndef __CUDA_ARCH__
#define U3D_HOST_ONLY 1
#endif

namespace u3d {

void set_error(const char* fmt, ...);

#define U3D_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::u3d::set_error(__VA_ARGS__);      \
      return U3D_EINVAL;                  \
    }                                     \
  } while (0)

#define U3D_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::u3d::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                    \
      return U3D_ECUDA;                                                        \
    }                                                                          \
  } while (0)

// every kernel launch of this library goes through U3D_LAUNCH_CHECK: it also feeds the
// launch counter behind u3d_launch_count() (bench.py reports it as `gpu_launches`).
void count_launch();
#define U3D_LAUNCH_CHECK()           \
  do {                               \
    ::u3d::count_launch();           \
    U3D_CUDA(cudaGetLastError());    \
  } while (0)

constexpr int kNumSMs = 148;  // B200
constexpr int kSlotEmpty = 0x7f7f7f7f;  // memset(0x7f) pattern; > any point index

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- VoxelMap ----
// One uint2 per 32 linear cells: .x occupancy bits, .y exclusive prefix popcount.
struct MapView {
  const uint2* words;
  const int32_t* perm;  // rank -> row, or nullptr (identity)
};

__device__ __forceinline__ int map_rank_at(const uint2* __restrict__ words, uint32_t lin) {
  // number of set bits strictly below `lin` (valid for lin == total cells too: pad word)
  uint2 w = __ldg(&words[lin >> 5]);
  uint32_t bit = lin & 31u;
  return (int)w.y + __popc(w.x & ((1u << bit) - 1u));
}

__device__ __forceinline__ int map_lookup(const uint2* __restrict__ words,
                                          const int32_t* __restrict__ perm, uint32_t lin) {
  uint2 w = __ldg(&words[lin >> 5]);
  uint32_t bit = lin & 31u;
  if (!((w.x >> bit) & 1u)) return -1;
  int r = (int)w.y + __popc(w.x & ((1u << bit) - 1u));
  return perm ? __ldg(&perm[r]) : r;
}

// set a bit with warp-aggregated atomics: lanes hitting the same word merge their bits
// and one lane issues the atomicOr.
__device__ __forceinline__ void map_set_bit_aggregated(uint2* words, bool valid, uint32_t lin) {
  uint32_t word = valid ? (lin >> 5) : 0xffffffffu;
  uint32_t bits = valid ? (1u << (lin & 31u)) : 0u;
  // must be called by all 32 lanes of the warp (callers keep the warp converged)
  unsigned peers = __match_any_sync(0xffffffffu, word);
  int leader = __ffs(peers) - 1;
  int lane = threadIdx.x & 31;
  uint32_t merged = __reduce_or_sync(peers, bits);
  if (valid && lane == leader) atomicOr(&words[word].x, merged);
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix, `total`
// receives the block sum. smem must hold 33 ints. All threads must call.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  total = smem[32];
  return smem[warp] + inc - v;
}

// -------------------------------------------------------------- dtype utils ---
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the requirement grows: launches (and
// CUDA-graph captures) after the first one of a kernel/size make no attribute call at all
template <typename Kern>
static inline cudaError_t ensure_dynamic_smem(Kern kern, size_t bytes, int* cur) {
  if ((int)bytes <= *cur) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) *cur = (int)bytes;
  return e;
}

static inline size_t dtype_size(int dtype) { return dtype == U3D_BF16 ? 2 : 4; }

// internal launchers shared between translation units
int voxmap_scan(uint2* map, size_t words, int32_t* scratch, int32_t* total_out, cudaStream_t st);

int spconv_fwd_tc(const void* in, const int32_t* nbr, int nbr_stride,
                  const uint32_t* tile_mask, const int32_t* slot_row, const int32_t* n_out, int out_cap,
                  int K, const void* w,
                  const float* scale, const float* shift, const void* residual, int relu, void* out,
                  int Cin, int Cout, cudaStream_t st);
bool spconv_tc_supported(int Cin, int Cout, int dtype);
int spconv_fwd_tn(const void* in, const int32_t* nbr, int nbr_stride,
                  const uint32_t* tile_mask, const int32_t* slot_row, const int32_t* n_out, int out_cap,
                  int K, const void* w,
                  const float* scale, const float* shift, const void* residual, int relu, void* out,
                  int Cin, int Cout, cudaStream_t st);
int spconv_fwd_tn_ex(const void* in, const int32_t* nbr, int nbr_stride, const uint32_t* tile_mask,
                     const int32_t* slot_row, const int32_t* n_out, int out_cap, int K, const void* wpk,
                     const float* scale, const float* shift, const void* residual, int relu, void* out, int Cin,
                     int Cout, int x3, int in_ld, int out_ld, int cout_off, int cout_total, cudaStream_t st);
bool spconv_tn_supported(int Cin, int Cout, const int32_t* nbr);
bool mha_tc_supported(int seq_len, int ldq, int ldk, int ldv, const void* q, const void* k, const void* v,
                      const void* out);
bool mha_tc2_supported(int seq_len, int ldq, int ldk, int ldv, const void* q, const void* k, const void* v,
                       const void* out);
int mha_core_tc2(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int n_seq,
                 int seq_len, int heads, void* out, cudaStream_t st);
int mha_core_tc(const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int n_seq,
                int seq_len, int heads, void* out, cudaStream_t st);

}  // namespace u3d

// VoxelMap prefix scan + library-wide error plumbing.
//
// The VoxelMap (include/u3d.h) is the coordinate index every geometry kernel shares:
// occupancy bits + exclusive prefix popcount per 32 linear cells. Building it is
//   memset -> atomicOr of the occupied cells -> this scan.
// The scan is HBM/L2 streaming work: 8 B read + 4 B written per word, three launches
// (chunk partial sums, one-CTA scan of the partials, apply).
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace u3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanPasses = 16;                          // words per lane
constexpr int kScanChunk = kScanThreads * kScanPasses;   // 4096 words per CTA
constexpr int kWarpSpan = 32 * kScanPasses;              // 512 consecutive words per warp

__global__ void __launch_bounds__(kScanThreads)
k_scan_partials(const uint2* __restrict__ map, size_t words, int32_t* __restrict__ partial) {
  __shared__ int s_w[kScanWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)warp * kWarpSpan;
  int s = 0;
#pragma unroll
  for (int p = 0; p < kScanPasses; ++p) {
    size_t i = base + p * 32 + lane;
    if (i < words) s += __popc(__ldg(&map[i]).x);
  }
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) s_w[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) t += s_w[w];
    partial[blockIdx.x] = t;
  }
}

// exclusive scan of `n` partial sums in place; one CTA
__global__ void __launch_bounds__(1024)
k_scan_partials_scan(int32_t* __restrict__ partial, int n, int32_t* __restrict__ total_out) {
  __shared__ int smem[33];
  int running = 0;
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < n ? partial[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, smem, total);
    if (i < n) partial[i] = running + ex;
    running += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = running;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(uint2* __restrict__ map, size_t words, const int32_t* __restrict__ partial) {
  __shared__ int s_w[kScanWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)warp * kWarpSpan;
  uint32_t bits[kScanPasses];
  int s = 0;
#pragma unroll
  for (int p = 0; p < kScanPasses; ++p) {
    size_t i = base + p * 32 + lane;
    bits[p] = i < words ? map[i].x : 0u;
    s += __popc(bits[p]);
  }
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) s_w[warp] = s;
  __syncthreads();
  int running = partial[blockIdx.x];
  for (int w = 0; w < warp; ++w) running += s_w[w];
#pragma unroll
  for (int p = 0; p < kScanPasses; ++p) {
    int v = __popc(bits[p]);
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    size_t i = base + p * 32 + lane;
    if (i < words) map[i].y = (uint32_t)(running + inc - v);
    running += __shfl_sync(0xffffffffu, inc, 31);
  }
}

int voxmap_scan(uint2* map, size_t words, int32_t* scratch, int32_t* total_out, cudaStream_t st) {
  int nchunks = cdiv((long long)words, kScanChunk);
  k_scan_partials<<<nchunks, kScanThreads, 0, st>>>(map, words, scratch);
  U3D_LAUNCH_CHECK();
  k_scan_partials_scan<<<1, 1024, 0, st>>>(scratch, nchunks, total_out);
  U3D_LAUNCH_CHECK();
  k_scan_apply<<<nchunks, kScanThreads, 0, st>>>(map, words, scratch);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

}  // namespace u3d

extern "C" {

const char* u3d_last_error(void) { return u3d::g_err; }
int u3d_version(void) { return 200; }
#ifndef U3D_BUILD_ID_STR
#define U3D_BUILD_ID_STR "unversioned-----"
#endif
// sha256 prefix of the sources this binary was built from (uni3detr_b200/_lib.py source_hash());
// the loader refuses a library whose id differs from the sources next to it
const char* u3d_build_id(void) { return "U3D_BUILD_ID=" U3D_BUILD_ID_STR; }
unsigned long long u3d_launch_count(void) { return u3d::g_launches.load(std::memory_order_relaxed); }

size_t u3d_voxmap_words(int B, int D, int H, int W) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  unsigned long long cells = (unsigned long long)B * D * H * W;
  if (cells >= 0xffffffc0ull) return 0;  // 32-bit linear cell index
  return (size_t)(cells / 32 + 2);       // +pad so rank_at(total cells) is addressable
}

size_t u3d_scan_scratch_ints(size_t words) {
  return (words + u3d::kScanChunk - 1) / u3d::kScanChunk + 8;
}

}  // extern "C"

// Micro-benchmark: issue cost of back-to-back tcgen05.mma (kind::f16, bf16 operands, K = 16 per instruction) as a
// function of the instruction shape and of where the A operand lives (shared memory "SS" / tensor memory "TS").
// One thread per CTA issues `n_inst` MMAs over zeroed operands (4 K-steps of a 128-byte-swizzled K-major tile, cycling
// over two tiles), commits, waits on the mbarrier; cycles = clock64 delta / n_inst. Run on 1 CTA and on 148.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/ubench/mma_cost scripts/ubench/mma_cost.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {   // K-major, SWIZZLE_128B, SBO = 1024 B
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int kTS>
__global__ void __launch_bounds__(128, 1) k_cost(int M, int N, int n_inst, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (2 * 16384 + 2 * 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_s = smem_u32(base), b_s = smem_u32(base) + 2 * 16384;
    long long t0 = clock64();
    for (int i = 0; i < n_inst; ++i) {
      const int tile = (i >> 2) & 1, kk = i & 3;
      const uint64_t a_desc = desc128(a_s + tile * 16384) + (uint64_t)(kk * 2);
      const uint64_t b_desc = desc128(b_s + tile * 32768) + (uint64_t)(kk * 2);
      const uint32_t d = tmem + (uint32_t)(((i >> 3) & 1) * 256);
      if (kTS == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
      } else if (kTS == 1) {
        const uint32_t a_t = tmem + 256u - 0u + 0u;   // A columns: reuse columns [256+..] region's tail? keep separate: cols 480..511
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d & ~256u), "r"(tmem + 480u + (uint32_t)(kk * 8)),
                     "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
        (void)a_t;
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    out[2 * blockIdx.x] = t2 - t0;
    out[2 * blockIdx.x + 1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// Functional check of the accumulator layout of tcgen05.mma.ws M = 64 / 32 (and of the plain M = 64 form): two single
// K = 16 MMAs over integer operands; run 0 gives D[m][n] = m + 1, run 1 gives D[m][n] = n. The whole 128-lane x 256-
// column accumulator window is dumped so the host can print which (lane, column) holds which (m, n).
__global__ void __launch_bounds__(128, 1) k_layout(int M, int N, int ws, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(base);            // 2 runs x 16 KB
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(base + 32768);    // 2 runs x 32 KB
  for (int i = threadIdx.x; i < (2 * 16384 + 2 * 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int r = threadIdx.x; r < 256; r += blockDim.x) {
    const int off = (r * 128 + ((r & 7) << 4)) / 2;      // element (r, k = 0) of the swizzled K-major tile
    if (r < 128) { A[off] = __float2bfloat16((float)(r + 1)); A[8192 + off] = __float2bfloat16(1.f); }
    B[off] = __float2bfloat16(1.f);
    B[16384 + off] = __float2bfloat16((float)r);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  // clear both windows first (plain M = 128 N = 256 MMA over the zero K-step 1), then the MMAs under test
  if (threadIdx.x == 0) {
    const uint32_t idz = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int run = 0; run < 2; ++run) {
      const uint64_t a0 = desc128(smem_u32(base) + run * 16384), b0 = desc128(smem_u32(base) + 32768 + run * 32768);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + run * 256), "l"(a0 + 2), "l"(b0 + 2), "r"(idz), "r"(0u) : "memory");
      if (ws) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + run * 256), "l"(a0), "l"(b0), "r"(idesc), "r"(1u) : "memory");
      else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + run * 256), "l"(a0), "l"(b0), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < 512; c += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(size_t)(warp * 32 + lane) * 512 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static void layout_report(int M, int N, int ws, float* d_dump, size_t smem) {
  k_layout<<<1, 128, smem>>>(M, N, ws, d_dump);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("layout M=%d N=%d ws=%d: %s\n", M, N, ws, cudaGetErrorString(e)); exit(1); }
  static float h[128 * 512];
  cudaMemcpy(h, d_dump, sizeof(h), cudaMemcpyDeviceToHost);
  printf("layout %s M=%d N=%d: (lane, col) -> (m, n); columns used:", ws ? "WS" : "plain", M, N);
  int maxcol = -1;
  for (int l = 0; l < 128; ++l) for (int c = 0; c < 256; ++c) if (h[l * 512 + c] != 0.f && c > maxcol) maxcol = c;
  printf(" %d\n", maxcol + 1);
  const int lanes[] = {0, 1, 15, 16, 31, 32, 33, 48, 63, 64, 65, 96, 127};
  const int cols[] = {0, 1, 63, 64, 127, 128, 255};
  for (int l : lanes) {
    printf("  lane %3d:", l);
    for (int c : cols) {
      const float m1 = h[l * 512 + c], n = h[l * 512 + 256 + c];
      if (m1 == 0.f) printf("  c%-3d -> -      ", c); else printf("  c%-3d -> (%2d,%3d)", c, (int)m1 - 1, (int)n);
    }
    printf("\n");
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 2 * 148 * sizeof(long long));
  float* d_dump;
  cudaMalloc(&d_dump, 128 * 512 * sizeof(float));
  const size_t smem = 2 * 16384 + 2 * 32768 + 1024;
  cudaFuncSetAttribute(k_cost<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_cost<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_cost<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  layout_report(128, 256, 0, d_dump, smem);
  layout_report(64, 256, 0, d_dump, smem);
  layout_report(64, 256, 1, d_dump, smem);
  layout_report(32, 256, 1, d_dump, smem);
  layout_report(64, 128, 1, d_dump, smem);
  const int n_inst = 4096;
  const int Ms[] = {128, 64, 32};
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  const char* names[] = {"SS", "TS", "WS"};
  printf("mode M N grid cycles_per_mma(max over CTAs) issue_cycles_per_mma macs_per_cycle\n");
  for (int ts = 0; ts < 3; ++ts)
    for (int M : Ms)
      for (int N : Ns)
        for (int grid : {1, 148}) {
          if (ts == 1 && N > 224) continue;   // TS: D uses columns [0, N), A columns 480..511
          if (ts != 2 && M == 32) continue;
          if (ts == 2 && N != 64 && N != 128 && N != 256) continue;
          if (M == 128 && (N % 16)) continue;
          for (int rep = 0; rep < 2; ++rep) {
            if (ts == 2) k_cost<2><<<grid, 128, smem>>>(M, N, n_inst, d_out);
            else if (ts) k_cost<1><<<grid, 128, smem>>>(M, N, n_inst, d_out);
            else k_cost<0><<<grid, 128, smem>>>(M, N, n_inst, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s M=%d N=%d: %s\n", names[ts], M, N, cudaGetErrorString(e)); return 1; }
          }
          long long h[2 * 148];
          cudaMemcpy(h, d_out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0, is = 0;
          for (int i = 0; i < grid; ++i) { if (h[2 * i] > mx) mx = h[2 * i]; if (h[2 * i + 1] > is) is = h[2 * i + 1]; }
          const double c = (double)mx / n_inst;
          printf("%s %3d %3d %3d %8.1f %8.1f %8.0f\n", names[ts], M, N, grid, c, (double)is / n_inst, (double)M * N * 16 / c);
        }
  return 0;
}

// Tile scheduling for the tensor-core sparse conv: group output rows with similar neighbour
// patterns into the same 256-row tile.
//
// The conv (spconv_tn.cu) multiplies a whole tile for every kernel offset that feeds at least one of
// its rows. In the natural row order (ascending linear index) a tile mixes floor, wall and object
// surface voxels, the union of their neighbour sets is nearly all 27 offsets, and only 6 % (stage
// 0) .. 72 % (stage 3) of the multiplied (row, offset) slots carry a real pair. Rows are therefore
// bucketed by a 12-bit signature of their 27-bit neighbour mask - "does the (kz,ky) line have any
// neighbour" (9 bits) and "does the kx column have any" (3 bits) - which on the synthetic SUN RGB-D
// scenes raises the useful fraction to 45 % .. 91 % (better than sorting by the full mask, which
// splits similar rows over many tiny runs). Nothing of this exists in the reference (spconv
// multiplies pair lists); it is a pure scheduling permutation:
//   slot_row[s]            output row processed in slot s (a permutation of [0, n_out))
//   nbr_sorted[k][s]       = nbr[k][slot_row[s]]   (input rows keep their natural numbering)
//   tile_mask_sorted[t]    offsets that feed slots [128t, 128t+128)
// The conv writes (and reads the residual of) row slot_row[s], so feature matrices stay in the
// reference's row order and every row's accumulation order (ascending k) is unchanged: results are
// identical to the natural-order run (up to the sign of an exact zero).
//
// One counting-sort pass with device-side row count: per-block shared-memory histograms (warp
// aggregated), a 4096-bin scan, block-wise range reservation, ranks from a second shared-memory
// pass. The order inside a bucket depends on block scheduling (atomics); the conv result does not.
#include "common.cuh"

namespace u3d {

constexpr int kKeyBins = 4096;
constexpr int kSortThreads = 256;

__device__ __forceinline__ uint32_t tile_key_of_mask(uint32_t m) {
  uint32_t key = 0u;
#pragma unroll
  for (int l = 0; l < 9; ++l) key |= (((m >> (3 * l)) & 7u) ? 1u : 0u) << l;
  constexpr uint32_t kCol0 = 0x1249249u;   // bits 0,3,...,24: kx = 0 of every (kz,ky) line
#pragma unroll
  for (int c = 0; c < 3; ++c) key |= ((m & (kCol0 << c)) ? 1u : 0u) << (9 + c);
  return key;
}

__global__ void __launch_bounds__(256)
k_row_key(const int32_t* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ n_p, int K,
          int32_t* __restrict__ row_key) {
  const int n = *n_p;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    uint32_t m = 0u;
    for (int k = 0; k < K; ++k) m |= (__ldg(&nbr[(size_t)k * nbr_stride + o]) >= 0 ? 1u : 0u) << k;
    row_key[o] = (int32_t)tile_key_of_mask(m);
  }
}

// rows [start, end) of this block; every block takes a multiple of kSortThreads rows
__device__ __forceinline__ void block_range(int n, int& start, int& end) {
  int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  per = (per + kSortThreads - 1) / kSortThreads * kSortThreads;
  const long long s = (long long)blockIdx.x * per;
  start = s < n ? (int)s : n;
  end = s + per < n ? (int)(s + per) : n;
}

// shared-memory histogram of the block's keys, one atomic per distinct key per warp
__device__ __forceinline__ void block_histogram(const int32_t* __restrict__ row_key, int start, int end,
                                                int* s_hist) {
  const int lane = threadIdx.x & 31;
  for (int base = start; base < end; base += kSortThreads) {
    const int r = base + threadIdx.x;
    const bool valid = r < end;
    const int key = valid ? __ldg(&row_key[r]) : (kKeyBins + lane);     // distinct dummy per lane
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[key], __popc(peers));
  }
}

__global__ void __launch_bounds__(kSortThreads)
k_key_hist(const int32_t* __restrict__ row_key, const int32_t* __restrict__ n_p, int32_t* __restrict__ hist) {
  __shared__ int s_hist[kKeyBins];
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads) s_hist[b] = 0;
  __syncthreads();
  int start, end;
  block_range(*n_p, start, end);
  block_histogram(row_key, start, end, s_hist);
  __syncthreads();
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads)
    if (s_hist[b]) atomicAdd(&hist[b], s_hist[b]);
}

// exclusive scan of the 4096 bins (one block of 1024 threads, 4 bins each)
__global__ void __launch_bounds__(1024)
k_key_scan(const int32_t* __restrict__ hist, int32_t* __restrict__ cursor) {
  __shared__ int smem[33];
  const int b0 = threadIdx.x * 4;
  int v[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = hist[b0 + i]; sum += v[i]; }
  int total;
  int ex = block_exclusive_scan(sum, smem, total);
#pragma unroll
  for (int i = 0; i < 4; ++i) { cursor[b0 + i] = ex; ex += v[i]; }
}

__global__ void __launch_bounds__(kSortThreads)
k_key_scatter(const int32_t* __restrict__ row_key, const int32_t* __restrict__ n_p,
              int32_t* __restrict__ cursor, int32_t* __restrict__ slot_row) {
  __shared__ int s_cnt[kKeyBins];    // block histogram, then the block's running rank per key
  __shared__ int s_base[kKeyBins];   // first slot of the block's range of each key
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads) s_cnt[b] = 0;
  __syncthreads();
  int start, end;
  block_range(*n_p, start, end);
  block_histogram(row_key, start, end, s_cnt);
  __syncthreads();
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads) {
    const int c = s_cnt[b];
    if (c) s_base[b] = atomicAdd(&cursor[b], c);   // reserve c consecutive slots of bucket b
    s_cnt[b] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int base = start; base < end; base += kSortThreads) {
    const int r = base + threadIdx.x;
    const bool valid = r < end;
    const int key = valid ? __ldg(&row_key[r]) : (kKeyBins + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    int first = 0;
    if (valid && lane == leader) first = atomicAdd(&s_cnt[key], __popc(peers));
    first = __shfl_sync(0xffffffffu, first, leader);
    if (valid) slot_row[s_base[key] + first + __popc(peers & ((1u << lane) - 1u))] = r;
  }
}

__global__ void __launch_bounds__(256)
k_nbr_permute(const int32_t* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ slot_row,
              const int32_t* __restrict__ n_p, int K, int32_t* __restrict__ sorted, int sorted_stride,
              uint32_t* __restrict__ tile_mask) {
  const int n = *n_p;
  const int per_round = gridDim.x * blockDim.x;
  const int nrounds = (n + per_round - 1) / per_round;
  for (int r = 0; r < nrounds; ++r) {
    const int s = r * per_round + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0u;
    if (s < n) {
      const int o = __ldg(&slot_row[s]);
      for (int k = 0; k < K; ++k) {
        const int v = __ldg(&nbr[(size_t)k * nbr_stride + o]);
        sorted[(size_t)k * sorted_stride + s] = v;
        m |= (v >= 0 ? 1u : 0u) << k;
      }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if (m && (threadIdx.x & 31) == 0) atomicOr(&tile_mask[s >> 7], m);
  }
}

// ------------------------------------------------------------------------------------------------
// Grouped variant (validated in round 2; reached through u3d_rulebook_sort_tiles_grouped and the n_groups > 1 form of
// u3d_rulebook_subm_sorted / U3D_SORT_GROUP; measured neutral on the step, not the default): bucket by
// signature INSIDE groups of consecutive scenes, so that the gathers of a tile stay within a few
// scenes' feature rows (L2 resident on the wide levels at batch 32, where the global order loses:
// 64->64 0.348 -> 0.404 ms). CPU estimate of the useful slot fraction on 8 scenes, stage 2:
// natural 0.65, per-scene buckets 0.88, global 0.92 (scripts/tile_padding_stats.py).
// Rows are scene-major at every level, so a group is a contiguous row segment seg[g]..seg[g+1].

// seg[g] = first row whose batch index is >= g * scenes_per_group (binary search), seg[G] = n
__global__ void k_seg_bounds(const int32_t* __restrict__ coors, const int32_t* __restrict__ n_p, int G,
                             int scenes_per_group, int32_t* __restrict__ seg) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  const int n = *n_p;
  const int want = g * scenes_per_group;
  int lo = 0, hi = n;                       // first row with batch >= want
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&coors[(size_t)mid * 4]) < want) lo = mid + 1; else hi = mid;
  }
  seg[g] = g == G ? n : lo;
}

__device__ __forceinline__ void seg_block_range(const int32_t* __restrict__ seg, int& start, int& end) {
  const int s0 = seg[blockIdx.y], s1 = seg[blockIdx.y + 1];
  const int n = s1 - s0;
  int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  per = (per + kSortThreads - 1) / kSortThreads * kSortThreads;
  const long long s = (long long)blockIdx.x * per;
  start = s0 + (s < n ? (int)s : n);
  end = s0 + (s + per < n ? (int)(s + per) : n);
}

__global__ void __launch_bounds__(kSortThreads)
k_key_hist_seg(const int32_t* __restrict__ row_key, const int32_t* __restrict__ seg, int32_t* __restrict__ hist) {
  __shared__ int s_hist[kKeyBins];
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads) s_hist[b] = 0;
  __syncthreads();
  int start, end;
  seg_block_range(seg, start, end);
  block_histogram(row_key, start, end, s_hist);
  __syncthreads();
  int32_t* h = hist + (size_t)blockIdx.y * kKeyBins;
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads)
    if (s_hist[b]) atomicAdd(&h[b], s_hist[b]);
}

// exclusive scan over the concatenated (group-major) histograms: G * 4096 bins, one block
__global__ void __launch_bounds__(1024)
k_key_scan_seg(const int32_t* __restrict__ hist, int total_bins, int32_t* __restrict__ cursor) {
  __shared__ int smem[33];
  const int per = (total_bins + 1023) / 1024;
  const int b0 = threadIdx.x * per;
  int sum = 0;
  for (int i = 0; i < per; ++i)
    if (b0 + i < total_bins) sum += hist[b0 + i];
  int total;
  int ex = block_exclusive_scan(sum, smem, total);
  for (int i = 0; i < per; ++i)
    if (b0 + i < total_bins) { cursor[b0 + i] = ex; ex += hist[b0 + i]; }
}

__global__ void __launch_bounds__(kSortThreads)
k_key_scatter_seg(const int32_t* __restrict__ row_key, const int32_t* __restrict__ seg,
                  int32_t* __restrict__ cursor, int32_t* __restrict__ slot_row) {
  __shared__ int s_cnt[kKeyBins];
  __shared__ int s_base[kKeyBins];
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads) s_cnt[b] = 0;
  __syncthreads();
  int start, end;
  seg_block_range(seg, start, end);
  block_histogram(row_key, start, end, s_cnt);
  __syncthreads();
  int32_t* cur = cursor + (size_t)blockIdx.y * kKeyBins;
  for (int b = threadIdx.x; b < kKeyBins; b += kSortThreads) {
    const int c = s_cnt[b];
    if (c) s_base[b] = atomicAdd(&cur[b], c);
    s_cnt[b] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int base = start; base < end; base += kSortThreads) {
    const int r = base + threadIdx.x;
    const bool valid = r < end;
    const int key = valid ? __ldg(&row_key[r]) : (kKeyBins + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    int first = 0;
    if (valid && lane == leader) first = atomicAdd(&s_cnt[key], __popc(peers));
    first = __shfl_sync(0xffffffffu, first, leader);
    if (valid) slot_row[s_base[key] + first + __popc(peers & ((1u << lane) - 1u))] = r;
  }
}


// ------------------------------------------------------------------------------------------------
// Sorted SubM rulebook straight from the coordinates (round 2). u3d_rulebook_sort_tiles reads the natural
// 27 x rows table twice (signature, permute) and writes it once more - 3 x 108 B per row on top of the
// 108 B the natural build wrote. A SubM level whose convs all take the sorted table never needs the
// natural one: pass 1 derives the signature from 9 - 12 VoxelMap occupancy words per row, the counting
// sort yields slot_row, pass 2 looks the neighbours of row slot_row[s] up again (with ranks) and writes
// the slot-ordered table directly: one 108 B/row write, two passes over the 16 B/row coordinates.

// the 27 neighbour lookups of one output row (SubM: stride 1, pad 1, the level's own map; strided conv: the input
// level's map): emit(k, input row or -1).
// kRanks = false: occupancy only (emit gets 0 / -1), no rank arithmetic, no perm load.
struct ConvGeo { int D, H, W, sz, sy, sx, pz, py, px; };   // input grid, stride, padding (3x3x3 kernel)

template <bool kRanks, class Emit>
__device__ __forceinline__ void subm_neighbours(const int4 c, const uint2* __restrict__ map,
                                                const int32_t* __restrict__ perm, const ConvGeo g, Emit&& emit) {
  const int D = g.D, H = g.H, W = g.W;
  const int z0 = c.y * g.sz - g.pz, y0 = c.z * g.sy - g.py, x0 = c.w * g.sx - g.px;
#pragma unroll
  for (int kz = 0; kz < 3; ++kz) {
    const int z = z0 + kz;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = y0 + ky;
      const bool line_ok = z >= 0 && z < D && y >= 0 && y < H;
      const uint32_t lin0 = line_ok ? (uint32_t)((((size_t)c.x * D + z) * H + y) * W) : 0u;
      uint32_t cached = 0xffffffffu;
      uint2 w = make_uint2(0u, 0u);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int x = x0 + kx;
        int row = -1;
        if (line_ok && x >= 0 && x < W) {
          const uint32_t lin = lin0 + (uint32_t)x;
          const uint32_t wi = lin >> 5, bit = lin & 31u;
          if (wi != cached) { w = __ldg(&map[wi]); cached = wi; }
          if ((w.x >> bit) & 1u) {
            if (kRanks) {
              const int rank = (int)w.y + __popc(w.x & ((1u << bit) - 1u));
              row = perm ? __ldg(&perm[rank]) : rank;
            } else {
              row = 0;
            }
          }
        }
        emit((kz * 3 + ky) * 3 + kx, row);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_row_key_map(const int32_t* __restrict__ coors, const int32_t* __restrict__ n_p, const uint2* __restrict__ map,
              const ConvGeo geo, int32_t* __restrict__ row_key) {
  const int n = *n_p;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(coors) + o);
    uint32_t m = 0u;
    subm_neighbours<false>(c, map, nullptr, geo, [&](int k, int row) { m |= (row >= 0 ? 1u : 0u) << k; });
    row_key[o] = (int32_t)tile_key_of_mask(m);
  }
}

__global__ void __launch_bounds__(256)
k_nbr_build_slots(const int32_t* __restrict__ coors, const int32_t* __restrict__ n_p, const uint2* __restrict__ map,
                  const int32_t* __restrict__ perm, const ConvGeo geo, const int32_t* __restrict__ slot_row,
                  int32_t* __restrict__ sorted, int sorted_stride, uint32_t* __restrict__ tile_mask) {
  const int n = *n_p;
  const int per_round = gridDim.x * blockDim.x;
  const int nrounds = (n + per_round - 1) / per_round;   // uniform trip count: whole warps reach the reduction
  for (int r = 0; r < nrounds; ++r) {
    const int s = r * per_round + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0u;
    if (s < n) {
      const int o = __ldg(&slot_row[s]);
      const int4 c = __ldg(reinterpret_cast<const int4*>(coors) + o);
      int32_t* dst = sorted + s;
      subm_neighbours<true>(c, map, perm, geo, [&](int k, int row) {
        dst[(size_t)k * sorted_stride] = row;
        m |= (row >= 0 ? 1u : 0u) << k;
      });
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if (m && (threadIdx.x & 31) == 0) atomicOr(&tile_mask[s >> 7], m);
  }
}

}  // namespace u3d

using namespace u3d;

extern "C" size_t u3d_tile_sort_scratch_ints(int cap) {
  return (size_t)(cap > 0 ? cap : 1) + 2 * (size_t)kKeyBins;   // row keys + histogram + cursors
}

extern "C" int u3d_rulebook_sort_tiles(const int32_t* nbr, int nbr_stride, const int32_t* n_out, int cap,
                                       int K, int32_t* scratch, int32_t* slot_row, int32_t* nbr_sorted,
                                       int sorted_stride, uint32_t* tile_mask_sorted, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(nbr && n_out && scratch && slot_row && nbr_sorted && tile_mask_sorted,
                "u3d_rulebook_sort_tiles: null buffer");
  U3D_CHECK_ARG(K >= 1 && K <= 27 && cap >= 0 && nbr_stride >= cap && sorted_stride >= cap,
                "u3d_rulebook_sort_tiles: bad shape (K=%d cap=%d strides %d/%d)", K, cap, nbr_stride,
                sorted_stride);
  if (cap == 0) return U3D_OK;
  int32_t* row_key = scratch;
  int32_t* hist = scratch + cap;
  int32_t* cursor = hist + kKeyBins;
  U3D_CUDA(cudaMemsetAsync(hist, 0, 2 * (size_t)kKeyBins * sizeof(int32_t), st));
  U3D_CUDA(cudaMemsetAsync(tile_mask_sorted, 0, (size_t)cdiv(cap, 128) * sizeof(uint32_t), st));
  int g = cdiv(cap, 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  k_row_key<<<g, 256, 0, st>>>(nbr, nbr_stride, n_out, K, row_key);
  U3D_LAUNCH_CHECK();
  int gs = cdiv(cap, 8 * kSortThreads);      // >= 2048 rows per block at full capacity
  if (gs < 1) gs = 1;
  if (gs > kNumSMs * 4) gs = kNumSMs * 4;
  k_key_hist<<<gs, kSortThreads, 0, st>>>(row_key, n_out, hist);
  U3D_LAUNCH_CHECK();
  k_key_scan<<<1, 1024, 0, st>>>(hist, cursor);
  U3D_LAUNCH_CHECK();
  k_key_scatter<<<gs, kSortThreads, 0, st>>>(row_key, n_out, cursor, slot_row);
  U3D_LAUNCH_CHECK();
  k_nbr_permute<<<g, 256, 0, st>>>(nbr, nbr_stride, slot_row, n_out, K, nbr_sorted, sorted_stride,
                                   tile_mask_sorted);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" size_t u3d_tile_sort_grouped_scratch_ints(int cap, int n_groups) {
  return (size_t)(cap > 0 ? cap : 1) + (size_t)(n_groups + 1) + 2 * (size_t)kKeyBins * (size_t)(n_groups > 0 ? n_groups : 1);
}

// u3d_rulebook_sort_tiles with the signature buckets kept
// inside groups of `scenes_per_group` consecutive scenes. coors: (cap,4) int32 [b,z,y,x] of the OUTPUT rows
// (scene-major); n_groups = ceil(B / scenes_per_group).
extern "C" int u3d_rulebook_sort_tiles_grouped(const int32_t* nbr, int nbr_stride, const int32_t* coors,
                                               const int32_t* n_out, int cap, int K, int n_groups,
                                               int scenes_per_group, int32_t* scratch, int32_t* slot_row,
                                               int32_t* nbr_sorted, int sorted_stride,
                                               uint32_t* tile_mask_sorted, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(nbr && coors && n_out && scratch && slot_row && nbr_sorted && tile_mask_sorted,
                "u3d_rulebook_sort_tiles_grouped: null buffer");
  U3D_CHECK_ARG(K >= 1 && K <= 27 && cap >= 0 && nbr_stride >= cap && sorted_stride >= cap && n_groups >= 1 &&
                    n_groups <= 1024 && scenes_per_group >= 1,
                "u3d_rulebook_sort_tiles_grouped: bad shape (K=%d cap=%d groups=%d)", K, cap, n_groups);
  if (cap == 0) return U3D_OK;
  int32_t* row_key = scratch;
  int32_t* seg = scratch + cap;
  int32_t* hist = seg + (n_groups + 1);
  int32_t* cursor = hist + (size_t)kKeyBins * n_groups;
  U3D_CUDA(cudaMemsetAsync(hist, 0, 2 * (size_t)kKeyBins * n_groups * sizeof(int32_t), st));
  U3D_CUDA(cudaMemsetAsync(tile_mask_sorted, 0, (size_t)cdiv(cap, 128) * sizeof(uint32_t), st));
  int g = cdiv(cap, 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  k_row_key<<<g, 256, 0, st>>>(nbr, nbr_stride, n_out, K, row_key);
  U3D_LAUNCH_CHECK();
  k_seg_bounds<<<cdiv(n_groups + 1, 128), 128, 0, st>>>(coors, n_out, n_groups, scenes_per_group, seg);
  U3D_LAUNCH_CHECK();
  int per_group = cdiv(cdiv(cap, n_groups), 8 * kSortThreads);   // >= 2048 rows per block at full capacity
  if (per_group < 1) per_group = 1;
  if (per_group * n_groups > kNumSMs * 8) per_group = cdiv(kNumSMs * 8, n_groups);
  dim3 gs(per_group, n_groups);
  k_key_hist_seg<<<gs, kSortThreads, 0, st>>>(row_key, seg, hist);
  U3D_LAUNCH_CHECK();
  k_key_scan_seg<<<1, 1024, 0, st>>>(hist, kKeyBins * n_groups, cursor);
  U3D_LAUNCH_CHECK();
  k_key_scatter_seg<<<gs, kSortThreads, 0, st>>>(row_key, seg, cursor, slot_row);
  U3D_LAUNCH_CHECK();
  k_nbr_permute<<<g, 256, 0, st>>>(nbr, nbr_stride, slot_row, n_out, K, nbr_sorted, sorted_stride,
                                   tile_mask_sorted);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

// Sorted SubM rulebook straight from the coordinates (see k_row_key_map / k_nbr_build_slots): the outputs of
// u3d_rulebook_subm + u3d_rulebook_sort_tiles(_grouped) without the natural-order table in between.
// n_groups <= 1: one global set of buckets; else buckets inside groups of scenes_per_group scenes.
// scratch: u3d_tile_sort_grouped_scratch_ints(cap, max(n_groups, 1)) int32.
static int sorted_from_coords(const int32_t* coors, const int32_t* n_rows, int cap, const void* map, const int32_t* perm,
                              const ConvGeo geo, int n_groups, int scenes_per_group, int32_t* scratch,
                              int32_t* slot_row, int32_t* nbr_sorted, int sorted_stride, uint32_t* tile_mask_sorted,
                              cudaStream_t st) {
  if (cap == 0) return U3D_OK;
  const int G = n_groups > 1 ? n_groups : 1;
  int32_t* row_key = scratch;
  int32_t* seg = scratch + cap;
  int32_t* hist = seg + (G + 1);
  int32_t* cursor = hist + (size_t)kKeyBins * G;
  U3D_CUDA(cudaMemsetAsync(hist, 0, 2 * (size_t)kKeyBins * G * sizeof(int32_t), st));
  U3D_CUDA(cudaMemsetAsync(tile_mask_sorted, 0, (size_t)cdiv(cap, 128) * sizeof(uint32_t), st));
  int g = cdiv(cap, 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  k_row_key_map<<<g, 256, 0, st>>>(coors, n_rows, (const uint2*)map, geo, row_key);
  U3D_LAUNCH_CHECK();
  if (G > 1) {
    k_seg_bounds<<<cdiv(G + 1, 128), 128, 0, st>>>(coors, n_rows, G, scenes_per_group, seg);
    U3D_LAUNCH_CHECK();
    int per_group = cdiv(cdiv(cap, G), 8 * kSortThreads);
    if (per_group < 1) per_group = 1;
    if (per_group * G > kNumSMs * 8) per_group = cdiv(kNumSMs * 8, G);
    dim3 gs(per_group, G);
    k_key_hist_seg<<<gs, kSortThreads, 0, st>>>(row_key, seg, hist);
    U3D_LAUNCH_CHECK();
    k_key_scan_seg<<<1, 1024, 0, st>>>(hist, kKeyBins * G, cursor);
    U3D_LAUNCH_CHECK();
    k_key_scatter_seg<<<gs, kSortThreads, 0, st>>>(row_key, seg, cursor, slot_row);
    U3D_LAUNCH_CHECK();
  } else {
    int gs = cdiv(cap, 8 * kSortThreads);
    if (gs < 1) gs = 1;
    if (gs > kNumSMs * 4) gs = kNumSMs * 4;
    k_key_hist<<<gs, kSortThreads, 0, st>>>(row_key, n_rows, hist);
    U3D_LAUNCH_CHECK();
    k_key_scan<<<1, 1024, 0, st>>>(hist, cursor);
    U3D_LAUNCH_CHECK();
    k_key_scatter<<<gs, kSortThreads, 0, st>>>(row_key, n_rows, cursor, slot_row);
    U3D_LAUNCH_CHECK();
  }
  k_nbr_build_slots<<<g, 256, 0, st>>>(coors, n_rows, (const uint2*)map, perm, geo, slot_row, nbr_sorted, sorted_stride,
                                       tile_mask_sorted);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_rulebook_subm_sorted(const int32_t* coors, const int32_t* n_rows, int cap, const void* map,
                                        const int32_t* perm, int B, int D, int H, int W, int n_groups,
                                        int scenes_per_group, int32_t* scratch, int32_t* slot_row,
                                        int32_t* nbr_sorted, int sorted_stride, uint32_t* tile_mask_sorted,
                                        void* stream) {
  U3D_CHECK_ARG(coors && n_rows && map && scratch && slot_row && nbr_sorted && tile_mask_sorted,
                "u3d_rulebook_subm_sorted: null buffer");
  U3D_CHECK_ARG(cap >= 0 && sorted_stride >= cap && B >= 1 && D >= 1 && H >= 1 && W >= 1 && n_groups <= 1024,
                "u3d_rulebook_subm_sorted: bad shape (cap=%d stride=%d groups=%d)", cap, sorted_stride, n_groups);
  const ConvGeo geo{D, H, W, 1, 1, 1, 1, 1, 1};
  return sorted_from_coords(coors, n_rows, cap, map, perm, geo, n_groups, scenes_per_group, scratch, slot_row, nbr_sorted,
                            sorted_stride, tile_mask_sorted, (cudaStream_t)stream);
}

// The same for the table of a strided SparseConv3d: out_coors / n_out are the output rows u3d_rulebook_down emitted
// (call it with nbr = NULL to skip its natural-order table), in_map / in_perm / in_dims describe the INPUT level.
extern "C" int u3d_rulebook_down_sorted(const int32_t* out_coors, const int32_t* n_out, int out_cap, const void* in_map,
                                        const int32_t* in_perm, int B, const int32_t* in_dims, const int32_t* stride,
                                        const int32_t* pad, int n_groups, int scenes_per_group, int32_t* scratch,
                                        int32_t* slot_row, int32_t* nbr_sorted, int sorted_stride,
                                        uint32_t* tile_mask_sorted, void* stream) {
  U3D_CHECK_ARG(out_coors && n_out && in_map && in_dims && stride && pad && scratch && slot_row && nbr_sorted &&
                    tile_mask_sorted, "u3d_rulebook_down_sorted: null buffer");
  U3D_CHECK_ARG(out_cap >= 0 && sorted_stride >= out_cap && B >= 1 && n_groups <= 1024,
                "u3d_rulebook_down_sorted: bad shape (cap=%d stride=%d groups=%d)", out_cap, sorted_stride, n_groups);
  const ConvGeo geo{in_dims[0], in_dims[1], in_dims[2], stride[0], stride[1], stride[2], pad[0], pad[1], pad[2]};
  return sorted_from_coords(out_coors, n_out, out_cap, in_map, in_perm, geo, n_groups, scenes_per_group, scratch,
                            slot_row, nbr_sorted, sorted_stride, tile_mask_sorted, (cudaStream_t)stream);
}

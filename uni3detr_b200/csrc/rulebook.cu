// K2 — sparse-conv rulebook build (SubMConv3d / strided SparseConv3d index generation).
//
// Reference semantics: SURVEY.md A.3; the convs are instantiated at
// projects/mmdet3d_plugin/models/pts_encoder/sparse_encoder_hd.py:71-104,161-213 and the
// index generation itself is spconv's get_indice_pairs (third-party, not vendored).
//
// The rulebook is kept output-stationary: nbr[k][o] = input row feeding output row o
// through kernel offset k (or -1). That is the form the gather-GEMM consumes without
// atomics; u3d_rulebook_pairs converts it to spconv-1.x pair lists for API parity.
// All lookups go through the VoxelMap (an 8-byte load + popcount, shared by x-neighbours), output
// coordinates of a strided conv come out in ascending linear index for free.
// Integer/HBM-bound: reads 16 B/row coords, writes 4 B per (row, offset).
#include "common.cuh"

namespace u3d {

struct Dims3 { int d[3]; };  // z,y,x

// One thread per OUTPUT row: the coordinate row is loaded once, the 27 lookups run out of
// registers (the three kx neighbours of a (kz,ky) line usually share one 8-byte map word, so a
// row costs ~9-12 map loads instead of 27), each k-plane of the table is written coalesced, and
// the warp's 27-bit "which offsets feed these rows" mask costs ONE atomicOr per warp.
__global__ void __launch_bounds__(256)
k_nbr_build(const int32_t* __restrict__ out_coors, const int32_t* __restrict__ n_out_p,
            const uint2* __restrict__ in_map, const int32_t* __restrict__ in_perm, Dims3 in_dims,
            Dims3 stride, Dims3 pad, int32_t* __restrict__ nbr, int nbr_stride,
            uint32_t* __restrict__ tile_mask) {
  const int n_out = *n_out_p;
  const int D = in_dims.d[0], H = in_dims.d[1], W = in_dims.d[2];
  // uniform trip count: whole warps stay converged for the reduction behind the tile masks
  const int per_round = gridDim.x * blockDim.x;
  const int nrounds = (n_out + per_round - 1) / per_round;
  for (int r = 0; r < nrounds; ++r) {
    const int o = r * per_round + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t rowmask = 0u;
    if (o < n_out) {
      const int4 c = __ldg(reinterpret_cast<const int4*>(out_coors) + o);  // b,z,y,x
      const int z0 = c.y * stride.d[0] - pad.d[0];
      const int y0 = c.z * stride.d[1] - pad.d[1];
      const int x0 = c.w * stride.d[2] - pad.d[2];
      int32_t* dst = nbr + o;
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
            const int k = (kz * 3 + ky) * 3 + kx;
            const int x = x0 + kx;
            int row = -1;
            if (line_ok && x >= 0 && x < W) {
              const uint32_t lin = lin0 + (uint32_t)x;
              const uint32_t wi = lin >> 5, bit = lin & 31u;
              if (wi != cached) { w = __ldg(&in_map[wi]); cached = wi; }
              if ((w.x >> bit) & 1u) {
                const int rank = (int)w.y + __popc(w.x & ((1u << bit) - 1u));
                row = in_perm ? __ldg(&in_perm[rank]) : rank;
              }
            }
            dst[(size_t)k * nbr_stride] = row;
            rowmask |= (row >= 0 ? 1u : 0u) << k;
          }
        }
      }
    }
    // bit k of tile_mask[t] = "offset k feeds at least one of output rows [128t, 128t+128)";
    // a warp covers 32 consecutive rows of one tile
    const uint32_t m = __reduce_or_sync(0xffffffffu, rowmask);
    if (tile_mask && m && (threadIdx.x & 31) == 0) atomicOr(&tile_mask[o >> 7], m);
  }
}

// every input row marks its candidate output cells: per axis only the offsets k with
// (c + pad - k) divisible by the stride qualify (at most 2 of 3 for stride 2), so a row issues
// <= 8 atomicOr instead of testing 27 combinations.
__global__ void __launch_bounds__(256)
k_down_mark(const int32_t* __restrict__ in_coors, const int32_t* __restrict__ n_in_p,
            Dims3 out_dims, Dims3 stride, Dims3 pad, uint2* __restrict__ out_map) {
  const int n_in = *n_in_p;
  const int D = out_dims.d[0], H = out_dims.d[1], W = out_dims.d[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(in_coors) + i);
    int oz[3], oy[3], ox[3];   // output coordinate per kernel offset, or -1
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int z = c.y + pad.d[0] - k, y = c.z + pad.d[1] - k, x = c.w + pad.d[2] - k;
      oz[k] = (z >= 0 && z % stride.d[0] == 0 && z / stride.d[0] < D) ? z / stride.d[0] : -1;
      oy[k] = (y >= 0 && y % stride.d[1] == 0 && y / stride.d[1] < H) ? y / stride.d[1] : -1;
      ox[k] = (x >= 0 && x % stride.d[2] == 0 && x / stride.d[2] < W) ? x / stride.d[2] : -1;
    }
#pragma unroll
    for (int kz = 0; kz < 3; ++kz) {
      if (oz[kz] < 0) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        if (oy[ky] < 0) continue;
        const uint32_t line = (uint32_t)((((size_t)c.x * D + oz[kz]) * H + oy[ky]) * W);
        // the (up to 3) x candidates of one line fall into one or two map words: merge them
        uint32_t w0 = 0xffffffffu, b0 = 0u, w1 = 0xffffffffu, b1 = 0u;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if (ox[kx] < 0) continue;
          const uint32_t lin = line + (uint32_t)ox[kx];
          const uint32_t wi = lin >> 5, bit = 1u << (lin & 31u);
          if (w0 == 0xffffffffu || w0 == wi) { w0 = wi; b0 |= bit; }
          else { w1 = wi; b1 |= bit; }
        }
        if (b0) atomicOr(&out_map[w0].x, b0);
        if (b1) atomicOr(&out_map[w1].x, b1);
      }
    }
  }
}

// VoxelMap from an arbitrary coordinate list (SparseConvTensor built by the caller)
__global__ void __launch_bounds__(256)
k_coor_mark(const int32_t* __restrict__ coors, const int32_t* __restrict__ n_p, int B, Dims3 dims,
            uint2* __restrict__ map) {
  const int n = *n_p;
  const int D = dims.d[0], H = dims.d[1], W = dims.d[2];
  const int per_round = gridDim.x * blockDim.x;
  const int nrounds = (n + per_round - 1) / per_round;
  for (int r = 0; r < nrounds; ++r) {
    int i = r * per_round + blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    uint32_t lin = 0xffffffffu;
    if (i < n) {
      int4 c = __ldg(reinterpret_cast<const int4*>(coors) + i);
      valid = c.x >= 0 && c.x < B && c.y >= 0 && c.y < D && c.z >= 0 && c.z < H && c.w >= 0 && c.w < W;
      if (valid) lin = (uint32_t)((((size_t)c.x * D + c.y) * H + c.z) * W + c.w);
    }
    map_set_bit_aggregated(map, valid, lin);
  }
}

__global__ void __launch_bounds__(256)
k_coor_perm(const int32_t* __restrict__ coors, const int32_t* __restrict__ n_p, int B, Dims3 dims,
            const uint2* __restrict__ map, int32_t* __restrict__ perm) {
  const int n = *n_p;
  const int D = dims.d[0], H = dims.d[1], W = dims.d[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = __ldg(reinterpret_cast<const int4*>(coors) + i);
    bool valid = c.x >= 0 && c.x < B && c.y >= 0 && c.y < D && c.z >= 0 && c.z < H && c.w >= 0 && c.w < W;
    if (!valid) continue;
    uint32_t lin = (uint32_t)((((size_t)c.x * D + c.y) * H + c.z) * W + c.w);
    perm[map_rank_at(map, lin)] = i;
  }
}

// decode every set bit of a scanned map into its coordinate row (rank order)
__global__ void __launch_bounds__(256)
k_map_emit_coors(const uint2* __restrict__ map, size_t words, int D, int H, int W,
                 int32_t* __restrict__ coors, int cap) {
  for (size_t wi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; wi < words;
       wi += (size_t)gridDim.x * blockDim.x) {
    uint2 w = __ldg(&map[wi]);
    uint32_t bits = w.x;
    int row = (int)w.y;
    while (bits) {
      int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      uint32_t t = (uint32_t)(wi * 32 + bit);
      int x = t % W; t /= W;
      int y = t % H; t /= H;
      int z = t % D; t /= D;
      if (row < cap) reinterpret_cast<int4*>(coors)[row] = make_int4((int)t, z, y, x);
      ++row;
    }
  }
}

// nbr -> spconv 1.x pair lists; one CTA per kernel offset, stable in output-row order
__global__ void __launch_bounds__(1024)
k_pairs(const int32_t* __restrict__ nbr, int nbr_stride, const int32_t* __restrict__ n_out_p,
        int32_t* __restrict__ pairs_in, int32_t* __restrict__ pairs_out, int pair_stride,
        int32_t* __restrict__ pair_num) {
  __shared__ int smem[33];
  const int k = blockIdx.x;
  const int n_out = *n_out_p;
  int running = 0;
  for (int start = 0; start < n_out; start += blockDim.x) {
    int o = start + threadIdx.x;
    int src = o < n_out ? nbr[(size_t)k * nbr_stride + o] : -1;
    int flag = src >= 0;
    int total;
    int ex = block_exclusive_scan(flag, smem, total);
    if (flag) {
      int pos = running + ex;
      if (pos < pair_stride) {
        pairs_in[(size_t)k * pair_stride + pos] = src;
        pairs_out[(size_t)k * pair_stride + pos] = o;
      }
    }
    running += total;
  }
  for (int pos = running + threadIdx.x; pos < pair_stride; pos += blockDim.x) {
    pairs_in[(size_t)k * pair_stride + pos] = -1;
    pairs_out[(size_t)k * pair_stride + pos] = -1;
  }
  if (threadIdx.x == 0) pair_num[k] = running;
}

static inline int grid_x_for(long long n, int threads, int max_ctas) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > max_ctas) g = max_ctas;
  return (int)g;
}

}  // namespace u3d

using namespace u3d;

extern "C" int u3d_voxmap_build(const int32_t* coors, const int32_t* n_rows, int cap, int B, int D,
                                int H, int W, void* map_, int32_t* scan_scratch, int32_t* perm,
                                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(coors && n_rows && map_ && scan_scratch && perm, "u3d_voxmap_build: null buffer");
  size_t words = u3d_voxmap_words(B, D, H, W);
  if (!words) {
    set_error("u3d_voxmap_build: B*D*H*W exceeds the 32-bit cell index; split the batch");
    return U3D_ERANGE;
  }
  uint2* map = (uint2*)map_;
  Dims3 dims{{D, H, W}};
  U3D_CUDA(cudaMemsetAsync(map, 0, words * sizeof(uint2), st));
  int gx = grid_x_for(cap, 256, kNumSMs * 8);
  k_coor_mark<<<gx, 256, 0, st>>>(coors, n_rows, B, dims, map);
  U3D_LAUNCH_CHECK();
  int rc = voxmap_scan(map, words, scan_scratch, nullptr, st);
  if (rc) return rc;
  k_coor_perm<<<gx, 256, 0, st>>>(coors, n_rows, B, dims, map, perm);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_rulebook_subm(const int32_t* coors, const int32_t* n_rows, int cap,
                                 const void* map, const int32_t* perm, int B, int D, int H, int W,
                                 int32_t* nbr, int nbr_stride, uint32_t* tile_mask, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(coors && n_rows && map && nbr, "u3d_rulebook_subm: null buffer");
  U3D_CHECK_ARG(nbr_stride >= cap && cap >= 0, "u3d_rulebook_subm: nbr_stride < cap");
  U3D_CHECK_ARG(u3d_voxmap_words(B, D, H, W) != 0, "u3d_rulebook_subm: bad grid");
  Dims3 dims{{D, H, W}}, one{{1, 1, 1}};
  if (tile_mask) U3D_CUDA(cudaMemsetAsync(tile_mask, 0, (size_t)cdiv(cap > 0 ? cap : 1, 128) * 4, st));
  const int grid = grid_x_for(cap, 256, kNumSMs * 8);
  k_nbr_build<<<grid, 256, 0, st>>>(coors, n_rows, (const uint2*)map, perm, dims, one, one, nbr,
                                    nbr_stride, tile_mask);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_rulebook_down(const int32_t* in_coors, const int32_t* n_in, int in_cap,
                                 const void* in_map, const int32_t* in_perm, int B,
                                 const int32_t* in_dims, const int32_t* out_dims,
                                 const int32_t* stride, const int32_t* pad, void* out_map_,
                                 int32_t* scan_scratch, int32_t* out_coors, int32_t* n_out,
                                 int out_cap, int32_t* nbr, int nbr_stride, uint32_t* tile_mask,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(in_coors && n_in && in_map && out_map_ && scan_scratch && out_coors && n_out,
                "u3d_rulebook_down: null buffer");
  U3D_CHECK_ARG(nbr == nullptr || nbr_stride >= out_cap, "u3d_rulebook_down: nbr_stride < out_cap");
  Dims3 id{{in_dims[0], in_dims[1], in_dims[2]}}, od{{out_dims[0], out_dims[1], out_dims[2]}};
  Dims3 s{{stride[0], stride[1], stride[2]}}, p{{pad[0], pad[1], pad[2]}};
  for (int i = 0; i < 3; ++i) {
    U3D_CHECK_ARG(s.d[i] >= 1 && p.d[i] >= 0, "u3d_rulebook_down: bad stride/pad");
    int expect = (id.d[i] + 2 * p.d[i] - 3) / s.d[i] + 1;
    U3D_CHECK_ARG(od.d[i] == expect, "u3d_rulebook_down: out_dims[%d]=%d, expected %d", i, od.d[i],
                  expect);
  }
  size_t words = u3d_voxmap_words(B, od.d[0], od.d[1], od.d[2]);
  U3D_CHECK_ARG(words != 0, "u3d_rulebook_down: bad out grid");
  uint2* out_map = (uint2*)out_map_;
  U3D_CUDA(cudaMemsetAsync(out_map, 0, words * sizeof(uint2), st));
  const int gmark = grid_x_for(in_cap, 256, kNumSMs * 8);
  k_down_mark<<<gmark, 256, 0, st>>>(in_coors, n_in, od, s, p, out_map);
  U3D_LAUNCH_CHECK();
  int rc = voxmap_scan(out_map, words, scan_scratch, n_out, st);
  if (rc) return rc;
  k_map_emit_coors<<<grid_x_for((long long)words, 256, kNumSMs * 8), 256, 0, st>>>(
      out_map, words, od.d[0], od.d[1], od.d[2], out_coors, out_cap);
  U3D_LAUNCH_CHECK();
  if (nbr == nullptr) return U3D_OK;      // output set only: the caller builds a slot-ordered table (u3d_rulebook_down_sorted)
  if (tile_mask) U3D_CUDA(cudaMemsetAsync(tile_mask, 0, (size_t)cdiv(out_cap > 0 ? out_cap : 1, 128) * 4, st));
  const int gn = grid_x_for(out_cap, 256, kNumSMs * 8);
  k_nbr_build<<<gn, 256, 0, st>>>(out_coors, n_out, (const uint2*)in_map, in_perm, id, s, p, nbr,
                                  nbr_stride, tile_mask);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

extern "C" int u3d_rulebook_pairs(const int32_t* nbr, int nbr_stride, const int32_t* n_out, int K,
                                  int32_t* pairs_in, int32_t* pairs_out, int pair_stride,
                                  int32_t* pair_num, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(nbr && n_out && pairs_in && pairs_out && pair_num && K >= 1,
                "u3d_rulebook_pairs: bad argument");
  k_pairs<<<K, 1024, 0, st>>>(nbr, nbr_stride, n_out, pairs_in, pairs_out, pair_stride, pair_num);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

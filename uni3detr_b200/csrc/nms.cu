// "Next" row 1: per-class rotated-BEV-IoU NMS on the device, batched over scenes.
//
// Reference: the per-class python loop over mmcv.ops.nms3d in Uni3DETRHead.get_bboxes,
// projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:847-871 (nms3d at :861). mmcv's
// nms3d sorts a class's boxes by score, computes the BEV IoU of the rotated rectangles
// (x, y, dx, dy, heading) and greedily suppresses boxes whose IoU with a kept box exceeds the
// threshold (iou3d_nms3d_forward; source not vendored - the IoU here is the exact area of the
// convex polygon intersection, computed by clipping one rectangle against the other's 4 edges).
//
// Here ONE launch pair handles every class of every scene: the caller passes each scene's boxes
// sorted by (label, score desc); kernel 1 fills the upper-triangular suppression bit matrix for
// same-label pairs (64 x 64 box blocks, column boxes staged in shared memory), kernel 2 sweeps it
// with one warp per scene. No per-class launch, no host synchronisation, fixed-size outputs.
#include "bev_geom.cuh"

namespace u3d {

// grid (col blocks, row blocks, scenes), 64 threads: thread i = row box, 64 column boxes in smem.
// Boxes are sorted by label, so a 64 x 64 block whose row and column label ranges do not meet has
// nothing to do (with 10 classes ~90 % of the blocks): it writes zeros and leaves.
__global__ void __launch_bounds__(64)
k_nms_mask(const float* __restrict__ boxes, const int32_t* __restrict__ labels,
           const uint8_t* __restrict__ valid, int N, float thr, unsigned long long* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y, s = blockIdx.z;
  if (cb < rb) return;  // strictly-lower blocks are never read
  const int nblk = (N + 63) >> 6;
  __shared__ float s_box[64][8];
  __shared__ int s_lab[64];
  const int r = rb * 64 + threadIdx.x;
  const int c = cb * 64 + threadIdx.x;
  const int lab = (r < N && valid[(size_t)s * N + r]) ? labels[(size_t)s * N + r] : -2;
  const int clab = (c < N && valid[(size_t)s * N + c]) ? labels[(size_t)s * N + c] : -1;
  s_lab[threadIdx.x] = clab;
  // label ranges of the block's rows / columns (valid entries only)
  const int rmin = __reduce_min_sync(0xffffffffu, lab >= 0 ? lab : 0x7fffffff);
  const int rmax = __reduce_max_sync(0xffffffffu, lab);
  const int cmin = __reduce_min_sync(0xffffffffu, clab >= 0 ? clab : 0x7fffffff);
  const int cmax = __reduce_max_sync(0xffffffffu, clab);
  __shared__ int s_rng[2][4];
  if ((threadIdx.x & 31) == 0) {
    int* q = s_rng[threadIdx.x >> 5];
    q[0] = rmin; q[1] = rmax; q[2] = cmin; q[3] = cmax;
  }
  __syncthreads();
  const int Rmin = min(s_rng[0][0], s_rng[1][0]), Rmax = max(s_rng[0][1], s_rng[1][1]);
  const int Cmin = min(s_rng[0][2], s_rng[1][2]), Cmax = max(s_rng[0][3], s_rng[1][3]);
  const bool overlap = Rmax >= 0 && Cmax >= 0 && Rmin <= Cmax && Cmin <= Rmax;   // block-uniform
  unsigned long long bits = 0ull;
  if (overlap) {
    if (c < N) {
#pragma unroll
      for (int j = 0; j < 7; ++j) s_box[threadIdx.x][j] = boxes[((size_t)s * N + c) * 7 + j];
    }
    __syncthreads();
    if (lab >= 0) {
      float a[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) a[j] = boxes[((size_t)s * N + r) * 7 + j];
      const int j0 = (cb == rb) ? threadIdx.x + 1 : 0;
      for (int j = j0; j < 64; ++j)
        if (s_lab[j] == lab && bev_iou(a, s_box[j]) > thr) bits |= 1ull << j;
    }
  }
  if (r < N) mask[((size_t)s * N + r) * nblk + cb] = bits;
}

// one warp per scene, block-wise greedy sweep in (label, score) order: the 64 boxes of a block are
// resolved serially against each other with register-only work on lane 0 (their 64 diagonal words are
// loaded up front, off the serial chain); the kept boxes' rows are then OR-ed into the removed-bits
// of the later blocks by all lanes in parallel.
__global__ void __launch_bounds__(32)
k_nms_sweep(const unsigned long long* __restrict__ mask, const uint8_t* __restrict__ valid, int N,
            uint8_t* __restrict__ keep) {
  const int s = blockIdx.x, lane = threadIdx.x;
  const int nblk = (N + 63) >> 6;
  constexpr int kW = 4;                 // lane l owns removed-bits words l, l+32, ... (N <= 8192)
  unsigned long long remv[kW];
#pragma unroll
  for (int w = 0; w < kW; ++w) remv[w] = 0ull;
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_kept;
  for (int b = 0; b < nblk; ++b) {
    const int i0 = b * 64;
    // diagonal words of this block's rows + validity bits
    unsigned long long vbits = 0ull;
    for (int j = lane; j < 64; j += 32) {
      const int i = i0 + j;
      s_diag[j] = i < N ? mask[((size_t)s * N + i) * nblk + b] : 0ull;
    }
    const unsigned v0 = __ballot_sync(0xffffffffu, i0 + lane < N && valid[(size_t)s * N + i0 + lane]);
    const unsigned v1 = __ballot_sync(0xffffffffu, i0 + 32 + lane < N && valid[(size_t)s * N + i0 + 32 + lane]);
    vbits = (unsigned long long)v0 | ((unsigned long long)v1 << 32);
    unsigned long long rw = 0ull;
#pragma unroll
    for (int w = 0; w < kW; ++w)
      if ((b >> 5) == w) rw = remv[w];
    rw = __shfl_sync(0xffffffffu, rw, b & 31);
    __syncwarp();
    if (lane == 0) {
      unsigned long long kept = 0ull;
      for (int j = 0; j < 64; ++j) {
        if (((vbits >> j) & 1ull) && !((rw >> j) & 1ull)) {
          kept |= 1ull << j;
          rw |= s_diag[j];
        }
      }
      s_kept = kept;
    }
    __syncwarp();
    const unsigned long long kept = s_kept;
    for (int j = lane; j < 64; j += 32)
      if (i0 + j < N) keep[(size_t)s * N + i0 + j] = (kept >> j) & 1ull;
    // rows of the kept boxes -> removed bits of the later blocks: lane l loads the rows of boxes l
    // and l+32 (independent, pipelined loads), one warp OR-reduction per later word
    const bool k0 = (kept >> lane) & 1ull, k1 = (kept >> (lane + 32)) & 1ull;
    const unsigned long long* row0 = mask + ((size_t)s * N + i0 + lane) * nblk;
    const unsigned long long* row1 = mask + ((size_t)s * N + i0 + lane + 32) * nblk;
    for (int wb = b + 1; wb < nblk; ++wb) {
      unsigned long long v = 0ull;
      if (k0) v |= row0[wb];
      if (k1) v |= row1[wb];
      const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
      const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
      const unsigned long long r = (unsigned long long)lo | ((unsigned long long)hi << 32);
#pragma unroll
      for (int w = 0; w < kW; ++w)
        if ((wb >> 5) == w && (wb & 31) == lane) remv[w] |= r;
    }
    __syncwarp();
  }
}

}  // namespace u3d

using namespace u3d;

extern "C" size_t u3d_nms3d_mask_words(int N) { return N <= 0 ? 0 : (size_t)N * (size_t)((N + 63) / 64); }

extern "C" int u3d_nms3d_bev(const float* boxes, const int32_t* labels, const uint8_t* valid, int B, int N,
                             float iou_threshold, unsigned long long* mask, uint8_t* keep, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  U3D_CHECK_ARG(boxes && labels && valid && mask && keep, "u3d_nms3d_bev: null buffer");
  U3D_CHECK_ARG(B >= 0 && N >= 0 && N <= 64 * 32 * 4, "u3d_nms3d_bev: N=%d (max 8192)", N);
  if (B == 0 || N == 0) return U3D_OK;
  const int nblk = (N + 63) / 64;
  k_nms_mask<<<dim3(nblk, nblk, B), 64, 0, st>>>(boxes, labels, valid, N, iou_threshold, mask);
  U3D_LAUNCH_CHECK();
  k_nms_sweep<<<B, 32, 0, st>>>(mask, valid, N, keep);
  U3D_LAUNCH_CHECK();
  return U3D_OK;
}

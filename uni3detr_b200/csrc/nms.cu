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
#include "common.cuh"

namespace u3d {

struct P2 { float x, y; };

__device__ __forceinline__ void rect_corners(const float* b, P2* c) {
  // b = [x, y, z, dx, dy, dz, heading]; corners counter-clockwise
  const float cs = cosf(b[6]), sn = sinf(b[6]);
  const float hx = 0.5f * b[3], hy = 0.5f * b[4];
  const float lx[4] = {hx, -hx, -hx, hx}, ly[4] = {hy, hy, -hy, -hy};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i].x = b[0] + lx[i] * cs - ly[i] * sn;
    c[i].y = b[1] + lx[i] * sn + ly[i] * cs;
  }
}

// area of (rect a) ∩ (rect b): Sutherland-Hodgman clipping of a's polygon by b's 4 half-planes
__device__ float rect_intersection(const float* a, const float* b) {
  P2 poly[8], tmp[8], cb[4];
  rect_corners(a, poly);
  rect_corners(b, cb);
  int n = 4;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const P2 p0 = cb[e], p1 = cb[(e + 1) & 3];
    const float ex = p1.x - p0.x, ey = p1.y - p0.y;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const P2 s = poly[i], t = poly[(i + 1 == n) ? 0 : i + 1];
      const float ds = ex * (s.y - p0.y) - ey * (s.x - p0.x);   // >= 0: inside (left of the ccw edge)
      const float dt = ex * (t.y - p0.y) - ey * (t.x - p0.x);
      if (ds >= 0.f) tmp[m++] = s;
      if ((ds >= 0.f) != (dt >= 0.f)) {
        const float u = ds / (ds - dt);
        tmp[m].x = s.x + u * (t.x - s.x);
        tmp[m].y = s.y + u * (t.y - s.y);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) poly[i] = tmp[i];
    if (n == 0) return 0.f;
  }
  float area = 0.f;
  for (int i = 0; i < n; ++i) {
    const P2 s = poly[i], t = poly[(i + 1 == n) ? 0 : i + 1];
    area += s.x * t.y - s.y * t.x;
  }
  return 0.5f * fabsf(area);
}

__device__ __forceinline__ float bev_iou(const float* a, const float* b) {
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  const float so = rect_intersection(a, b);
  return so / fmaxf(sa + sb - so, 1e-8f);
}

// grid (col blocks, row blocks, scenes), 64 threads: thread i = row box, 64 column boxes in smem
__global__ void __launch_bounds__(64)
k_nms_mask(const float* __restrict__ boxes, const int32_t* __restrict__ labels,
           const uint8_t* __restrict__ valid, int N, float thr, unsigned long long* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y, s = blockIdx.z;
  if (cb < rb) return;  // strictly-lower blocks are never read
  const int nblk = (N + 63) >> 6;
  __shared__ float s_box[64][7];
  __shared__ int s_lab[64];
  const int c = cb * 64 + threadIdx.x;
  if (c < N) {
#pragma unroll
    for (int j = 0; j < 7; ++j) s_box[threadIdx.x][j] = boxes[((size_t)s * N + c) * 7 + j];
    s_lab[threadIdx.x] = valid[(size_t)s * N + c] ? labels[(size_t)s * N + c] : -1;
  } else {
    s_lab[threadIdx.x] = -1;
  }
  __syncthreads();
  const int r = rb * 64 + threadIdx.x;
  if (r >= N) return;
  unsigned long long bits = 0ull;
  const int lab = valid[(size_t)s * N + r] ? labels[(size_t)s * N + r] : -2;
  if (lab >= 0) {
    float a[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) a[j] = boxes[((size_t)s * N + r) * 7 + j];
    const int j0 = (cb == rb) ? threadIdx.x + 1 : 0;
    for (int j = j0; j < 64; ++j)
      if (s_lab[j] == lab && bev_iou(a, s_box[j]) > thr) bits |= 1ull << j;
  }
  mask[((size_t)s * N + r) * nblk + cb] = bits;
}

// one warp per scene: greedy sweep in (label, score) order
__global__ void __launch_bounds__(32)
k_nms_sweep(const unsigned long long* __restrict__ mask, const uint8_t* __restrict__ valid, int N,
            uint8_t* __restrict__ keep) {
  const int s = blockIdx.x, lane = threadIdx.x;
  const int nblk = (N + 63) >> 6;
  // removed-bits words: lane l owns words l, l+32, ... (N <= 64*32*kW)
  constexpr int kW = 4;
  unsigned long long remv[kW];
#pragma unroll
  for (int w = 0; w < kW; ++w) remv[w] = 0ull;
  for (int i = 0; i < N; ++i) {
    const int word = i >> 6;
    unsigned long long rw = 0ull;
#pragma unroll
    for (int w = 0; w < kW; ++w)
      if ((word >> 5) == w) rw = remv[w];
    rw = __shfl_sync(0xffffffffu, rw, word & 31);
    const bool k = valid[(size_t)s * N + i] && !((rw >> (i & 63)) & 1ull);
    if (lane == 0) keep[(size_t)s * N + i] = k ? 1 : 0;
    if (k) {
      const unsigned long long* row = mask + ((size_t)s * N + i) * nblk;
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        const int b = w * 32 + lane;
        if (b < nblk && b >= word) remv[w] |= row[b];
      }
    }
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

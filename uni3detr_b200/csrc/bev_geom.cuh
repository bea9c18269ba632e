// Rotated-rectangle BEV geometry shared by the NMS (nms.cu) and the training-side IoU kernels (train.cu):
// exact intersection area of two rotated rectangles by Sutherland-Hodgman clipping.
#pragma once
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
static __device__ float rect_intersection(const float* a, const float* b) {
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
  // bounding circles apart -> the rectangles cannot meet (exact: IoU is 0 either way)
  const float dx = a[0] - b[0], dy = a[1] - b[1];
  const float ra = 0.5f * sqrtf(a[3] * a[3] + a[4] * a[4]), rb = 0.5f * sqrtf(b[3] * b[3] + b[4] * b[4]);
  if (dx * dx + dy * dy > (ra + rb) * (ra + rb)) return 0.f;
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  const float so = rect_intersection(a, b);
  return so / fmaxf(sa + sb - so, 1e-8f);
}

}  // namespace u3d

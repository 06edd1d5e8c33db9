// The ONE definition of the DCNv2 sampling position and of its bilinear corner set, shared by every kernel that
// samples (dcn_l1 / dcn_hr in dcn.cu, dcn_tc / dcn_tc3 / dcn_tc3_ws in dcn_tc.cu, the index dump, the backward kernels in
// bwd.cu), so that integer sampling indices agree between all of them by construction.
//
// Semantics (dcn_v2 / torchvision deform_conv2d; call site /root/reference/model/CRFP.py:350):
//   p = (base - 1 + tap) + offset   evaluated as ONE fp32 add of an exact integer and the fp32 offset;
//   the sample is 0 if p <= -1 or p >= size; otherwise 4-corner bilinear, corners outside the image contribute 0.
// Plain C++ (floorf only): also compiled by g++ for the CPU host emulation of bwd.cu (tests/tools/hostemu).
#pragma once
#include <math.h>

namespace crfp {

__device__ __forceinline__ float dcn_pos(int base, int tap, float offset) { return (float)(base - 1 + tap) + offset; }

__device__ __forceinline__ void dcn_corner_w(float py, float px, int H, int W, int& y0, int& x0, float& w00, float& w01,
                                             float& w10, float& w11) {
  const float fy = floorf(py), fx = floorf(px);
  y0 = (int)fy; x0 = (int)fx;
  const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
  const bool inside = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
  const bool vy0 = inside && y0 >= 0, vy1 = inside && (y0 + 1 <= H - 1);
  const bool vx0 = x0 >= 0, vx1 = (x0 + 1 <= W - 1);
  w00 = (vy0 && vx0) ? hy * hx : 0.f;
  w01 = (vy0 && vx1) ? hy * lx : 0.f;
  w10 = (vy1 && vx0) ? ly * hx : 0.f;
  w11 = (vy1 && vx1) ? ly * lx : 0.f;
}

}  // namespace crfp

// aug_geom.cuh -- the reference's resampling geometry, shared by the dense gather (augment.cu) and the fused
// sparse tile kernel (lidar_aug_tile.cuh).  Exact restatements of OpenCV's coordinate arithmetic:
//   cv2.warpAffine INTER_LINEAR (reference utils.py:430-433): imgwarp.cpp WarpAffineInvoker, AB_BITS 10, INTER_BITS 5
//   cv2.resize INTER_LINEAR     (reference utils.py:462):     resize.cpp, float32 coordinates and weights
#pragma once

#include "bevgen_common.cuh"

namespace bevgen {

// warpAffine: the source coordinate of destination (x, y) on the 1/1024-pixel grid is X0(y) + adelta(x), Y0(y) + bdelta(x)
// with M = the inverted 2x3 matrix (dst -> src); the +16 is round_delta = AB_SCALE / INTER_TAB_SIZE / 2.
__device__ __forceinline__ int rot_X0(const double* __restrict__ M, int y) {
    return __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
}
__device__ __forceinline__ int rot_Y0(const double* __restrict__ M, int y) {
    return __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
}
__device__ __forceinline__ int rot_adelta(const double* __restrict__ M, int x) {
    return __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
}
__device__ __forceinline__ int rot_bdelta(const double* __restrict__ M, int x) {
    return __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
}
// (X, Y) = coordinates on the 1/32-pixel grid -> the 4 float32 bilinear weights of remapBilinear's table
__device__ __forceinline__ void rot_weights(int X, int Y, float (&w)[4]) {
    const float fx = __fmul_rn((float)(X & 31), 0.03125f), fy = __fmul_rn((float)(Y & 31), 0.03125f);
    const float tx0 = __fsub_rn(1.f, fx), ty0 = __fsub_rn(1.f, fy);
    w[0] = __fmul_rn(ty0, tx0); w[1] = __fmul_rn(ty0, fx); w[2] = __fmul_rn(fy, tx0); w[3] = __fmul_rn(fy, fx);
}
// 4-tap sum, left to right, no FMA
__device__ __forceinline__ float rot_sum(float v0, float v1, float v2, float v3, const float (&w)[4]) {
    float r = __fmul_rn(v0, w[0]);
    r = __fadd_rn(r, __fmul_rn(v1, w[1]));
    r = __fadd_rn(r, __fmul_rn(v2, w[2]));
    r = __fadd_rn(r, __fmul_rn(v3, w[3]));
    return r;
}

// cv2.resize INTER_LINEAR source taps for destination index d (resize.cpp): f = (float)((d+0.5)*scale-0.5)
__device__ __forceinline__ void resize_tap(int d, int dsize, int ssize, int& s0, int& s1, float& w0, float& w1) {
    const double inv_scale = __ddiv_rn((double)dsize, (double)ssize);
    const double scale = __ddiv_rn(1.0, inv_scale);
    float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    s0 = s;
    s1 = s + 1 < ssize ? s + 1 : ssize - 1;
    w0 = __fsub_rn(1.f, f);
    w1 = f;
}
// horizontal pass then vertical pass, float32, no FMA (resize.cpp HResize / VResize)
__device__ __forceinline__ float resize_sum(float t00, float t01, float t10, float t11, float wx0, float wx1, float wy0, float wy1) {
    const float top = __fadd_rn(__fmul_rn(t00, wx0), __fmul_rn(t01, wx1));
    const float bot = __fadd_rn(__fmul_rn(t10, wx0), __fmul_rn(t11, wx1));
    return __fadd_rn(__fmul_rn(top, wy0), __fmul_rn(bot, wy1));
}

// the resized image's placement inside the 400 x 720 output: utils.py:457 (new size), :464-471 (centre crop / zero pad)
struct ScalePlacement { int nh, nw, h0, w0; bool crop; };
__device__ __forceinline__ ScalePlacement scale_placement(double scale) {
    ScalePlacement p;
    p.nh = (int)__dmul_rn((double)kH, scale);
    p.nw = (int)__dmul_rn((double)kW, scale);
    p.crop = scale > 1.0;
    p.h0 = p.crop ? (p.nh - kH) / 2 : (kH - p.nh) / 2;
    p.w0 = p.crop ? (p.nw - kW) / 2 : (kW - p.nw) / 2;
    return p;
}

}  // namespace bevgen

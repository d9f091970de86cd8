// augment.cu -- train-time BEV augmentation: flip -> rotate -> scale -> dropout in ONE gather pass (sm_100a).
//
// Replaces the image part of reference utils.py:394-498 (augment_bev, utils.py:500-517, applies them in this
// order).  The reference materialises up to four full copies of the 299-channel grids (np.flip, 299 x
// cv2.warpAffine + np.stack, 299 x cv2.resize + np.stack); here every output pixel is produced once:
//
//   out(y,x) = 0                                   if (y,x) is inside a dropout rectangle       utils.py:496
//            = S(y,x),  S = resize+crop/pad of R   (cv2.resize INTER_LINEAR, utils.py:457-471)  if do_scale
//            = R(y,x),  R = warpAffine of F        (cv2.warpAffine INTER_LINEAR, utils.py:430-433) if do_rotate
//            = F(y,x) = src(y, 719-x) or src(y,x)  (np.flip axis 2, utils.py:399)
//
// Exactness: R is a deterministic float32 function of F (fixed-point coordinates on a 1/32-pixel grid,
// float32 weights, 4-tap sum evaluated left to right without FMA), so recomputing the <= 4 R values a
// scaled pixel needs reproduces the reference's intermediate rounding bit for bit -- no intermediate
// image, 1 read + 1 write of the grids instead of 3 + 3.
// Geometry (tap offsets, weights) depends only on (y,x), not on the channel: it is computed once per
// thread and reused for all C channels of the frame.
#include "aug_geom.cuh"

namespace bevgen {
namespace {

constexpr int kAugThreads = 240;                 // 3 blocks per 720-pixel row

struct RotTap {                                  // 4 taps of cv2's remapBilinear for one destination pixel
    int off[4];                                  // offset into the source plane, -1 = outside (BORDER_CONSTANT 0)
    float w[4];
};

__device__ __forceinline__ int flip_x(int x, bool flip) { return flip ? kW - 1 - x : x; }

// Source of the gather: a float32 grid, or the map rasterizer's bit planes (values 0 / 1; an "offset" is then a bit index
// in the 768-bit-row layout) -- same arithmetic downstream, no intermediate float grid for the map channels.
template <bool BITS> struct AugSrc;
template <> struct AugSrc<false> {
    using T = float;
    static constexpr int kRow = kW, kPlaneElems = kPlane;
    static __device__ __forceinline__ float ld(const float* __restrict__ plane, int off) { return __ldg(plane + off); }
};
template <> struct AugSrc<true> {
    using T = uint32_t;
    static constexpr int kRow = kMapWordsPerRow * 32, kPlaneElems = kMapPlaneWords;
    static __device__ __forceinline__ float ld(const uint32_t* __restrict__ plane, int off) {
        return ((__ldg(plane + (off >> 5)) >> (off & 31)) & 1u) ? 1.0f : 0.0f;
    }
};

// cv2.warpAffine coordinate scheme (aug_geom.cuh)
template <int ROWSTRIDE>
__device__ __forceinline__ RotTap rot_taps(const double* __restrict__ M, int y, int x, bool flip) {
    const int X = (rot_X0(M, y) + rot_adelta(M, x)) >> 5, Y = (rot_Y0(M, y) + rot_bdelta(M, x)) >> 5;
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx)); sy = max(-32768, min(32767, sy));      // saturate_cast<short>
    RotTap t;
    rot_weights(X, Y, t.w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = sx + (k & 1), yy = sy + (k >> 1);
        t.off[k] = (xx >= 0 && xx < kW && yy >= 0 && yy < kH) ? yy * ROWSTRIDE + flip_x(xx, flip) : -1;
    }
    return t;
}

template <bool BITS>
__device__ __forceinline__ float rot_eval(const typename AugSrc<BITS>::T* __restrict__ plane, const RotTap& t) {
    const float v0 = t.off[0] >= 0 ? AugSrc<BITS>::ld(plane, t.off[0]) : 0.f;
    const float v1 = t.off[1] >= 0 ? AugSrc<BITS>::ld(plane, t.off[1]) : 0.f;
    const float v2 = t.off[2] >= 0 ? AugSrc<BITS>::ld(plane, t.off[2]) : 0.f;
    const float v3 = t.off[3] >= 0 ? AugSrc<BITS>::ld(plane, t.off[3]) : 0.f;
    return rot_sum(v0, v1, v2, v3, t.w);
}

struct FrameList {                               // frames of one class, passed by value
    int n;
    int use_flip, use_rects;                     // 0 = ignore the frame's flip / dropout in this pass (two-pass rotate+scale)
    short src[60], dst[60], par[60];             // frame slot read, frame slot written, index into params
};

// One thread per output pixel (x fastest), looping over the C channels of its frame.  ROT / SCL are uniform per
// launch (the host groups frames by class), so the channel loop carries no dead code: 1, 4, 4 or 16 taps per value.
template <bool ROT, bool SCL, bool BITS>
__global__ void __launch_bounds__(kAugThreads) k_augment(const typename AugSrc<BITS>::T* __restrict__ src, float* __restrict__ dst, int C,
                                                         const bevgen_aug_params* __restrict__ params, FrameList fl) {
    using Src = AugSrc<BITS>;
    using SrcT = typename Src::T;
    const int x = blockIdx.x * kAugThreads + threadIdx.x, y = blockIdx.y;
    const int bsrc = fl.src[blockIdx.z], bdst = fl.dst[blockIdx.z];
    __shared__ bevgen_aug_params P;
    {
        const int* g = reinterpret_cast<const int*>(params + fl.par[blockIdx.z]);
        int* s = reinterpret_cast<int*>(&P);
        for (int i = threadIdx.x; i < (int)(sizeof(bevgen_aug_params) / 4); i += kAugThreads) s[i] = g[i];
    }
    __syncthreads();
    const bool flip = P.do_flip != 0 && fl.use_flip;
    int nrect = fl.use_rects ? P.n_rects : 0;
    nrect = nrect > kMaxRects ? kMaxRects : (nrect < 0 ? 0 : nrect);
    const bool dropped = nrect && in_rects(P.rects, nrect, y, x);

    // ---- geometry: which R (or F) values does this output pixel need, with which weights -------------
    constexpr int NQ = SCL ? 4 : 1;
    int qy[NQ], qx[NQ];
    qy[0] = y; qx[0] = x;
    float wx0 = 1.f, wx1 = 0.f, wy0 = 1.f, wy1 = 0.f;
    bool outside = false;                          // scaled-down frame: zero padding around the resized image
    if (SCL) {
        const ScalePlacement sp_ = scale_placement(P.scale);                                          // utils.py:457, :464-471
        const int nh = sp_.nh, nw = sp_.nw, h0 = sp_.h0, w0 = sp_.w0;
        const bool crop = sp_.crop;
        const int ry = crop ? y + h0 : y - h0, rx = crop ? x + w0 : x - w0;
        int sx0 = 0, sx1 = 0, sy0 = 0, sy1 = 0;
        if (ry < 0 || ry >= nh || rx < 0 || rx >= nw) {
            outside = true;
        } else {
            resize_tap(rx, nw, kW, sx0, sx1, wx0, wx1);
            resize_tap(ry, nh, kH, sy0, sy1, wy0, wy1);
        }
        qy[0] = sy0; qx[0] = sx0;
        if (NQ == 4) { qy[1] = sy0; qx[1] = sx1; qy[2] = sy1; qx[2] = sx0; qy[3] = sy1; qx[3] = sx1; }
    }
    const size_t pix = (size_t)y * kW + x;
    const SrcT* sp = src + (size_t)bsrc * C * Src::kPlaneElems;
    float* dp = dst + (size_t)bdst * C * kPlane + pix;
    if (dropped || outside) {
#pragma unroll 8
        for (int c = 0; c < C; ++c) __stcs(dp + (size_t)c * kPlane, 0.f);
        return;
    }
    RotTap tap[ROT ? NQ : 1];
    int direct[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        direct[q] = qy[q] * Src::kRow + flip_x(qx[q], flip);
        if (ROT) tap[q] = rot_taps<Src::kRow>(P.minv, qy[q], qx[q], flip);
    }

    // ---- channel loop: U channels per iteration, all loads issued before the first use ------------------
    constexpr int U = (ROT && SCL) ? 2 : 4;
    auto value = [&](const float (&r)[NQ]) -> float {
        if (!SCL) return r[0];
        return resize_sum(r[0], r[NQ > 1 ? 1 : 0], r[NQ > 2 ? 2 : 0], r[NQ > 3 ? 3 : 0], wx0, wx1, wy0, wy1);
    };
    int c = 0;
    for (; c + U <= C; c += U) {
        float v[U][NQ][ROT ? 4 : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const SrcT* plane = sp + (size_t)(c + u) * Src::kPlaneElems;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (ROT) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[u][q][k] = tap[q].off[k] >= 0 ? Src::ld(plane, tap[q].off[k]) : 0.f;
                } else {
                    v[u][q][0] = Src::ld(plane, direct[q]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float r[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (ROT) {
                    r[q] = rot_sum(v[u][q][0], v[u][q][1], v[u][q][2], v[u][q][3], tap[q].w);
                } else {
                    r[q] = v[u][q][0];
                }
            }
            __stcs(dp + (size_t)(c + u) * kPlane, value(r));
        }
    }
    for (; c < C; ++c) {
        const SrcT* plane = sp + (size_t)c * Src::kPlaneElems;
        float r[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) r[q] = ROT ? rot_eval<BITS>(plane, tap[q]) : Src::ld(plane, direct[q]);
        __stcs(dp + (size_t)c * kPlane, value(r));
    }
}

}  // namespace
}  // namespace bevgen

// map bit planes -> augmented float32 map grids for the frames that rotate / scale (one fused gather per class; the
// 16-tap form for rotate + scale: 9 channels of 0 / 1 values need no scratch pass)
int bevgen::launch_augment_bitplanes(const uint32_t* planes, float* dst, int B, const bevgen_aug_params* params,
                                     const bevgen_aug_params* params_host, cudaStream_t st) {
    if (B <= 0 || !planes || !dst || !params || !params_host) return BEVGEN_E_INVALID_ARG;
    if (B > 32767) return BEVGEN_E_UNSUPPORTED;
    for (int cls = 1; cls < 4; ++cls) {
        FrameList fl;
        fl.n = 0; fl.use_flip = 1; fl.use_rects = 1;
        auto flush = [&]() -> int {
            if (fl.n == 0) return 0;
            const dim3 grid(kW / kAugThreads, kH, fl.n);
            switch (cls) {
                case 1:  BEVGEN_LAUNCH("k_augment", st, k_augment<true, false, true><<<grid, kAugThreads, 0, st>>>(planes, dst, kMapC, params, fl)); break;
                case 2:  BEVGEN_LAUNCH("k_augment", st, k_augment<false, true, true><<<grid, kAugThreads, 0, st>>>(planes, dst, kMapC, params, fl)); break;
                default: BEVGEN_LAUNCH("k_augment", st, k_augment<true, true, true><<<grid, kAugThreads, 0, st>>>(planes, dst, kMapC, params, fl)); break;
            }
            fl.n = 0;
            return 0;
        };
        for (int b = 0; b < B; ++b) {
            if (((params_host[b].do_rotate != 0) | ((params_host[b].do_scale != 0) << 1)) == cls) {
                fl.src[fl.n] = fl.dst[fl.n] = fl.par[fl.n] = (short)b;
                if (++fl.n == 60) { const int rc = flush(); if (rc) return rc; }
            }
        }
        const int rc = flush();
        if (rc) return rc;
    }
    return 0;
}

using namespace bevgen;

extern "C" int bevgen_augment(const float* src, float* dst, int B, int C, const bevgen_aug_params* params,
                              const bevgen_aug_params* params_host, void* workspace, size_t workspace_bytes,
                              bevgen_stream_t stream) {
    if (B <= 0 || C <= 0 || !src || !dst || !params || !params_host || src == dst) return BEVGEN_E_INVALID_ARG;
    if (B > 32767) return BEVGEN_E_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t frame_bytes = (size_t)C * kPlane * sizeof(float);
    // rotate+scale frames: two 4-tap passes through a scratch frame are faster than the fused 16-tap gather (which is
    // L1-throughput bound); the fused kernel stays as the exact fallback when no scratch is supplied.
    const int scratch_slots = workspace ? (int)(workspace_bytes / frame_bytes) : 0;
    float* ws = static_cast<float*>(workspace);
    auto launch = [&](int cls, const float* s, float* d, const FrameList& fl) -> int {
        const dim3 grid(kW / kAugThreads, kH, fl.n);
        switch (cls) {
            case 3:  BEVGEN_LAUNCH("k_augment", st, k_augment<true, true, false><<<grid, kAugThreads, 0, st>>>(s, d, C, params, fl)); break;
            case 1:  BEVGEN_LAUNCH("k_augment", st, k_augment<true, false, false><<<grid, kAugThreads, 0, st>>>(s, d, C, params, fl)); break;
            case 2:  BEVGEN_LAUNCH("k_augment", st, k_augment<false, true, false><<<grid, kAugThreads, 0, st>>>(s, d, C, params, fl)); break;
            default: BEVGEN_LAUNCH("k_augment", st, k_augment<false, false, false><<<grid, kAugThreads, 0, st>>>(s, d, C, params, fl)); break;
        }
        return 0;
    };
    // group frames by class (bit0 rotate, bit1 scale) so each launch runs a specialised channel loop
    for (int cls = 0; cls < 4; ++cls) {
        const int chunk = (cls == 3 && scratch_slots > 0) ? (scratch_slots < 60 ? scratch_slots : 60) : 60;
        FrameList fl;
        fl.n = 0; fl.use_flip = 1; fl.use_rects = 1;
        auto flush = [&]() -> int {
            if (fl.n == 0) return 0;
            int rc;
            if (cls == 3 && scratch_slots > 0) {
                FrameList p1 = fl, p2 = fl;
                p1.use_rects = 0; p2.use_flip = 0;
                for (int j = 0; j < fl.n; ++j) { p1.dst[j] = (short)j; p2.src[j] = (short)j; }
                if ((rc = launch(1, src, ws, p1)) != 0) return rc;     // flip + rotate  -> scratch
                rc = launch(2, ws, dst, p2);                            // scale + dropout -> dst
            } else {
                rc = launch(cls, src, dst, fl);
            }
            fl.n = 0;
            return rc;
        };
        for (int b = 0; b < B; ++b) {
            if (((params_host[b].do_rotate != 0) | ((params_host[b].do_scale != 0) << 1)) == cls) {
                fl.src[fl.n] = fl.dst[fl.n] = fl.par[fl.n] = (short)b;
                if (++fl.n == chunk) { const int rc = flush(); if (rc) return rc; }
            }
        }
        const int rc = flush();
        if (rc) return rc;
    }
    return 0;
}

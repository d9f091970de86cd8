// lidar_aug_tile.cuh -- LiDAR generation FUSED with the resampling augmentations (rotate / scale), sm_100a.
//
// Replaces, for frames whose augmentation draw rotates and/or scales (reference utils.py:417-480), the sequence
//   create_intentnet_lidar_bev (utils.py:62-106) -> np.flip (:399) -> 290 x cv2.warpAffine (:430-433)
//   -> 290 x cv2.resize + crop/pad (:457-471) -> dropout rectangles (:496)
// without ever materialising the un-augmented 334 MB grid: the dense two-pass path writes it, reads it back through a
// 4-tap gather and writes the result (3 x 334 MB of HBM traffic per frame, LSU-bound); here the augmented grid is built
// directly from the point records of k_bin_fixed and written ONCE.
//
// The un-augmented grid is > 99 % zeros, and an output pixel of a bilinear resampler is non-zero only if one of its taps
// is.  The output is cut into 40 x 72 tiles; the source pixels a tile can read form its FOOTPRINT (bounding box of its
// taps, exact, from the endpoints of cv2's monotone coordinate tables).
//   k_rebin_aug   the row-band record bins of k_bin_fixed -> one bin per (sweep, height channel, output tile) holding the
//                 records inside that tile's footprint (a record lands in 1-4 tiles), including the channel's records in
//                 the sweep's overflow list.  One CTA per (frame, sweep, channel) is the only writer of its 100 bins:
//                 slots come from shared-memory counters, no global atomics.  A record is stored as the tile kernel's
//                 work-list entry: position inside the footprint + origin of the rotation's candidate window.
//   k_tile_aug    one CTA per (frame, sweep, output tile) walks the sweep's 29 height channels:
//     A  the tile's records of the channel (fetched two channels ahead with cp.async) -> max into a shared-memory copy T0
//        of the footprint (the mirrored column of a flipped frame is already in the record)
//     B  rotation: for every (record, one of 3 x 3 candidate destination pixels around the forward-mapped voxel) pair, in
//        parallel over the CTA: if the pixel's cv2 fixed-point tap set really contains the voxel it is evaluated from T0
//        with the reference's exact arithmetic (aug_geom.cuh).  Resize: the taps are separable and monotone, so an inverse
//        tap table gives exactly the output rows / columns a source pixel feeds (work items: 3 rows per record for
//        scale-only frames, 2 x 2 pixels per listed pixel after a rotation).  Results go into the shared output tile O;
//        rotate + scale goes through an intermediate tile T1 (the rotated image's float32 rounding is part of the
//        reference's result) whose non-zero pixels are listed as they are produced.
//     C  O -> HBM (40 row pieces of 288 B, 16-byte stores; measured faster here than 40 small bulk copies), then only the
//        touched entries of T0 / T1 are cleared again.
// Geometry tables (warpAffine's X0/Y0/adelta/bdelta, resize taps and weights) are built once per CTA and reused for all
// 29 channels.  Dense scenes: a (channel, tile) bin that exceeds its capacity (256 records) takes the SLOW path for that
// channel: scan the row-band bins and the sweep's overflow list, then scan T0 for its non-zeros.  Slower, same bits.
//
// Included by lidar.cu INSIDE namespace bevgen::{anonymous}, after the tiled path's bin layout (kBinCap, kTilesPerSweep,
// kTileElems, kFillRows, OvfRec); lidar.cu includes aug_geom.cuh at file scope first.
#pragma once

// Tile shape: 40 rows x 256 threads (4 CTAs / SM) measured against 20 rows x 128 threads (8 CTAs / SM) on B200, us per
// 290-channel frame: rotate 90 vs 98, scale 76 vs 73, rotate+scale 175 vs 187, re-binning 10 vs 13-16.
#ifndef BEVGEN_AT_H
#define BEVGEN_AT_H 40
#define BEVGEN_AT_THREADS 256
#endif
constexpr int kATW = 72, kATH = BEVGEN_AT_H;                 // output tile (columns x rows): 288 B row pieces
constexpr int kATBX = kW / kATW, kATBY = kH / kATH;          // 10 x 10 tiles per channel
constexpr int kATTiles = kATBX * kATBY;
constexpr int kATElems = kATW * kATH;                        // 2880 floats
constexpr int kATThreads = BEVGEN_AT_THREADS;
constexpr int kRebinThreads = 256;
constexpr int kATMaxBands = 8;                               // source row-bands under one footprint (host-checked)
constexpr int kATCap = kATThreads;                           // records per (channel, tile) bin: one per thread
constexpr int kBandsPerCh = kH / kFillRows;                  // 40 row-band bins per (sweep, channel)
static_assert(kW % kATW == 0 && kH % kATH == 0, "tiles must cover the grid");
static_assert(kATThreads >= kBinCap, "the slow path reads one row-band record per thread");
static_assert((kATW * 4) % 16 == 0, "row pieces must be 16-byte multiples");

template <bool ROT, bool SCL> struct AugTileCfg {
    // capacities (floats); the host checks a conservative bound per frame (aug_tile_supported), the kernel traps beyond
    static constexpr int kDW = SCL ? 80 : kATW;                    // rotation destination region (columns, rows)
    static constexpr int kDH = SCL ? (kATH == 40 ? 48 : 26) : kATH;
    // source footprint / rotated intermediate (= the resize's source region)
    static constexpr int kT0 = kATH == 40 ? (ROT ? (SCL ? 6400 : 5120) : 3840) : (ROT ? (SCL ? 4096 : 3584) : 2080);
    static constexpr int kT1 = (ROT && SCL) ? kDW * kDH : 0;
    // output tiles: the store of channel z is ONE asynchronous tensor-map copy, so a second tile lets channel z + 1 be built
    // while it drains; rotate + scale has no room for one (3 CTAs / SM) and does not need it: its O is only written in the
    // second half of a channel's work
    static constexpr int kOBufs = (ROT && SCL) ? 1 : 2;
    static constexpr int kSmem = (kOBufs * kATElems + kT0 + kT1) * 4 + kT1 * 2;   // O (x kOBufs), T0, T1, list of T1's non-zeros (u16)
    static constexpr int kMinBlocks = kATThreads == 256 ? ((ROT && SCL) ? 3 : 4) : ((ROT && SCL) ? 6 : 8);
};

struct AugTileFrames {                                       // frames of one class, passed by value
    int n;
    // frame index inside the binning launch (= params index), output slot, slot in the per-tile bins
    short src[60], dst[60], slot[60];
};

struct AugTileGeo {
    int dx0, dy0, DW, DH;            // rotation destination region (ROT); == the resize source region when ROT && SCL
    int qx0, qy0, QW, QH;            // resize source region (SCL)
    int fx0, fy0, FW, FH;            // T0 footprint, clipped to the grid (FW = 0: nothing of the tile maps into the grid)
};

struct AugTileTables {
    int X0[48], Y0[48], A[80], B[80];                        // warpAffine, over the rotation destination region
    short s0x[kATW], s1x[kATW], s0y[kATH], s1y[kATH];        // resize taps per output column / row (s0 < 0: outside the image)
    float wx0[kATW], wx1[kATW], wy0[kATH], wy1[kATH];
    // inverse of the resize taps: source column qx0 + i feeds the output columns [xf[i], xl[i]] of the tile (taps are
    // monotone, so it is a range; empty = (127, -1)); same for rows
    int xf[80], xl[80], yf[48], yl[48];
};

// forward (source -> destination) maps; approximate: they place the rotation's 3 x 3 candidate window and pick the
// candidate output tiles of a record, the exact tests decide
struct RotFwd { float f00, f01, f02, f10, f11, f12, hx, hy; };
struct SclFwd { float sfx, sfy, offx, offy; };
constexpr int kATCand = 9;                                   // rotation candidates per source voxel (3 x 3; kernel traps unless hx, hy < 1.5)
constexpr int kATOrgBias = 64;

__device__ __forceinline__ RotFwd make_rot_fwd(const double* __restrict__ M) {
    // inverse of minv (dst -> src); float is ample
    const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5];
    const double det = a * e - b * d, id = det != 0.0 ? 1.0 / det : 0.0;
    RotFwd F;
    F.f00 = (float)(e * id); F.f01 = (float)(-b * id); F.f10 = (float)(-d * id); F.f11 = (float)(a * id);
    F.f02 = (float)(-(e * id * c - b * id * f)); F.f12 = (float)(-(-d * id * c + a * id * f));
    // |dst - centre| <= |F| * 1.02 (taps within +-1 of the voxel, plus the fixed-point rounding of the coordinates)
    F.hx = (fabsf(F.f00) + fabsf(F.f01)) * 1.03f + 0.05f;
    F.hy = (fabsf(F.f10) + fabsf(F.f11)) * 1.03f + 0.05f;
    return F;
}
__device__ __forceinline__ SclFwd make_scl_fwd(const ScalePlacement& pl) {
    // destination x of source column u: (u + 0.5) nw/720 - 0.5 in resize coordinates, then the crop / pad shift
    SclFwd F;
    F.sfx = (float)pl.nw / (float)kW; F.sfy = (float)pl.nh / (float)kH;
    F.offx = pl.crop ? -(float)pl.w0 : (float)pl.w0; F.offy = pl.crop ? -(float)pl.h0 : (float)pl.h0;
    return F;
}

// The regions of output tile (ox0, oy0): exact, from the endpoints of cv2's monotone coordinate tables.  Both kernels
// call this same function, so the re-binning and the tile kernel agree on every footprint.
template <bool ROT, bool SCL>
__device__ __forceinline__ AugTileGeo aug_tile_geo(const bevgen_aug_params& P, int ox0, int oy0) {
    AugTileGeo G;
    G.dx0 = ox0; G.dy0 = oy0; G.DW = kATW; G.DH = kATH;
    G.qx0 = G.qy0 = 0; G.QW = G.QH = 0;
    G.fx0 = G.fy0 = 0; G.FW = G.FH = 0;
    if (SCL) {
        const ScalePlacement pl = scale_placement(P.scale);
        // output column ox0 + l reads resized column r = lo + l; valid while 0 <= r < nw (utils.py:464-471)
        const int lox = pl.crop ? ox0 + pl.w0 : ox0 - pl.w0, loy = pl.crop ? oy0 + pl.h0 : oy0 - pl.h0;
        const int c0 = max(0, -lox), c1 = min(kATW - 1, pl.nw - 1 - lox);
        const int r0 = max(0, -loy), r1 = min(kATH - 1, pl.nh - 1 - loy);
        if (c0 > c1 || r0 > r1) { if (ROT) { G.DW = G.DH = 0; } return G; }
        int sa, sb, dummy;
        float w0, w1;
        // taps are monotone in the destination index: the region is [first s0, last s1]
        resize_tap(lox + c0, pl.nw, kW, sa, dummy, w0, w1);
        resize_tap(lox + c1, pl.nw, kW, dummy, sb, w0, w1);
        G.qx0 = sa; G.QW = sb - sa + 1;
        resize_tap(loy + r0, pl.nh, kH, sa, dummy, w0, w1);
        resize_tap(loy + r1, pl.nh, kH, dummy, sb, w0, w1);
        G.qy0 = sa; G.QH = sb - sa + 1;
        if (ROT) { G.dx0 = G.qx0; G.dy0 = G.qy0; G.DW = G.QW; G.DH = G.QH; }
        else { G.fx0 = G.qx0; G.fy0 = G.qy0; G.FW = G.QW; G.FH = G.QH; }
    }
    if (ROT) {
        // X0(y) and adelta(x) are monotone: the extremes sit at the region's corners
        const int x0a = rot_X0(P.minv, G.dy0), x0b = rot_X0(P.minv, G.dy0 + G.DH - 1);
        const int y0a = rot_Y0(P.minv, G.dy0), y0b = rot_Y0(P.minv, G.dy0 + G.DH - 1);
        const int aa = rot_adelta(P.minv, G.dx0), ab = rot_adelta(P.minv, G.dx0 + G.DW - 1);
        const int ba = rot_bdelta(P.minv, G.dx0), bb = rot_bdelta(P.minv, G.dx0 + G.DW - 1);
        const int xa = min(x0a, x0b) + min(aa, ab), xb = max(x0a, x0b) + max(aa, ab);
        const int ya = min(y0a, y0b) + min(ba, bb), yb = max(y0a, y0b) + max(ba, bb);
        const int fx0 = max(xa >> 10, 0), fx1 = min((xb >> 10) + 1, kW - 1);
        const int fy0 = max(ya >> 10, 0), fy1 = min((yb >> 10) + 1, kH - 1);
        if (fx1 >= fx0 && fy1 >= fy0) { G.fx0 = fx0; G.fy0 = fy0; G.FW = fx1 - fx0 + 1; G.FH = fy1 - fy0 + 1; }
    }
    return G;
}

// A work-list entry packs the voxel's footprint position (7 + 7 bits: the footprint is at most 127 x 127, checked) and,
// for the rotation, the origin of its candidate window relative to the destination region (8 + 8 bits, biased).  The
// re-binning kernel writes the entry into the per-tile record, so the tile kernel only unpacks it.
__device__ __forceinline__ uint32_t at_pack_pos(int fx, int fy) { return ((uint32_t)fy << 7) | (uint32_t)fx; }
__device__ __forceinline__ int at_pos_x(uint32_t w) { return (int)(w & 127u); }
__device__ __forceinline__ int at_pos_y(uint32_t w) { return (int)((w >> 7) & 127u); }
__device__ __forceinline__ uint32_t rot_window_origin_at(float cx, float cy, const RotFwd& F, int dx0, int dy0) {
    const int xr = min(max((int)ceilf(cx - F.hx) - dx0 + kATOrgBias, 0), 255);   // clamped windows lie outside the region
    const int yr = min(max((int)ceilf(cy - F.hy) - dy0 + kATOrgBias, 0), 255);
    return ((uint32_t)yr << 22) | ((uint32_t)xr << 14);
}
__device__ __forceinline__ uint32_t rot_window_origin(int u, int v, const RotFwd& F, const AugTileGeo& G) {
    const float cx = fmaf(F.f00, (float)u, fmaf(F.f01, (float)v, F.f02)), cy = fmaf(F.f10, (float)u, fmaf(F.f11, (float)v, F.f12));
    return rot_window_origin_at(cx, cy, F, G.dx0, G.dy0);
}

// ---- re-binning: row-band bins -> per-(channel, output tile) bins ---------------------------------------------------
// grid (29 channels, S sweeps, resampled frames of ANY class: one launch per batch -- the kernel is one wave of latency-bound
// CTAs, three per-class launches cost three such waves); record = (work-list entry of the source pixel for that tile,
// intensity bits)
__global__ void __launch_bounds__(kRebinThreads, 4) k_rebin_aug(const uint32_t* __restrict__ cnt, const uint2* __restrict__ bins,
                                                          const uint32_t* __restrict__ ovf_cnt, const OvfRec* __restrict__ ovf,
                                                          const int32_t* __restrict__ sweep_off,
                                                          uint32_t* __restrict__ cnt2, uint2* __restrict__ bins2,
                                                          const bevgen_aug_params* __restrict__ params, AugTileFrames fl, int S,
                                                          int rec4) {
    __shared__ bevgen_aug_params P;
    __shared__ int4 box[kATTiles];                           // footprint of each output tile: x0, y0, x1, y1 (inclusive)
    __shared__ short2 dorg[kATTiles];                        // origin of its rotation destination region
    __shared__ uint32_t hist[kATTiles];
    __shared__ int s_over;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int zb = blockIdx.x, sweep = blockIdx.y;
    const int fsrc = fl.src[blockIdx.z], slot = fl.slot[blockIdx.z];
    {
        const int* g = reinterpret_cast<const int*>(params + fsrc);
        int* s = reinterpret_cast<int*>(&P);
        for (int i = tid; i < (int)(sizeof(bevgen_aug_params) / 4); i += kRebinThreads) s[i] = g[i];
    }
    if (tid == 0) s_over = 0;
    __syncthreads();
    const bool ROT = P.do_rotate != 0, SCL = P.do_scale != 0;     // the frame's class (uniform per CTA)
    for (int t = tid; t < kATTiles; t += kRebinThreads) {
        const int tx0 = (t % kATBX) * kATW, ty0 = (t / kATBX) * kATH;
        const AugTileGeo G = ROT ? (SCL ? aug_tile_geo<true, true>(P, tx0, ty0) : aug_tile_geo<true, false>(P, tx0, ty0))
                                 : aug_tile_geo<false, true>(P, tx0, ty0);
        box[t] = make_int4(G.fx0, G.fy0, G.fx0 + G.FW - 1, G.fy0 + G.FH - 1);
        dorg[t] = make_short2((short)G.dx0, (short)G.dy0);
        hist[t] = 0;
    }
    constexpr int kCandY = (36 + kATH - 1) / kATH;           // 1 tile row either side at 40-row tiles, 2 at 20-row tiles
    RotFwd RF = {};
    SclFwd SF = {};
    if (ROT) RF = make_rot_fwd(P.minv);
    if (SCL) SF = make_scl_fwd(scale_placement(P.scale));
    __syncthreads();
    const size_t bin_base = ((size_t)(fsrc * S + sweep) * kZ + zb) * kBandsPerCh;
    uint2* const out_base = bins2 + (((size_t)(slot * S + sweep) * kZ + zb) * kATTiles) * kATCap;
    // a warp owns the row bands warp, warp + 8, ...: all their counts, then all their records, are requested before the
    // first one is used (the kernel is one wave of latency-bound CTAs)
    constexpr int kWarps = kRebinThreads / 32, kPerWarp = (kBandsPerCh + kWarps - 1) / kWarps;   // 5
    uint32_t nb[kPerWarp];
#pragma unroll
    for (int j = 0; j < kPerWarp; ++j) {
        const int band = warp + j * kWarps;
        nb[j] = band < kBandsPerCh ? __ldg(cnt + bin_base + band) : 0u;
    }
    uint2 first[kPerWarp];                                   // records lane (< 32) of each band: most bands hold fewer than 64
#pragma unroll
    for (int j = 0; j < kPerWarp; ++j) {
        if (nb[j] > (uint32_t)kBinCap) { if (lane == 0) s_over = 1; nb[j] = kBinCap; }
        first[j] = make_uint2(0u, 0u);
        if ((uint32_t)lane < nb[j]) first[j] = band_rec_load(bins, (bin_base + warp + j * kWarps) * kBinCap + lane, rec4);
    }
    auto place = [&](int band, uint2 rec) {
        const int r = (int)(rec.x / (uint32_t)kW), u = (int)rec.x - r * kW, v = band * kFillRows + r;
        // where the voxel lands in the output (approximately): its tile and the ones around it (a footprint's
        // bounding box reaches < 36 rows / columns beyond the tile's own image, host-checked) are the candidates
        float px = (float)u, py = (float)v;
        if (ROT) { const float qx = fmaf(RF.f00, px, fmaf(RF.f01, py, RF.f02)); py = fmaf(RF.f10, px, fmaf(RF.f11, py, RF.f12)); px = qx; }
        const float rcx = px, rcy = py;                      // centre of the rotation's candidate window (rot_window_origin)
        if (SCL) { px = (px + 0.5f) * SF.sfx - 0.5f + SF.offx; py = (py + 0.5f) * SF.sfy - 0.5f + SF.offy; }
        const int tx = min(max((int)floorf(px * (1.0f / kATW)), 0), kATBX - 1);
        const int ty = min(max((int)floorf(py * (1.0f / kATH)), 0), kATBY - 1);
#pragma unroll
        for (int dy = -kCandY; dy <= kCandY; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int cx = tx + dx, cy = ty + dy;
                if ((unsigned)cx >= (unsigned)kATBX || (unsigned)cy >= (unsigned)kATBY) continue;
                const int t = cy * kATBX + cx;
                const int4 b = box[t];
                if (u < b.x || u > b.z || v < b.y || v > b.w) continue;
                const uint32_t s = atomicAdd(&hist[t], 1u);
                if (s < (uint32_t)kATCap) {
                    uint32_t w = at_pack_pos(u - b.x, v - b.y);
                    if (ROT) w |= rot_window_origin_at(rcx, rcy, RF, dorg[t].x, dorg[t].y);
                    out_base[(size_t)t * kATCap + s] = make_uint2(w, rec.y);
                }
            }
    };
#pragma unroll
    for (int j = 0; j < kPerWarp; ++j) {
        const int band = warp + j * kWarps;
        if ((uint32_t)lane < nb[j]) place(band, first[j]);
        for (uint32_t i = lane + 32; i < nb[j]; i += 32) place(band, band_rec_load(bins, (bin_base + band) * kBinCap + i, rec4));
    }
    __syncthreads();
    if (s_over) {
        // a row band of this channel spilled into the sweep's overflow list: its records of this channel are re-binned too
        const int bs = fsrc * S + sweep;
        const uint32_t n_ovf = ovf_cnt[bs];
        const OvfRec* list = ovf + (sweep_off[bs] - sweep_off[0]);
        for (uint32_t i = tid; i < n_ovf; i += kRebinThreads) {
            const OvfRec r = list[i];
            const uint32_t t = (uint32_t)(r.idx / kTileElems), cell = (uint32_t)(r.idx - (unsigned long long)t * kTileElems);
            const uint32_t bin_id = t - (uint32_t)bs * kTilesPerSweep;
            if ((int)(bin_id / kBandsPerCh) == zb) place((int)(bin_id % kBandsPerCh), make_uint2(cell, r.bits));
        }
        __syncthreads();
    }
    // a count above kATCap sends that (channel, tile) down the tile kernel's slow path
    for (int t = tid; t < kATTiles; t += kRebinThreads)
        cnt2[((size_t)(slot * S + sweep) * kZ + zb) * kATTiles + t] = hist[t];
}

// Candidate c (0..8) of the source voxel of work-list entry w (position + window origin): if that
// destination pixel's 4 cv2 taps include the voxel, evaluate it from T0 with the reference's arithmetic.  TO_OUT: D is the
// output tile (dropout rectangles apply); otherwise D is the rotated intermediate T1 and its non-zero pixels are
// appended (once each) to the list W2.
template <bool TO_OUT>
__device__ __forceinline__ void rot_candidate(uint32_t w_, int c, const AugTileGeo& G, const AugTileTables& T,
                                              const float* __restrict__ T0, float* __restrict__ D, const int* rects, int nrect,
                                              uint16_t* __restrict__ W2, uint32_t* nW2) {
    const int cyi = c / 3, cxi = c - 3 * cyi;
    const int lx = (int)((w_ >> 14) & 255u) - kATOrgBias + cxi, ly = (int)(w_ >> 22) - kATOrgBias + cyi;
    if ((unsigned)lx >= (unsigned)G.DW || (unsigned)ly >= (unsigned)G.DH) return;
    const int X = (T.X0[ly] + T.A[lx]) >> 5, Y = (T.Y0[ly] + T.B[lx]) >> 5;
    const int ix = (X >> 5) - G.fx0, iy = (Y >> 5) - G.fy0;                 // top-left tap, footprint coordinates
    const int fx = at_pos_x(w_), fy = at_pos_y(w_);
    if ((unsigned)(fx - ix) > 1u || (unsigned)(fy - iy) > 1u) return;       // the voxel is not one of this pixel's taps
    if (TO_OUT && nrect && in_rects(rects, nrect, G.dy0 + ly, G.dx0 + lx)) return;              // utils.py:496
    float w[4];
    rot_weights(X, Y, w);
    // taps outside the (grid-clipped) footprint are outside the grid: BORDER_CONSTANT 0
    const bool x0ok = (unsigned)ix < (unsigned)G.FW, x1ok = (unsigned)(ix + 1) < (unsigned)G.FW;
    const bool y0ok = (unsigned)iy < (unsigned)G.FH, y1ok = (unsigned)(iy + 1) < (unsigned)G.FH;
    const float* p = T0 + iy * G.FW + ix;
    const float v00 = (x0ok && y0ok) ? p[0] : 0.f, v01 = (x1ok && y0ok) ? p[1] : 0.f;
    const float v10 = (x0ok && y1ok) ? p[G.FW] : 0.f, v11 = (x1ok && y1ok) ? p[G.FW + 1] : 0.f;
    const float val = rot_sum(v00, v01, v10, v11, w);
    const int di = ly * G.DW + lx;
    if (TO_OUT) { D[di] = val; return; }
    // every writer of a pixel writes the same value; the one that finds it still zero lists it (a pixel whose value
    // is zero needs no further work: it equals an untouched one)
    const uint32_t old = atomicExch(reinterpret_cast<uint32_t*>(D) + di, __float_as_uint(val));
    if (old == 0u && __float_as_uint(val) != 0u) W2[atomicAdd(nW2, 1u)] = (uint16_t)((ly << 7) | lx);   // (DW <= 80, DH <= 48)
}

// Source pixel (u, v) of the resize (coordinates of the image being resized), work item q (0..3) of it: the output pixels
// whose 2 x 2 taps include the pixel are exactly the rows [yf, yl] x columns [xf, xl] of the inverse tap tables -- 2 x 2 of
// them as a rule, so item q = (row yf + (q >> 1), column xf + (q & 1)) evaluates ONE output pixel; the items with a set bit
// also take the rare third row / column.  src = tile holding the region [qx0, qx0+QW) x [qy0, qy0+QH), row stride QW.
constexpr int kATQuad = 4;
__device__ __forceinline__ void scl_quad(int u, int v, int q, const AugTileGeo& G, const AugTileTables& T,
                                         const float* __restrict__ src, float* __restrict__ O, int ox0, int oy0,
                                         const int* rects, int nrect) {
    const int ry = q >> 1, rx = q & 1;
    const int yl = T.yl[v - G.qy0], xl = T.xl[u - G.qx0];
    const int ly0 = T.yf[v - G.qy0] + ry, lx0 = T.xf[u - G.qx0] + rx;
    const int ly1 = ry ? yl : min(ly0, yl), lx1 = rx ? xl : min(lx0, xl);
    for (int ly = ly0; ly <= ly1; ++ly) {
        const int b0 = T.s0y[ly], b1 = T.s1y[ly];
        const float wy0 = T.wy0[ly], wy1 = T.wy1[ly];
        const float* r0 = src + (b0 - G.qy0) * G.QW - G.qx0;
        const float* r1 = src + (b1 - G.qy0) * G.QW - G.qx0;
        for (int lx = lx0; lx <= lx1; ++lx) {
            if (nrect && in_rects(rects, nrect, oy0 + ly, ox0 + lx)) continue;                  // utils.py:496
            const int a0 = T.s0x[lx], a1 = T.s1x[lx];
            O[ly * kATW + lx] = resize_sum(r0[a0], r0[a1], r1[a0], r1[a1], T.wx0[lx], T.wx1[lx], wy0, wy1);
        }
    }
}

// The same by rows: r-th (0..2) of the output rows source pixel (u, v) feeds, all its columns.  Measured per class on B200:
// scale-only frames are faster with 3 row items per record (73 vs 77 us per frame), rotate + scale frames with 4 quad
// items per listed T1 pixel (159 vs 176 us).
constexpr int kATRows = 3;
__device__ __forceinline__ void scl_row(int u, int v, int r, const AugTileGeo& G, const AugTileTables& T,
                                        const float* __restrict__ src, float* __restrict__ O, int ox0, int oy0,
                                        const int* rects, int nrect) {
    const int ly = T.yf[v - G.qy0] + r;
    if (ly > T.yl[v - G.qy0]) return;
    const int x0 = T.xf[u - G.qx0], x1 = T.xl[u - G.qx0];
    const int b0 = T.s0y[ly], b1 = T.s1y[ly];
    const float wy0 = T.wy0[ly], wy1 = T.wy1[ly];
    const float* r0 = src + (b0 - G.qy0) * G.QW - G.qx0;
    const float* r1 = src + (b1 - G.qy0) * G.QW - G.qx0;
    for (int lx = x0; lx <= x1; ++lx) {
        if (nrect && in_rects(rects, nrect, oy0 + ly, ox0 + lx)) continue;                      // utils.py:496
        const int a0 = T.s0x[lx], a1 = T.s1x[lx];
        O[ly * kATW + lx] = resize_sum(r0[a0], r0[a1], r1[a0], r1[a1], T.wx0[lx], T.wx1[lx], wy0, wy1);
    }
}

// Slow path, dense scenes: the row-band bins under the footprint (+ the sweep's overflow list when one of them spilled)
// -> max into T0.  Kept out of line so that its registers do not count against the common path.
struct AugSlowArgs {                                         // parked in shared memory: only the slow path reads them
    const uint32_t* cnt; const uint2* bins; const uint32_t* ovf_cnt; const OvfRec* ovf; const int32_t* sweep_off;
    int bs, rec4;
};
static __device__ __noinline__ void at_slow_fill(const AugSlowArgs* a, int z, int fx0, int fy0, int FW, int FH, float* __restrict__ T0) {
    const uint32_t* __restrict__ cnt = a->cnt;
    const uint2* __restrict__ bins = a->bins;
    const uint32_t* __restrict__ ovf_cnt = a->ovf_cnt;
    const OvfRec* __restrict__ ovf = a->ovf;
    const int32_t* __restrict__ sweep_off = a->sweep_off;
    const int bs = a->bs;
    const int tid = threadIdx.x;
    bool spilled = false;
    const int band0 = fy0 / kFillRows, band1 = (fy0 + FH - 1) / kFillRows;
    for (int band = band0; band <= band1; ++band) {
        const size_t t = ((size_t)bs * kZ + z) * kBandsPerCh + band;
        uint32_t nb = __ldg(cnt + t);
        if (nb > (uint32_t)kBinCap) { spilled = true; nb = kBinCap; }
        if ((uint32_t)tid < nb) {
            const uint2 rec = band_rec_load(bins, t * kBinCap + tid, a->rec4);
            const int r = (int)(rec.x / (uint32_t)kW), fx = (int)rec.x - r * kW - fx0, fy = band * kFillRows + r - fy0;
            if ((unsigned)fx < (unsigned)FW && (unsigned)fy < (unsigned)FH)
                atomicMax(reinterpret_cast<uint32_t*>(T0) + fy * FW + fx, rec.y);
        }
    }
    if (!spilled) return;
    const uint32_t n_ovf = ovf_cnt[bs];
    const OvfRec* list = ovf + (sweep_off[bs] - sweep_off[0]);
    for (uint32_t i = tid; i < n_ovf; i += kATThreads) {
        const OvfRec r = list[i];
        const uint32_t t = (uint32_t)(r.idx / kTileElems), cell = (uint32_t)(r.idx - (unsigned long long)t * kTileElems);
        const uint32_t bin_id = t - (uint32_t)bs * kTilesPerSweep;
        if ((int)(bin_id / kBandsPerCh) != z) continue;
        const int rr = (int)(cell / (uint32_t)kW), fx = (int)cell - rr * kW - fx0;
        const int fy = (int)(bin_id % kBandsPerCh) * kFillRows + rr - fy0;
        if ((unsigned)fx < (unsigned)FW && (unsigned)fy < (unsigned)FH)
            atomicMax(reinterpret_cast<uint32_t*>(T0) + fy * FW + fx, r.bits);
    }
}

#ifdef BEVGEN_AT_PROFILE
__device__ unsigned long long g_at_prof[16];
#define AT_TICK(k) do { if (tid == PROF_TID) { const long long t_ = clock64(); prof[k] += t_ - tprev; tprev = t_; } } while (0)
#else
#define AT_TICK(k) do {} while (0)
#endif

template <bool ROT, bool SCL>
__global__ void __launch_bounds__(kATThreads, AugTileCfg<ROT, SCL>::kMinBlocks)
k_tile_aug(const uint32_t* __restrict__ cnt, const uint2* __restrict__ bins, const uint32_t* __restrict__ ovf_cnt,
           const OvfRec* __restrict__ ovf, const int32_t* __restrict__ sweep_off, const uint32_t* __restrict__ cnt2,
           const uint2* __restrict__ bins2, const __grid_constant__ CUtensorMap out_map,
           const bevgen_aug_params* __restrict__ params, AugTileFrames fl, int S, int l2_hints, int rec4) {
    static_assert(ROT || SCL, "frames that neither rotate nor scale go through k_tile_store");
    using Cfg = AugTileCfg<ROT, SCL>;
    extern __shared__ __align__(128) unsigned char at_smem[];
    float* const O0 = reinterpret_cast<float*>(at_smem);
    float* const T0 = O0 + Cfg::kOBufs * kATElems;
    static_assert((kATElems * 4) % 128 == 0, "every output tile buffer must be 128-byte aligned (TMA source)");
    float* const T1 = T0 + Cfg::kT0;
    uint16_t* const W2 = reinterpret_cast<uint16_t*>(T1 + Cfg::kT1);   // indices of T1's non-zeros (ROT && SCL)
    __shared__ bevgen_aug_params P;
    __shared__ AugTileTables T;
    __shared__ AugTileGeo G;
    __shared__ RotFwd RF;
    __shared__ uint32_t W[kATCap];                           // this channel's records: at_pack_pos | rot_window_origin
    __shared__ uint32_t nW2;
    __shared__ uint2 Rs[2][kATCap];                          // cp.async staging, thread-private slots: channel z lands in Rs[z & 1]
    __shared__ uint32_t cN[kZ + 3];                          // this tile's record count per channel (0 past the last one)
    __shared__ AugSlowArgs SA;

    const int tid = threadIdx.x;
#ifdef BEVGEN_AT_PROFILE
    const long long t_entry = clock64();
    long long t_s1 = 0, t_s2 = 0;
#endif
    const int tile = blockIdx.x, bx = tile % kATBX, by = tile / kATBX, sweep = blockIdx.y;
    const int fsrc = fl.src[blockIdx.z], fdst = fl.dst[blockIdx.z], slot = fl.slot[blockIdx.z];
    const int ox0 = bx * kATW, oy0 = by * kATH;
    // everything the start-up needs from global memory is requested at once: the parameters, this tile's record counts of
    // all 29 channels and the first two channels' whole bins (4 KB once per CTA)
    const size_t bin2_0 = ((size_t)(slot * S + sweep) * kZ) * kATTiles + tile;
    if (tid < kZ + 3) cN[tid] = tid < kZ ? __ldg(cnt2 + bin2_0 + (size_t)tid * kATTiles) : 0u;
    const uint2* binp = bins2 + bin2_0 * kATCap + tid;
    cp_async_8(&Rs[0][tid], binp);                           // channels 0 and 1: whole bins, their counts are not known yet
    cp_async_commit();
    cp_async_8(&Rs[1][tid], binp + (size_t)kATTiles * kATCap);
    cp_async_commit();
    binp += 2 * (size_t)kATTiles * kATCap;
    {
        const int* g = reinterpret_cast<const int*>(params + fsrc);
        int* s = reinterpret_cast<int*>(&P);
        for (int i = tid; i < (int)(sizeof(bevgen_aug_params) / 4); i += kATThreads) s[i] = g[i];
    }
    if (SCL && tid < 80) { T.xf[tid] = 127; T.xl[tid] = -1; if (tid < 48) { T.yf[tid] = 127; T.yl[tid] = -1; } }
    static_assert(kATThreads >= kATW + kATH + 32 && kATThreads >= 80, "one thread per table entry");
    __syncthreads();
#ifdef BEVGEN_AT_PROFILE
    t_s1 = clock64();
#endif
    int nrect = P.n_rects;
    nrect = nrect > kMaxRects ? kMaxRects : (nrect < 0 ? 0 : nrect);

    // ---- geometry, once per CTA ---------------------------------------------------------------------------------
    if (tid == 0) {
        SA.cnt = cnt; SA.bins = bins; SA.ovf_cnt = ovf_cnt; SA.ovf = ovf; SA.sweep_off = sweep_off; SA.bs = fsrc * S + sweep; SA.rec4 = rec4;
        G = aug_tile_geo<ROT, SCL>(P, ox0, oy0);
        nW2 = 0;
    }
    if (ROT && tid == kATThreads - 32) RF = make_rot_fwd(P.minv);
    if (SCL && tid >= 32 && tid < 32 + kATW + kATH) {
        const ScalePlacement pl = scale_placement(P.scale);
        const int t = tid - 32;
        const bool col = t < kATW;
        const int l = col ? t : t - kATW;
        const int o = col ? ox0 + l : oy0 + l, off = col ? pl.w0 : pl.h0, n = col ? pl.nw : pl.nh;
        const int r = pl.crop ? o + off : o - off;
        int s0 = -1, s1 = -1;
        float w0 = 0.f, w1 = 0.f;
        if (r >= 0 && r < n) resize_tap(r, n, col ? kW : kH, s0, s1, w0, w1);
        if (col) { T.s0x[l] = (short)s0; T.s1x[l] = (short)s1; T.wx0[l] = w0; T.wx1[l] = w1; }
        else     { T.s0y[l] = (short)s0; T.s1y[l] = (short)s1; T.wy0[l] = w0; T.wy1[l] = w1; }
    }
    __syncthreads();
#ifdef BEVGEN_AT_PROFILE
    t_s2 = clock64();
#endif
    if (G.FW * G.FH > Cfg::kT0 || G.FW > 127 || G.FH > 127 || G.QW > 80 || G.QH > 48 || (ROT && (G.DW > Cfg::kDW || G.DH > Cfg::kDH)) ||
        (ROT && SCL && G.DW * G.DH > Cfg::kT1) || (ROT && (RF.hx >= 1.5f || RF.hy >= 1.5f)))
        __trap();
    if (SCL && tid < kATW + kATH && G.QW > 0) {
        // inverse tap tables: which output columns / rows of the tile read source column / row i
        const bool col = tid < kATW;
        const int l = col ? tid : tid - kATW;
        const int s0 = col ? T.s0x[l] : T.s0y[l], s1 = col ? T.s1x[l] : T.s1y[l];
        if (s0 >= 0) {
            int* f = col ? T.xf : T.yf;
            int* la = col ? T.xl : T.yl;
            const int q0 = col ? G.qx0 : G.qy0;
            atomicMin(f + s0 - q0, l); atomicMax(la + s0 - q0, l);
            atomicMin(f + s1 - q0, l); atomicMax(la + s1 - q0, l);
        }
    }
    if (ROT) {
        if (tid < G.DW) { T.A[tid] = rot_adelta(P.minv, G.dx0 + tid); T.B[tid] = rot_bdelta(P.minv, G.dx0 + tid); }
        else if (tid >= 80 && tid - 80 < G.DH) { T.X0[tid - 80] = rot_X0(P.minv, G.dy0 + tid - 80); T.Y0[tid - 80] = rot_Y0(P.minv, G.dy0 + tid - 80); }
    }
    for (int i = tid; i < (Cfg::kOBufs * kATElems + Cfg::kT0 + Cfg::kT1) / 4; i += kATThreads) reinterpret_cast<float4*>(at_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) prefetch_tensormap(&out_map);
    __syncthreads();
    if (SCL && !ROT && tid < G.QH && T.yl[tid] - T.yf[tid] >= kATRows) __trap();   // (scale factors >= 1.45; the host refuses them)
    const int FW = G.FW, FH = G.FH, fx0 = G.fx0, fy0 = G.fy0;

    // records of channel z: one per thread, fetched TWO channels ahead straight into shared memory (cp.async: no registers
    // held across the stages; one channel of lead does not cover the read latency under the kernel's own write stream);
    // each thread reads back only the slots it filled itself, so its own wait_group is all the ordering needed
    auto load_records = [&](int z) {
        const uint32_t nz = cN[z];
        if ((uint32_t)tid < nz && nz <= (uint32_t)kATCap) cp_async_8(&Rs[z & 1][tid], binp);
        cp_async_commit();                                   // (an empty group keeps the wait_group arithmetic uniform)
        binp += (size_t)kATTiles * kATCap;
    };
    const int plane0 = (fdst * S + sweep) * kZ;                 // plane index of channel 0 in the [planes, 400, 720] view
    auto zero_tile = [&](float* Ot) {
        for (int i = tid; i < kATElems / 4; i += kATThreads) reinterpret_cast<float4*>(Ot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    };

#ifdef BEVGEN_AT_PROFILE
    constexpr int PROF_TID = BEVGEN_AT_PROFILE;
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
    const long long t_loop = tprev;
#endif
#pragma unroll 1
    for (int z = 0; z < kZ; ++z) {
        float* const O = O0 + (Cfg::kOBufs == 2 ? (z & 1) * kATElems : 0);
        // the tile this channel builds was last stored two channels ago (kOBufs == 2; thread 0 waited for that copy's smem
        // reads before the last barrier) -- clear it now, B starts after the next barrier
        if (Cfg::kOBufs == 2 && z >= 2) zero_tile(O);
        const uint32_t n = cN[z];
        const bool slow = n > (uint32_t)kATCap;              // uniform: over-full bin (dense scene)
        const uint32_t nw = slow ? 0u : n;
        // ---- A: records -> T0 (max) and the work list W --------------------------------------------------------------
        cp_async_wait_group<1>();                            // channel z has landed (z + 1 may still be in flight)
        AT_TICK(0);
        if ((uint32_t)tid < nw) {
            const uint2 rec = Rs[z & 1][tid];                // (work-list entry, intensity bits), inside the footprint by construction
            atomicMax(reinterpret_cast<uint32_t*>(T0) + at_pos_y(rec.x) * FW + at_pos_x(rec.x), rec.y);
            W[tid] = rec.x;
        }
        if (slow && FW > 0) at_slow_fill(&SA, z, fx0, fy0, FW, FH, T0);
        AT_TICK(1);
        if (Cfg::kOBufs == 1 && z > 0 && tid == 0) bulk_wait_read<0>();   // single tile: the previous channel's copy has read it
        __syncthreads();
        if (Cfg::kOBufs == 1 && z > 0) zero_tile(O);            // (not written before the barrier in front of the resize stage)
        AT_TICK(2);
        load_records(z + 2);                                  // (Rs[z & 1] is free: every thread has read its slot)
        AT_TICK(3);
        // ---- B: every non-zero source voxel -> the destination pixels it feeds -----------------------------------------
#ifdef BEVGEN_AT_SKIP_B
        constexpr bool kDoB = false;                          // timing experiment only: leaves the output wrong
#else
        constexpr bool kDoB = true;
#endif
        if (!kDoB) {
        } else if (ROT) {
            float* const D = SCL ? T1 : O;
            if (!slow) {
                for (uint32_t i = tid; i < nw * kATCand; i += kATThreads) {
                    const uint32_t r = i / kATCand;
                    rot_candidate<!SCL>(W[r], (int)(i - r * kATCand), G, T, T0, D, P.rects, nrect, W2, &nW2);
                }
            } else {
                for (int i = tid; i < FW * FH; i += kATThreads)
                    if (__float_as_uint(T0[i]) != 0u) {
                        const int fy = i / FW, fx = i - fy * FW;
                        const uint32_t w = at_pack_pos(fx, fy) | rot_window_origin(fx0 + fx, fy0 + fy, RF, G);
                        for (int c = 0; c < kATCand; ++c) rot_candidate<!SCL>(w, c, G, T, T0, D, P.rects, nrect, W2, &nW2);
                    }
            }
        } else {
            if (!slow) {
                for (uint32_t i = tid; i < nw * kATRows; i += kATThreads) {
                    const uint32_t r = i / kATRows, w = W[r];
                    scl_row(fx0 + at_pos_x(w), fy0 + at_pos_y(w), (int)(i - r * kATRows), G, T, T0, O, ox0, oy0, P.rects, nrect);
                }
            } else {
                for (int i = tid; i < FW * FH; i += kATThreads)
                    if (__float_as_uint(T0[i]) != 0u)
                        for (int r = 0; r < kATRows; ++r) scl_row(fx0 + i % FW, fy0 + i / FW, r, G, T, T0, O, ox0, oy0, P.rects, nrect);
            }
        }
        uint32_t nw2 = 0;
        if (ROT && SCL) {
            __syncthreads();
            nw2 = nW2;                                        // T1's non-zero pixels, each listed once
            for (uint32_t i = tid; i < nw2 * kATQuad; i += kATThreads) {
                const int w = W2[i / kATQuad];
                scl_quad(G.dx0 + (w & 127), G.dy0 + (w >> 7), (int)(i % kATQuad), G, T, T1, O, ox0, oy0, P.rects, nrect);
            }
        }
        AT_TICK(4);
        fence_proxy_async_smem();                            // this thread's writes to O, before the async-proxy read below
        __syncthreads();
        AT_TICK(5);
        // ---- C: O -> HBM as ONE tensor-map bulk copy (40 x 72 box of plane z), asynchronous: the next channel is built
        //      while it drains; clear the touched entries of T0 / T1 ----------------------------------------------------
        if (tid == 0) {
            if (l2_hints) tensor_store_3d_hint(&out_map, ox0, oy0, plane0 + z, O, l2_policy_evict_first());   // records stay in L2
            else tensor_store_3d(&out_map, ox0, oy0, plane0 + z, O);
            bulk_commit();
            if (Cfg::kOBufs == 2) bulk_wait_read<1>();       // the copy of channel z - 1 has read its tile: free for z + 1
        }
        if (!slow) {
            if ((uint32_t)tid < nw) { const uint32_t w = W[tid]; T0[at_pos_y(w) * FW + at_pos_x(w)] = 0.f; }
        } else {
            for (int i = tid; i < Cfg::kT0 / 4; i += kATThreads) reinterpret_cast<float4*>(T0)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (ROT && SCL) {
            for (uint32_t i = tid; i < nw2; i += kATThreads) { const int w = W2[i]; T1[(w >> 7) * G.DW + (w & 127)] = 0.f; }
            if (tid == 0) nW2 = 0;                           // every thread took its copy (nw2) before the barrier above
        }
        AT_TICK(6);
        __syncthreads();
        AT_TICK(7);
    }
    if (tid == 0) bulk_wait_read<0>();                        // shared memory must outlive the last copies' reads
#ifdef BEVGEN_AT_PROFILE
    if (tid == PROF_TID) {
        for (int k = 0; k < 8; ++k) atomicAdd(&g_at_prof[k], (unsigned long long)prof[k]);
        atomicAdd(&g_at_prof[8], 1ull);
        // set-up: entry -> parameters / counts loaded (first barrier) -> geometry + taps (second) -> tables + cleared tiles (loop)
        atomicAdd(&g_at_prof[9], (unsigned long long)(t_s1 - t_entry));
        atomicAdd(&g_at_prof[10], (unsigned long long)(t_s2 - t_s1));
        atomicAdd(&g_at_prof[11], (unsigned long long)(t_loop - t_s2));
        atomicAdd(&g_at_prof[12], (unsigned long long)(clock64() - t_entry));
    }
#endif
}

// Host: can k_tile_aug handle this frame's parameters?  Conservative real-valued bounds on the region sizes the kernel
// derives exactly (the reference's own draws -- |angle| <= 15 degrees, 0.95 <= scale <= 1.05 -- always pass).
inline bool aug_tile_supported(const bevgen_aug_params& p) {
    const bool rot = p.do_rotate != 0, scl = p.do_scale != 0;
    if (!rot && !scl) return true;
    auto bands_ok = [](int fh) { return (fh + kFillRows - 2) / kFillRows + 1 <= kATMaxBands; };   // fh rows span at most this many bands
    double dw = kATW, dh = kATH;
    if (scl) {
        if (!(p.scale > 0.8 && p.scale < 1.25)) return false;
        const double sf = (double)(int)(kW * p.scale) / kW, sg = (double)(int)(kH * p.scale) / kH;
        if (!(sf > 0.0 && sg > 0.0)) return false;
        dw = kATW / sf + 3.0; dh = kATH / sg + 3.0;                       // resize source region of one tile
        const int qw = (int)dw + 1, qh = (int)dh + 1;
        if (!rot) return qw * qh <= AugTileCfg<false, true>::kT0 && qw <= 80 && qh <= 48 && bands_ok(qh);
        if (qw > AugTileCfg<true, true>::kDW || qh > AugTileCfg<true, true>::kDH || qw * qh > AugTileCfg<true, true>::kT1) return false;
    }
    for (int k = 0; k < 6; ++k) if (!(p.minv[k] == p.minv[k]) || fabs(p.minv[k]) > 1e6) return false;
    const double det = p.minv[0] * p.minv[4] - p.minv[1] * p.minv[3];
    if (!(fabs(det) > 0.8 && fabs(det) < 1.25)) return false;
    // a record's candidate output tiles are those within 36 rows / columns of where it maps: the footprint's
    // bounding-box slack (region extent x |sin| + taps) must stay below that
    if (fabs(p.minv[1]) * (dw > dh ? dw : dh) + 4.0 > 36.0 || fabs(p.minv[3]) * (dw > dh ? dw : dh) + 4.0 > 36.0) return false;
    // source footprint of the rotation's destination region (+2 for the taps, +2 slack for the fixed-point rounding)
    const int fw = (int)(fabs(p.minv[0]) * (dw - 1) + fabs(p.minv[1]) * (dh - 1) + 4.0) + 1;
    const int fh = (int)(fabs(p.minv[3]) * (dw - 1) + fabs(p.minv[4]) * (dh - 1) + 4.0) + 1;
    return fw * fh <= (scl ? AugTileCfg<true, true>::kT0 : AugTileCfg<true, false>::kT0) && fw <= 127 && fh <= 127 && bands_ok(fh);
}


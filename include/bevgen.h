/*
 * bevgen.h -- C ABI of libbevgen_b200.so: B200 (sm_100a) BEV input generation.
 *
 * Drop-in boundary for the ONE hot path of Nadeem202020/VisionTransformer-Intention-Prediction:
 * the BEV input generation its Dataset does per sample in CPU numpy/OpenCV.  The reference
 * has no FFI of its own: its boundary is the set of Python names dataset.py:21-23 imports
 * from utils.py and calls at dataset.py:340,347,348,353.  Each entry point below names the
 * reference function (file:line, relative to the reference repo) whose arithmetic it
 * replaces; the Python host layer (bev_ops.py) keeps the reference's names and signatures
 * on top of these calls, and INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller unless the parameter is
 *     documented as "host"; nothing here allocates device memory
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*)
 *   - return value: 0 = ok, > 0 = cudaError_t from a launch, < 0 = BEVGEN_E_* below
 *   - grids are float32, C-contiguous [C,400,720] per frame, frames stacked [B,C,400,720]
 *     (the layout dataset.py:146-147 stacks and train_vit.py:159 feeds to the model)
 */
#ifndef BEVGEN_H_
#define BEVGEN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: column (AV2) and sweep-store entry points, bevgen_store_put_cols, bevgen_lidar_overflow_total added; max_sweep_points is a
 *    sizing hint and device-resident sweep offsets are clamped (round 2).  Everything declared in version 1 is unchanged. */
#define BEVGEN_ABI_VERSION 2

/* grid constants (reference constants.py:28-47) */
#define BEVGEN_GRID_H 400
#define BEVGEN_GRID_W 720
#define BEVGEN_ZBINS 29
#define BEVGEN_MAP_CHANNELS 9
#define BEVGEN_MAX_DROP_RECTS 5      /* utils.py:485 num_patches_range = (1, 5) */
#define BEVGEN_MAX_POLYLINE_VERTS 1024

#define BEVGEN_ALGO_TILED 0
#define BEVGEN_ALGO_ATOMIC 1
#define BEVGEN_ALGO_MEMSET_ATOMIC 2
/* (the Python layer adds ALGO_AUTO = 3: TILED, switching to ATOMIC while the scenes keep overflowing the tile bins) */

#define BEVGEN_E_INVALID_ARG (-1)
#define BEVGEN_E_WORKSPACE   (-2)
#define BEVGEN_E_UNSUPPORTED (-3)

typedef void *bevgen_stream_t;       /* cudaStream_t */

int bevgen_abi_version(void);
/* out[10] (host) = GRID_H, GRID_W, VOXEL, OFF_X, OFF_Y, Z_MIN, Z_MAX, ZBINS, MAP_CH, Z_MAX-Z_MIN */
void bevgen_constants(double *out);
const char *bevgen_error_string(int code);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t bevgen_launch_count(void);
/* optional per-kernel timing: when enabled every launch is bracketed by CUDA events on its stream;
 * collect() synchronises them and writes "name count total_ms" lines into buf (host), then resets. */
void bevgen_timing_enable(int on);
int bevgen_timing_collect(char *buf, size_t cap);

/* ---------------------------------------------------------------------------------------
 * transform_points -- utils.py:27-33.  out[i,:] = (T . [x y z 1]^T)[:3] in float64, FMA chain
 * in k order (what numpy's dgemm does).  T is a HOST pointer to 16 doubles (row-major 4x4).
 * `stride` = elements per input row (>= 3).  out is float64 [n,3].
 * ------------------------------------------------------------------------------------- */
int bevgen_transform_points_f32(const float *pts, int64_t n, int64_t stride, const double *T_host,
                                double *out, bevgen_stream_t stream);
int bevgen_transform_points_f64(const double *pts, int64_t n, int64_t stride, const double *T_host,
                                double *out, bevgen_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * create_intentnet_lidar_bev -- utils.py:62-106 (and, for the fused form, the per-sweep
 * transform of dataset.py:339-340 in front of it).
 *
 * Frames b = 0..B-1, sweep slots s = 0..S-1 (oldest first); sweep (b,s) owns the point rows
 * [sweep_offsets[b*S+s], sweep_offsets[b*S+s+1]).  An empty range is the reference's
 * `None` / missing sweep: its 29 channels stay zero (utils.py:75-76).  Output channel =
 * s*29 + z_bin (utils.py:97,102); out is [B, S*29, 400, 720] and is fully written (zero
 * fill included -- the caller does not pre-clear it).
 *
 *   fused:   pts4   float32 [N,4] = x,y,z,intensity in the SWEEP frame
 *            rel_tf float64 [B*S,16] = ego_SE3_world @ sweep_SE3 (dataset.py:339), device
 *   ego_f64: pts3   float64 [N,3] already in the ego frame (what dataset.py:340 produces)
 *            intensity float32 [N]
 *
 * Optional train-time fusions (NULL = off), exact by construction:
 *   flip        uint8 [B]   1 = mirror W (utils.py:399)
 *   drop_rects  int32 [B, 5, 4] = (y0, x0, ph, pw); drop_counts int32 [B] (utils.py:496)
 *               -- only meaningful when no rotate/scale follows for that frame
 *
 * max_sweep_points: max over sweeps of the row count -- a SIZING HINT for the launch grid (the host usually knows it): a
 *                   sweep longer than the hint is still processed completely, only slower.
 * sweep_offsets live in device memory and cannot be validated by the host without a synchronisation: the kernels clamp
 * every range to [0, total_points) and treat a range that is not ordered after sweep_offsets[0] as empty.
 * algo: BEVGEN_ALGO_TILED          one binning pass into per-tile record bins; per 28.8 KB tile one CTA builds the tile
 *                                  in shared memory (warp-aggregated atomicMax) and writes it with ONE TMA bulk store
 *                                  [default of the Python layer; fastest except on dense / skewed scenes]
 *       BEVGEN_ALGO_ATOMIC         TMA zero-fill kernel + global RED.MAX scatter kernel
 *       BEVGEN_ALGO_MEMSET_ATOMIC  cudaMemsetAsync + the same scatter kernel (library fill; comparison point)
 *       All produce identical bits; a workspace is needed for TILED only.
 * ------------------------------------------------------------------------------------- */
size_t bevgen_lidar_workspace_bytes(int64_t total_points, int B, int S);

/* records the last BEVGEN_ALGO_TILED launch that used `workspace` (same total_points, B, S) sent to its overflow lists, summed
 * into the device word out_dev on `stream`: dense / skewed scenes (most records overflowing their 128-slot tile bins) run faster
 * on BEVGEN_ALGO_ATOMIC, and this is what a caller needs to notice (the Python layer's ALGO_AUTO does) */
int bevgen_lidar_overflow_total(const void *workspace, size_t workspace_bytes, int64_t total_points, int B, int S,
                                uint32_t *out_dev, bevgen_stream_t stream);

int bevgen_lidar_bev_fused(const float *pts4, const int32_t *sweep_offsets, const double *rel_tf,
                           int B, int S, int64_t total_points, int max_sweep_points,
                           const uint8_t *flip, const int32_t *drop_rects, const int32_t *drop_counts,
                           float *out, void *workspace, size_t workspace_bytes, int algo,
                           bevgen_stream_t stream);

/* same as bevgen_lidar_bev_fused with the points packed as 4 x fp16 (x, y, z, intensity), 8 B per point: the precision
 * AV2 sweeps are stored in (dataset.py:311-321 reads float16 x,y,z and uint8 intensity); fp16 -> fp64 widening is exact,
 * so the result equals the reference's on the same values while halving the host->device bytes. */
int bevgen_lidar_bev_fused_f16(const void *pts_h4, const int32_t *sweep_offsets, const double *rel_tf,
                               int B, int S, int64_t total_points, int max_sweep_points,
                               const uint8_t *flip, const int32_t *drop_rects, const int32_t *drop_counts,
                               float *out, void *workspace, size_t workspace_bytes, int algo,
                               bevgen_stream_t stream);

/* same with the points in AV2's on-disk COLUMN form -- what dataset.py:311 reads from each sweep's feather file: x, y, z as
 * float16 arrays and intensity as a uint8 array (7 B per point, no row interleaving on the host: the Arrow column buffers can
 * be copied as they are).  Sweeps are concatenated inside each column and share sweep_offsets.  Each column pointer must be
 * 16-byte aligned and readable up to the next multiple of 8 rows past total_points (threads read whole 8-row vectors; the
 * padding rows are never used as points).  fp16 -> fp64 and uint8 -> fp32 are exact, so the result equals the reference's
 * on the same values (dataset.py:320-321 widens exactly these dtypes). */
int bevgen_lidar_bev_fused_cols(const void *x_f16, const void *y_f16, const void *z_f16, const uint8_t *intensity_u8,
                                const int32_t *sweep_offsets, const double *rel_tf,
                                int B, int S, int64_t total_points, int max_sweep_points,
                                const uint8_t *flip, const int32_t *drop_rects, const int32_t *drop_counts,
                                float *out, void *workspace, size_t workspace_bytes, int algo,
                                bevgen_stream_t stream);

/* Sweep STORE form of the column entry point.  The four column arrays are a device-resident store of `store_rows` rows
 * holding many sweeps (e.g. every sweep of the logs in use: 0.7 MB per 100 k-point sweep, a whole AV2 split fits in a B200's
 * HBM); a launch names the sweep of slot k = b*S + s by its first row, sweep_start[k], and sweep_offsets[k+1] - sweep_offsets[k]
 * is its length (sweep_offsets is the exclusive prefix sum of the lengths, total_points = its last entry).  Frames may
 * share sweeps -- the reference's samples are 10-sweep sliding windows (dataset.py:174-211), so consecutive samples share 9
 * of their 10 sweeps -- and a step then uploads only the sweeps that are not resident yet instead of every frame's 10.
 * Same arithmetic and bits as bevgen_lidar_bev_fused_cols. */
int bevgen_lidar_bev_store_cols(const void *x_f16, const void *y_f16, const void *z_f16, const uint8_t *intensity_u8,
                                int64_t store_rows, const int32_t *sweep_start, const int32_t *sweep_offsets,
                                const double *rel_tf, int B, int S, int64_t total_points, int max_sweep_points,
                                const uint8_t *flip, const int32_t *drop_rects, const int32_t *drop_counts,
                                float *out, void *workspace, size_t workspace_bytes, int algo,
                                bevgen_stream_t stream);

int bevgen_lidar_bev_ego_f64(const double *pts3, const float *intensity, const int32_t *sweep_offsets,
                             int B, int S, int64_t total_points, int max_sweep_points,
                             const uint8_t *flip, const int32_t *drop_rects, const int32_t *drop_counts,
                             float *out, void *workspace, size_t workspace_bytes, int algo,
                             bevgen_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * rasterize_map_ego_centric -- utils.py:108-182 (with world_to_bev_pixel utils.py:47-60,
 * cv2.fillPoly / cv2.polylines utils.py:155-180) on pre-parsed maps.
 *
 * Map table (device, may hold several maps back to back; see maps.py for the layout):
 *   verts  float64 [V,2]; pl_off int32 [P+1]; lanes int32 [L,4] (left pl, right pl, flags, 0);
 *   cws int32 [CW]
 * frame_map int32 [B,4] = lane_begin, lane_count, cw_begin, cw_count (rows into lanes / cws)
 * ego_T     float64 [B,9] = 3x3 world->ego matrix of utils.py:35-45 per frame
 * max_prims = max over frames of lane_count + cw_count
 * max_polyline_verts = longest polyline in the table (<= BEVGEN_MAX_POLYLINE_VERTS; sizes the per-warp scratch)
 * out float32 [B,9,400,720], fully written.  flip / drop_* as above.
 * ------------------------------------------------------------------------------------- */
size_t bevgen_map_workspace_bytes(int B);

int bevgen_map_raster(const double *verts, const int32_t *pl_off, const int32_t *lanes,
                      const int32_t *cws, const int32_t *frame_map, const double *ego_T,
                      int B, int max_prims, int max_polyline_verts,
                      const uint8_t *flip, const int32_t *drop_rects, const int32_t *drop_counts,
                      float *out, void *workspace, size_t workspace_bytes, bevgen_stream_t stream);

/* world_to_bev_pixel -- utils.py:47-60: xy float64 [m,2] -> int64 [m,2] (px,py); T_host 9 doubles */
int bevgen_world_to_bev_pixel(const double *xy, int64_t m, const double *T_host, int64_t *out,
                              bevgen_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Augmentation -- utils.py:394-498, image part (GT-box arithmetic stays on the host).
 * Per-frame parameters live in a device array of bevgen_aug_params, drawn on the host in the
 * reference's RNG order (utils.py:398,421-422,455-456,488-494).
 *   src, dst: [B, C, 400, 720]; dst != src.  Order: flip -> rotate -> scale -> dropout
 *   (utils.py:508-511).  One fused gather per output pixel reproduces the reference's
 *   intermediate float32 roundings exactly (rotate: cv2.warpAffine fixed-point bilinear,
 *   utils.py:430-433; scale: cv2.resize separable bilinear, utils.py:462).
 * ------------------------------------------------------------------------------------- */
typedef struct bevgen_aug_params {
    double minv[6];        /* inverted 2x3 of cv2.getRotationMatrix2D (dst->src), used if do_rotate */
    double scale;          /* s of utils.py:456, used if do_scale */
    int32_t do_flip, do_rotate, do_scale, n_rects;
    int32_t rects[BEVGEN_MAX_DROP_RECTS * 4];   /* y0, x0, ph, pw */
} bevgen_aug_params;

/* params = device copy, params_host = the same B structs in host memory (the launcher groups frames by
 * (rotate?, scale?) class from the host copy and launches one specialised kernel per class).
 * workspace (optional, device): scratch of k * C*400*720*4 bytes, k >= 1; when present, frames that rotate AND scale
 * run as two 4-tap passes through it (faster); without it they use the fused 16-tap gather.  Same bits either way. */
int bevgen_augment(const float *src, float *dst, int B, int C, const bevgen_aug_params *params,
                   const bevgen_aug_params *params_host, void *workspace, size_t workspace_bytes,
                   bevgen_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Generation FUSED with augmentation: out[b] = augment_bev's image part (utils.py:508-511: flip -> rotate ->
 * scale -> dropout) applied to create_intentnet_lidar_bev of frame b (utils.py:62-106, with the per-sweep transform
 * of dataset.py:339-340 in front), for a whole batch, with every output byte written exactly once: the un-augmented
 * 334 MB grid of a frame that rotates or scales is never materialised (the dense route -- bevgen_lidar_bev_fused into a
 * scratch grid, then bevgen_augment -- moves 3x the bytes).  Same bits as that route.
 *   pts / pts_f16   float32 [N,4] (pts_f16 = 0) or 4 x fp16 per point (pts_f16 = 1); sweep_offsets, rel_tf as above
 *   params          device array of B bevgen_aug_params; params_host = the same structs in host memory
 *   workspace       bevgen_lidar_aug_workspace_bytes(total_points, B, S, n_resampled), n_resampled = number of
 *                   frames with do_rotate or do_scale (59 MB of per-tile record bins per such frame at S = 10)
 * Returns BEVGEN_E_UNSUPPORTED when a frame's parameters are outside what the fused kernel's shared-memory tiles
 * cover (never for the reference's own draws: |angle| <= 15 deg, 0.95 <= scale <= 1.05); query first with
 * bevgen_aug_fused_supported (1 = every frame supported) and take the dense route otherwise.
 * ------------------------------------------------------------------------------------- */
int bevgen_aug_fused_supported(const bevgen_aug_params *params_host, int B);
size_t bevgen_lidar_aug_workspace_bytes(int64_t total_points, int B, int S, int n_resampled);

int bevgen_lidar_bev_augmented(const void *pts, int pts_f16, const int32_t *sweep_offsets, const double *rel_tf,
                               int B, int S, int64_t total_points, int max_sweep_points,
                               const bevgen_aug_params *params, const bevgen_aug_params *params_host,
                               float *out, void *workspace, size_t workspace_bytes, bevgen_stream_t stream);

/* bevgen_lidar_bev_augmented on AV2 column arrays (see bevgen_lidar_bev_fused_cols for the layout rules) */
int bevgen_lidar_bev_augmented_cols(const void *x_f16, const void *y_f16, const void *z_f16, const uint8_t *intensity_u8,
                                    const int32_t *sweep_offsets, const double *rel_tf,
                                    int B, int S, int64_t total_points, int max_sweep_points,
                                    const bevgen_aug_params *params, const bevgen_aug_params *params_host,
                                    float *out, void *workspace, size_t workspace_bytes, bevgen_stream_t stream);

/* Upload one sweep's columns into a sweep store: dst_* = device addresses of the sweep's first row in the store's x / y / z /
 * intensity columns, src_* = HOST arrays of n float16 / float16 / float16 / uint8 values (asynchronous on `stream` when they
 * are pinned).  This is the sweep-decode boundary of dataset.py:311-321: the four Arrow column buffers go over as they are. */
int bevgen_store_put_cols(void *dst_x, void *dst_y, void *dst_z, void *dst_i, const void *src_x, const void *src_y,
                          const void *src_z, const void *src_i, int64_t n, bevgen_stream_t stream);

/* bevgen_lidar_bev_augmented on a sweep store (see bevgen_lidar_bev_store_cols) */
int bevgen_lidar_bev_store_augmented_cols(const void *x_f16, const void *y_f16, const void *z_f16, const uint8_t *intensity_u8,
                                          int64_t store_rows, const int32_t *sweep_start, const int32_t *sweep_offsets,
                                          const double *rel_tf, int B, int S, int64_t total_points, int max_sweep_points,
                                          const bevgen_aug_params *params, const bevgen_aug_params *params_host,
                                          float *out, void *workspace, size_t workspace_bytes, bevgen_stream_t stream);

/* rasterize_map_ego_centric followed by the image part of augment_bev (utils.py:508-511) on the 9 map channels, per-frame
 * parameters as in bevgen_augment (device copy + host copy): frames that neither rotate nor scale get flip / dropout
 * fused into the bit-plane expand; the others are resampled straight from the bit planes (no intermediate float grid).
 * Same bits as bevgen_map_raster followed by bevgen_augment. */
int bevgen_map_raster_augmented(const double *verts, const int32_t *pl_off, const int32_t *lanes,
                                const int32_t *cws, const int32_t *frame_map, const double *ego_T,
                                int B, int max_prims, int max_polyline_verts,
                                const bevgen_aug_params *params, const bevgen_aug_params *params_host,
                                float *out, void *workspace, size_t workspace_bytes, bevgen_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVGEN_H_ */

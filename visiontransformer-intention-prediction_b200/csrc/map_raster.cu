// placeholder until the rasterizer lands (next commit)
#include "bevgen_common.cuh"
extern "C" size_t bevgen_map_workspace_bytes(int) { return 0; }
extern "C" int bevgen_map_raster(const double*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const double*,
                                 int, int, const uint8_t*, const int32_t*, const int32_t*, float*, void*, size_t,
                                 bevgen_stream_t) { return BEVGEN_E_UNSUPPORTED; }
extern "C" int bevgen_world_to_bev_pixel(const double*, int64_t, const double*, int64_t*, bevgen_stream_t) { return BEVGEN_E_UNSUPPORTED; }

// tmabox_bench.cu -- dev microbenchmark: write rate of tensor-map (2-D box) TMA stores into a [planes, 400, 720] float32
// grid as a function of the box shape, for long-lived CTAs that walk 29 planes (the structure of k_tile_aug) and for
// short-lived CTAs (one box each).  Question answered: does the 288-byte row piece of a 72-column tile (4.5 DRAM atoms of
// 64 B, alternately mis-aligned by 32 B) cost bandwidth against 64-byte-aligned widths (48, 80, 144, 240 columns)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmabox_bench tools/tmabox_bench.cu && tools/tmabox_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kW = 720, kH = 400, kZ = 29;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// long-lived: CTA (tile, sweep, frame) stores its box of planes (frame*S + sweep)*29 + z, z = 0..28, double-buffered
// variant: 0 = planes in order; 1 = every CTA starts at a different plane (z rotated by its tile index); 2 = wait for full
// completion of the previous store (not only its smem read); 3 = tile position rotates with the plane
__global__ void __launch_bounds__(256) k_long(const __grid_constant__ CUtensorMap map, int bw, int bh, int tiles_x, int S, int barriers, int zc, int variant) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int elems = bw * bh;
    for (int i = threadIdx.x; i < 2 * elems / 4; i += 256) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int nchunk = (kZ + zc - 1) / zc;                   // the 29 planes of a sweep are split over nchunk CTAs
    const int sweep = blockIdx.y / nchunk, chunk = blockIdx.y % nchunk;
    const int plane0 = (blockIdx.z * S + sweep) * kZ;
    const int z1 = min(kZ, (chunk + 1) * zc);
    const int ntile = gridDim.x;
    for (int zi = chunk * zc; zi < z1; ++zi) {
        for (int b = 0; b < barriers; ++b) __syncthreads();
        if (threadIdx.x == 0) {
            const int z = variant == 1 ? (zi + (int)blockIdx.x * 7 + (int)blockIdx.y * 3) % kZ : zi;
            int ttx = tx, tty = ty;
            if (variant == 3) { const int t = ((int)blockIdx.x + zi * 37) % ntile; ttx = t % tiles_x; tty = t / tiles_x; }
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         :: "l"(reinterpret_cast<uint64_t>(&map)), "r"(ttx * bw), "r"(tty * bh), "r"(plane0 + z),
                            "r"(smem_u32(sm + (zi & 1) * elems * 4)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (variant == 2) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// plane-major: CTA (tile group, plane) stores `k` ADJACENT tiles of ONE plane one after the other (double-buffered), so the
// stream of stores is address-ordered across the grid like the short-lived form, with k iterations of per-CTA work
__global__ void __launch_bounds__(256) k_plane_major(const __grid_constant__ CUtensorMap map, int bw, int bh, int tiles_x, int k, int barriers) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int elems = bw * bh;
    for (int i = threadIdx.x; i < 2 * elems / 4; i += 256) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        for (int b = 0; b < barriers; ++b) __syncthreads();
        if (threadIdx.x == 0) {
            const int t = blockIdx.x * k + j, tx = t % tiles_x, ty = t / tiles_x;
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         :: "l"(reinterpret_cast<uint64_t>(&map)), "r"(tx * bw), "r"(ty * bh), "r"((int)blockIdx.y),
                            "r"(smem_u32(sm + (j & 1) * elems * 4)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// short-lived: one box per CTA (blockIdx.y = plane)
__global__ void __launch_bounds__(128) k_short(const __grid_constant__ CUtensorMap map, int bw, int bh, int tiles_x) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < bw * bh / 4; i += 128) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 :: "l"(reinterpret_cast<uint64_t>(&map)), "r"(tx * bw), "r"(ty * bh), "r"((int)blockIdx.y), "r"(smem_u32(sm)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int F = 8, S = 10;                                  // 8 frames x 290 planes = 2.67 GB
    const size_t planes = (size_t)F * S * kZ, bytes = planes * kW * kH * 4;
    float* buf; CK(cudaMalloc(&buf, bytes));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = reinterpret_cast<EncodeFn>(p);
    CK(cudaFuncSetAttribute(k_long, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    CK(cudaFuncSetAttribute(k_short, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    CK(cudaFuncSetAttribute(k_plane_major, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int shapes[][2] = {{72, 40}};      // tile bytes must be multiples of 128
    for (auto& sh : shapes) {
        const int bw = sh[0], bh = sh[1];
        if (kW % bw || kH % bh) continue;
        CUtensorMap map;
        const cuuint64_t gdim[3] = {kW, kH, planes}, gstr[2] = {kW * 4, (cuuint64_t)kW * kH * 4};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed %dx%d\n", bw, bh); continue; }
        const int tiles_x = kW / bw, tiles = tiles_x * (kH / bh);
        const int tile_bytes = bw * bh * 4;
        for (int zcv : {29, 129, 4, 1}) {
            const int zc = zcv % 100, variant = zcv / 100, barriers = 3;
            auto run = [&] { k_long<<<dim3(tiles, S * ((kZ + zc - 1) / zc), F), 256, 2 * tile_bytes>>>(map, bw, bh, tiles_x, S, barriers, zc, variant); };
            for (int i = 0; i < 2; ++i) run();
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a); for (int i = 0; i < 5; ++i) run(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
            printf("box %3d x %2d (%5d B rows %4d B) planes/CTA %2d variant %d : %7.3f ms  %7.1f GB/s  %5.1f us/frame\n", bw, bh, tile_bytes, bw * 4,
                   zc, variant, ms, bytes / ms / 1e6, ms * 1e3 / F);
        }
        for (int k : {2, 5, 10, 20, 50, 100}) {
            if (tiles % k) continue;
            auto run = [&] { k_plane_major<<<dim3(tiles / k, (unsigned)planes), 256, 2 * tile_bytes>>>(map, bw, bh, tiles_x, k, 3); };
            for (int i = 0; i < 2; ++i) run();
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a); for (int i = 0; i < 5; ++i) run(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
            printf("plane-major box %3d x %2d, %3d adjacent tiles of one plane per CTA : %7.3f ms  %7.1f GB/s  %5.1f us/frame\n", bw, bh, k, ms,
                   bytes / ms / 1e6, ms * 1e3 / F);
        }
        auto run2 = [&] { k_short<<<dim3(tiles, (unsigned)planes), 128, tile_bytes>>>(map, bw, bh, tiles_x); };
        for (int i = 0; i < 2; ++i) run2();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(a); for (int i = 0; i < 5; ++i) run2(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
        printf("short-lived box %3d x %2d                                          : %7.3f ms  %7.1f GB/s  %5.1f us/frame\n", bw, bh, ms, bytes / ms / 1e6, ms * 1e3 / F);
    }
    printf("done\n");
    return 0;
}

// redbench.cu -- dev microbenchmark: do atomics (RED.MAX) hit the L2 on lines that were just zero-filled?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_fill_st(uint4* __restrict__ out, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = make_uint4(0, 0, 0, 0);
}
__global__ void k_fill_tma(unsigned char* __restrict__ out, size_t total_bytes, int chunk) {
    extern __shared__ __align__(128) unsigned char z[];
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4*>(z)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = total_bytes / chunk;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(z);
    for (size_t g = blockIdx.x; g < nchunks; g += gridDim.x) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + g * chunk), "r"(saddr), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// n random REDs into region [0, words)
__global__ void k_red(uint32_t* __restrict__ region, size_t words, int n, uint32_t seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t h = (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull + seed; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    atomicMax(region + (size_t)(h % words), (uint32_t)(i | 1));
}

int main() {
    const size_t total = (size_t)2 << 30;
    unsigned char* buf; CK(cudaMalloc(&buf, total));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    CK(cudaFuncSetAttribute(k_fill_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int n = 500000;
    for (size_t regionMB : {8, 32, 64, 128, 334}) {
        const size_t rbytes = regionMB << 20;
        for (int mode = 0; mode < 4; ++mode) {   // 0 memset, 1 st.global, 2 tma, 3 cold (fill then flush L2 with 600 MB of other writes)
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                if (mode == 0) CK(cudaMemsetAsync(buf, 0, rbytes));
                else if (mode == 1 || mode == 3) k_fill_st<<<sms * 4, 256>>>((uint4*)buf, rbytes / 16);
                else k_fill_tma<<<sms * 4, 128, 28800>>>(buf, rbytes - rbytes % 28800, 28800);
                if (mode == 3) k_fill_st<<<sms * 4, 256>>>((uint4*)(buf + (1ull << 30)), (600ull << 20) / 16);
                cudaEventRecord(a);
                k_red<<<(n + 255) / 256, 256>>>((uint32_t*)buf, rbytes / 4, n, rep * 7919u);
                cudaEventRecord(b);
                CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            const char* nm[] = {"after cudaMemset", "after st.global fill", "after TMA fill", "cold (L2 flushed)"};
            printf("region %4zu MB  %-22s  %d REDs: %7.1f us  (%.1f G RED/s)\n", regionMB, nm[mode], n, best * 1e3, n / best / 1e6);
        }
    }
    return 0;
}

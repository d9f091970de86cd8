// fillbench.cu -- dev microbenchmark: how fast can B200 write zeros?  (cudaMemset vs st.global.v4 vs TMA bulk store)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_fill_st(uint4* __restrict__ out, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = make_uint4(0, 0, 0, 0);
}
__global__ void k_fill_st_v8(uint4* __restrict__ out, size_t n32) {   // 256-bit stores (sm_100+)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride)
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(out + 2 * i), "r"(0) : "memory");
}
// each block owns a contiguous span (like a memset): block b writes [b*span, (b+1)*span)
__global__ void k_fill_st_span(uint4* __restrict__ out, size_t n16, size_t span16) {
    const size_t beg = (size_t)blockIdx.x * span16, end = beg + span16 < n16 ? beg + span16 : n16;
    for (size_t i = beg + threadIdx.x; i < end; i += blockDim.x) out[i] = make_uint4(0, 0, 0, 0);
}
__global__ void k_fill_st_cs(uint4* __restrict__ out, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(out + i, make_uint4(0, 0, 0, 0));
}

// TMA bulk store of a zero smem tile; `prod` issuing threads per CTA (one per warp), depth stores in flight each
template <int HINT>
__global__ void k_fill_tma(unsigned char* __restrict__ out, size_t total_bytes, int chunk, int prod, int depth) {
    extern __shared__ __align__(128) unsigned char z[];
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4*>(z)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= prod || lane != 0) return;
    uint64_t pol = 0;
    if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (HINT == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t nchunks = total_bytes / chunk;
    const size_t stride = (size_t)gridDim.x * prod;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(z);
    int inflight = 0;
    for (size_t g = (size_t)blockIdx.x * prod + warp; g < nchunks; g += stride) {
        if (HINT == 0)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + g * chunk), "r"(saddr), "r"(chunk) : "memory");
        else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" :: "l"(out + g * chunk), "r"(saddr), "r"(chunk), "l"(pol) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++inflight >= depth) {
            if (depth == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            else if (depth == 2) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
            else if (depth == 4) asm volatile("cp.async.bulk.wait_group 3;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
            --inflight;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// short-lived CTAs: each zero-inits a small smem tile and bulk-stores `per` consecutive chunks, then exits
__global__ void k_fill_tma_short(unsigned char* __restrict__ out, size_t total_bytes, int chunk, int per) {
    extern __shared__ __align__(128) unsigned char z[];
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4*>(z)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = total_bytes / chunk;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(z);
    for (int k = 0; k < per; ++k) {
        const size_t g = (size_t)blockIdx.x * per + k;
        if (g >= nchunks) break;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + g * chunk), "r"(saddr), "r"(chunk) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <typename F>
float timeit(F f, int n = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < n; ++i) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / n;
}

int main() {
    const size_t bytes = (size_t)8 * 290 * 400 * 720 * 4;   // 8 frames = 2.67 GB
    unsigned char* buf; CK(cudaMalloc(&buf, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char* name, float ms) { printf("%-46s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    rep("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(buf, 0, bytes); }));
    for (int occ : {4, 8}) {
        char nm[96]; snprintf(nm, 96, "st.global.v4 grid=%dxSM", occ);
        rep(nm, timeit([&] { k_fill_st<<<sms * occ, 256>>>((uint4*)buf, bytes / 16); }));
        snprintf(nm, 96, "st.global.cs.v4 grid=%dxSM", occ);
        rep(nm, timeit([&] { k_fill_st_cs<<<sms * occ, 256>>>((uint4*)buf, bytes / 16); }));
    }
    for (int bs : {128, 512, 1024}) for (int occ : {1, 2, 4, 16}) {
        char nm[96]; snprintf(nm, 96, "st.global.v4 block=%d grid=%dxSM", bs, occ);
        rep(nm, timeit([&] { k_fill_st<<<sms * occ, bs>>>((uint4*)buf, bytes / 16); }));
    }
    for (int occ : {2, 4, 8}) {
        char nm[96]; snprintf(nm, 96, "st.global.v8 grid=%dxSM", occ);
        rep(nm, timeit([&] { k_fill_st_v8<<<sms * occ, 256>>>((uint4*)buf, bytes / 32); }));
    }
    for (size_t spanKB : {64, 256, 1024, 4096}) {
        const size_t span16 = spanKB * 1024 / 16, nb = (bytes / 16 + span16 - 1) / span16;
        char nm[96]; snprintf(nm, 96, "st.global.v4 contiguous span=%zuKB blocks=%zu", spanKB, nb);
        rep(nm, timeit([&] { k_fill_st_span<<<(unsigned)nb, 256>>>((uint4*)buf, bytes / 16, span16); }));
    }
    CK(cudaFuncSetAttribute(k_fill_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_fill_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_fill_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    for (int chunk : {28800, 115200})
        for (int occ : {1, 2, 4})
            for (int prod : {1, 4})
                for (int depth : {2, 8}) {
                    if ((size_t)chunk * occ > 200000) continue;
                    char nm[96]; snprintf(nm, 96, "tma chunk=%d cta/SM=%d prod=%d depth=%d", chunk, occ, prod, depth);
                    rep(nm, timeit([&] { k_fill_tma<0><<<sms * occ, 128, chunk>>>(buf, bytes, chunk, prod, depth); }, 5));
                }
    for (int chunk : {7200, 14400, 28800})
        for (int per : {1, 2, 4, 8, 16}) {
            const size_t nchunks = bytes / chunk;
            char nm[96]; snprintf(nm, 96, "tma short-lived chunk=%d per=%d (span %d KB)", chunk, per, chunk * per / 1024);
            rep(nm, timeit([&] { k_fill_tma_short<<<(unsigned)((nchunks + per - 1) / per), 128, chunk>>>(buf, bytes, chunk, per); }, 5));
        }
    rep("tma 28800 occ4 prod4 depth2 evict_last", timeit([&] { k_fill_tma<1><<<sms * 4, 128, 28800>>>(buf, bytes, 28800, 4, 2); }, 5));
    rep("tma 28800 occ4 prod4 depth2 evict_first", timeit([&] { k_fill_tma<2><<<sms * 4, 128, 28800>>>(buf, bytes, 28800, 4, 2); }, 5));
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}

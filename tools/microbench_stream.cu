// microbench_stream.cu — design probe (dev tool): which read pattern saturates B200 HBM for a pure
// coalesced u32 stream (the link-code stream of the day kernel)?  Sums the words so nothing is elided.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_stream microbench_stream.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 1) v = __ldg(p);
    if (MODE == 2) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// A: grid-stride, U loads in flight per thread
template <int MODE, int U>
__global__ void k_gridstride(const uint4* __restrict__ x, size_t n, unsigned long long* out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (i0 + u * stride < n) ? ld<MODE>(x + i0 + u * stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

// D: every warp owns contiguous blocks of BLK uint4 per lane-group: block b -> x[b*32*U .. ), lane reads U consecutive rows
template <int MODE, int U>
__global__ void k_warpblock(const uint4* __restrict__ x, size_t n, unsigned long long* out) {
    uint32_t acc = 0;
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t n_blocks = n / (32 * U);
    for (size_t b = warp; b < n_blocks; b += n_warps) {
        const uint4* p = x + b * 32 * U + lane;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld<MODE>(p + u * 32);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

// E: TMA 1-D bulk copies into a shared-memory ring, consumed by all threads
template <int STAGES, int CHUNK>   // CHUNK bytes per stage
__global__ void __launch_bounds__(1024, 1) k_bulk(const uint4* __restrict__ x, size_t n, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    const uint32_t tid = threadIdx.x;
    const size_t n_chunks = n * 16 / CHUNK;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            uint32_t a = (uint32_t)__cvta_generic_to_shared(&full[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t chunk, int s) {
        uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + (size_t)s * CHUNK);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"((const uint8_t*)x + chunk * CHUNK), "r"(CHUNK), "r"(bar) : "memory");
    };
    size_t c = blockIdx.x, issued = 0;
    if (tid == 0)
        for (int s = 0; s < STAGES && c + (size_t)s * gridDim.x < n_chunks; ++s) { issue(c + (size_t)s * gridDim.x, s); }
    uint32_t acc = 0;
    uint32_t phase = 0;
    int s = 0;
    for (; c < n_chunks; c += gridDim.x) {
        uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        const uint4* sp = (const uint4*)(smem + (size_t)s * CHUNK);
        for (uint32_t k = tid; k < CHUNK / 16; k += blockDim.x) { uint4 v = sp[k]; acc += v.x ^ v.y ^ v.z ^ v.w; }
        __syncthreads();
        const size_t nxt = c + (size_t)STAGES * gridDim.x;
        if (tid == 0 && nxt < n_chunks) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(nxt, s); }
        if (++s == STAGES) { s = 0; phase ^= 1; }
    }
    (void)issued;
    if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

template <class F>
void timeit(const char* name, size_t bytes, F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    printf("%-46s %8.3f ms  %7.1f GB/s\n", name, best, bytes / best / 1e6);
}

int main() {
    const size_t bytes = 2ull << 30;                       // 2 GiB stream (>> L2)
    const size_t n = bytes / 16;
    uint4* x; unsigned long long* out;
    CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(x, 1, bytes)); CK(cudaMemset(out, 0, 8));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d, stream %.0f MB\n", sms, bytes / 1e6);
#define GS(MODE, U, CTAS, THR) timeit("gridstride mode" #MODE " U" #U " " #CTAS "x" #THR, bytes, [&] { k_gridstride<MODE, U><<<sms * CTAS, THR>>>(x, n, out); });
#define WB(MODE, U, CTAS, THR) timeit("warpblock  mode" #MODE " U" #U " " #CTAS "x" #THR, bytes, [&] { k_warpblock<MODE, U><<<sms * CTAS, THR>>>(x, n, out); });
    GS(0, 4, 1, 1024) GS(1, 4, 1, 1024) GS(2, 4, 1, 1024)
    GS(0, 4, 2, 512) GS(0, 4, 4, 256) GS(0, 4, 8, 256) GS(0, 8, 8, 256) GS(0, 2, 16, 128)
    GS(0, 8, 1, 1024) GS(0, 16, 1, 1024) GS(0, 1, 2, 1024) GS(0, 4, 2, 1024) GS(0, 4, 32, 256)
    WB(0, 4, 1, 1024) WB(0, 8, 1, 1024) WB(0, 4, 2, 1024) WB(0, 8, 4, 256) WB(1, 8, 1, 1024)
    CK(cudaFuncSetAttribute(k_bulk<4, 32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768));
    timeit("bulk TMA 4 stages x 32 KB, 1x1024", bytes, [&] { k_bulk<4, 32768><<<sms, 1024, 4 * 32768>>>(x, n, out); });
    CK(cudaFuncSetAttribute(k_bulk<6, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16384));
    timeit("bulk TMA 6 stages x 16 KB, 1x1024", bytes, [&] { k_bulk<6, 16384><<<sms, 1024, 6 * 16384>>>(x, n, out); });
    CK(cudaFuncSetAttribute(k_bulk<3, 65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536));
    timeit("bulk TMA 3 stages x 64 KB, 1x1024", bytes, [&] { k_bulk<3, 65536><<<sms, 1024, 3 * 65536>>>(x, n, out); });
    // reference point: device-to-device memcpy (read + write)
    uint4* y; CK(cudaMalloc(&y, bytes / 2));
    timeit("cudaMemcpy D2D 1 GiB (read+write bytes)", bytes, [&] { CK(cudaMemcpyAsync(y, x, bytes / 2, cudaMemcpyDeviceToDevice)); });
    return 0;
}

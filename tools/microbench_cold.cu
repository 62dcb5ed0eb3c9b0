// microbench_cold.cu — design probe (dev tool, not product): which path resolves the COLD gathers of the day kernel
// fastest?  Cold = the target's word of yesterday's 2-bit mode vector is not staged in shared memory and has to come
// from the L2-resident global vector (2 MB for 8M agents, 12.5 MB for 50M).  One 1024-thread block per SM with 200 KB
// of dynamic shared memory, as in the day kernel (so L1 is ~28 KB); link codes are streamed coalesced (uint4).
//   L    __ldg word gather                      (what the day kernel does today)
//   CG   ld.global.cg word gather               (bypass L1)
//   T    tex1Dfetch<uint32_t> from a linear texture object over the same vector (TEX pipe instead of the LSU pipe)
//   B    one 16-byte cp.async.bulk per lane into shared memory, then LDS (TMA as the gather engine)
//   H    all gathers from shared memory (the hot path, for scale)
//   HL / HT   70 % hot (LDS) + 30 % cold through L / T in the same loop: do the two pipes overlap?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_cold microbench_cold.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t n_agents) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t s = (i + 1) * 0x9E3779B97F4A7C15ull;
        s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
        idx[i] = (uint32_t)(s % n_agents);
    }
}

enum { kL = 0, kCG = 1, kT = 2, kB = 3, kH = 4, kHL = 5, kHT = 6 };

// idx entries are agent numbers; word = j >> 4, field = j & 15
template <int V, int U>
__global__ void __launch_bounds__(1024, 1)
gather_kernel(const uint4* __restrict__ idx, size_t n_vec, const uint32_t* __restrict__ vec, cudaTextureObject_t tex,
              uint32_t n_smem_words, unsigned long long* out) {
    extern __shared__ __align__(128) uint32_t s_mode[];
    __shared__ uint64_t s_bar[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (V == kH || V == kHL || V == kHT) {
        for (uint32_t k = threadIdx.x; k < n_smem_words; k += blockDim.x) s_mode[k] = vec[k];
    }
    // B: every warp owns a landing zone of 32 lanes x U*4 gathers x 16 bytes at the start of shared memory
    uint32_t* zone = s_mode + (size_t)warp * 32 * U * 4 * 4;
    uint32_t phase = 0;
    if (V == kB && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[warp])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * U) {
        uint4 vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) vv[u] = (i0 + u * stride < n_vec) ? ldg_stream(idx + i0 + u * stride) : make_uint4(0, 0, 0, 0);
        uint32_t w[U * 4], j[U * 4];
#pragma unroll
        for (int u = 0; u < U; ++u) { j[4 * u] = vv[u].x; j[4 * u + 1] = vv[u].y; j[4 * u + 2] = vv[u].z; j[4 * u + 3] = vv[u].w; }
        if (V == kB) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar[warp])), "r"(32u * U * 4u * 16u) : "memory");
            __syncwarp();
#pragma unroll
            for (int k = 0; k < U * 4; ++k)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                             ::"r"(smem_u32(zone + ((size_t)k * 32 + lane) * 4)), "l"(vec + ((j[k] >> 4) & ~3u)), "r"(smem_u32(&s_bar[warp])) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_u32(&s_bar[warp])), "r"(phase) : "memory");
            phase ^= 1u;
#pragma unroll
            for (int k = 0; k < U * 4; ++k) w[k] = zone[((size_t)k * 32 + lane) * 4 + ((j[k] >> 4) & 3u)];
            __syncwarp();
        } else {
#pragma unroll
            for (int k = 0; k < U * 4; ++k) {
                const uint32_t wd = j[k] >> 4;
                if (V == kL) w[k] = __ldg(vec + wd);
                if (V == kCG) w[k] = ld_cg(vec + wd);
                if (V == kT) w[k] = tex1Dfetch<uint32_t>(tex, (int)wd);
                if (V == kH) w[k] = s_mode[wd % n_smem_words];
                if (V == kHL || V == kHT) {
                    // 30 % cold: decided per link (k-position pattern 3 of 10 would be compile-time; use the index bits so
                    // that a warp's row is uniformly hot or cold, as in the SELL layout where cold lists are separate rows)
                    const bool cold = ((k * 7 + 3) % 10) < 3;
                    if (!cold) w[k] = s_mode[wd % n_smem_words];
                    else w[k] = V == kHL ? __ldg(vec + wd) : tex1Dfetch<uint32_t>(tex, (int)wd);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < U * 4; ++k) acc += 1u << (8u * ((w[k] >> ((j[k] & 15u) * 2u)) & 3u));
    }
    if (acc == 0xFFFFFFFFu) atomicAdd(out, 1ull);
    atomicAdd(out + 1, (unsigned long long)(acc & 0xFF));
}

template <int V, int U>
void run(const char* name, int grid, size_t smem, const uint4* idx, size_t n_vec, const uint32_t* vec, cudaTextureObject_t tex,
         uint32_t words, unsigned long long* out) {
    CK(cudaFuncSetAttribute(gather_kernel<V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        gather_kernel<V, U><<<grid, 1024, smem>>>(idx, n_vec, vec, tex, words, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    const double gathers = (double)n_vec * 4;
    printf("  %-28s U=%d %8.3f ms  %7.2f Ggather/s  %5.2f gathers/clk/SM@1.9GHz\n", name, U, best, gathers / best / 1e6,
           gathers / (best * 1e-3) / grid / 1.9e9);
    fflush(stdout);
}

int main() {
    const size_t E = 240ull * 1000 * 1000;               // 960 MB of link indices (> L2)
    const size_t n_vec = E / 4;
    uint32_t* d_idx; unsigned long long* d_out;
    CK(cudaMalloc(&d_idx, E * 4)); CK(cudaMalloc(&d_out, 16)); CK(cudaMemset(d_out, 0, 16));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t smem = 200 * 1024;
    const uint32_t words = smem / 4;
    for (uint32_t n_agents : {8000000u, 50000000u}) {
        const size_t vec_bytes = (size_t)n_agents / 4 + 256;
        uint32_t* d_vec; CK(cudaMalloc(&d_vec, vec_bytes)); CK(cudaMemset(d_vec, 0x1B, vec_bytes));
        fill_idx<<<sms * 8, 256>>>(d_idx, E, n_agents);
        CK(cudaDeviceSynchronize());
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d_vec;
        rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>(); rd.res.linear.sizeInBytes = vec_bytes / 4 * 4;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        printf("%u agents: mode vector %.2f MB, %zu gathers, %d SMs\n", n_agents, vec_bytes / 1e6, E, sms);
        const uint4* idx = (const uint4*)d_idx;
#define ROW(U) \
        run<kL, U>("L  __ldg", sms, smem, idx, n_vec, d_vec, tex, words, d_out); \
        run<kCG, U>("CG ld.global.cg", sms, smem, idx, n_vec, d_vec, tex, words, d_out); \
        run<kT, U>("T  tex1Dfetch", sms, smem, idx, n_vec, d_vec, tex, words, d_out); \
        if (U <= 2) run<kB, (U <= 2 ? U : 1)>("B  cp.async.bulk 16 B + LDS", sms, smem, idx, n_vec, d_vec, tex, words, d_out); \
        run<kH, U>("H  shared memory", sms, smem, idx, n_vec, d_vec, tex, words, d_out); \
        run<kHL, U>("HL 70% LDS + 30% __ldg", sms, smem, idx, n_vec, d_vec, tex, words, d_out); \
        run<kHT, U>("HT 70% LDS + 30% tex", sms, smem, idx, n_vec, d_vec, tex, words, d_out);
        ROW(1) ROW(2) ROW(4)
        CK(cudaDestroyTextureObject(tex)); CK(cudaFree(d_vec));
    }
    return 0;
}

// microbench_push.cu — design probe (dev tool, not product): "push" instead of "pull" for the links whose target is
// not staged in shared memory.  Pull (the day kernel today): lane = source agent, a divergent L2 gather per link,
// bound by the L1TEX wavefront rate at ~0.5 gathers per clock per SM (microbench_cold).  Push: a block owns a chunk of
// source agents whose packed counters live in shared memory; the chunk's cold links are stored SORTED BY TARGET, so the
// 32 gathers of a warp fall into a few 128-byte lines (few wavefronts), and each gathered mode is added to the source
// agent's counter with a shared-memory reduction (red.shared.add.u32 of 1 << 8*mode).
//   P  push: u32 target code + u16 source per link, sorted by target (mean gap `gap` agents), red.shared
//   G  gathers only (sorted), no reduction        R  reductions only (random sources), no gather
//   U  pull baseline: unsorted gathers, no reduction
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_push microbench_push.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u16(const uint16_t* p) {
    uint16_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_shared(uint32_t* p, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t mix(uint64_t s) {
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32; s *= 0x94D049BB133111EBull; s ^= s >> 29;
    return s;
}

// Per block b: links_per_block links; link i targets agent ~ i * gap + jitter (sorted), or a random agent (unsorted).
__global__ void fill(uint32_t* code, uint16_t* src, size_t links_per_block, uint32_t n_agents, uint32_t chunk_agents, int sorted) {
    const size_t n = links_per_block * gridDim.y;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < links_per_block; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = (size_t)blockIdx.y * links_per_block + i;
        const uint64_t h = mix(g + 1);
        uint32_t agent;
        if (sorted) {
            const double pos = ((double)i + (double)(h & 0xFFFF) / 65536.0) / (double)links_per_block;
            agent = (uint32_t)(pos * n_agents);
        } else agent = (uint32_t)(h % n_agents);
        if (agent >= n_agents) agent = n_agents - 1;
        code[g] = ((agent >> 4) * 4u) << 5 | (30u - 2u * (agent & 15u));
        src[g] = (uint16_t)((h >> 32) % chunk_agents);
    }
    (void)n;
}

enum { kP = 0, kG = 1, kR = 2, kU = 3 };

template <int V, int U>
__global__ void __launch_bounds__(1024, 1)
push_kernel(const uint32_t* __restrict__ code, const uint16_t* __restrict__ src, size_t links_per_block,
            const uint32_t* __restrict__ vec, uint32_t chunk_agents, unsigned long long* out) {
    extern __shared__ __align__(128) uint32_t s_cnt[];
    for (uint32_t k = threadIdx.x; k < chunk_agents; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const uint32_t* c = code + (size_t)blockIdx.x * links_per_block;
    const uint16_t* s = src + (size_t)blockIdx.x * links_per_block;
    uint32_t acc = 0;
    // warp w takes rows w, w + 32, ... of 32 consecutive links; U rows in flight
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const size_t rows = links_per_block / 32;
    for (size_t r0 = warp; r0 < rows; r0 += 32 * U) {
        uint32_t cc[U], ss[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t r = r0 + 32 * (size_t)u;
            cc[u] = r < rows ? ld_stream_u32(c + r * 32 + lane) : 0u;
            ss[u] = (V == kP || V == kR) ? (r < rows ? ld_stream_u16(s + r * 32 + lane) : 0u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            w[u] = (V == kR) ? cc[u] : ld_cg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(vec) + (cc[u] >> 5)));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t m = __funnelshift_l(0u, w[u], cc[u]) >> 30;
            if (V == kP || V == kR) red_shared(&s_cnt[ss[u]], 1u << (8u * m));
            else acc += 1u << (8u * m);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < chunk_agents; k += blockDim.x) acc += s_cnt[k];
    if (acc == 0xFFFFFFFFu) atomicAdd(out, 1ull);
    atomicAdd(out + 1, (unsigned long long)(acc & 0xFF));
}

template <int V, int U>
void run(const char* name, int grid, size_t smem, const uint32_t* code, const uint16_t* src, size_t lpb, const uint32_t* vec,
         uint32_t chunk, unsigned long long* out) {
    CK(cudaFuncSetAttribute(push_kernel<V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        push_kernel<V, U><<<grid, 1024, smem>>>(code, src, lpb, vec, chunk, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    const double links = (double)lpb * grid;
    printf("  %-34s U=%d %8.3f ms  %7.2f Glink/s  %5.2f links/clk/SM@1.9GHz\n", name, U, best, links / best / 1e6,
           links / (best * 1e-3) / grid / 1.9e9);
    fflush(stdout);
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint32_t chunk = 42 * 1024;                    // source agents per block: 168 KB of packed counters
    const size_t smem = (size_t)chunk * 4;
    const size_t lpb_max = 1600 * 1024;
    const size_t E = lpb_max * sms;
    uint32_t* d_code; uint16_t* d_src; unsigned long long* d_out;
    CK(cudaMalloc(&d_code, E * 4)); CK(cudaMalloc(&d_src, E * 2)); CK(cudaMalloc(&d_out, 16)); CK(cudaMemset(d_out, 0, 16));
    // {agents, links per block}: 50M / 416k = the social links of one block's share of 50M agents on 8 GPUs (gap 120);
    // 5M / 512k = the cold neighbour links of that share inside a 5M-agent neighbourhood (gap 10); 50M / 1.6M (gap 31)
    const uint32_t cfg[3][2] = {{50000000u, 416u * 1024u}, {5000000u, 512u * 1024u}, {50000000u, 1600u * 1024u}};
    for (int ci = 0; ci < 3; ++ci) {
        const uint32_t n_agents = cfg[ci][0];
        const size_t lpb = cfg[ci][1];
        const size_t vec_bytes = (size_t)n_agents / 4 + 256;
        uint32_t* d_vec; CK(cudaMalloc(&d_vec, vec_bytes)); CK(cudaMemset(d_vec, 0x1B, vec_bytes));
        printf("%u agents (vector %.2f MB), %zu links per block (mean gap %.1f agents when sorted), %d blocks, %u source agents per block\n",
               n_agents, vec_bytes / 1e6, lpb, (double)n_agents / lpb, sms, chunk);
        for (int sorted = 1; sorted >= 0; --sorted) {
            fill<<<dim3(64, sms), 256>>>(d_code, d_src, lpb, n_agents, chunk, sorted);
            CK(cudaDeviceSynchronize());
            if (sorted) {
                run<kP, 2>("P push (sorted gathers + red.shared)", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
                run<kP, 4>("P push (sorted gathers + red.shared)", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
                run<kP, 8>("P push (sorted gathers + red.shared)", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
                run<kG, 4>("G sorted gathers only", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
                run<kR, 4>("R red.shared only", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
            } else {
                run<kU, 4>("U unsorted gathers only (pull)", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
                run<kU, 8>("U unsorted gathers only (pull)", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
                run<kP, 4>("unsorted gathers + red.shared", sms, smem, d_code, d_src, lpb, d_vec, chunk, d_out);
            }
        }
        CK(cudaFree(d_vec));
    }
    return 0;
}

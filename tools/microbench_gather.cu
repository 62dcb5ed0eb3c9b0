// microbench_gather.cu — design probe (dev tool, not product): how fast can an SM resolve the random
// "yesterday's mode of agent j" gathers that dominate the day step, depending on where the mode vector lives?
//   V0 stream only        : read the u32 link indices coalesced (uint4), no gather -> streaming ceiling
//   V1 global u8          : mode bytes in global memory (1 MB, L2/L1)         -> __ldg gather
//   V2 global 2-bit       : packed 2-bit vector in global (250 KB, L2/L1)     -> __ldg gather + shift
//   V3 smem 2-bit         : packed vector staged in shared memory (<= 227 KB) -> LDS gather + shift
//   V4 cluster-2 DSMEM    : packed vector split over the 2 CTAs of a cluster   -> ld.shared::cluster gather
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_gather microbench_gather.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int V, int U>
__global__ void __launch_bounds__(1024, 1)
gather_kernel(const uint4* __restrict__ idx, size_t n_vec, const uint8_t* __restrict__ mode8,
              const uint32_t* __restrict__ mode2, uint32_t n_agents, uint32_t n_smem_words, unsigned long long* out) {
    extern __shared__ uint32_t s_mode[];
    uint32_t* remote = nullptr;
    uint32_t half_words = 0, my_rank = 0;
    if (V == 3 || V == 5) {
        for (uint32_t k = threadIdx.x; k < n_smem_words; k += blockDim.x) s_mode[k] = mode2[k];
        __syncthreads();
    }
    if (V == 4) {
        cg::cluster_group cluster = cg::this_cluster();
        my_rank = cluster.block_rank();
        half_words = n_smem_words;                        // words per CTA
        for (uint32_t k = threadIdx.x; k < half_words; k += blockDim.x) s_mode[k] = mode2[my_rank * half_words + k];
        cluster.sync();
        remote = cluster.map_shared_rank(s_mode, my_rank ^ 1u);
    }
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * U) {
        uint4 vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) vv[u] = (i0 + u * stride < n_vec) ? ldg_stream(idx + i0 + u * stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t j[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t m;
                if (V == 0) m = j[k] & 3u;
                if (V == 1) m = __ldg(mode8 + j[k]);
                if (V == 2) m = (__ldg(mode2 + (j[k] >> 4)) >> ((j[k] & 15u) * 2u)) & 3u;
                if (V == 3) {
                    const uint32_t jj = j[k] < n_smem_words * 16u ? j[k] : j[k] - n_smem_words * 16u;   // fold the tail
                    m = (s_mode[jj >> 4] >> ((jj & 15u) * 2u)) & 3u;
                }
                if (V == 4) {
                    const uint32_t wd = j[k] >> 4;
                    const uint32_t owner = wd >= half_words;
                    const uint32_t* base = owner == my_rank ? s_mode : remote;
                    m = (base[wd - owner * half_words] >> ((j[k] & 15u) * 2u)) & 3u;
                }
                if (V == 5) {      // hybrid: hot part in shared memory, cold tail from global (L2)
                    const uint32_t wd = j[k] >> 4;
                    const uint32_t w = wd < n_smem_words ? s_mode[wd] : __ldg(mode2 + wd);
                    m = (w >> ((j[k] & 15u) * 2u)) & 3u;
                }
                acc += 1u << (8u * m);
            }
        }
    }
    if (V == 4) cg::this_cluster().sync();
    if (acc == 0xFFFFFFFFu) atomicAdd(out, 1ull);
    atomicAdd(out + 1, (unsigned long long)(acc & 0xFF));
}

template <int V, int U>
float run(const char* name, int grid, size_t smem, const uint4* idx, size_t n_vec, const uint8_t* m8, const uint32_t* m2,
          uint32_t n_agents, uint32_t words, unsigned long long* out, int cluster) {
    CK(cudaFuncSetAttribute(gather_kernel<V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0));
        if (cluster > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, gather_kernel<V, U>, idx, n_vec, m8, m2, n_agents, words, out));
        } else {
            gather_kernel<V, U><<<grid, 1024, smem>>>(idx, n_vec, m8, m2, n_agents, words, out);
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    const double gathers = (double)n_vec * 4;
    printf("%-22s %8.3f ms  idx stream %7.1f GB/s  %7.2f Ggather/s  %6.2f gathers/clk/SM@1.9GHz\n", name, best,
           n_vec * 16.0 / best / 1e6, gathers / best / 1e6, gathers / (best * 1e-3) / 148 / 1.9e9);
    return best;
}

int main() {
    const uint32_t N = 1000000;
    const size_t E = 30ull * N * 8;                      // 8 "replica-days" of link indices = 960 MB (> L2)
    const size_t n_vec = E / 4;
    std::vector<uint32_t> h(E);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < E; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % N); }
    uint32_t* d_idx; uint8_t* d_m8; uint32_t* d_m2; unsigned long long* d_out;
    CK(cudaMalloc(&d_idx, E * 4)); CK(cudaMalloc(&d_m8, N)); CK(cudaMalloc(&d_m2, N / 4 + 64)); CK(cudaMalloc(&d_out, 16));
    CK(cudaMemcpy(d_idx, h.data(), E * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_m8, 1, N)); CK(cudaMemset(d_m2, 0x55, N / 4 + 64)); CK(cudaMemset(d_out, 0, 16));
    int dev_sms = 0; CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d, links %zu (%.0f MB)\n", dev_sms, E, E * 4 / 1e6);
    const uint4* idx = (const uint4*)d_idx;
    const uint32_t words3 = 57000;                        // 228000 B: 912k agents resident
    const uint32_t words4 = N / 16 / 2;                   // 31250 words = 125 KB per CTA
#define ROW(U) \
    printf("--- %d uint4 loads in flight per thread\n", U); \
    run<0, U>("V0 stream only", dev_sms, 0, idx, n_vec, d_m8, d_m2, N, 0, d_out, 1); \
    run<1, U>("V1 global u8", dev_sms, 0, idx, n_vec, d_m8, d_m2, N, 0, d_out, 1); \
    run<2, U>("V2 global 2-bit", dev_sms, 0, idx, n_vec, d_m8, d_m2, N, 0, d_out, 1); \
    run<3, U>("V3 smem 2-bit (fold)", dev_sms, words3 * 4, idx, n_vec, d_m8, d_m2, N, words3, d_out, 1); \
    run<5, U>("V5 hot smem+cold L2", dev_sms, words3 * 4, idx, n_vec, d_m8, d_m2, N, words3, d_out, 1);
    ROW(1) ROW(2) ROW(4) ROW(8)
    run<4, 4>("V4 cluster2 DSMEM", dev_sms, words4 * 4, idx, n_vec, d_m8, d_m2, N, words4, d_out, 2);
    return 0;
}

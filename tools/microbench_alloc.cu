// microbench_alloc.cu — how long cudaMalloc / cudaFree take for the allocation pattern of mvb_create_batch
// (64 replicas: 64 x 130 MB + 64 x 26 MB) versus one large allocation (dev tool).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaFree(0);
    for (int rep = 0; rep < 3; ++rep) {
        std::vector<void*> p;
        double t0 = now();
        for (int r = 0; r < 64; ++r) { void* a; void* b; cudaMalloc(&a, 130u << 20); cudaMalloc(&b, 26u << 20); p.push_back(a); p.push_back(b); }
        double t1 = now();
        for (void* q : p) cudaMemsetAsync(q, 1, 26u << 20);
        cudaDeviceSynchronize();
        double t2 = now();
        for (void* q : p) cudaFree(q);
        double t3 = now();
        void* big;
        cudaMalloc(&big, (size_t)64 * (156u << 20));
        double t4 = now();
        cudaMemsetAsync(big, 1, (size_t)64 * (156u << 20)); cudaDeviceSynchronize();
        double t5 = now();
        cudaFree(big);
        double t6 = now();
        printf("128 buffers: malloc %.3f s, free %.3f s | one 10 GB buffer: malloc %.3f s, free %.3f s (memset %.3f / %.3f)\n",
               t1 - t0, t3 - t2, t4 - t3, t6 - t5, t2 - t1, t5 - t4);
    }
    return 0;
}

// mvb_internal.h — host-side internals of libmotivate_b200 (not part of the ABI).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/motivate_b200.h"

namespace mvb {

void set_error(const char* fmt, ...);

// SplitMix64: the seeded stream used by every synthetic generator here (the
// reference uses thread_rng / HC-128, which are not reproducible / not restated).
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline double f64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }      // [0,1)
    inline float  f32() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }              // [0,1)
    inline uint64_t below(uint64_t n) {   // unbiased enough for n << 2^64 (multiply-shift)
        return (uint64_t)(((unsigned __int128)next() * n) >> 64);
    }
};

}  // namespace mvb

#define MVB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            mvb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                   \
                           cudaGetErrorString(e_));                                             \
            return MVB_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

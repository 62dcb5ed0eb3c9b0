// synth.h — the object behind mvb_synth (internal): a synthetic population whose arrays live either in host vectors
// (mvb_synth_create, synth.cpp) or in one block of device memory (mvb_synth_create_device, synth_device.cuh).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/motivate_b200.h"

struct mvb_synth {
    mvb_population pop{};              // pointers into the vectors below, or into d_block
    std::vector<uint8_t> sub, commute, car, bike, mode, norm;
    std::vector<uint16_t> nbhd;
    std::vector<float> ws, habit;
    std::vector<uint64_t> s_rowptr, n_rowptr;
    std::vector<uint32_t> s_col, n_col;
    int device = -1;                   // >= 0: the arrays are device memory on this GPU
    void* d_block = nullptr;           // per-agent arrays + row pointers
    void* d_block2 = nullptr;          // column arrays
    uint64_t Es = 0, En = 0;           // link counts (= rowptr[N]) of a device-resident population
};

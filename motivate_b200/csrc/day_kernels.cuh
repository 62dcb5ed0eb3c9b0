// day_kernels.cuh — device side of the weekday step.
//
// Per-day record of a replica (u32 counters, one row per emitted statistics row):
//     hist[H][4]   residents' current_mode histogram  -> tomorrow's congestion input
//                  (neighbourhood.rs:68-75) and the per-neighbourhood statistics columns
//     joint[4]     index = active*2 + active_norm  (statistics.rs:14-57)
//     commute[3]   active by commute length        (statistics.rs:62-70)
//     sub[S]       active by subculture            (statistics.rs:75-83)
// A day's kernel reads yesterday's record (hist) and accumulates today's; records are
// zeroed once when allocated, never per day.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "decide.cuh"

namespace mvb {

struct RepDev {                     // device-side descriptor of one replica
    uint32_t N, S, H, rec_off;
    const uint32_t* stat;           // packed static word per agent (decide.cuh)
    const float* ws;                // weather_sensitivity
    float4* habit;                  // dense habit, updated in place
    uint8_t* mode[2];               // current_mode, double buffered by day parity
    uint8_t* norm;
    const float* pa;                // per-agent weights [5][N] or nullptr (uniform)
    const uint64_t* s_rowptr; const uint32_t* s_col;
    const uint64_t* n_rowptr; const uint32_t* n_col;
    const float* supp; const float* desir;       // [H][4], [S][4]
    const uint32_t* cap; const uint32_t* nres;   // [H][4], [H]
    float* cong;                    // [H][4] modifier used by the last step (for mvb_get_state)
};

__host__ __device__ inline uint32_t rec_width(uint32_t S, uint32_t H) { return 4 * H + 7 + S; }
// shared-memory words a block needs: cong[H][4] + cost_base[H][3][4] + desir[S][4] + record
__host__ __device__ inline uint32_t smem_words(uint32_t S, uint32_t H) {
    return 4 * H + 12 * H + 4 * S + rec_width(S, H);
}

// Block-level tables: congestion from yesterday's histogram, pre-weather cost, desirability.
__device__ __forceinline__ void build_block_tables(const RepDev& r, const uint32_t* rec_prev, float* s_cong,
                                                   float* s_cost, float* s_des, uint32_t* s_rec,
                                                   bool publish_cong) {
    const uint32_t H = r.H, S = r.S;
    for (uint32_t k = threadIdx.x; k < 4 * H; k += blockDim.x) {
        const uint32_t h = k >> 2;
        const float c = congestion_entry(rec_prev[r.rec_off + k], r.cap[k], r.nres[h]);
        s_cong[k] = c;
        if (publish_cong) r.cong[k] = c;
    }
    for (uint32_t k = threadIdx.x; k < 12 * H; k += blockDim.x) {
        const uint32_t h = k / 12, c = (k >> 2) % 3, m = k & 3;
        s_cost[k] = cost_base_entry(r.supp[4 * h + m], c, m);
    }
    for (uint32_t k = threadIdx.x; k < 4 * S; k += blockDim.x) s_des[k] = r.desir[k];
    for (uint32_t k = threadIdx.x; k < rec_width(S, H); k += blockDim.x) s_rec[k] = 0;
}

__device__ __forceinline__ void tally(uint32_t* s_rec, uint32_t H, uint32_t stat, uint32_t mode, uint32_t norm) {
    const uint32_t a = mode >= 2u, an = norm >= 2u;
    atomicAdd(&s_rec[4 * st_nbhd(stat) + mode], 1u);
    atomicAdd(&s_rec[4 * H + a * 2 + an], 1u);
    if (a) {
        atomicAdd(&s_rec[4 * H + 4 + st_commute(stat)], 1u);
        atomicAdd(&s_rec[4 * H + 7 + st_sub(stat)], 1u);
    }
}

// Census of the current state: fills one record row (day-0 row, mvb_stats_row).
__global__ void census_kernel(const RepDev* reps, uint32_t parity, uint32_t* rec) {
    extern __shared__ uint32_t smem[];
    const RepDev r = reps[blockIdx.y];
    if (blockIdx.x * blockDim.x >= r.N) return;
    const uint32_t W = rec_width(r.S, r.H);
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x) smem[k] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r.N) tally(smem, r.H, r.stat[i], r.mode[parity][i], r.norm[i]);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x)
        if (smem[k]) atomicAdd(&rec[r.rec_off + k], smem[k]);
}

// Correctness-first weekday kernel: one thread per agent, link lists walked straight
// from the CSR, yesterday's modes gathered from global memory (L2).  simulation.rs:116-128.
template <bool PER_AGENT>
__global__ void __launch_bounds__(256)
day_kernel_simple(const RepDev* reps, uint32_t parity, const uint32_t* rec_prev, uint32_t* rec_next,
                  uint32_t weather, uint32_t changed, AgentWeights uw) {
    extern __shared__ uint32_t smem[];
    const RepDev r = reps[blockIdx.y];
    if (blockIdx.x * blockDim.x >= r.N) return;
    float* s_cong = reinterpret_cast<float*>(smem);
    float* s_cost = s_cong + 4 * r.H;
    float* s_des  = s_cost + 12 * r.H;
    uint32_t* s_rec = smem + 16 * r.H + 4 * r.S;
    build_block_tables(r, rec_prev, s_cong, s_cost, s_des, s_rec, blockIdx.x == 0);
    __syncthreads();

    const uint8_t* __restrict__ prev_mode = r.mode[parity];
    uint8_t* __restrict__ next_mode = r.mode[parity ^ 1u];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r.N) {
        const uint32_t stat = r.stat[i];
        const uint32_t last = prev_mode[i];                         // update_habit: last = current
        AgentWeights w = uw;
        if (PER_AGENT && r.pa) {
            w.consistency = r.pa[i];                 w.social_c = r.pa[(size_t)r.N + i];
            w.subculture_c = r.pa[2 * (size_t)r.N + i]; w.neighbour_c = r.pa[3 * (size_t)r.N + i];
            w.average_weight = r.pa[4 * (size_t)r.N + i];
        }
        float4 hv = r.habit[i];
        float h[4] = {hv.x, hv.y, hv.z, hv.w};
        habit_update(h, last, w.average_weight);
        r.habit[i] = make_float4(h[0], h[1], h[2], h[3]);

        uint32_t cs[4] = {0, 0, 0, 0}, cn[4] = {0, 0, 0, 0};
        const uint64_t sb = r.s_rowptr[i], se = r.s_rowptr[i + 1];
        for (uint64_t e = sb; e < se; ++e) {
            const uint32_t m = prev_mode[r.s_col[e]];
            cs[0] += m == 0; cs[1] += m == 1; cs[2] += m == 2; cs[3] += m == 3;
        }
        const uint64_t nb = r.n_rowptr[i], ne = r.n_rowptr[i + 1];
        for (uint64_t e = nb; e < ne; ++e) {
            const uint32_t m = prev_mode[r.n_col[e]];
            cn[0] += m == 0; cn[1] += m == 1; cn[2] += m == 2; cn[3] += m == 3;
        }
        uint32_t mode, norm;
        decide(cs, (uint32_t)(se - sb), cn, (uint32_t)(ne - nb), w, h, s_des + 4 * st_sub(stat),
               s_cong + 4 * st_nbhd(stat), s_cost + 12 * st_nbhd(stat) + 4 * st_commute(stat),
               st_commute(stat), st_car(stat), st_bike(stat), r.ws[i], last, weather, changed, mode, norm);
        next_mode[i] = (uint8_t)mode;
        r.norm[i] = (uint8_t)norm;
        tally(s_rec, r.H, stat, mode, norm);
    }
    __syncthreads();
    const uint32_t W = rec_width(r.S, r.H);
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x)
        if (s_rec[k]) atomicAdd(&rec_next[r.rec_off + k], s_rec[k]);
}

}  // namespace mvb

// day_kernels.cuh — device side of the weekday step (src/simulation.rs:116-135) for sm_100a.
//
// One launch per simulated weekday steps EVERY replica resident on the GPU.  The grid is one
// persistent 1024-thread block per SM.  Work is cut into R*P "parts" (replica r, class p of P);
// a block takes parts blockIdx.x, blockIdx.x + gridDim.x, ...  For each part it
//   1. stages the hot part of yesterday's 2-bit mode vector of replica r in shared memory with
//      TMA bulk copies (cp.async.bulk + mbarrier) and builds the per-replica tables
//      (congestion from yesterday's histogram, pre-weather cost, desirability);
//   2. hub slices p, p+P, ... : one WARP per hub agent streams that agent's link codes
//      (uint4 per lane, coalesced), gathers the targets' modes from shared memory, and leaves 8
//      counts per hub in shared memory; warp 0 then finishes the 32 hubs lane-per-agent;
//   3. SELL slices (taken from a block-wide counter, longest first): one LANE per agent, link
//      codes read k-major (uint4 body + u32 tail, next row prefetched while the current one is
//      gathered), modes collected 2 bits at a time and counted with popc;
//   4. per agent (agent.rs:244-316): habit EMA, shares, norm, budget/cost ranks, choice; the
//      warp packs its 32 new modes into one u64 of tomorrow's vector; statistics are reduced
//      with warp REDUX and a handful of shared-memory atomics, flushed once per part.
//
// Per-day record of a replica (u32 counters, one row per emitted statistics row):
//     hist[H][4]   residents' current_mode histogram  -> tomorrow's congestion input
//                  (neighbourhood.rs:68-75) and the per-neighbourhood statistics columns
//     joint[4]     index = active*2 + active_norm  (statistics.rs:14-57)
//     commute[3]   active by commute length        (statistics.rs:62-70)
//     sub[S]       active by subculture            (statistics.rs:75-83)
// A day's kernel reads yesterday's record (hist) and accumulates today's; records are zeroed
// once when allocated, never per day.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "decide.cuh"
#include "layout.h"

namespace mvb {

constexpr uint32_t kThreads = 1024;
constexpr uint32_t kWarps = kThreads / 32;
constexpr uint32_t kCodeAddrMask = 0x07FFFFFCu;     // byte offset of the vector word (shared or global)
constexpr uint32_t kStatValid = 0x80000000u;        // bit 31 of the packed static word: a real agent

struct RepDev {                     // device-side descriptor of one replica
    uint32_t N, N_pad, S, H, rec_off;
    uint32_t n_hub_slices, n_sell_slices, n_ranges, hot_words;
    const uint32_t* stat;           // [N_pad] packed static word (decide.cuh) | kStatValid
    const float* ws;                // [N_pad] weather_sensitivity
    float4* habit;                  // [N_pad] dense habit, updated in place
    const float* pa;                // per-agent weights [5][N_pad] or nullptr (uniform)
    uint2* vec[2];                  // [N_pad/32] packed current_mode per slice, double buffered by day parity
    uint2* normvec;                 // [N_pad/32] packed norm per slice
    const uint32_t* deg;            // [N_pad] SELL agents: ds | dn << 16
    const StageRange* ranges;
    const uint64_t* hub_rowptr; const uint32_t* hub_ds; const uint32_t* hub_dn; const uint32_t* hub_codes;
    const uint64_t* sell_off; const uint32_t* sell_K; const uint32_t* sell_order; const uint32_t* sell_codes;
    const float* supp; const float* desir;       // [H][4], [S][4]
    const uint32_t* cap; const uint32_t* nres;   // [H][4], [H]
    float* cong;                    // [H][4] modifier used by the last step (for mvb_get_state)
};

__host__ __device__ inline uint32_t rec_width(uint32_t S, uint32_t H) { return 4 * H + 7 + S; }
// words of shared memory a block needs besides the staged vector:
// mbarrier(2) + counter(2) + hub counts [2][32][8] + replica descriptor + cong[H][4] + cost_base[H][3][4] + desir[S][4] + record
constexpr uint32_t kRepWords = 64;                   // sizeof(RepDev) rounded up, in u32
static_assert(sizeof(RepDev) <= kRepWords * 4, "RepDev outgrew its shared-memory slot");
__host__ __device__ inline uint32_t tail_words(uint32_t S, uint32_t H) {
    return 4 + 512 + kRepWords + 4 * H + 12 * H + 4 * S + rec_width(S, H);
}

// ---- small PTX helpers ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p) {       // read-once data: do not pollute L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- gathering yesterday's modes ---------------------------------------------------------------
// Link code (layout.h): bits 31-27 = 2*position in the word, bits 26-2 = word index, bit 1 = cold.
__device__ __forceinline__ uint32_t gather_hot(const uint32_t* s_vec, uint32_t code) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(s_vec) + (code & kCodeAddrMask));
    return (w >> (code >> 27)) & 3u;
}
__device__ __forceinline__ uint32_t gather_any(const uint32_t* s_vec, const uint32_t* g_vec, uint32_t code) {
    uint32_t w;
    if (code & 2u) w = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(g_vec) + (code & kCodeAddrMask)));
    else w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(s_vec) + (code & kCodeAddrMask));
    return (w >> (code >> 27)) & 3u;
}

// `bits` holds up to 16 gathered modes, field f (bits 2f+1..2f) = link number base+f of a list whose first `ds`
// links are social and the following are neighbour links, `d` in total (fields at or past d are padding).  Adds the per-mode counts to the packed
// 8-bit accumulators (social, neighbour).  count_in_subgroup's grouping, agent.rs:325-329.
template <int NF>   // NF = fields of `bits` that were filled (16 for a SELL group, 4 for one uint4 of a hub row)
__device__ __forceinline__ void count_fields(uint32_t bits, int base, int ds, int d, uint32_t& acc_s, uint32_t& acc_n) {
    const int nv = min(max(d - base, 0), NF), ns = min(max(ds - base, 0), NF);
    const uint32_t vmask = (nv >= 16 ? 0xFFFFFFFFu : ((1u << (2 * nv)) - 1u)) & 0x55555555u;
    const uint32_t smask = (ns >= 16 ? 0xFFFFFFFFu : ((1u << (2 * ns)) - 1u)) & 0x55555555u;
    const uint32_t nmask = vmask & ~smask;
    const uint32_t lo = bits & 0x55555555u, hi = (bits >> 1) & 0x55555555u;
    const uint32_t m3 = hi & lo, m2 = hi & ~lo, m1 = ~hi & lo, m0 = ~(hi | lo);
    acc_s += __popc(m0 & smask) | __popc(m1 & smask) << 8 | __popc(m2 & smask) << 16 | __popc(m3 & smask) << 24;
    acc_n += __popc(m0 & nmask) | __popc(m1 & nmask) << 8 | __popc(m2 & nmask) << 16 | __popc(m3 & nmask) << 24;
}

// ---- per-warp statistics -------------------------------------------------------------------------
struct WarpStats {                  // 16-bit packed running sums (two u32 per class group), flushed per part
    uint32_t joint_lo, joint_hi, commute_lo, commute_hi, sub_lo, sub_hi;
    uint32_t slices;
};

// Everything that happens once the link counts of 32 agents (one per lane) are known.
template <bool PER_AGENT>
__device__ __forceinline__ void finish_slice(const RepDev& r, uint32_t slice, uint32_t lane, bool uniform_nbhd,
                                             const uint32_t cs[4], uint32_t ds, const uint32_t cn[4], uint32_t dn,
                                             uint32_t stat, float wsv, float4 hv, uint32_t last,
                                             const AgentWeights& uw, const float* s_cong, const float* s_cost,
                                             const float* s_des, uint32_t* s_rec, uint32_t weather, uint32_t changed,
                                             uint32_t parity, WarpStats& st) {
    const uint32_t agent = slice * 32 + lane;
    const bool valid = (stat & kStatValid) != 0;
    AgentWeights w = uw;
    if (PER_AGENT && r.pa) {
        const size_t np = r.N_pad;
        w.consistency = r.pa[agent];             w.social_c = r.pa[np + agent];
        w.subculture_c = r.pa[2 * np + agent];   w.neighbour_c = r.pa[3 * np + agent];
        w.average_weight = r.pa[4 * np + agent];
    }
    float h[4] = {hv.x, hv.y, hv.z, hv.w};
    habit_update(h, last, w.average_weight);
    uint32_t mode, norm;
    decide(cs, ds, cn, dn, w, h, s_des + 4 * st_sub(stat), s_cong + 4 * st_nbhd(stat),
           s_cost + 12 * st_nbhd(stat) + 4 * st_commute(stat), st_commute(stat), st_car(stat), st_bike(stat), wsv,
           last, weather, changed, mode, norm);
    if (!valid) { mode = 0; norm = 0; }
    if (valid) r.habit[agent] = make_float4(h[0], h[1], h[2], h[3]);
    // 32 agents x 2 bits -> one u64 (lanes 0-15 in .x, 16-31 in .y)
    const uint32_t sh = (lane & 15u) * 2u;
    const uint32_t mlo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? mode << sh : 0u);
    const uint32_t mhi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : mode << sh);
    const uint32_t nlo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? norm << sh : 0u);
    const uint32_t nhi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : norm << sh);
    if (lane == 0) r.vec[parity ^ 1u][slice] = make_uint2(mlo, mhi);
    if (lane == 1) r.normvec[slice] = make_uint2(nlo, nhi);
    // statistics (statistics.rs:14-93)
    const uint32_t a = valid && mode >= 2u, an = valid && norm >= 2u;
    if (uniform_nbhd) {
        const uint32_t hsum = __reduce_add_sync(0xFFFFFFFFu, valid ? 1u << (8u * mode) : 0u);   // 4 counts <= 32 each
        const uint32_t nb = __shfl_sync(0xFFFFFFFFu, st_nbhd(stat), 0);   // lane 0 of a non-hub slice is always valid
        if (lane < 4) {
            const uint32_t c = (hsum >> (8u * lane)) & 0xFFu;
            if (c) atomicAdd(&s_rec[4 * nb + lane], c);
        }
    } else if (valid) {
        atomicAdd(&s_rec[4 * st_nbhd(stat) + mode], 1u);
    }
    const uint32_t jsum = __reduce_add_sync(0xFFFFFFFFu, valid ? 1u << (8u * (a * 2u + an)) : 0u);
    const uint32_t csum = __reduce_add_sync(0xFFFFFFFFu, a ? 1u << (8u * st_commute(stat)) : 0u);
    st.joint_lo += jsum & 0x00FF00FFu;   st.joint_hi += (jsum >> 8) & 0x00FF00FFu;
    st.commute_lo += csum & 0x00FF00FFu; st.commute_hi += (csum >> 8) & 0x00FF00FFu;
    if (r.S <= 4) {
        const uint32_t ssum = __reduce_add_sync(0xFFFFFFFFu, a ? 1u << (8u * st_sub(stat)) : 0u);
        st.sub_lo += ssum & 0x00FF00FFu; st.sub_hi += (ssum >> 8) & 0x00FF00FFu;
    } else if (a) {
        atomicAdd(&s_rec[4 * r.H + 7 + st_sub(stat)], 1u);
    }
    st.slices++;
}

__device__ __forceinline__ void flush_warp_stats(const RepDev& r, uint32_t lane, uint32_t* s_rec, WarpStats& st) {
    // 16-bit fields: lo = classes 0,2 ; hi = classes 1,3
    if (lane < 4) {
        const uint32_t j = ((lane & 1u) ? st.joint_hi : st.joint_lo) >> (16u * (lane >> 1)) & 0xFFFFu;
        if (j) atomicAdd(&s_rec[4 * r.H + lane], j);
        const uint32_t c = ((lane & 1u) ? st.commute_hi : st.commute_lo) >> (16u * (lane >> 1)) & 0xFFFFu;
        if (c && lane < 3) atomicAdd(&s_rec[4 * r.H + 4 + lane], c);
        const uint32_t s = ((lane & 1u) ? st.sub_hi : st.sub_lo) >> (16u * (lane >> 1)) & 0xFFFFu;
        if (s && r.S <= 4 && lane < r.S) atomicAdd(&s_rec[4 * r.H + 7 + lane], s);
    }
    st = WarpStats{0, 0, 0, 0, 0, 0, 0};
}

// ---- the weekday kernel ----------------------------------------------------------------------------
template <bool PER_AGENT>
__global__ void __launch_bounds__(kThreads, 1)
day_kernel(const RepDev* __restrict__ reps, uint32_t R, uint32_t P, uint32_t parity,
           const uint32_t* __restrict__ rec_prev, uint32_t* __restrict__ rec_next, uint32_t weather,
           uint32_t changed, AgentWeights uw, uint32_t vec_cap_words) {
    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_vec = smem;
    uint32_t* s_tail = smem + vec_cap_words;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tail);
    uint32_t* s_next = s_tail + 2;
    uint32_t* s_hub = s_tail + 4;                     // [2][32][8]
    uint32_t* s_rep = s_hub + 512;                    // RepDev of the current part
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) mbar_init(s_bar, 1);
    uint32_t phase = 0;

    for (uint32_t part = blockIdx.x; part < R * P; part += gridDim.x) {
        const uint32_t p = part % P;
        __syncthreads();                              // previous part is done with shared memory
        if (tid < sizeof(RepDev) / 4) s_rep[tid] = reinterpret_cast<const uint32_t*>(reps + part / P)[tid];
        __syncthreads();
        const RepDev& r = *reinterpret_cast<const RepDev*>(s_rep);
        const uint32_t H = r.H, S = r.S;
        float* s_cong = reinterpret_cast<float*>(s_rep + kRepWords);
        float* s_cost = s_cong + 4 * H;
        float* s_des = s_cost + 12 * H;
        uint32_t* s_rec = reinterpret_cast<uint32_t*>(s_des + 4 * S);
        const uint32_t* __restrict__ g_vec = reinterpret_cast<const uint32_t*>(r.vec[parity]);

        if (tid == 0) {                               // 1. stage yesterday's modes (TMA bulk copies)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(s_bar, r.hot_words * 4u);
            for (uint32_t k = 0; k < r.n_ranges; ++k) {
                const StageRange g = r.ranges[k];
                for (uint32_t o = 0; o < g.n_words; o += 16384) {          // <= 64 KB per copy
                    const uint32_t n = min(g.n_words - o, 16384u);
                    bulk_g2s(s_vec + g.dst_word + o, g_vec + g.src_word + o, n * 4u, s_bar);
                }
            }
            *s_next = 0;
        }
        for (uint32_t k = tid; k < 4 * H; k += kThreads) {                 // neighbourhood.rs:57-89
            const float c = congestion_entry(rec_prev[r.rec_off + k], r.cap[k], r.nres[k >> 2]);
            s_cong[k] = c;
            if (p == 0) r.cong[k] = c;
        }
        for (uint32_t k = tid; k < 12 * H; k += kThreads) {                // agent.rs:185-196
            const uint32_t hh = k / 12, c = (k >> 2) % 3, m = k & 3;
            s_cost[k] = cost_base_entry(r.supp[4 * hh + m], c, m);
        }
        for (uint32_t k = tid; k < 4 * S; k += kThreads) s_des[k] = r.desir[k];
        for (uint32_t k = tid; k < rec_width(S, H); k += kThreads) s_rec[k] = 0;
        __syncthreads();
        mbar_wait(s_bar, phase);
        phase ^= 1u;

        WarpStats st{0, 0, 0, 0, 0, 0, 0};

        // 2. hub slices: one warp per hub agent
        uint32_t hbuf = 0;
        for (uint32_t hs = p; hs < r.n_hub_slices; hs += P, hbuf ^= 1u) {
            const uint32_t hub = hs * 32 + warp;
            const uint64_t b = r.hub_rowptr[hub], e = r.hub_rowptr[hub + 1];
            const int ds = (int)r.hub_ds[hub], d = ds + (int)r.hub_dn[hub];
            uint32_t cs[4] = {0, 0, 0, 0}, cn[4] = {0, 0, 0, 0};
            uint32_t acc_s = 0, acc_n = 0, rows = 0;
            const uint4* codes = reinterpret_cast<const uint4*>(r.hub_codes + b);
            const uint32_t n4 = (uint32_t)((e - b) >> 2);
            for (uint32_t q0 = 0; q0 < n4; q0 += 64) {                     // two rows of 32 uint4 per iteration
                const uint32_t qa = q0 + lane, qb = q0 + 32 + lane;
                uint4 va = make_uint4(0, 0, 0, 0), vb = va;
                if (qa < n4) va = ld_stream_v4(codes + qa);
                if (qb < n4) vb = ld_stream_v4(codes + qb);
                uint32_t bits = gather_any(s_vec, g_vec, va.x) | gather_any(s_vec, g_vec, va.y) << 2 |
                                gather_any(s_vec, g_vec, va.z) << 4 | gather_any(s_vec, g_vec, va.w) << 6;
                count_fields<4>(bits, (int)(qa * 4), ds, d, acc_s, acc_n);
                bits = gather_any(s_vec, g_vec, vb.x) | gather_any(s_vec, g_vec, vb.y) << 2 |
                       gather_any(s_vec, g_vec, vb.z) << 4 | gather_any(s_vec, g_vec, vb.w) << 6;
                count_fields<4>(bits, (int)(qb * 4), ds, d, acc_s, acc_n);
                if (++rows == 30) {                                        // 8 links per iteration: < 256 per field
#pragma unroll
                    for (int m = 0; m < 4; ++m) { cs[m] += (acc_s >> (8 * m)) & 0xFFu; cn[m] += (acc_n >> (8 * m)) & 0xFFu; }
                    acc_s = acc_n = 0; rows = 0;
                }
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                cs[m] = __reduce_add_sync(0xFFFFFFFFu, cs[m] + ((acc_s >> (8 * m)) & 0xFFu));
                cn[m] = __reduce_add_sync(0xFFFFFFFFu, cn[m] + ((acc_n >> (8 * m)) & 0xFFu));
            }
            if (lane < 8) {
                const uint32_t v = lane == 0 ? cs[0] : lane == 1 ? cs[1] : lane == 2 ? cs[2] : lane == 3 ? cs[3]
                                 : lane == 4 ? cn[0] : lane == 5 ? cn[1] : lane == 6 ? cn[2] : cn[3];
                s_hub[hbuf * 256 + warp * 8 + lane] = v;
            }
            __syncthreads();
            if (warp == 0) {
                const uint32_t agent = hs * 32 + lane;
                uint32_t hcs[4], hcn[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) { hcs[m] = s_hub[hbuf * 256 + lane * 8 + m]; hcn[m] = s_hub[hbuf * 256 + lane * 8 + 4 + m]; }
                const uint32_t stat = r.stat[agent];
                const uint2 own = r.vec[parity][hs];
                const uint32_t last = ((lane < 16 ? own.x : own.y) >> ((lane & 15u) * 2u)) & 3u;
                finish_slice<PER_AGENT>(r, hs, lane, false, hcs, r.hub_ds[agent], hcn, r.hub_dn[agent], stat, r.ws[agent],
                                        r.habit[agent], last, uw, s_cong, s_cost, s_des, s_rec, weather, changed, parity, st);
            }
        }

        // 3. SELL slices: one lane per agent, slices handed out longest-first from a block-wide counter
        for (;;) {
            uint32_t q = 0;
            if (lane == 0) q = atomicAdd(s_next, 1u);
            q = __shfl_sync(0xFFFFFFFFu, q, 0);
            const uint64_t idx = (uint64_t)p + (uint64_t)P * q;
            if (idx >= r.n_sell_slices) break;
            const uint32_t s = r.sell_order[idx];
            const uint32_t slice = r.n_hub_slices + s, agent = slice * 32 + lane;
            const uint32_t K = r.sell_K[s], K4 = K >> 2;
            const uint4* body = reinterpret_cast<const uint4*>(r.sell_codes + r.sell_off[s]) + lane;
            // the first rows of link codes and the agent's own state go out together
            uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
            if (K4 > 0) cur = ld_stream_v4(body);
            if (K4 > 1) nxt = ld_stream_v4(body + 32);
            const uint32_t stat = r.stat[agent];
            const uint32_t dg = r.deg[agent];
            const float wsv = r.ws[agent];
            const float4 hv = r.habit[agent];
            const uint2 own = r.vec[parity][slice];
            const int ds = (int)(dg & 0xFFFFu), d = ds + (int)(dg >> 16);
            uint32_t acc_s = 0, acc_n = 0, bits = 0;
            for (uint32_t q4 = 0; q4 < K4; ++q4) {
                uint4 pre = make_uint4(0, 0, 0, 0);
                if (q4 + 2 < K4) pre = ld_stream_v4(body + (size_t)(q4 + 2) * 32);       // keep two rows in flight
                const uint32_t f = (q4 & 3u) * 8u;
                if (__builtin_expect(((cur.x | cur.y | cur.z | cur.w) & 2u) == 0u, 1)) {
                    bits |= (gather_hot(s_vec, cur.x) | gather_hot(s_vec, cur.y) << 2 | gather_hot(s_vec, cur.z) << 4 |
                             gather_hot(s_vec, cur.w) << 6) << f;
                } else {
                    bits |= (gather_any(s_vec, g_vec, cur.x) | gather_any(s_vec, g_vec, cur.y) << 2 |
                             gather_any(s_vec, g_vec, cur.z) << 4 | gather_any(s_vec, g_vec, cur.w) << 6) << f;
                }
                if ((q4 & 3u) == 3u) { count_fields<16>(bits, (int)(q4 >> 2) * 16, ds, d, acc_s, acc_n); bits = 0; }
                cur = nxt; nxt = pre;
            }
            {   // tail: K % 4 single codes per lane, after the last complete group of the body
                const uint32_t* tail = r.sell_codes + r.sell_off[s] + (size_t)K4 * 128 + lane;
                const uint32_t Kt = K & 3u;
                uint32_t f = (K4 & 3u) * 8u;
                for (uint32_t t = 0; t < Kt; ++t, f += 2u) bits |= gather_any(s_vec, g_vec, ld_stream_u32(tail + t * 32)) << f;
                if ((K4 & 3u) != 0u || Kt != 0u) count_fields<16>(bits, (int)(K4 >> 2) * 16, ds, d, acc_s, acc_n);
            }
            uint32_t cs[4], cn[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) { cs[m] = (acc_s >> (8 * m)) & 0xFFu; cn[m] = (acc_n >> (8 * m)) & 0xFFu; }
            const uint32_t last = ((lane < 16 ? own.x : own.y) >> ((lane & 15u) * 2u)) & 3u;
            finish_slice<PER_AGENT>(r, slice, lane, true, cs, (uint32_t)ds, cn, (uint32_t)(d - ds), stat, wsv, hv, last, uw,
                                    s_cong, s_cost, s_des, s_rec, weather, changed, parity, st);
            if (st.slices >= 1024) flush_warp_stats(r, lane, s_rec, st);
        }
        flush_warp_stats(r, lane, s_rec, st);
        __syncthreads();
        for (uint32_t k = tid; k < rec_width(S, H); k += kThreads)
            if (s_rec[k]) atomicAdd(&rec_next[r.rec_off + k], s_rec[k]);
    }
}

// Census of the current state: fills one record row (day-0 row, mvb_stats_row).  Runs once per create.
__global__ void census_kernel(const RepDev* reps, uint32_t parity, uint32_t* rec) {
    extern __shared__ uint32_t s_rec[];
    const RepDev& r = reps[blockIdx.y];
    if (blockIdx.x * blockDim.x >= r.N_pad) return;
    const uint32_t W = rec_width(r.S, r.H);
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x) s_rec[k] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stat = i < r.N_pad ? r.stat[i] : 0u;
    if (stat & kStatValid) {
        const uint2 mv = r.vec[parity][i >> 5], nv = r.normvec[i >> 5];
        const uint32_t l = i & 31u, sh = (l & 15u) * 2u;
        const uint32_t mode = ((l < 16 ? mv.x : mv.y) >> sh) & 3u, norm = ((l < 16 ? nv.x : nv.y) >> sh) & 3u;
        const uint32_t a = mode >= 2u, an = norm >= 2u;
        atomicAdd(&s_rec[4 * st_nbhd(stat) + mode], 1u);
        atomicAdd(&s_rec[4 * r.H + a * 2 + an], 1u);
        if (a) {
            atomicAdd(&s_rec[4 * r.H + 4 + st_commute(stat)], 1u);
            atomicAdd(&s_rec[4 * r.H + 7 + st_sub(stat)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x)
        if (s_rec[k]) atomicAdd(&rec[r.rec_off + k], s_rec[k]);
}

}  // namespace mvb

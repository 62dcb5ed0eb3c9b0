// day_kernels.cuh — device side of the weekday step (src/simulation.rs:116-135) for sm_100a.
//
// One launch per simulated weekday steps EVERY replica resident on the GPU.  The grid is one
// persistent 1024-thread block per SM.  Work is cut into R*P "parts" (replica r, class p of P);
// a block takes parts blockIdx.x, blockIdx.x + gridDim.x, ...  For each part it
//   1. stages the hot part of yesterday's 2-bit mode vector of replica r in shared memory with
//      TMA bulk copies (cp.async.bulk + mbarrier) and builds the per-replica tables
//      (congestion from yesterday's histogram, pre-weather cost, desirability);
//   2. hub slices p, p+P, ... : one WARP per hub agent; the hub's link list is cut into 32
//      contiguous chunks, one per lane, and goes through the same routine as a SELL slice
//      (accumulate_block); the warp then adds its lanes up and leaves 8 counts per hub in shared
//      memory; warp 0 finishes the 32 hubs lane-per-agent;
//   3. SELL slices (taken from a block-wide counter, longest first): one LANE per agent, link
//      codes read k-major (uint4 body + u32 tail, two rows prefetched ahead of the one being
//      gathered), each target's 2 bits pulled to the top of its word and pushed into a 32-bit
//      collector by two funnel shifts, 16 links at a time counted with popc;
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

constexpr uint32_t kCodeAddrMask = 0x07FFFFFCu;     // byte offset of the vector word (shared or global)
constexpr uint32_t kStatValid = 0x80000000u;        // bit 31 of the packed static word: a real agent

// Agent-partitioned mode with the exchange fused into the day kernel (sim.cu: exchange over peer memory).  Every rank
// maps every other rank's buffers (CUDA IPC over NVLink / NVSwitch); the day kernel stores each new mode word into ALL
// ranks' vectors, the last block to finish copies the rank's partial record into every rank's landing zone and then
// raises this rank's flag on every rank; exchange_finish_kernel waits for all flags and adds the partial records up.
constexpr uint32_t kMaxPeers = 16;
struct PeerExchange {
    uint32_t world, rank, rec_stride, pad_;
    uint2* vec[2][kMaxPeers];       // every rank's mode vectors by day parity ([..][rank] = this rank's own)
    uint2* normvec[kMaxPeers];
    uint32_t* xrec[kMaxPeers];      // every rank's landing zone for partial records: [2 (epoch parity)][world][rec_stride]
    uint32_t* flags[kMaxPeers];     // every rank's flags [world]: flags[p][q] = last epoch rank q has finished, as seen on rank p
    uint32_t* done;                 // this rank: blocks of the current launch that have finished
    uint32_t* error;                // this rank: set if a wait for the other ranks timed out
};

struct RepDev {                     // device-side descriptor of one replica
    uint32_t N, N_pad, S, H, rec_off;
    uint32_t n_hub_slices, n_sell_slices, n_ranges, hot_words;
    uint32_t hub_begin, hub_end, sell_begin, sell_end;   // slices this rank updates (everything unless agent-partitioned)
    uint32_t* state;                // [N_pad / 32] slice records (layout.h: static word, list lengths, weather
                                    // sensitivity, habit of the slice's 32 agents), the habit updated in place
    const float* pa;                // per-agent weights [5][N_pad] or nullptr (uniform)
    uint2* vec[2];                  // [N_pad/32] packed current_mode per slice, double buffered by day parity
    uint2* normvec;                 // [N_pad/32] packed norm per slice
    const StageRange* ranges;
    const uint64_t* hub_off; const uint4* hub_deg; const uint32_t* hub_codes;   // per hub: block offset, (ds_hot, dn_hot, ds_cold, dn_cold)
    const uint4* hub_units; const uint32_t* hub_unit_begin;   // units (<= 1 024 links) of every hub slice, layout.h
    uint32_t* hub_cnt;              // [n_hub_slices*32][8] scratch, zero between days: [1..3] social, [5..7] neighbour counts of modes 1..3
    const uint4* sell_meta; const uint32_t* sell_codes;       // per slice {off_lo, off_hi, K | Kc << 16, 0}
    const float* supp; const float* desir;       // [H][4], [S][4]
    const uint32_t* cap; const uint32_t* nres;   // [H][4], [H]
    float* cong;                    // [H][4] modifier used by the last step (for mvb_get_state)
    const uint32_t* int2ext;        // [N_pad] internal slot -> caller's agent id (kInvalid = padding); not read by the day kernel
    const uint32_t* nb_slice_begin; // [H + 1] SELL slices of every neighbourhood (LOCAL layout: parts are neighbourhood-pure)
    uint32_t local_mode, cold_cg;   // layout.h LayoutMode; cold gathers bypass L1 (ld_cold)
    const uint4* part_table;        // LOCAL layout: part p of a launch = {neighbourhood, class q, classes Q of that neighbourhood, 0}
};
// What only a handle with ONE population needs (push mode, agent-partitioned mode): a kernel parameter of its own, so that
// the 64 replica descriptors of the headline launches stay as small as they were (the parameter block is copied per launch).
struct SoloExtra {
    const uint2* push_cnt;          // push mode (layout.h): per-mode counts {social, neighbour} of an agent's cold links, filled by
                                    // push_kernel before the day kernel, indexed by internal slot (pointer pre-offset); else nullptr
    const PeerExchange* px;         // agent-partitioned mode, exchange over peer memory (day_kernel<.., XCHG = true>), else nullptr
    uint32_t px_epoch, px_norms;    // this launch's epoch (flag value raised when the rank has finished); also store the norm words
};

__host__ __device__ inline uint32_t rec_width(uint32_t S, uint32_t H) { return 4 * H + 7 + S; }
// words of shared memory a block needs besides the staged vector:
// mbarrier(2) + work counter(2) + good-weather cost ranks[H][3] + cong[H][4] + cost_base[H][3][4] + desir[S][4] + record
// The day kernel receives the replica descriptors BY VALUE (constant bank, uniform registers): up to
// kMaxRepsPerLaunch per launch; more replicas take several launches per weekday.
constexpr uint32_t kMaxRepsPerLaunch = 64;
struct RepArray { RepDev r[kMaxRepsPerLaunch]; };
static_assert(sizeof(RepArray) <= 24 * 1024, "kernel parameter space");
__host__ __device__ inline uint32_t tail_words(uint32_t S, uint32_t H) {
    return 4 + 4 * H + 12 * H + 3 * H + 4 * S + rec_width(S, H);
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
__device__ __forceinline__ float ld_stream_f32(const float* p) { return __uint_as_float(ld_stream_u32(reinterpret_cast<const uint32_t*>(p))); }
__device__ __forceinline__ uint2 ld_stream_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
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
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {     // TMA prefetch of a byte range into L2
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// System-scope flag traffic of the peer-to-peer exchange.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {              // from L2: written by other blocks / other GPUs
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Shared-memory atomics issued by ONE lane of a warp.  Plain atomicAdd makes the compiler wrap the instruction in its
// warp-aggregation sequence (vote, find-leader, popc, shuffle: ~12 instructions) because it cannot know that.
__device__ __forceinline__ uint32_t smem_fetch_add(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v));
    return old;
}
__device__ __forceinline__ void smem_add(uint32_t* p, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v));
}

// ---- gathering yesterday's modes ---------------------------------------------------------------
// Link code (layout.h): code >> 5 = byte offset of the target's word, low 5 bits = left shift that brings the
// target's 2 bits to bits 31:30.  __funnelshift_l wraps the shift amount, so the code is used as it is.
__device__ __forceinline__ uint32_t top2_hot(const uint32_t* s_vec, uint32_t code) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(s_vec) + (code >> 5));
    return __funnelshift_l(0u, w, code);                       // w << (code & 31)
}
// The word a cold code names, from the L2-resident global vector.  Divergent gathers that miss L1 are bound by the
// L1TEX wavefront rate (one 128-byte line per ~2 clocks, profiles/r02_microbench_cold.txt: 0.46 gathers per clock per SM
// through ld.global.nc, 0.52 through ld.global.cg, the same through the texture path or 16-byte TMA copies), so when the
// cold part of the vector is larger than what L1 keeps next to 227 KB of shared memory the loads bypass L1 (`cg`, a
// launch-uniform flag: -4.5 % on 8M agents); a cold part of a few KB (1M agents) stays on the L1-cached path.
__device__ __forceinline__ uint32_t ld_cold(const uint32_t* g_vec, uint32_t code, bool cg) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(g_vec) + (code >> 5));
    uint32_t w;
    if (cg) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w) : "l"(p));
    else w = __ldg(p);
    return w;
}
__device__ __forceinline__ uint32_t push2(uint32_t bits, uint32_t top) {      // (bits << 2) | (top >> 30)
    return __funnelshift_l(top, bits, 2u);
}
__device__ __forceinline__ uint32_t push_row(const uint32_t* s_vec, uint32_t bits, const uint4& c) {
    bits = push2(bits, top2_hot(s_vec, c.x));
    bits = push2(bits, top2_hot(s_vec, c.y));
    bits = push2(bits, top2_hot(s_vec, c.z));
    return push2(bits, top2_hot(s_vec, c.w));
}
__device__ __forceinline__ uint32_t top_fields(int n) {      // mask of the low bit of the n highest 2-bit fields
    return ~__funnelshift_rc(0xFFFFFFFFu, 0u, 2u * (uint32_t)n) & 0x55555555u;
}

// `bits` holds 16 gathered modes, link number base+e in field 15-e (the first link pushed ends up highest), of a
// list whose first `ds` links are social and the rest neighbour links.  Fields past the end of the list hold 0: the
// padding codes of a block point at a shared-memory word that is always zero (kPadWord, layout.h), and 0 is Car,
// which is not counted here but derived from the list length (car_counts).  Adds the counts of modes 1..3 to bits
// 8-31 of the packed 8-bit accumulators: acc_t over the whole group, acc_s over its social links
// (count_in_subgroup's grouping, agent.rs:325-329).
__device__ __forceinline__ void count_fields(uint32_t bits, int base, int ds, uint32_t& acc_s, uint32_t& acc_t) {
    const uint32_t smask = top_fields(min(max(ds - base, 0), 16));
    const uint32_t lo = bits & 0x55555555u, hi = (bits >> 1) & 0x55555555u;
    const uint32_t m3 = hi & lo, m2 = hi & ~lo, m1 = ~hi & lo;
    acc_s += __popc(m1 & smask) << 8;
    acc_s += __popc(m2 & smask) << 16;
    acc_s += __popc(m3 & smask) << 24;
    acc_t += __popc(m1) << 8;
    acc_t += __popc(m2) << 16;
    acc_t += __popc(m3) << 24;
}
// Fill in the Car count of a list of n links whose other three counts sit in bits 8-31 of acc.
__device__ __forceinline__ uint32_t car_counts(uint32_t acc, uint32_t n) {
    return acc | (n - __dp4a(acc, 0x01010100u, 0u));            // n minus the sum of bytes 1..3
}

// One lane's part of a block of link codes laid out k-major for the warp (layout.h): K hot codes per lane
// (floor(K/4) uint4 rows, then K%4 u32 rows) followed by Kc cold codes per lane.  The first ds_h links of the lane's
// hot list are social (its end needs no bound: padding gathers 0, see count_fields); the cold list has d_c links of
// which the first ds_c are social.  Per-mode counts are added to acc_s / acc_n (8 bits per mode: the caller keeps
// every list below 256 links).
__device__ __forceinline__ void accumulate_block(const uint32_t* __restrict__ blk, uint32_t K, uint32_t Kc, uint32_t lane,
                                                 int ds_h, int ds_c, int d_c, const uint32_t* s_vec,
                                                 const uint32_t* __restrict__ g_vec, bool cold_cg, uint32_t& acc_s, uint32_t& acc_n) {
    uint32_t hs = 0, ht = 0;                                    // hot list: social / all
    const uint32_t K4 = K >> 2, Kt = K & 3u;
    const uint4* body = reinterpret_cast<const uint4*>(blk) + lane;
    const uint32_t* tail = blk + (size_t)K4 * 128 + lane;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    uint4 r0 = zero, r1 = zero, r2 = zero, r3 = zero;
    uint32_t t0 = 0, t1 = 0, t2 = 0;
    if (K4 > 0) r0 = ld_stream_v4(body);
    if (K4 > 1) r1 = ld_stream_v4(body + 32);
    if (Kt > 0) t0 = ld_stream_u32(tail);                       // the short tail goes out with the first rows
    if (Kt > 1) t1 = ld_stream_u32(tail + 32);
    if (Kt > 2) t2 = ld_stream_u32(tail + 64);
    // cold list (targets outside shared memory, gathered from L2): its first two links go out with the first rows
    // and their gathers are issued before the hot loop, so both round trips overlap the shared-memory work
    const uint32_t* cold = blk + (size_t)K * 32 + lane;
    uint32_t c0 = 0, c1 = 0, w0 = 0, w1 = 0;
    if (Kc > 0) c0 = ld_stream_u32(cold);
    if (Kc > 1) c1 = ld_stream_u32(cold + 32);
    if (Kc > 0) w0 = ld_cold(g_vec, c0, cold_cg);
    if (Kc > 1) w1 = ld_cold(g_vec, c1, cold_cg);
    uint32_t q = 0, bits = 0;
    for (; q + 4 <= K4; q += 4) {                               // 16 links per lane, two rows kept in flight
        r2 = ld_stream_v4(body + (size_t)(q + 2) * 32);
        bits = push_row(s_vec, bits, r0);
        r3 = ld_stream_v4(body + (size_t)(q + 3) * 32);
        bits = push_row(s_vec, bits, r1);
        if (q + 4 < K4) r0 = ld_stream_v4(body + (size_t)(q + 4) * 32);
        bits = push_row(s_vec, bits, r2);
        if (q + 5 < K4) r1 = ld_stream_v4(body + (size_t)(q + 5) * 32);
        bits = push_row(s_vec, bits, r3);
        count_fields(bits, (int)(q * 4), ds_h, hs, ht);
        bits = 0;
    }
    const uint32_t rem = K4 - q;                                // 0..3 rows left, then the tail
    if (rem + Kt) {
        if (rem > 2) r2 = ld_stream_v4(body + (size_t)(q + 2) * 32);
        if (rem > 0) bits = push_row(s_vec, bits, r0);
        if (rem > 1) bits = push_row(s_vec, bits, r1);
        if (rem > 2) bits = push_row(s_vec, bits, r2);
        if (Kt > 0) bits = push2(bits, top2_hot(s_vec, t0));
        if (Kt > 1) bits = push2(bits, top2_hot(s_vec, t1));
        if (Kt > 2) bits = push2(bits, top2_hot(s_vec, t2));
        bits <<= 2u * (16u - (rem * 4u + Kt));                  // first link of the group back to field 15
        count_fields(bits, (int)(q * 4), ds_h, hs, ht);
    }
    acc_s += hs;
    acc_n += ht - hs;                                           // byte-wise: every count of ht is >= that of hs
    if (0 < d_c) {
        const uint32_t inc = (1u << ((__funnelshift_l(0u, w0, c0) >> 30) * 8u)) & ~0xFFu;      // Car is derived, see count_fields
        if (0 < ds_c) acc_s += inc; else acc_n += inc;
    }
    if (1 < d_c) {
        const uint32_t inc = (1u << ((__funnelshift_l(0u, w1, c1) >> 30) * 8u)) & ~0xFFu;
        if (1 < ds_c) acc_s += inc; else acc_n += inc;
    }
    for (uint32_t k = 2; k < Kc; k += 8) {                      // the rest of the cold list, eight links at a time:
        uint32_t c[8], w[8];                                    // all eight codes, then all eight gathers, then the counts
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = k + u < Kc ? ld_stream_u32(cold + (size_t)(k + u) * 32) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = ld_cold(g_vec, c[u], cold_cg);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if ((int)(k + u) < d_c) {
                const uint32_t inc = (1u << ((__funnelshift_l(0u, w[u], c[u]) >> 30) * 8u)) & ~0xFFu;
                if ((int)(k + u) < ds_c) acc_s += inc; else acc_n += inc;
            }
    }
}

// Per-lane statistics of the agents a lane has finished since the last flush: four 8-bit counters per word.
struct LaneStats {
    uint32_t hist;      // by mode, for neighbourhood `nbhd` (SELL slices are neighbourhood-uniform)
    uint32_t joint;     // by active*2 + active_norm     (statistics.rs:14-57)
    uint32_t commute;   // active by commute length      (statistics.rs:62-70)
    uint32_t sub;       // active by subculture, S <= 4  (statistics.rs:75-83)
    uint32_t nbhd, n;   // neighbourhood of `hist`, agents counted since the last flush
};
__device__ __forceinline__ void flush_hist(uint32_t lane, uint32_t* s_rec, LaneStats& st) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, (st.hist >> (8 * k)) & 0xFFu);
        if (lane == 0 && c) atomicAdd(&s_rec[4 * st.nbhd + k], c);
    }
    st.hist = 0;
}
__device__ __forceinline__ void flush_stats(uint32_t lane, uint32_t H, uint32_t S, uint32_t* s_rec, LaneStats& st) {
    flush_hist(lane, s_rec, st);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t j = __reduce_add_sync(0xFFFFFFFFu, (st.joint >> (8 * k)) & 0xFFu);
        const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, (st.commute >> (8 * k)) & 0xFFu);
        const uint32_t u = __reduce_add_sync(0xFFFFFFFFu, (st.sub >> (8 * k)) & 0xFFu);
        if (lane == 0) {
            if (j) atomicAdd(&s_rec[4 * H + k], j);
            if (c && k < 3) atomicAdd(&s_rec[4 * H + 4 + k], c);
            if (u && (uint32_t)k < S) atomicAdd(&s_rec[4 * H + 7 + k], u);
        }
    }
    st.joint = st.commute = st.sub = 0;
    st.n = 0;
}

// Everything that happens once the link counts of 32 agents (one per lane) are known.
template <bool PER_AGENT, bool XCHG>
__device__ __forceinline__ void finish_slice(const RepDev& r, const SoloExtra& solo, uint32_t slice, uint32_t lane, bool uniform_nbhd,
                                             const uint32_t cs[4], uint32_t ds, const uint32_t cn[4], uint32_t dn,
                                             uint32_t stat, float wsv, float4 hv, uint32_t last,
                                             const AgentWeights& uw, const float* s_cong, const float* s_cost,
                                             const uint32_t* s_crank, const float* s_des, uint32_t* s_rec,
                                             uint32_t weather, uint32_t changed, uint32_t parity, LaneStats& st) {
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
    const uint32_t nb = st_nbhd(stat), cm = st_commute(stat);
    decide<!PER_AGENT>(cs, ds, cn, dn, w, h, s_des + 4 * st_sub(stat), s_cong + 4 * nb, s_cost + 12 * nb + 4 * cm,
                       s_crank[3 * nb + cm], cm, st_car(stat), st_bike(stat), wsv, last, weather, changed, mode, norm);
    if (!valid) { mode = 0; norm = 0; }
    if (valid) reinterpret_cast<float4*>(r.state + (size_t)slice * kRecWords + kRecHabit)[lane] = make_float4(h[0], h[1], h[2], h[3]);
    // 32 agents x 2 bits -> one u64 (lanes 0-15 in .x, 16-31 in .y)
    const uint32_t sh = (lane & 15u) * 2u;
    const uint32_t mlo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? mode << sh : 0u);
    const uint32_t mhi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : mode << sh);
    const uint32_t nlo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? norm << sh : 0u);
    const uint32_t nhi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : norm << sh);
    if (XCHG) {
        // exchange fused into the update: lane p stores the word into rank p's vector (its own included), one store
        // instruction for all ranks; norm words only travel when asked (no rank's update reads another agent's norm)
        const PeerExchange* px = solo.px;
        const uint32_t world = px->world;
        if (lane < world) px->vec[parity ^ 1u][lane][slice] = make_uint2(mlo, mhi);
        if (lane >= 16u && lane - 16u < world && (solo.px_norms || lane - 16u == px->rank)) px->normvec[lane - 16u][slice] = make_uint2(nlo, nhi);
    } else {
        if (lane == 0) r.vec[parity ^ 1u][slice] = make_uint2(mlo, mhi);
        if (lane == 1) r.normvec[slice] = make_uint2(nlo, nhi);
    }
    // statistics (statistics.rs:14-93): every lane counts its own agent; the warp adds its lanes up at a flush
    const uint32_t a = valid && mode >= 2u, an = valid && norm >= 2u;
    if (uniform_nbhd) {
        const uint32_t nb0 = __shfl_sync(0xFFFFFFFFu, nb, 0);              // lane 0 of a non-hub slice is valid if any lane is
        if (nb0 != st.nbhd) { flush_hist(lane, s_rec, st); st.nbhd = nb0; }
        st.hist += valid ? 1u << (8u * mode) : 0u;
    } else if (valid) {
        atomicAdd(&s_rec[4 * nb + mode], 1u);
    }
    st.joint += valid ? 1u << (8u * (a * 2u + an)) : 0u;
    st.commute += a << (8u * cm);
    if (r.S <= 4) st.sub += a << (8u * st_sub(stat));
    else if (a) atomicAdd(&s_rec[4 * r.H + 7 + st_sub(stat)], 1u);
    if (++st.n == 255) flush_stats(lane, r.H, r.S, s_rec, st);
}

// Debug build only (make stamps: -DMVB_DEBUG_STAMPS, libmotivate_b200_stamps.so; mvbx_launch_stamps, tools/launch_gaps.py):
// {start of the first block, end of the last block} per launch, in ns.  Not in the product library: the two calls keep
// rec_next alive across the kernel and cost it 92 bytes of spills at its 64-register budget.
// The slot of a launch follows from the record row it writes, so the kernel needs no extra parameter;
// g_stamps.slots is null unless the hook has been called.
#ifdef MVB_DEBUG_STAMPS
struct StampCfg { unsigned long long* slots; const uint32_t* rec_base; uint32_t rec_stride, first_row, n; };
__device__ StampCfg g_stamps = {nullptr, nullptr, 0, 0, 0};
__device__ __forceinline__ void stamp_launch(const uint32_t* rec_next, int end) {
    const StampCfg c = g_stamps;
    const uint32_t row = (uint32_t)((rec_next - c.rec_base) / c.rec_stride);
    if (row < c.first_row || row - c.first_row >= c.n) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (end) atomicMax(&c.slots[2 * (size_t)(row - c.first_row) + 1], t); else atomicMin(&c.slots[2 * (size_t)(row - c.first_row)], t);
}
#define MVB_STAMP(end) do { if (tid == 0 && g_stamps.slots) stamp_launch(rec_next, end); } while (0)
#else
#define MVB_STAMP(end) do { } while (0)
#endif

// ---- the weekday kernel ----------------------------------------------------------------------------
// SOLO: the handle holds ONE population and uses what only such handles have: push mode, cold gathers that bypass L1, the
// peer-to-peer exchange (XCHG).  A compile-time switch: on the headline launches (many replicas of 1M agents) the run-time
// tests of these features cost 3 % (8 x 1M agents 252 -> 260 us per weekday), so those launches get the kernel without them.
template <bool PER_AGENT, int THREADS, uint32_t G, bool SOLO = false, bool XCHG = false>
__global__ void __launch_bounds__(THREADS, 1)
day_kernel(const __grid_constant__ RepArray reps, uint32_t R, uint32_t P, uint32_t parity,
           const uint32_t* __restrict__ rec_prev, uint32_t* __restrict__ rec_next, uint32_t weather,
           uint32_t changed, AgentWeights uw, uint32_t vec_cap_words, const SoloExtra solo) {
    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_vec = smem;
    uint32_t* s_tail = smem + vec_cap_words;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tail);
    uint32_t* s_next = s_tail + 2;                    // work-item counter of the current part
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // Programmatic dependent launch (the on-device face of the day loop, simulation.rs:101-136): tomorrow's launch may
    // start its blocks as soon as every block of today's has passed this line and an SM is free, i.e. while today's last
    // blocks still run.  Tomorrow's blocks do everything that does not depend on today (barrier set-up, static tables, their
    // first slice records, L2 prefetches of link codes) and then wait at griddepcontrol.wait (below) for today to be
    // complete and visible.  Without the launch attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    MVB_STAMP(0);                                   // debug build: when this launch's first block started
    if (tid == 0) { mbar_init(s_bar, 1); s_vec[vec_cap_words - 1] = 0; }   // the word padding codes point at (pad_code)
    uint32_t phase = 0;
    bool first_part = true;

    for (uint32_t part = blockIdx.x; part < R * P; part += gridDim.x) {
        const uint32_t p = part % P;
        const RepDev& r = reps.r[part / P];
        const uint32_t H = r.H, S = r.S;
        uint32_t* s_crank = s_tail + 4;                // [H][3] cost ranks on a Good-weather day
        float* s_cong = reinterpret_cast<float*>(s_crank + 3 * H);
        __syncthreads();                              // previous part is done with shared memory
        float* s_cost = s_cong + 4 * H;
        float* s_des = s_cost + 12 * H;
        uint32_t* s_rec = reinterpret_cast<uint32_t*>(s_des + 4 * S);
        const uint32_t* __restrict__ g_vec = reinterpret_cast<const uint32_t*>(r.vec[parity]);

        // LOCAL layout: part p = (a neighbourhood, class q of Q of its slices on this rank: part_table); the block stages the hub region and
        // that neighbourhood's own prefix.  GLOBAL: class p of all slices; every range is staged.
        uint32_t Q = P, q = p, sell_b = r.sell_begin, sell_e = r.sell_end, nb_part = 0;
        if (r.local_mode) {
            const uint4 pt = r.part_table[p];
            nb_part = pt.x; q = pt.y; Q = pt.z;
            sell_b = max(sell_b, r.nb_slice_begin[nb_part]); sell_e = max(sell_b, min(sell_e, r.nb_slice_begin[nb_part + 1]));
        }
        for (uint32_t k = tid; k < 12 * H; k += THREADS) {                // agent.rs:185-196
            const uint32_t hh = k / 12, c = (k >> 2) % 3, m = k & 3;
            s_cost[k] = cost_base_entry(r.supp[4 * hh + m], c, m);
        }
        for (uint32_t k = tid; k < 3 * H; k += THREADS) {                 // agent.rs:224-232 without a weather factor
            const uint32_t hh = k / 3, c = k % 3;
            float cc[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) cc[m] = cost_base_entry(r.supp[4 * hh + m], c, m);
            if (c == 2u) { cc[2] = __int_as_float(0x7f800000); cc[3] = cc[2]; }
            s_crank[k] = cost_ranks(cc);
        }
        // 1. stage yesterday's modes (TMA bulk copies) and build the congestion table from yesterday's histogram.  A block's
        // FIRST part does this further down, after griddepcontrol.wait: until then only launch-independent data is touched.
        auto stage_and_congestion = [&]() {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // the ranges every block stages, then (LOCAL layout) the piece of this part's own neighbourhood behind them
                const uint32_t n_stage = r.n_ranges + (r.local_mode ? 1u : 0u);
                const StageRange own = r.local_mode ? r.ranges[r.n_ranges + nb_part] : StageRange{0, 0, 0, 0};
                mbar_expect_tx(s_bar, (r.hot_words + own.n_words) * 4u);
                for (uint32_t k = 0; k < n_stage; ++k) {
                    const StageRange g = k < r.n_ranges ? r.ranges[k] : own;
                    for (uint32_t o = 0; o < g.n_words; o += 16384) {          // <= 64 KB per copy
                        const uint32_t n = min(g.n_words - o, 16384u);
                        bulk_g2s(s_vec + g.dst_word + o, g_vec + g.src_word + o, n * 4u, s_bar);
                    }
                }
            }
            for (uint32_t k = tid; k < 4 * H; k += THREADS) {                 // neighbourhood.rs:57-89
                const float c = congestion_entry(rec_prev[r.rec_off + k], r.cap[k], r.nres[k >> 2]);
                s_cong[k] = c;                                                // (read after the block barrier that ends the hub phase)
                if (p == 0) r.cong[k] = c;
            }
        };
        if (tid == 0) *s_next = 0;
        if (!first_part) stage_and_congestion();
        for (uint32_t k = tid; k < 4 * S; k += THREADS) s_des[k] = r.desir[k];
        for (uint32_t k = tid; k < rec_width(S, H); k += THREADS) s_rec[k] = 0;
        __syncthreads();

        // While the staging copies are in flight: every warp claims its first two work items (below), reads their slice
        // records and pulls the first one's link codes and agent state into L2; likewise the first hub unit of the warp.
        const uint32_t n_own_hub = r.hub_end - r.hub_begin;
        const uint32_t n_my_hub_slices = n_own_hub > p ? (n_own_hub - p + P - 1) / P : 0;
        const uint32_t n_sell = sell_e - sell_b;
        const uint32_t n_my_sell = n_sell > q ? (n_sell - q + Q - 1) / Q : 0;
        const uint32_t n_items = n_my_hub_slices + n_my_sell;
        // Items are claimed in groups of G consecutive ones (one shared-memory atomic and one shuffle per group: the
        // claim sits on the critical path of every slice it is made for).  G is a compile-time constant (a run-time
        // one costs more than it saves); the host picks the instantiation: 4 when a warp gets many items, else 2.
        uint32_t claim = 0;                            // first item of the group (a multiple of G) | items of it still to come
        auto next_item = [&]() -> uint32_t {
            if (claim & (G - 1u)) { const uint32_t it = (claim & ~(G - 1u)) + G - (claim & (G - 1u)); --claim; return it; }
            uint32_t q = 0;
            if (lane == 0) q = atomicAdd(s_next, G);
            q = __shfl_sync(0xFFFFFFFFu, q, 0);
            claim = q | (G - 1u);
            return q;
        };
        auto meta_of = [&](uint32_t item) -> uint4 {   // SELL items only; zero for hub items / past the end
            if (item < n_my_hub_slices || item >= n_items) return make_uint4(0, 0, 0, 0);
            return ld_stream_v4(r.sell_meta + (sell_b + q + (uint64_t)Q * (item - n_my_hub_slices)));
        };
        auto prefetch_item = [&](uint32_t item, const uint4& m) {
            if (item >= n_items || item < n_my_hub_slices) return;
            const uint32_t s1 = sell_b + q + Q * (item - n_my_hub_slices);
            const uint32_t nrow = (m.z & 0xFFFFu) + (m.z >> 16);
            if (lane == 0 && nrow) prefetch_l2(r.sell_codes + (((uint64_t)m.y << 32) | m.x), nrow * 128u);
            if (lane == 1) prefetch_l2(r.state + (size_t)(r.n_hub_slices + s1) * kRecWords, kRecWords * 4u);
        };
        uint32_t it0 = next_item(), it1 = next_item();
        uint4 meta = meta_of(it0), meta1 = meta_of(it1);
        prefetch_item(it0, meta);
        // this warp's hub units of the part: unit records are read two units ahead, a unit's link codes are pulled
        // into L2 one unit ahead (the first two records and the first prefetch go out while the staging copies fly)
        uint32_t hk = 0, hu = 0, hu_end = 0;
        auto seek_unit = [&] {                         // (hk, hu) -> this warp's next unit at or after it, or hk == n_my_hub_slices
            while (hk < n_my_hub_slices && hu >= hu_end) {
                if (++hk < n_my_hub_slices) {
                    const uint32_t hs = r.hub_begin + p + P * hk;
                    hu = r.hub_unit_begin[hs] + warp; hu_end = r.hub_unit_begin[hs + 1];
                }
            }
        };
        auto load_unit = [&]() -> uint4 {              // record of the unit at the cursor (zero = none left), cursor moves on
            uint4 rec = make_uint4(0, 0, 0, 0);
            if (hk < n_my_hub_slices) { rec = ld_stream_v4(r.hub_units + hu); hu += THREADS / 32; seek_unit(); }
            return rec;
        };
        auto prefetch_unit = [&](const uint4& rec) {
            const uint32_t rows = (rec.z & 0xFFu) + ((rec.z >> 8) & 0xFFu);
            if (lane == 0 && rows) prefetch_l2(r.hub_codes + (uint64_t)rec.x * 32, rows * 128u);
        };
        if (n_my_hub_slices) {
            const uint32_t hs = r.hub_begin + p;
            hu = r.hub_unit_begin[hs] + warp; hu_end = r.hub_unit_begin[hs + 1];
            seek_unit();
        }
        uint4 unit0 = load_unit(), unit1 = load_unit();
        prefetch_unit(unit0);
        if (first_part) {                             // everything from here on reads what yesterday's launch wrote
            asm volatile("griddepcontrol.wait;" ::: "memory");
            stage_and_congestion();
            first_part = false;
        }
        mbar_wait(s_bar, phase);
        phase ^= 1u;


        // 2. hubs of this part's hub slices (hub_begin + p, + P, ...).  A hub's lists are cut into units of at most
        // 1 024 links (layout.h); a warp takes a unit, its 32 lanes take 32 contiguous chunks of it, and the unit's
        // counts of modes 1..3 are added to the hub's entry of a small global scratch.  The hub slices are finished
        // below like any other slice, so no warp waits for another here and the biggest hubs spread over many warps.
        while (unit0.z) {
            const uint4 unit2 = load_unit();
            prefetch_unit(unit1);
            const uint32_t K = unit0.z & 0xFFu, Kc = (unit0.z >> 8) & 0xFFu, hub = unit0.y;
            const uint32_t* blk = r.hub_codes + (uint64_t)unit0.x * 32;
            // social hot / social cold / cold links left at the start of the unit -> this lane's chunk of them
            const int oh = (int)(lane * K), oc = (int)(lane * Kc);
            uint32_t acc_s = 0, acc_n = 0;
            accumulate_block(blk, K, Kc, lane, min(max((int)(unit0.z >> 16) - oh, 0), (int)K),
                             min(max((int)(unit0.w & 0xFFFFu) - oc, 0), (int)Kc), min(max((int)(unit0.w >> 16) - oc, 0), (int)Kc),
                             s_vec, g_vec, SOLO && r.cold_cg != 0, acc_s, acc_n);
            uint32_t mine = 0;                         // lane 1..3: social count of mode lane; lane 5..7: neighbour count
#pragma unroll
            for (int m = 1; m < 4; ++m) {
                const uint32_t cs_m = __reduce_add_sync(0xFFFFFFFFu, (acc_s >> (8 * m)) & 0xFFu);
                const uint32_t cn_m = __reduce_add_sync(0xFFFFFFFFu, (acc_n >> (8 * m)) & 0xFFu);
                if (lane == (uint32_t)m) mine = cs_m;
                if (lane == (uint32_t)m + 4u) mine = cn_m;
            }
            if (mine) atomicAdd(&r.hub_cnt[(size_t)hub * 8 + lane], mine);
            unit0 = unit1; unit1 = unit2;
        }
        __syncthreads();                              // the hub counts written above are read by other warps below

        // 3. work items handed out by a block-wide counter: first this part's hub slices (finish only), then its
        // SELL slices p + P*k: one lane per agent.  The slice record is read two items ahead, and the next slice is
        // pulled into L2 (its link codes and its agents' state) by TMA prefetches while this one is processed.
        LaneStats st{0, 0, 0, 0, 0, 0};
        while (it0 < n_items) {
            const uint32_t it2 = next_item();
            const uint4 meta2 = meta_of(it2);
            prefetch_item(it1, meta1);
            if (it0 < n_my_hub_slices) {              // finish one hub slice: counts come from the scratch
                const uint32_t hs = r.hub_begin + p + P * it0, agent = hs * 32 + lane;
                uint4* cnt = reinterpret_cast<uint4*>(r.hub_cnt + (size_t)agent * 8);
                const uint4 ca = cnt[0], cb = cnt[1];
                cnt[0] = make_uint4(0, 0, 0, 0); cnt[1] = make_uint4(0, 0, 0, 0);      // ready for tomorrow
                const uint4 ad = r.hub_deg[agent];
                // Car = what is left of each list (count_fields)
                const uint32_t hcs[4] = {ad.x + ad.z - ca.y - ca.z - ca.w, ca.y, ca.z, ca.w};
                const uint32_t hcn[4] = {ad.y + ad.w - cb.y - cb.z - cb.w, cb.y, cb.z, cb.w};
                const uint32_t* rec = r.state + (size_t)hs * kRecWords;
                const uint32_t stat = rec[kRecStat + lane];
                const uint2 own = r.vec[parity][hs];
                const uint32_t last = ((lane < 16 ? own.x : own.y) >> ((lane & 15u) * 2u)) & 3u;
                finish_slice<PER_AGENT, XCHG>(r, solo, hs, lane, false, hcs, ad.x + ad.z, hcn, ad.y + ad.w, stat, __uint_as_float(rec[kRecWs + lane]),
                                        reinterpret_cast<const float4*>(rec + kRecHabit)[lane], last, uw, s_cong, s_cost, s_crank, s_des, s_rec, weather, changed, parity, st);
            } else {
            const uint32_t s = sell_b + q + Q * (it0 - n_my_hub_slices);
            const uint32_t slice = r.n_hub_slices + s;
            const uint32_t KK = meta.z;
            const uint32_t* blk = r.sell_codes + (((uint64_t)meta.y << 32) | meta.x);
            // the agent's own state goes out with the first rows of link codes
            // (all streaming loads bypass L1, which is left to the cold part of the mode vector)
            const uint32_t* rec = r.state + (size_t)slice * kRecWords + lane;
            const uint32_t stat = ld_stream_u32(rec + kRecStat);
            const uint32_t dg = ld_stream_u32(rec + kRecLens);
            const float wsv = __uint_as_float(ld_stream_u32(rec + kRecWs));
            const uint4 hraw = ld_stream_v4(reinterpret_cast<const uint4*>(rec + kRecHabit - lane) + lane);
            const float4 hv = make_float4(__uint_as_float(hraw.x), __uint_as_float(hraw.y), __uint_as_float(hraw.z), __uint_as_float(hraw.w));
            const uint2 own = ld_stream_v2(r.vec[parity] + slice);
            const int ds_h = (int)(dg & 0xFFu), dn_h = (int)((dg >> 8) & 0xFFu), ds_c = (int)((dg >> 16) & 0xFFu), dn_c = (int)(dg >> 24);
            uint32_t acc_s = 0, acc_n = 0;
            int d_c = ds_c + dn_c;
            if (SOLO && solo.push_cnt) {                       // the cold links were counted by push_kernel; the block has hot rows only
                const uint2 pc = ld_stream_v2(solo.push_cnt + (size_t)slice * 32 + lane);
                acc_s = pc.x; acc_n = pc.y; d_c = 0;
            }
            accumulate_block(blk, KK & 0xFFFFu, KK >> 16, lane, ds_h, ds_c, d_c, s_vec, g_vec, SOLO && r.cold_cg != 0, acc_s, acc_n);
            acc_s = car_counts(acc_s, (uint32_t)(ds_h + ds_c));
            acc_n = car_counts(acc_n, (uint32_t)(dn_h + dn_c));
            uint32_t cs[4], cn[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) { cs[m] = (acc_s >> (8 * m)) & 0xFFu; cn[m] = (acc_n >> (8 * m)) & 0xFFu; }
            const uint32_t last = ((lane < 16 ? own.x : own.y) >> ((lane & 15u) * 2u)) & 3u;
            finish_slice<PER_AGENT, XCHG>(r, solo, slice, lane, true, cs, (uint32_t)(ds_h + ds_c), cn, (uint32_t)(dn_h + dn_c), stat, wsv, hv,
                                    last, uw, s_cong, s_cost, s_crank, s_des, s_rec, weather, changed, parity, st);
            }
            it0 = it1; it1 = it2; meta = meta1; meta1 = meta2;
        }
        flush_stats(lane, H, S, s_rec, st);
        __syncthreads();
        for (uint32_t k = tid; k < rec_width(S, H); k += THREADS)
            if (s_rec[k]) atomicAdd(&rec_next[r.rec_off + k], s_rec[k]);
    }
    if (XCHG) {
        // Peer-to-peer exchange, second half.  The mode words are already on their way (finish_slice); when the LAST block
        // of this rank has finished, it hands the rank's partial record to every rank and raises the rank's flag there.
        const PeerExchange* px = solo.px;
        const uint32_t epoch = solo.px_epoch, world = px->world, W = px->rec_stride;
        uint32_t* s_last = s_tail + 3;                 // (a free word next to the work counter: no static shared memory here)
        __threadfence_system();                        // this thread's stores to the other ranks have been performed
        __syncthreads();
        if (tid == 0) *s_last = atomicAdd(px->done, 1u) == gridDim.x - 1u;
        __syncthreads();
        if (*s_last) {
            __threadfence();                           // the other blocks' additions to rec_next
            for (uint32_t k = tid; k < world * W; k += THREADS) {
                const uint32_t p = k / W, j = k - p * W;
                px->xrec[p][((epoch & 1u) * world + px->rank) * W + j] = ld_cg_u32(rec_next + j);
            }
            __threadfence_system();
            __syncthreads();
            if (tid < world) st_release_sys(px->flags[tid] + px->rank, epoch);
            if (tid == 0) *px->done = 0;               // the next launch's blocks get here only after this one is complete
        }
    }
    MVB_STAMP(1);                                   // ... and when its last block ended
}

// Peer-to-peer exchange, closing step of a weekday (one small block, stream-ordered behind the day kernel): wait until
// every rank has raised its flag for this epoch (all mode words and partial records have then landed in this rank's
// memory), then add the partial records up into the day's record row, which tomorrow's congestion and the statistics read.
// A rank that never arrives (a dead process) must not hang the GPU: after ~20 s the wait gives up and sets px->error.
__global__ void exchange_finish_kernel(const PeerExchange* px, uint32_t epoch, uint32_t* rec_row) {
    const uint32_t world = px->world, W = px->rec_stride, tid = threadIdx.x;
    if (tid < world) {
        const uint32_t* f = px->flags[px->rank] + tid;
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 20000000000ull) { *px->error = 1u; break; }
        }
    }
    __syncthreads();
    const uint32_t* part = px->xrec[px->rank] + (size_t)(epoch & 1u) * world * W;
    for (uint32_t j = tid; j < W; j += blockDim.x) {
        uint32_t sum = 0;
        for (uint32_t p = 0; p < world; ++p) sum += ld_cg_u32(part + (size_t)p * W + j);
        rec_row[j] = sum;
    }
}

// ---- push mode: the cold links of the SELL agents, turned round (layout.h) ---------------------------------------
// One launch per weekday BEFORE the day kernel.  A block takes chunks b, b + grid, ...; a chunk = up to kPushChunkMax
// consecutive SELL agents this handle updates, their packed counters {social, neighbour} (bytes 1..3 = modes 1..3, as in
// accumulate_block) in shared memory.  The chunk's cold links come as two coalesced streams sorted by target (code u32,
// source u16): a warp's 32 gathers fall into a few 128-byte lines of the L2-resident vector, and each gathered mode is
// added to its source's counter by a shared-memory reduction (Car, mode 0, is derived from the list length and skipped).
struct PushDesc {
    const uint32_t* code; const uint16_t* src;      // [n cold links], chunk by chunk, sorted by target inside a chunk
    const uint32_t* chunk_begin;                    // [n_chunks + 1] first list entry of every chunk
    const uint32_t* chunk_first;                    // [n_chunks + 1] first owned SELL slot of every chunk (chunks hold about the same number of links)
    uint2* cnt;                                     // [n_own_slots] out: counts of every owned SELL agent
    uint32_t n_chunks, cold_cg;
};
__global__ void __launch_bounds__(1024, 1) push_kernel(const PushDesc d, const uint32_t* __restrict__ g_vec) {
    extern __shared__ __align__(16) uint32_t s_cnt[];                    // [chunk_agents][2]
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t U = 8;                               // (8M agents 0.489 -> 0.479 ms per weekday against 4; 64 registers)
    for (uint32_t c = blockIdx.x; c < d.n_chunks; c += gridDim.x) {
        const uint32_t first = d.chunk_first[c], n_ag = d.chunk_first[c + 1] - first;
        for (uint32_t k = tid; k < 2 * n_ag; k += 1024) s_cnt[k] = 0;
        __syncthreads();
        const uint32_t b = d.chunk_begin[c], e = d.chunk_begin[c + 1];
        // software pipeline: the next U entries of both streams are on their way from HBM while the current ones are
        // gathered and reduced (a thread has one dependent chain code -> gather -> reduction per entry otherwise)
        uint32_t code_n[U], src_n[U];
        auto fetch = [&](uint32_t i0) {
#pragma unroll
            for (uint32_t u = 0; u < U; ++u) {
                const uint32_t i = i0 + 1024 * u;
                const bool ok = i < e;
                code_n[u] = ok ? ld_stream_u32(d.code + i) : 0u;
                src_n[u] = ok ? (uint32_t)__ldg(d.src + i) : 0xFFFFFFFFu;
            }
        };
        if (b + tid < e) fetch(b + tid);
        for (uint32_t i0 = b + tid; i0 < e; i0 += 1024 * U) {
            uint32_t code[U], src[U], w[U];
#pragma unroll
            for (uint32_t u = 0; u < U; ++u) { code[u] = code_n[u]; src[u] = src_n[u]; }
            if (i0 + 1024 * U < e) fetch(i0 + 1024 * U);
#pragma unroll
            for (uint32_t u = 0; u < U; ++u) w[u] = src[u] != 0xFFFFFFFFu ? ld_cold(g_vec, code[u], d.cold_cg != 0) : 0u;
#pragma unroll
            for (uint32_t u = 0; u < U; ++u) {
                const uint32_t m = __funnelshift_l(0u, w[u], code[u]) >> 30;
                if (m) smem_add(&s_cnt[2u * (src[u] & 0x7FFFu) + ((src[u] >> 15) & 1u)], 1u << (8u * m));
            }
        }
        __syncthreads();
        for (uint32_t k = tid; k < n_ag; k += 1024) d.cnt[first + k] = make_uint2(s_cnt[2 * k], s_cnt[2 * k + 1]);
        __syncthreads();
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
    const uint32_t stat = i < r.N_pad ? r.state[rec_word(i, kRecStat)] : 0u;
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

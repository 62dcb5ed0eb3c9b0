// fill_kernels.cuh — create-time kernels that write the link-code arrays of the day kernel's layout (layout.h)
// in HBM, straight from the caller's two CSR graphs.  Setup only: they run once per distinct graph in
// mvb_create_batch, so the 120 MB of link indices cross PCIe once, as they are, and the k-major re-encoding
// (60 M random lookups per 1M-agent population) happens at HBM speed instead of on a host core.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "decide.cuh"
#include "layout.h"

namespace mvb {

struct FillArgs {
    const uint64_t* s_rowptr; const uint32_t* s_col;     // caller's social CSR (agent.social_network)
    const uint64_t* n_rowptr; const uint32_t* n_col;     // caller's neighbour CSR (agent.neighbours)
    const uint32_t* tcode;                               // [N] code of a target | 1 if hot
    const uint32_t* int2ext;                             // [N_pad] internal slot -> caller's id or kInvalid
    uint32_t n_hub_slices, n_sell_slices;
    const uint4* sell_meta; uint32_t* sell_codes; uint32_t* deg;
    const uint64_t* hub_off; const uint4* hub_deg; uint32_t* hub_codes;
    uint32_t own_hub_begin, own_hub_end, own_sell_begin, own_sell_end;   // slices whose codes this handle stores (all, unless agent-partitioned)
    const uint32_t* tcold;                               // LOCAL layout: [N] cold code of a target (nullptr: GLOBAL layout)
    const uint8_t* cls;                                  // [N] 0 never staged, 1 staged for every block, 2 for blocks of its own neighbourhood
    const uint16_t* nbhd;                                // [N] caller's neighbourhood array
    // push mode (layout.h): the cold links of owned SELL agents go to a list instead of the block's cold rows
    uint32_t* push_code; uint32_t* push_val;             // [n cold links] target code; chunk << 16 | neighbour-list bit << 15 | agent index in the chunk
    const uint32_t* push_off;                            // [owned SELL slots] first list entry of the agent
    const uint32_t* push_first; uint32_t push_chunks;    // [push_chunks + 1] first owned SELL slot of every chunk
};
// Code of the link src -> t (| 1 if hot).  LOCAL layout: a target staged with its neighbourhood's prefix is hot only for
// sources of that neighbourhood that are not hubs (hub slices mix neighbourhoods); everybody else reads it from L2.
__device__ __forceinline__ uint32_t link_code(const FillArgs& a, uint32_t src_nbhd, bool src_is_hub, uint32_t t) {
    uint32_t c = a.tcode[t];
    if (a.tcold && (c & 1u) && a.cls[t] == 2 && (src_is_hub || a.nbhd[t] != src_nbhd)) c = a.tcold[t];
    return c;
}

// chunk of every owned SELL slot's agent: c << 16 | index in the chunk
__device__ __forceinline__ uint32_t push_chunk_of(const uint32_t* __restrict__ first, uint32_t n_chunks, uint32_t own_k) {
    uint32_t lo = 0, hi = n_chunks;                                              // the last chunk starting at or before own_k
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (first[mid] <= own_k) lo = mid; else hi = mid; }
    return lo << 16 | (own_k - first[lo]);
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint64_t n, uint32_t v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// One warp per agent: how many of its social / neighbour links point at agents whose mode will sit in shared memory
// (layout.h, layout_count_hot).  This is the one step of the layout that touches every link, so it runs here, on
// the CSR the fill kernels need anyway.  Also the range check of the link targets: any target >= N sets *bad.
__global__ void count_hot_kernel(uint32_t N, const uint64_t* __restrict__ s_rowptr, const uint32_t* __restrict__ s_col,
                                 const uint64_t* __restrict__ n_rowptr, const uint32_t* __restrict__ n_col,
                                 const uint8_t* __restrict__ is_hot, const uint16_t* __restrict__ nbhd, uint32_t* __restrict__ nsh,
                                 uint32_t* __restrict__ nnh, uint32_t* __restrict__ bad) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps) {
        uint32_t a = 0, b = 0, oob = 0;
        // class 1 = staged for every block; class 2 (LOCAL layout) = staged for the blocks of the target's own
        // neighbourhood: hot for sources that live there and are not hubs (layout.h)
        const uint32_t my_nb = nbhd[i];
        const bool local_ok = (s_rowptr[i + 1] - s_rowptr[i]) + (n_rowptr[i + 1] - n_rowptr[i]) <= kHubDegree;
        auto hot = [&](uint32_t t) -> uint32_t { const uint32_t c = is_hot[t]; return c == 1u || (c == 2u && local_ok && nbhd[t] == my_nb); };
        for (uint64_t x = s_rowptr[i] + lane; x < s_rowptr[i + 1]; x += 32) {
            const uint32_t t = s_col[x];
            if (t < N) a += hot(t); else oob = 1;
        }
        for (uint64_t x = n_rowptr[i] + lane; x < n_rowptr[i + 1]; x += 32) {
            const uint32_t t = n_col[x];
            if (t < N) b += hot(t); else oob = 1;
        }
        a = __reduce_add_sync(0xFFFFFFFFu, a);
        b = __reduce_add_sync(0xFFFFFFFFu, b);
        oob = __reduce_or_sync(0xFFFFFFFFu, oob);
        if (lane == 0) { nsh[i] = a; nnh[i] = b; if (oob) atomicOr(bad, 1u); }
    }
}

// Link code of every possible target (layout.h: word byte offset << 5 | shift, | 1 if hot): a hot agent is addressed
// by its place in the staged vector (ranges), a cold one by its internal slot.  Same rule as layout_finish.
__global__ void tcode_kernel(uint32_t N_pad, const uint32_t* __restrict__ int2ext, const StageRange* __restrict__ ranges,
                             uint32_t n_ranges, const uint8_t* __restrict__ is_hot, const uint16_t* __restrict__ nbhd_local,
                             uint32_t H, uint32_t hub_pad, uint32_t* __restrict__ tcode, uint32_t* __restrict__ tcold) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N_pad) return;
    const uint32_t e = int2ext[t];
    if (e == kInvalid) return;
    uint32_t pos = t, hot = is_hot[e];
    if (nbhd_local) {      // LOCAL layout: ranges[0] = hub region, [1 + h] = common prefix of group h, [1 + H + h] = own piece of neighbourhood h
        tcold[e] = (t >> 4) << 7 | (30u - 2u * (t & 15u));
        if (hot && t >= hub_pad) { const StageRange r = ranges[(hot == 2u ? 1u + H : 1u) + nbhd_local[e]]; pos = r.dst_word * 16u + (t - r.src_word * 16u); }
        tcode[e] = ((pos >> 4) << 7 | (30u - 2u * (pos & 15u))) | (hot ? 1u : 0u);
        return;
    }
    if (hot) {
        uint32_t lo = 0, hi = n_ranges;                       // ranges are in increasing src_word order: last one starting at or before t
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (ranges[mid].src_word * 16u <= t) lo = mid; else hi = mid; }
        pos = ranges[lo].dst_word * 16u + (t - ranges[lo].src_word * 16u);
    }
    tcode[e] = ((pos >> 4) << 7 | (30u - 2u * (pos & 15u))) | hot;
}

__device__ __forceinline__ void put_hot(uint32_t* blk, uint32_t K, uint32_t lane, uint32_t k, uint32_t code) {
    const uint32_t K4 = K >> 2;
    if (k < 4 * K4) blk[(size_t)(k >> 2) * 128 + lane * 4 + (k & 3u)] = code;
    else blk[(size_t)K4 * 128 + (size_t)(k - 4 * K4) * 32 + lane] = code;
}
__device__ __forceinline__ void put_cold(uint32_t* blk, uint32_t K, uint32_t lane, uint32_t k, uint32_t code) {
    blk[(size_t)K * 32 + (size_t)k * 32 + lane] = code;
}

// One thread per agent of a SELL slice (lane = position in the slice): hot links first by list (social, then
// neighbour), cold links likewise; also writes the agent's four list lengths.
__global__ void fill_sell_kernel(FillArgs a) {
    const uint32_t slot_in_sell = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = slot_in_sell >> 5, lane = slot_in_sell & 31u;
    if (s >= a.n_sell_slices || s < a.own_sell_begin || s >= a.own_sell_end) return;
    const uint32_t slot = (a.n_hub_slices + s) * 32 + lane;
    const uint32_t e = a.int2ext[slot];
    if (e == kInvalid) { a.deg[slot] = 0; return; }
    const uint4 meta = a.sell_meta[s];
    uint32_t* blk = a.sell_codes + (((uint64_t)meta.y << 32) | meta.x);
    const uint32_t K = meta.z & 0xFFFFu;
    uint32_t kh = 0, kc = 0;
    const uint32_t my_nb = a.nbhd[e];
    const uint32_t own_k = slot_in_sell - a.own_sell_begin * 32;                  // push mode: place among the owned SELL slots
    const uint32_t pv = a.push_code ? push_chunk_of(a.push_first, a.push_chunks, own_k) : 0u;
    const uint32_t po = a.push_code ? a.push_off[own_k] : 0u;
    auto cold = [&](uint32_t c, uint32_t list) {
        if (a.push_code) { a.push_code[po + kc] = c; a.push_val[po + kc] = pv | list << 15; }
        else put_cold(blk, K, lane, kc, c);
        ++kc;
    };
    for (uint64_t x = a.s_rowptr[e]; x < a.s_rowptr[e + 1]; ++x) {
        const uint32_t c = link_code(a, my_nb, false, a.s_col[x]);
        if (c & 1u) put_hot(blk, K, lane, kh++, c & ~1u); else cold(c, 0u);
    }
    const uint32_t c_sh = kh, c_sc = kc;
    for (uint64_t x = a.n_rowptr[e]; x < a.n_rowptr[e + 1]; ++x) {
        const uint32_t c = link_code(a, my_nb, false, a.n_col[x]);
        if (c & 1u) put_hot(blk, K, lane, kh++, c & ~1u); else cold(c, 1u);
    }
    a.deg[slot] = c_sh | (kh - c_sh) << 8 | c_sc << 16 | (kc - c_sc) << 24;
}

// One warp per hub: the hub's hot (cold) list is the sequence of its hot (cold) links in list order; link number i
// of the list goes to block `seg`, lane (i - start_of_seg) / K, position (i - start_of_seg) % K (hub_segment).
// Lane l of the warp handles list positions by ballot-compaction, so the order of the lists is the link order.
__global__ void fill_hub_kernel(FillArgs a, uint32_t n_hubs) {
    const uint32_t hub = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (hub >= n_hubs || (hub >> 5) < a.own_hub_begin || (hub >> 5) >= a.own_hub_end) return;
    const uint32_t e = a.int2ext[hub];
    if (e == kInvalid) return;
    const uint4 hd = a.hub_deg[hub];
    const uint32_t n_hot = hd.x + hd.y, n_cold = hd.z + hd.w;
    uint32_t* const base = a.hub_codes + a.hub_off[hub];
    // walker over the hub's blocks for one list (hot or cold): position i -> (block pointer, K, Kc, offset in block)
    struct Walker {
        uint32_t* blk; uint32_t dh, dc, K, Kc, n_hot, n_cold;
        __device__ void init(uint32_t* b, uint32_t nh, uint32_t nc) { blk = b; dh = dc = 0; n_hot = nh; n_cold = nc; hub_segment(nh, nc, K, Kc); }
        __device__ void advance() {
            blk += (size_t)32 * (K + Kc); dh += 32 * K; dc += 32 * Kc;
            hub_segment(n_hot - min(dh, n_hot), n_cold - min(dc, n_cold), K, Kc);
        }
    };
    for (int pass = 0; pass < 2; ++pass) {                 // 0: hot list, 1: cold list
        Walker w; w.init(base, n_hot, n_cold);
        uint32_t pos = 0;                                   // links of this list placed so far (warp-uniform)
        for (int g = 0; g < 2; ++g) {                       // social links, then neighbour links
            const uint64_t* rp = g ? a.n_rowptr : a.s_rowptr;
            const uint32_t* cl = g ? a.n_col : a.s_col;
            const uint64_t b = rp[e], end = rp[e + 1];
            for (uint64_t x0 = b; x0 < end; x0 += 32) {
                const uint64_t x = x0 + lane;
                uint32_t c = 0; bool mine = false;
                if (x < end) { c = link_code(a, 0u, true, cl[x]); mine = ((c & 1u) != 0u) == (pass == 0); }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
                if (mine) {
                    uint32_t i = pos + __popc(m & ((1u << lane) - 1u));
                    Walker v = w;                            // private copy: this link may be a few blocks ahead
                    while (i >= (pass == 0 ? v.dh + 32 * v.K : v.dc + 32 * v.Kc)) v.advance();
                    const uint32_t local = i - (pass == 0 ? v.dh : v.dc), per = pass == 0 ? v.K : v.Kc;
                    if (pass == 0) put_hot(v.blk, v.K, local / per, local % per, c & ~1u);
                    else put_cold(v.blk, v.K, local / per, local % per, c);
                }
                pos += __popc(m);
                while ((pass == 0 ? w.K : w.Kc) && pos >= (pass == 0 ? w.dh + 32 * w.K : w.dc + 32 * w.Kc) &&
                       (w.dh + 32 * w.K < n_hot || w.dc + 32 * w.Kc < n_cold)) w.advance();
            }
        }
    }
}

// Per-agent state of one replica from the caller's arrays (caller order) into internal slot order: packed static
// word, weather sensitivity, habit, optional per-agent weights, and the 2-bit packed current_mode / norm vectors.
struct InitArgs {
    const uint32_t* int2ext; uint32_t N, N_pad;
    const uint8_t* sub; const uint16_t* nbhd; const uint8_t* commute; const uint8_t* car; const uint8_t* bike;
    const float* ws; const float4* habit; const uint8_t* mode; const uint8_t* norm;
    const float* pa_src[5]; float pa_def[5];              // per-agent weights (nullptr = uniform default), agent.rs:44-57
    const uint32_t* lens;                                 // [N_pad] list lengths of the SELL agents (written by fill_sell_kernel)
    uint32_t* state; float* pa_out;                       // slice records (layout.h)
    uint2* vec0; uint2* vec1; uint2* normvec;
};
__global__ void init_state_kernel(InitArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    if (i >= a.N_pad) return;                             // N_pad is a multiple of 64: whole warps
    const uint32_t e = a.int2ext[i];
    const bool valid = e != kInvalid;
    uint32_t mode = 0, norm = 0;
    float4* habit_out = reinterpret_cast<float4*>(a.state + rec_habit_word(i));
    a.state[rec_word(i, kRecLens)] = a.lens[i];
    if (valid) {
        a.state[rec_word(i, kRecStat)] = 0x80000000u | pack_static(a.nbhd[e], a.sub[e], a.commute[e], a.car[e] != 0, a.bike[e] != 0);
        a.state[rec_word(i, kRecWs)] = __float_as_uint(a.ws[e]);
        *habit_out = a.habit[e];
        mode = a.mode[e] & 3u; norm = a.norm[e] & 3u;
    } else {
        a.state[rec_word(i, kRecStat)] = 0; a.state[rec_word(i, kRecWs)] = 0; *habit_out = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (a.pa_out)
        for (int k = 0; k < 5; ++k)
            a.pa_out[(size_t)k * a.N_pad + i] = (valid && a.pa_src[k]) ? a.pa_src[k][e] : a.pa_def[k];
    const uint32_t sh = (lane & 15u) * 2u;
    const uint32_t mlo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? mode << sh : 0u);
    const uint32_t mhi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : mode << sh);
    const uint32_t nlo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? norm << sh : 0u);
    const uint32_t nhi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : norm << sh);
    if (lane == 0) { a.vec0[i >> 5] = make_uint2(mlo, mhi); a.vec1[i >> 5] = make_uint2(mlo, mhi); a.normvec[i >> 5] = make_uint2(nlo, nhi); }
}

// Ownership part of `intervene` (simulation.rs:383-463): rewrite the car / bike bits of the static words.
__global__ void set_ownership_kernel(const uint32_t* int2ext, uint32_t N_pad, const uint8_t* car, const uint8_t* bike, uint32_t* state) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N_pad) return;
    const uint32_t e = int2ext[i];
    if (e == kInvalid) return;
    uint32_t w = state[rec_word(i, kRecStat)];
    if (car)  w = (w & ~(1u << 26)) | ((car[e] ? 1u : 0u) << 26);
    if (bike) w = (w & ~(1u << 27)) | ((bike[e] ? 1u : 0u) << 27);
    state[rec_word(i, kRecStat)] = w;
}

// Agent-partitioned exchange (sim.cu launch_day): a rank's piece of the all-gather buffer is [M mode words | M norm
// words] of its own slice range [b, e); after the all-gather the other ranks' pieces go back into the vectors.
__global__ void pack_exchange_kernel(const uint2* __restrict__ vec, const uint2* __restrict__ normvec, uint32_t b, uint32_t e, uint32_t M,
                                     uint2* __restrict__ send) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (b + k < e) { send[k] = vec[b + k]; if (normvec) send[M + k] = normvec[b + k]; }
}
__global__ void unpack_exchange_kernel(const uint2* __restrict__ recv, const uint32_t* __restrict__ bounds, uint32_t world, uint32_t rank,
                                       uint32_t M, uint32_t piece, uint32_t n_slices, uint2* __restrict__ vec, uint2* __restrict__ normvec) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    uint32_t r = 0;
    while (r + 1 < world && bounds[r + 1] <= s) ++r;
    if (r == rank) return;
    const uint2* p = recv + (size_t)r * piece;             // rank r's piece: [M mode words | M norm words (if sent)]
    vec[s] = p[s - bounds[r]];
    if (normvec) normvec[s] = p[M + (s - bounds[r])];
}

// `intervene` (simulation.rs:304-464) as ONE launch for every replica whose intervention falls on the day: the host
// has resolved it to new tables and ownership bits (caller's agent order), staged in device scratch before the day
// loop starts; blockIdx.y picks the entry.  Tables: supportiveness[H][4] | desirability[S][4] and capacity[H][4].
struct IvDev {
    uint32_t replica, has_tables;
    const uint8_t* car; const uint8_t* bike;              // [N] or nullptr (= unchanged)
    const float* tf; const uint32_t* tu;                  // staged tables (has_tables)
};
struct IvRep {                                            // what the kernel needs of a replica
    const uint32_t* int2ext; uint32_t* state; float* tables_f; uint32_t* tables_u; uint32_t N_pad, S, H;
};
__global__ void apply_interventions_kernel(const IvDev* __restrict__ ivs, const IvRep* __restrict__ reps) {
    const IvDev iv = ivs[blockIdx.y];
    const IvRep r = reps[iv.replica];
    if (iv.has_tables && blockIdx.x == 0) {
        for (uint32_t k = threadIdx.x; k < 4 * (r.H + r.S); k += blockDim.x) r.tables_f[k] = iv.tf[k];
        for (uint32_t k = threadIdx.x; k < 4 * r.H; k += blockDim.x) r.tables_u[k] = iv.tu[k];
    }
    if (!iv.car && !iv.bike) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < r.N_pad; i += gridDim.x * blockDim.x) {
        const uint32_t e = r.int2ext[i];
        if (e == kInvalid) continue;
        uint32_t w = r.state[rec_word(i, kRecStat)];
        if (iv.car)  w = (w & ~(1u << 26)) | ((iv.car[e] ? 1u : 0u) << 26);
        if (iv.bike) w = (w & ~(1u << 27)) | ((iv.bike[e] ? 1u : 0u) << 27);
        r.state[rec_word(i, kRecStat)] = w;
    }
}

}  // namespace mvb

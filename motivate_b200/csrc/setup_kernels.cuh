// setup_kernels.cuh — create-time device code: the layout of layout.h built ON THE GPU from the caller's arrays.
//
// Round 1 built the orders, slices and offsets on host threads (layout.cpp: ~34 ms per 1M-agent graph per core),
// which capped every end-to-end number once several ranks shared one box's cores.  Here the same layout — the
// SAME one, entry for entry: layout.cpp stays as the specification and the tests compare the two — comes out of a
// handful of kernels: degree histogram -> (host: the O(H) plan) -> stable LSD radix sorts for the two orders ->
// per-slice maxima -> scans for the offsets.  Only a few scalars and the H x 253 histogram cross PCIe.
//
// Primitives (hand-written; nothing here calls a library):
//   exclusive scan (u32 / u64), three launches, any length
//   stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass, three launches per pass
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "layout.h"

namespace mvb {

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kScanThreads = 1024, kScanItems = 4, kScanTile = kScanThreads * kScanItems;

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp, T& block_total) {      // 1024 threads
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        T w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T y = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= (uint32_t)o) w += y;
        }
        s_warp[lane] = w;                                    // inclusive over warps
    }
    __syncthreads();
    block_total = s_warp[31];
    const T before = warp ? s_warp[warp - 1] : T(0);
    __syncthreads();
    return before + x - v;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const T* __restrict__ in, uint64_t n, T* __restrict__ sums) {
    __shared__ T s_warp[32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    T v = 0;
#pragma unroll
    for (uint32_t k = 0; k < kScanItems; ++k) if (base + k < n) v += in[base + k];
    T total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(T* __restrict__ sums, uint32_t nb, T* __restrict__ total_out) {
    __shared__ T s_warp[32];
    T carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += kScanThreads) {
        const uint32_t i = b0 + threadIdx.x;
        const T v = i < nb ? sums[i] : T(0);
        T total;
        const T ex = block_exclusive_scan(v, s_warp, total);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const T* __restrict__ in, T* __restrict__ out, uint64_t n,
                                                                  const T* __restrict__ sums) {
    __shared__ T s_warp[32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    T it[kScanItems];
    T v = 0;
#pragma unroll
    for (uint32_t k = 0; k < kScanItems; ++k) { it[k] = base + k < n ? in[base + k] : T(0); v += it[k]; }
    T total;
    T ex = block_exclusive_scan(v, s_warp, total) + sums[blockIdx.x];
#pragma unroll
    for (uint32_t k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = ex; ex += it[k]; }
}
inline uint32_t scan_blocks(uint64_t n) { return (uint32_t)((n + kScanTile - 1) / kScanTile); }
// out[i] = sum of in[0..i); *total_out (device, optional) = sum of all.  `sums` = scratch of scan_blocks(n) entries.
// in == out is allowed.
template <class T>
inline cudaError_t exclusive_scan(const T* in, T* out, uint64_t n, T* sums, T* total_out, cudaStream_t st) {
    const uint32_t nb = scan_blocks(n);
    if (nb == 0) { if (total_out) return cudaMemsetAsync(total_out, 0, sizeof(T), st); return cudaSuccess; }
    scan_reduce_kernel<T><<<nb, kScanThreads, 0, st>>>(in, n, sums);
    scan_sums_kernel<T><<<1, kScanThreads, 0, st>>>(sums, nb, total_out);
    scan_apply_kernel<T><<<nb, kScanThreads, 0, st>>>(in, out, n, sums);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// stable LSD radix sort, (key, value) pairs, 8 bits per pass
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kRsThreads = 256, kRsWarps = kRsThreads / 32, kRsRounds = 8, kRsTile = kRsThreads * kRsRounds;   // 2 048 per tile

// Element index order inside a tile = (warp, round, lane): warp w owns 256 consecutive elements, 32 per round.
__global__ void __launch_bounds__(kRsThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift,
                                                                uint32_t ntiles, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kRsTile;
#pragma unroll
    for (uint32_t r = 0; r < kRsRounds; ++r) {
        const uint32_t i = base + r * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = s_h[threadIdx.x];       // bin-major: one scan gives every base
}
__global__ void __launch_bounds__(kRsThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                   uint32_t n, uint32_t shift, uint32_t ntiles,
                                                                   const uint32_t* __restrict__ bases, uint32_t* __restrict__ keys_out,
                                                                   uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t s_cnt[kRsWarps][256];               // per warp: elements of each digit (then: running position)
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t k = threadIdx.x; k < kRsWarps * 256; k += kRsThreads) (&s_cnt[0][0])[k] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * kRsTile + warp * (32 * kRsRounds);
    uint32_t key[kRsRounds], dig[kRsRounds];
    // pass A: count this warp's digits
#pragma unroll
    for (uint32_t r = 0; r < kRsRounds; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys[i] : 0u;
        dig[r] = ok ? (key[r] >> shift) & 255u : 256u + lane;          // out-of-range lanes match nobody
        const uint32_t m = __match_any_sync(0xFFFFFFFFu, dig[r]);
        if (ok && (m & ((1u << lane) - 1u)) == 0) s_cnt[warp][dig[r]] += __popc(m);   // leader of its digit group
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan over the warps of every digit, plus the tile's global base of the digit
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = bases[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (uint32_t w = 0; w < kRsWarps; ++w) { const uint32_t c = s_cnt[w][d]; s_cnt[w][d] = run; run += c; }
    }
    __syncthreads();
    // pass B: same walk, stable positions
#pragma unroll
    for (uint32_t r = 0; r < kRsRounds; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t m = __match_any_sync(0xFFFFFFFFu, dig[r]);
        const uint32_t before = __popc(m & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (ok) pos = s_cnt[warp][dig[r]] + before;
        __syncwarp();
        if (ok && before == 0) s_cnt[warp][dig[r]] += __popc(m);
        __syncwarp();
        if (ok) { keys_out[pos] = key[r]; vals_out[pos] = vals[i]; }
    }
}
inline uint32_t radix_tiles(uint32_t n) { return (n + kRsTile - 1) / kRsTile; }
inline size_t radix_scratch_u32(uint32_t n) { const size_t c = (size_t)256 * radix_tiles(n); return c + scan_blocks(c) + 8; }
// Sorts by bits [shift0, shift0 + bits) of the key, stable.  keys/vals are ping-pong buffers [2][n]; returns in *result
// which half holds the output (0 or 1; `start` = the half that holds the input).  scratch: radix_scratch_u32(n) words.
inline cudaError_t radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, uint32_t bits, uint32_t* scratch,
                                    int* result, cudaStream_t st, uint32_t shift0 = 0, int start = 0) {
    int cur = start;
    if (n) {
        const uint32_t ntiles = radix_tiles(n);
        const size_t c = (size_t)256 * ntiles;
        uint32_t* counts = scratch;
        uint32_t* sums = scratch + c;
        for (uint32_t shift = shift0; shift < shift0 + bits; shift += 8) {
            radix_hist_kernel<<<ntiles, kRsThreads, 0, st>>>(keys[cur], n, shift, ntiles, counts);
            cudaError_t e = exclusive_scan<uint32_t>(counts, counts, c, sums, nullptr, st);
            if (e != cudaSuccess) return e;
            radix_scatter_kernel<<<ntiles, kRsThreads, 0, st>>>(keys[cur], vals[cur], n, shift, ntiles, counts, keys[cur ^ 1], vals[cur ^ 1]);
            cur ^= 1;
        }
    }
    *result = cur;
    return cudaGetLastError();
}
inline uint32_t bits_for(uint32_t max_value) { uint32_t b = 1; while (b < 32 && (max_value >> b)) ++b; return b; }

// ---------------------------------------------------------------------------------------------------------------
// layout kernels (same rules as layout.cpp; comments there)
// ---------------------------------------------------------------------------------------------------------------
// Error word of a set-up: first error wins.  code | index
enum SetupError : uint32_t { kErrNone = 0, kErrSubculture = 1, kErrNeighbourhood = 2, kErrCode = 3, kErrRowptr = 4, kErrDegree = 5, kErrTarget = 6 };
struct SetupFlag { uint32_t code, index; };
__device__ __forceinline__ void setup_fail(SetupFlag* f, uint32_t code, uint32_t index) {
    if (atomicCAS(&f->code, 0u, code) == 0u) f->index = index;
}

// Per agent: degree, hub or not, key of the first order; histogram cnt[h][252 - d] of the non-hubs, residents per
// neighbourhood (hubs included), number of hubs.  Also the range checks of the graph-side arrays.
__global__ void degree_kernel(uint32_t N, uint32_t H, const uint64_t* __restrict__ srp, const uint64_t* __restrict__ nrp,
                              uint64_t Es, uint64_t En, const uint16_t* __restrict__ nbhd, uint32_t* __restrict__ deg,
                              uint32_t* __restrict__ key1, uint32_t* __restrict__ ids, uint32_t* __restrict__ cnt,
                              uint32_t* __restrict__ n_hubs, SetupFlag* flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint64_t s0 = srp[i], s1 = srp[i + 1], n0 = nrp[i], n1 = nrp[i + 1];
    const uint32_t h = nbhd[i];
    bool ok = true;
    if (s1 < s0 || n1 < n0 || s1 > Es || n1 > En || (i == 0 && (s0 | n0))) { setup_fail(flag, kErrRowptr, i); ok = false; }
    if (h >= H) { setup_fail(flag, kErrNeighbourhood, i); ok = false; }
    const uint64_t d64 = ok ? (s1 - s0) + (n1 - n0) : 0;
    if (d64 > 0xFFFFFFFFull) { setup_fail(flag, kErrDegree, i); ok = false; }
    const uint32_t d = ok ? (uint32_t)d64 : 0u;
    deg[i] = d;
    ids[i] = i;
    if (!ok) { key1[i] = 0; return; }
    if (d > kHubDegree) {
        key1[i] = H << 8;                                   // after every neighbourhood group
        atomicAdd(n_hubs, 1u);
    } else {
        key1[i] = h << 8 | (kHubDegree - d);
        atomicAdd(&cnt[(size_t)h * (kHubDegree + 1) + (kHubDegree - d)], 1u);
    }
}
// Range checks of the per-agent state arrays of one replica + residents per neighbourhood (neighbourhood.rs:71).
__global__ void validate_state_kernel(uint32_t N, uint32_t S, uint32_t H, const uint8_t* __restrict__ sub, const uint16_t* __restrict__ nbhd,
                                      const uint8_t* __restrict__ commute, const uint8_t* __restrict__ mode, const uint8_t* __restrict__ norm,
                                      uint32_t* __restrict__ nres, SetupFlag* flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (sub[i] >= S) setup_fail(flag, kErrSubculture, i);
    if (commute[i] > 2 || mode[i] > 3 || norm[i] > 3) setup_fail(flag, kErrCode, i);
    const uint32_t h = nbhd[i];
    if (h >= H) { setup_fail(flag, kErrNeighbourhood, i); return; }
    const uint32_t m = __match_any_sync(__activemask(), h);
    if ((m & ((1u << (threadIdx.x & 31u)) - 1u)) == 0) atomicAdd(&nres[h], __popc(m));
}

// ~degree of the hubs (they sit at the end of the first order), for the (degree desc, id asc) sort of the hubs
__global__ void hub_keys_kernel(uint32_t n_hubs, const uint32_t* __restrict__ hub_ids, const uint32_t* __restrict__ deg,
                                uint32_t* __restrict__ keys) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_hubs) keys[j] = ~deg[hub_ids[j]];
}
// hubs dealt round the hub slices: hub number j (biggest first) -> slot (j % n_hs) * 32 + j / n_hs
__global__ void place_hubs_kernel(uint32_t n_hubs, uint32_t n_hs, uint32_t staged_hub_slots, const uint32_t* __restrict__ hubs_sorted,
                                  uint32_t* __restrict__ int2ext, uint8_t* __restrict__ is_hot) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_hubs) return;
    const uint32_t slot = (j % n_hs) * 32 + j / n_hs, e = hubs_sorted[j];
    int2ext[slot] = e;
    is_hot[e] = slot < staged_hub_slots;
}
// first order (neighbourhood, degree desc, id asc), compact position q: hot if its rank in the group is below take[h]
__global__ void mark_hot_kernel(uint32_t n2, const uint32_t* __restrict__ key1_sorted, const uint32_t* __restrict__ order1,
                                const uint32_t* __restrict__ gofs, const uint32_t* __restrict__ take,
                                const uint32_t* __restrict__ take_common, uint8_t* __restrict__ is_hot) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n2) return;
    const uint32_t h = key1_sorted[q] >> 8, r = q - gofs[h];
    // 1 = staged for every block, 2 = for blocks of its own neighbourhood (LOCAL layout: behind the common prefix)
    is_hot[order1[q]] = r < take[h] ? ((take_common && r >= take_common[h]) ? 2 : 1) : 0;
}
// key of the second order: (group, hot part first, hot-list length desc); values keep the first order
__global__ void key2_kernel(uint32_t n2, const uint32_t* __restrict__ key1_sorted, const uint32_t* __restrict__ order1,
                            const uint32_t* __restrict__ gofs, const uint32_t* __restrict__ take, const uint32_t* __restrict__ take_common,
                            const uint32_t* __restrict__ nsh, const uint32_t* __restrict__ nnh, uint32_t* __restrict__ key2) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n2) return;
    const uint32_t h = key1_sorted[q] >> 8, e = order1[q], r = q - gofs[h];
    // three pieces per group, each ordered by hot-list length: common prefix | rest of the staged prefix | never staged
    const uint32_t part = r < take[h] ? ((take_common && r >= take_common[h]) ? 1u : 0u) : 2u;
    key2[q] = (4u * h + part) << 8 | (kHubDegree - (nsh[e] + nnh[e]));
}
// compact position q of the second order -> internal slot gstart[h] + (q - gofs[h])
__global__ void place_agents_kernel(uint32_t n2, const uint32_t* __restrict__ key2_sorted, const uint32_t* __restrict__ order2,
                                    const uint32_t* __restrict__ gofs, const uint32_t* __restrict__ gstart, uint32_t* __restrict__ int2ext) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n2) return;
    const uint32_t h = key2_sorted[q] >> 10;
    int2ext[gstart[h] + (q - gofs[h])] = order2[q];
}
// one warp per SELL slice: K = longest hot list, Kc = longest cold list of its 32 agents
__global__ void slice_dims_kernel(uint32_t n_hub_slices, uint32_t n_sell_slices, const uint32_t* __restrict__ int2ext,
                                  const uint32_t* __restrict__ deg, const uint32_t* __restrict__ nsh, const uint32_t* __restrict__ nnh,
                                  uint32_t push, uint32_t* __restrict__ sell_K, uint64_t* __restrict__ sell_len) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (s >= n_sell_slices) return;
    const uint32_t e = int2ext[(n_hub_slices + s) * 32 + lane];
    uint32_t hot = 0, cold = 0;
    if (e != kInvalid) { hot = nsh[e] + nnh[e]; cold = deg[e] - hot; }
    const uint32_t K = __reduce_max_sync(0xFFFFFFFFu, hot), Kc = __reduce_max_sync(0xFFFFFFFFu, cold);
    if (lane == 0) { sell_K[s] = K | Kc << 16; sell_len[s] = (uint64_t)32 * (K + (push ? 0u : Kc)); }   // push mode: hot rows only (layout.h)
}
// what the day kernel reads per slice (+ 64 zero records of slack); `rebase` is subtracted (agent-partitioned handles
// store only their own slices' codes)
__global__ void sell_meta_kernel(uint32_t n_sell_slices, const uint64_t* __restrict__ sell_off, const uint32_t* __restrict__ sell_K,
                                 uint64_t rebase, uint32_t own_begin, uint32_t own_end, uint32_t push, uint4* __restrict__ meta) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sell_slices + 64) return;
    uint4 m = make_uint4(0, 0, 0, 0);
    if (s >= own_begin && s < own_end) { const uint64_t o = sell_off[s] - rebase; m = make_uint4((uint32_t)o, (uint32_t)(o >> 32), push ? sell_K[s] & 0xFFFFu : sell_K[s], 0); }
    meta[s] = m;
}
// per hub slot: the four list lengths, how many units its lists are cut into and how many code words they take
__global__ void hub_dims_kernel(uint32_t n_hub_rows, const uint32_t* __restrict__ int2ext, const uint64_t* __restrict__ srp,
                                const uint64_t* __restrict__ nrp, const uint32_t* __restrict__ nsh, const uint32_t* __restrict__ nnh,
                                uint4* __restrict__ hub_deg, uint32_t* __restrict__ n_units, uint64_t* __restrict__ code_len) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_hub_rows) return;
    const uint32_t e = int2ext[k];
    uint4 dg = make_uint4(0, 0, 0, 0);
    uint32_t nu = 0;
    uint64_t len = 0;
    if (e != kInvalid) {
        const uint32_t ds = (uint32_t)(srp[e + 1] - srp[e]), dn = (uint32_t)(nrp[e + 1] - nrp[e]);
        dg = make_uint4(nsh[e], nnh[e], ds - nsh[e], dn - nnh[e]);
        const uint32_t nh = dg.x + dg.y, nc = dg.z + dg.w;
        for (uint32_t dh = 0, dc = 0; dh < nh || dc < nc;) {
            uint32_t K, Kc;
            hub_segment(nh - min(dh, nh), nc - min(dc, nc), K, Kc);
            len += (uint64_t)32 * (K + Kc);
            dh += 32 * K; dc += 32 * Kc;
            ++nu;
        }
    }
    hub_deg[k] = dg; n_units[k] = nu; code_len[k] = len;
}
__global__ void hub_units_kernel(uint32_t n_hub_rows, const uint4* __restrict__ hub_deg, const uint32_t* __restrict__ unit_base,
                                 const uint64_t* __restrict__ hub_off, uint64_t rebase, uint4* __restrict__ units,
                                 uint32_t* __restrict__ hub_unit_begin, uint32_t total_units) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_hub_rows) return;
    if (k == n_hub_rows) { hub_unit_begin[k >> 5] = total_units; units[total_units] = make_uint4(0, 0, 0, 0); return; }
    if ((k & 31u) == 0) hub_unit_begin[k >> 5] = unit_base[k];
    const uint4 dg = hub_deg[k];
    const uint32_t nh = dg.x + dg.y, nc = dg.z + dg.w;
    uint64_t hoff = hub_off[k] - rebase;
    uint32_t u = unit_base[k];
    for (uint32_t dh = 0, dc = 0; dh < nh || dc < nc; ++u) {
        uint32_t K, Kc;
        hub_segment(nh - min(dh, nh), nc - min(dc, nc), K, Kc);
        const uint32_t l0 = dg.x > dh ? min(dg.x - dh, 1024u) : 0u, l2 = dg.z > dc ? min(dg.z - dc, 1024u) : 0u,
                       lc = nc > dc ? min(nc - dc, 1024u) : 0u;
        units[u] = make_uint4((uint32_t)(hoff / 32), k, K | Kc << 8 | l0 << 16, l2 | lc << 16);
        hoff += (uint64_t)32 * (K + Kc);
        dh += 32 * K; dc += 32 * Kc;
    }
}
// ---- push mode (layout.h): the cold links of the SELL agents a handle updates, as chunk-wise target-sorted lists ----
// cold links of every owned SELL slot (own_slot0 = first owned slot, counted from the start of the SELL region)
__global__ void push_count_kernel(uint32_t n_hub_slices, uint32_t own_slot0, uint32_t n_own_slots, const uint32_t* __restrict__ int2ext,
                                  const uint32_t* __restrict__ deg, const uint32_t* __restrict__ nsh, const uint32_t* __restrict__ nnh,
                                  uint32_t* __restrict__ cold_cnt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_own_slots) return;
    const uint32_t e = int2ext[n_hub_slices * 32 + own_slot0 + k];
    cold_cnt[k] = e == kInvalid ? 0u : deg[e] - nsh[e] - nnh[e];
}
// The chunks: consecutive owned SELL slots with about `target_links` cold links each (so that every block of push_kernel
// gets the same work: the cold lists of a neighbourhood's first, high-degree agents are ten times longer than those of its
// last), at most max_agents agents (their counters must fit in shared memory), cut at slice boundaries.  One thread, a
// binary search in the exclusive prefix `emit_off` per chunk.  first[c] = first slot of chunk c, first[n_chunks] = n_own_slots.
__global__ void push_plan_kernel(uint32_t n_own_slots, const uint32_t* __restrict__ emit_off, const uint32_t* __restrict__ total,
                                 uint32_t grid, uint32_t max_agents, uint32_t cap, uint32_t* __restrict__ first,
                                 uint32_t* __restrict__ n_chunks_out) {
    if (blockIdx.x || threadIdx.x) return;
    const uint32_t tot = *total;
    auto prefix = [&](uint32_t k) { return k < n_own_slots ? emit_off[k] : tot; };
    // one greedy pass for a given number of rounds of the grid; returns the number of chunks (cap + 1: too many)
    auto pass = [&](uint64_t rounds, bool write) -> uint32_t {
        const uint32_t target = (uint32_t)(((uint64_t)tot + rounds * grid - 1) / (rounds * grid)) + 1;
        uint32_t pos = 0, c = 0;
        while (pos < n_own_slots) {
            if (c == cap) return cap + 1;
            if (write) first[c] = pos;
            const uint32_t begin = pos;
            ++c;
            const uint64_t want = (uint64_t)prefix(pos) + target;
            uint32_t lo = min(n_own_slots, pos + 32), hi = min(n_own_slots, pos + max_agents);   // end in [lo, hi], the largest with prefix(end) <= want
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo + 1) / 2;
                if (prefix(mid) <= want) lo = mid; else hi = mid - 1;
            }
            pos = min(n_own_slots, (lo + 31u) & ~31u);
            if (pos - begin > max_agents) pos = begin + max_agents;
        }
        if (write) first[c] = n_own_slots;
        return c;
    };
    // Chunks that stop at the agent limit hold fewer links than the target, so a pass may end with a few chunks more than
    // `rounds` per block, and the blocks that get one more set the pace: take more rounds (smaller targets) until the chunks
    // fill whole rounds of the grid again.
    uint64_t rounds = 1;
    while (((uint64_t)n_own_slots + rounds * grid - 1) / (rounds * grid) > max_agents) ++rounds;
    uint64_t best = rounds;
    for (uint64_t r = rounds; r < rounds + 12; ++r) {
        const uint32_t c = pass(r, false);
        if (c > cap) break;                                                      // (keep the last plan that fits the chunk table)
        best = r;
        if (c <= r * grid) break;
    }
    const uint32_t c = pass(best, true);
    *n_chunks_out = c > cap ? 0xFFFFFFFFu : c;                                  // (cap too small: the caller reports it)
}
// first list entry of every chunk (+ the total behind the last)
__global__ void push_chunk_begin_kernel(uint32_t n_chunks, uint32_t n_own_slots, const uint32_t* __restrict__ first,
                                        const uint32_t* __restrict__ emit_off, const uint32_t* __restrict__ total,
                                        uint32_t* __restrict__ chunk_begin) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    const uint32_t k = first[c];
    chunk_begin[c] = k >= n_own_slots ? *total : emit_off[k];
}
// the sorted (chunk << 16 | source) words -> the u16 the push kernel reads
__global__ void push_src_kernel(uint32_t n, const uint32_t* __restrict__ val, uint16_t* __restrict__ src) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) src[i] = (uint16_t)(val[i] & 0xFFFFu);
}

__global__ void rebase_u64_kernel(uint64_t* __restrict__ p, uint32_t n, uint64_t rebase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] -= rebase;
}

}  // namespace mvb

// synth_device.cuh — synthetic populations and Barabási–Albert networks generated ON THE GPU.
//
// The reference draws its populations sequentially (src/agent_generation.rs:114-325, src/social_network.rs:12-48:
// a WeightedChoice rebuilt per arriving node, O(N^2)); synth.cpp restates that with an endpoint list on one host
// core (1.3 s per 1M agents, minutes for 50M).  Here every agent and every link is its own thread:
//   * attributes from a counter-based hash stream (agent, field) -> 64 random bits;
//   * exactly n_cars / n_bikes owners: agents sorted by a hash key (radix sort of setup_kernels.cuh), first k win —
//     a uniform sample without replacement, like sample_slice_ref (agent_generation.rs:149-164);
//   * preferential attachment without the sequential dependency: link e of node i picks a uniformly random
//     position r of the endpoint list as it was BEFORE node i arrived (so weights are frozen per node, as
//     social_network.rs:30-34 does, and duplicates are kept, :37-41).  Position 2e' is the source of link e' (known
//     from the index alone), position 2e'+1 is ITS target — resolved by following the chain of picks until it ends at
//     an even position or in the initial clique.  The chain is short (each step at least halves the link index in
//     expectation) and every pick is a pure function of (seed, e), so all links resolve independently;
//   * CSR by degree count + scan + atomic cursors.  The order of a row's entries depends on the atomics; the update
//     only counts a list's modes (agent.rs:322-332), so nothing downstream depends on it.
// Same distributions as synth.cpp, not the same stream: a device population is its own seeded realisation.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "setup_kernels.cuh"
#include "synth.h"

namespace mvb {

__host__ __device__ inline uint64_t hash64(uint64_t seed, uint64_t stream, uint64_t index) {
    uint64_t z = seed + stream * 0xD1B54A32D192ED03ull + index * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t z) { return (double)(z >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ uint64_t below64(uint64_t z, uint64_t n) { return __umul64hi(z, n); }

enum SynthStream : uint64_t { kStSub = 1, kStNbhd, kStWs, kStGmmPick, kStGmmU1, kStGmmU2, kStMode, kStCar, kStBike, kStSocial, kStNeighbour = 64 };

struct SynthGmm { uint32_t n; double mean[8], sd[8], weight[8]; };

__global__ void synth_attr_kernel(uint32_t N, uint32_t S, uint32_t H, uint64_t seed, SynthGmm gmm, uint8_t* __restrict__ sub,
                                  uint16_t* __restrict__ nbhd, float* __restrict__ ws, uint8_t* __restrict__ commute,
                                  uint32_t* __restrict__ car_key, uint32_t* __restrict__ bike_key, uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    sub[i] = (uint8_t)below64(hash64(seed, kStSub, i), S);                 // uniform subculture / neighbourhood (:261-280)
    nbhd[i] = (uint16_t)below64(hash64(seed, kStNbhd, i), H);
    ws[i] = (float)(hash64(seed, kStWs, i) >> 40) * (1.0f / 16777216.0f);  // U[0,1) (:224)
    double total_w = 0;
    for (uint32_t g = 0; g < gmm.n; ++g) total_w += gmm.weight[g];
    const double r = u01(hash64(seed, kStGmmPick, i)) * total_w;           // |GMM sample| (:166-189, gaussian.rs:36-98)
    double cdf = 0;
    uint32_t g = gmm.n - 1;
    for (uint32_t k = 0; k < gmm.n; ++k) { cdf += gmm.weight[k]; if (r < cdf) { g = k; break; } }
    double u1 = u01(hash64(seed, kStGmmU1, i));
    const double u2 = u01(hash64(seed, kStGmmU2, i));
    if (u1 < 1e-300) u1 = 1e-300;
    const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    const double d = fabs(gmm.mean[g] + gmm.sd[g] * z);
    commute[i] = d < 4241.0 ? 0 : (d < 19457.0 ? 1 : 2);                    // (:182-188)
    car_key[i] = (uint32_t)(hash64(seed, kStCar, i) >> 32);
    bike_key[i] = (uint32_t)(hash64(seed, kStBike, i) >> 32);
    ids[i] = i;
}
__global__ void synth_mark_owners_kernel(uint32_t k, const uint32_t* __restrict__ sorted_ids, uint8_t* __restrict__ owns) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) owns[sorted_ids[j]] = 1;
}
__global__ void synth_mode_kernel(uint32_t N, uint64_t seed, const uint8_t* __restrict__ car, const uint8_t* __restrict__ bike,
                                  uint8_t* __restrict__ mode, uint8_t* __restrict__ norm, float4* __restrict__ habit) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double r = u01(hash64(seed, kStMode, i));                         // initial mode by ownership (:286-325)
    uint32_t m;
    if (car[i] && bike[i]) m = r < 0.4 ? 0 : (r < 0.7 ? 2 : (r < 0.85 ? 3 : 1));
    else if (car[i])       m = r < 0.57 ? 0 : (r < 0.79 ? 3 : 1);
    else if (bike[i])      m = r < 0.5 ? 2 : (r < 0.75 ? 3 : 1);
    else                   m = r < 0.5 ? 3 : 1;
    mode[i] = (uint8_t)m; norm[i] = (uint8_t)m;                             // habit = {mode: 1}, norm = mode (:198-200)
    habit[i] = make_float4(m == 0 ? 1.f : 0.f, m == 1 ? 1.f : 0.f, m == 2 ? 1.f : 0.f, m == 3 ? 1.f : 0.f);
}

// ---- preferential attachment, one thread per link -------------------------------------------------------------
// Links of a network of n nodes with minimum m: the first min(n, m) nodes form a clique (node i links to every k < i:
// link index i(i-1)/2 + k); every later node i adds m links, indices C + (i - m) m + k, C = m(m-1)/2.
__host__ __device__ inline uint64_t ba_link_count(uint32_t m, uint32_t n) {
    if (n <= m) return (uint64_t)n * (n - 1) / 2;
    return (uint64_t)m * (m - 1) / 2 + (uint64_t)(n - m) * m;
}
__host__ __device__ inline uint32_t ba_source(uint32_t m, uint64_t e) {     // the node that brought link e
    const uint64_t C = (uint64_t)m * (m - 1) / 2;
    if (e >= C) return m + (uint32_t)((e - C) / m);
    uint32_t i = 1;
    while ((uint64_t)(i + 1) * i / 2 <= e) ++i;                             // i(i-1)/2 <= e < (i+1)i/2
    return i;
}
__host__ __device__ inline uint32_t ba_target(uint32_t m, uint64_t e, uint64_t seed, uint64_t stream) {
    const uint64_t C = (uint64_t)m * (m - 1) / 2;
    for (;;) {
        if (e < C) { const uint32_t i = ba_source(m, e); return (uint32_t)(e - (uint64_t)i * (i - 1) / 2); }
        const uint64_t i = m + (e - C) / m;                                 // the arriving node
        const uint64_t len_before = 2 * (C + (i - m) * m);                  // endpoint list before node i (weights frozen per node)
        const uint64_t z = hash64(seed, stream, e);
#if defined(__CUDA_ARCH__)
        const uint64_t r = __umul64hi(z, len_before);
#else
        const uint64_t r = (uint64_t)(((unsigned __int128)z * len_before) >> 64);
#endif
        if ((r & 1u) == 0) return ba_source(m, r >> 1);
        e = r >> 1;                                                         // the target of that earlier link
    }
}
// One network over nodes 0..n-1 (social) or one per group (neighbour; node = position in the group's member list).
struct BaGroups {
    uint32_t n_groups;
    const uint64_t* link_begin;        // [n_groups + 1] first global link index of every group
    const uint32_t* member_begin;      // [n_groups + 1] first entry of every group in members (nullptr: one group, identity)
    const uint32_t* members;           // agent ids by group, id ascending inside a group
};
__global__ void ba_links_kernel(uint32_t m, uint64_t n_links, uint64_t seed, uint64_t stream, BaGroups gr,
                                uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t* __restrict__ deg) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_links) return;
    uint32_t a, b;
    if (!gr.members) {
        a = ba_source(m, e); b = ba_target(m, e, seed, stream);
    } else {
        uint32_t lo = 0, hi = gr.n_groups;                                   // group of link e: last one starting at or before it
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (gr.link_begin[mid] <= e) lo = mid; else hi = mid; }
        const uint64_t le = e - gr.link_begin[lo];
        const uint32_t* mem = gr.members + gr.member_begin[lo];
        a = mem[ba_source(m, le)]; b = mem[ba_target(m, le, seed, stream + lo)];
    }
    src[e] = a; dst[e] = b;
    atomicAdd(&deg[a], 1u); atomicAdd(&deg[b], 1u);
}
__global__ void widen_kernel(uint32_t n, const uint32_t* __restrict__ in, uint64_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void csr_fill_kernel(uint64_t n_links, const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst,
                                const uint64_t* __restrict__ rowptr, uint32_t* __restrict__ cursor, uint32_t* __restrict__ col) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_links) return;
    const uint32_t a = src[e], b = dst[e];
    col[rowptr[a] + atomicAdd(&cursor[a], 1u)] = b;
    col[rowptr[b] + atomicAdd(&cursor[b], 1u)] = a;
}
__global__ void iota_kernel(uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
__global__ void group_count_kernel(uint32_t N, const uint16_t* __restrict__ nbhd, uint32_t* __restrict__ count, uint32_t* __restrict__ keys,
                                   uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t h = nbhd[i];
    keys[i] = h; ids[i] = i;
    const uint32_t mm = __match_any_sync(__activemask(), h);
    if ((mm & ((1u << (threadIdx.x & 31u)) - 1u)) == 0) atomicAdd(&count[h], __popc(mm));
}

}  // namespace mvb

extern "C" {

int mvb_synth_create_device(const mvb_synth_spec* sp, int device, mvb_synth** out) {
    using namespace mvb;
    if (!sp || !out) { set_error("null argument"); return MVB_ERR_ARG; }
    *out = nullptr;
    const uint32_t N = sp->n_agents, S = sp->n_subcultures, H = sp->n_neighbourhoods;
    if (N == 0 || S == 0 || S > 256 || H == 0 || H > 65536 || sp->n_gmm == 0 || sp->n_gmm > 8) {
        set_error("bad synth spec (N=%u S=%u H=%u n_gmm=%u)", N, S, H, sp->n_gmm);
        return MVB_ERR_ARG;
    }
    if (sp->social_links_min < 2 || sp->neighbour_links_min < 2) { set_error("link minimums must be >= 2 (the reference panics for 1)"); return MVB_ERR_DOMAIN; }
    if (sp->n_cars > N || sp->n_bikes > N) { set_error("more cars/bikes than agents"); return MVB_ERR_DOMAIN; }
    int ndev = mvb_device_count();
    if (ndev <= 0) { set_error("no CUDA device available (mvb_synth_create builds the same kind of population on the host)"); return MVB_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(device));
    struct Stream { cudaStream_t s = nullptr; ~Stream() { if (s) cudaStreamDestroy(s); } } strm;
    MVB_CUDA(cudaStreamCreateWithFlags(&strm.s, cudaStreamNonBlocking));
    cudaStream_t st = strm.s;
    const uint32_t ms = sp->social_links_min, mn = sp->neighbour_links_min;
    const uint64_t n_social = ba_link_count(ms, N);

    // scratch: everything temporary in one allocation
    struct Scratch { void* p = nullptr; ~Scratch() { if (p) cudaFree(p); } } scr;
    auto al = [](size_t& at, size_t bytes) { const size_t o = at; at += (bytes + 255) & ~(size_t)255; return o; };
    // group sizes first (they fix the neighbour link count): attributes need only small scratch
    size_t a1 = 0;
    const size_t o_key0 = al(a1, 4 * (size_t)N), o_key1 = al(a1, 4 * (size_t)N), o_val0 = al(a1, 4 * (size_t)N), o_val1 = al(a1, 4 * (size_t)N),
                 o_bkey = al(a1, 4 * (size_t)N), o_rs = al(a1, 4 * radix_scratch_u32(N)), o_cnt = al(a1, 4 * (size_t)H),
                 o_mem = al(a1, 4 * (size_t)N);
    // persistent arrays except the columns, whose sizes come later
    std::unique_ptr<mvb_synth> sy(new (std::nothrow) mvb_synth());
    if (!sy) { set_error("out of memory"); return MVB_ERR_NOMEM; }
    sy->device = device;
    MVB_CUDA(cudaMalloc(&scr.p, a1));
    char* T = (char*)scr.p;
    uint32_t* key[2] = {(uint32_t*)(T + o_key0), (uint32_t*)(T + o_key1)};
    uint32_t* val[2] = {(uint32_t*)(T + o_val0), (uint32_t*)(T + o_val1)};
    uint32_t* bkey = (uint32_t*)(T + o_bkey);
    uint32_t* rs = (uint32_t*)(T + o_rs);
    uint32_t* gcount = (uint32_t*)(T + o_cnt);
    uint32_t* members = (uint32_t*)(T + o_mem);

    // per-agent arrays and row pointers
    size_t a2 = 0;
    const size_t o_sub = al(a2, N), o_nb = al(a2, 2 * (size_t)N), o_cm = al(a2, N), o_car = al(a2, N), o_bike = al(a2, N), o_ws = al(a2, 4 * (size_t)N),
                 o_hab = al(a2, 16 * (size_t)N), o_mode = al(a2, N), o_norm = al(a2, N), o_srp = al(a2, 8 * ((size_t)N + 1)), o_nrp = al(a2, 8 * ((size_t)N + 1));
    // (the columns are a second allocation: their sizes depend on the neighbourhood sizes drawn below)
    struct Block { void* p = nullptr; ~Block() { if (p) cudaFree(p); } } attr;
    MVB_CUDA(cudaMalloc(&attr.p, a2));
    char* A = (char*)attr.p;
    uint8_t* sub = (uint8_t*)(A + o_sub); uint16_t* nbhd = (uint16_t*)(A + o_nb); uint8_t* commute = (uint8_t*)(A + o_cm);
    uint8_t* car = (uint8_t*)(A + o_car); uint8_t* bike = (uint8_t*)(A + o_bike); float* ws = (float*)(A + o_ws);
    float4* habit = (float4*)(A + o_hab); uint8_t* mode = (uint8_t*)(A + o_mode); uint8_t* norm = (uint8_t*)(A + o_norm);
    uint64_t* srp = (uint64_t*)(A + o_srp); uint64_t* nrp = (uint64_t*)(A + o_nrp);

    SynthGmm gmm{};
    gmm.n = sp->n_gmm;
    for (uint32_t g = 0; g < sp->n_gmm; ++g) { gmm.mean[g] = sp->gmm[g][0]; gmm.sd[g] = sp->gmm[g][1]; gmm.weight[g] = sp->gmm[g][2]; }
    const uint64_t seed = hash64(sp->seed, 0, 0x5EED);
    const unsigned nb = (N + 255) / 256;
    synth_attr_kernel<<<nb, 256, 0, st>>>(N, S, H, seed, gmm, sub, nbhd, ws, commute, key[0], bkey, val[0]);
    MVB_CUDA(cudaGetLastError());
    MVB_CUDA(cudaMemsetAsync(car, 0, N, st)); MVB_CUDA(cudaMemsetAsync(bike, 0, N, st));
    int cur = 0;
    MVB_CUDA(radix_sort_pairs(key, val, N, 32, rs, &cur, st));                         // cars: the agents with the n_cars smallest keys
    if (sp->n_cars) synth_mark_owners_kernel<<<(sp->n_cars + 255) / 256, 256, 0, st>>>(sp->n_cars, val[cur], car);
    MVB_CUDA(cudaMemcpyAsync(key[0], bkey, 4 * (size_t)N, cudaMemcpyDeviceToDevice, st));
    iota_kernel<<<nb, 256, 0, st>>>(N, val[0]);
    MVB_CUDA(radix_sort_pairs(key, val, N, 32, rs, &cur, st));                         // bikes likewise, independent of the cars
    if (sp->n_bikes) synth_mark_owners_kernel<<<(sp->n_bikes + 255) / 256, 256, 0, st>>>(sp->n_bikes, val[cur], bike);
    synth_mode_kernel<<<nb, 256, 0, st>>>(N, seed, car, bike, mode, norm, habit);
    MVB_CUDA(cudaGetLastError());
    // members of every neighbourhood in agent order (simulation.rs:165-168): stable sort of the ids by neighbourhood
    MVB_CUDA(cudaMemsetAsync(gcount, 0, 4 * (size_t)H, st));
    group_count_kernel<<<nb, 256, 0, st>>>(N, nbhd, gcount, key[0], val[0]);
    MVB_CUDA(radix_sort_pairs(key, val, N, bits_for(H), rs, &cur, st));
    MVB_CUDA(cudaMemcpyAsync(members, val[cur], 4 * (size_t)N, cudaMemcpyDeviceToDevice, st));
    std::vector<uint32_t> gsize(H);
    MVB_CUDA(cudaMemcpyAsync(gsize.data(), gcount, 4 * (size_t)H, cudaMemcpyDeviceToHost, st));
    MVB_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t> link_begin((size_t)H + 1, 0);
    std::vector<uint32_t> member_begin((size_t)H + 1, 0);
    for (uint32_t h = 0; h < H; ++h) {
        link_begin[h + 1] = link_begin[h] + ba_link_count(mn, gsize[h]);
        member_begin[h + 1] = member_begin[h] + gsize[h];
    }
    const uint64_t n_neigh = link_begin[H];
    if (member_begin[H] != N) { set_error("internal: neighbourhood sizes do not add up"); return MVB_ERR_STATE; }
    if (2 * n_social >= (1ull << 40) || 2 * n_neigh >= (1ull << 40)) { set_error("too many links"); return MVB_ERR_ARG; }
    // columns + link scratch
    struct Block2 { void* p = nullptr; ~Block2() { if (p) cudaFree(p); } } cols, ls;
    size_t a3 = 0;
    const size_t o_scol = al(a3, 8 * n_social), o_ncol = al(a3, 8 * n_neigh);
    MVB_CUDA(cudaMalloc(&cols.p, std::max<size_t>(a3, 256)));
    uint32_t* scol = (uint32_t*)((char*)cols.p + o_scol); uint32_t* ncol = (uint32_t*)((char*)cols.p + o_ncol);
    const uint64_t n_max = std::max(n_social, n_neigh);
    size_t a4 = 0;
    const size_t o_src = al(a4, 4 * n_max), o_dst = al(a4, 4 * n_max), o_deg = al(a4, 4 * (size_t)N), o_deg64 = al(a4, 8 * ((size_t)N + 1)),
                 o_sums = al(a4, 8 * ((size_t)scan_blocks((uint64_t)N + 1) + 8)), o_lb = al(a4, 8 * ((size_t)H + 1)), o_mb = al(a4, 4 * ((size_t)H + 1));
    MVB_CUDA(cudaMalloc(&ls.p, a4));
    char* Lc = (char*)ls.p;
    uint32_t* src = (uint32_t*)(Lc + o_src); uint32_t* dst = (uint32_t*)(Lc + o_dst); uint32_t* deg = (uint32_t*)(Lc + o_deg);
    uint64_t* deg64 = (uint64_t*)(Lc + o_deg64); uint64_t* sums = (uint64_t*)(Lc + o_sums);
    uint64_t* d_lb = (uint64_t*)(Lc + o_lb); uint32_t* d_mb = (uint32_t*)(Lc + o_mb);
    MVB_CUDA(cudaMemcpyAsync(d_lb, link_begin.data(), 8 * ((size_t)H + 1), cudaMemcpyHostToDevice, st));
    MVB_CUDA(cudaMemcpyAsync(d_mb, member_begin.data(), 4 * ((size_t)H + 1), cudaMemcpyHostToDevice, st));
    for (int net = 0; net < 2; ++net) {
        const uint64_t n_links = net ? n_neigh : n_social;
        uint64_t* rowptr = net ? nrp : srp;
        uint32_t* col = net ? ncol : scol;
        MVB_CUDA(cudaMemsetAsync(deg, 0, 4 * (size_t)N, st));
        BaGroups gr{H, d_lb, net ? d_mb : nullptr, net ? members : nullptr};
        if (n_links)
            ba_links_kernel<<<(unsigned)((n_links + 255) / 256), 256, 0, st>>>(net ? mn : ms, n_links, seed, net ? (uint64_t)kStNeighbour : (uint64_t)kStSocial,
                                                                              gr, src, dst, deg);
        MVB_CUDA(cudaGetLastError());
        MVB_CUDA(cudaMemsetAsync(deg64, 0, 8 * ((size_t)N + 1), st));
        widen_kernel<<<nb, 256, 0, st>>>(N, deg, deg64);
        MVB_CUDA(exclusive_scan<uint64_t>(deg64, rowptr, (uint64_t)N + 1, sums, nullptr, st));
        MVB_CUDA(cudaMemsetAsync(deg, 0, 4 * (size_t)N, st));
        if (n_links) csr_fill_kernel<<<(unsigned)((n_links + 255) / 256), 256, 0, st>>>(n_links, src, dst, rowptr, deg, col);
        MVB_CUDA(cudaGetLastError());
    }
    MVB_CUDA(cudaStreamSynchronize(st));
    mvb_population& p = sy->pop;
    p.n_agents = N;
    p.subculture = sub; p.neighbourhood = nbhd; p.commute = commute; p.owns_car = car; p.owns_bike = bike; p.weather_sensitivity = ws;
    p.consistency = p.social_connectivity = p.subculture_connectivity = nullptr;
    p.neighbourhood_connectivity = p.average_weight = nullptr;
    p.habit = (const float*)habit; p.current_mode = mode; p.norm = norm;
    p.social_rowptr = srp; p.social_col = scol; p.neighbour_rowptr = nrp; p.neighbour_col = ncol;
    sy->Es = 2 * n_social; sy->En = 2 * n_neigh;
    sy->d_block = attr.p; attr.p = nullptr;
    sy->d_block2 = cols.p; cols.p = nullptr;
    *out = sy.release();
    return MVB_OK;
}

int mvb_synth_copy_to_host(const mvb_synth* d, mvb_synth** out) {
    using namespace mvb;
    if (!d || !out) { set_error("null argument"); return MVB_ERR_ARG; }
    *out = nullptr;
    if (d->device < 0) { set_error("the population is already in host memory"); return MVB_ERR_STATE; }
    MVB_CUDA(cudaSetDevice(d->device));
    std::unique_ptr<mvb_synth> h(new (std::nothrow) mvb_synth());
    if (!h) { set_error("out of memory"); return MVB_ERR_NOMEM; }
    const mvb_population& s = d->pop;
    const size_t N = s.n_agents;
    h->sub.resize(N); h->nbhd.resize(N); h->commute.resize(N); h->car.resize(N); h->bike.resize(N); h->ws.resize(N); h->habit.resize(4 * N);
    h->mode.resize(N); h->norm.resize(N); h->s_rowptr.resize(N + 1); h->n_rowptr.resize(N + 1); h->s_col.resize(d->Es); h->n_col.resize(d->En);
    MVB_CUDA(cudaMemcpy(h->sub.data(), s.subculture, N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->nbhd.data(), s.neighbourhood, 2 * N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->commute.data(), s.commute, N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->car.data(), s.owns_car, N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->bike.data(), s.owns_bike, N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->ws.data(), s.weather_sensitivity, 4 * N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->habit.data(), s.habit, 16 * N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->mode.data(), s.current_mode, N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->norm.data(), s.norm, N, cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->s_rowptr.data(), s.social_rowptr, 8 * (N + 1), cudaMemcpyDeviceToHost));
    MVB_CUDA(cudaMemcpy(h->n_rowptr.data(), s.neighbour_rowptr, 8 * (N + 1), cudaMemcpyDeviceToHost));
    if (d->Es) MVB_CUDA(cudaMemcpy(h->s_col.data(), s.social_col, 4 * d->Es, cudaMemcpyDeviceToHost));
    if (d->En) MVB_CUDA(cudaMemcpy(h->n_col.data(), s.neighbour_col, 4 * d->En, cudaMemcpyDeviceToHost));
    mvb_population& p = h->pop;
    p.n_agents = s.n_agents;
    p.subculture = h->sub.data(); p.neighbourhood = h->nbhd.data(); p.commute = h->commute.data();
    p.owns_car = h->car.data(); p.owns_bike = h->bike.data(); p.weather_sensitivity = h->ws.data();
    p.consistency = p.social_connectivity = p.subculture_connectivity = nullptr;
    p.neighbourhood_connectivity = p.average_weight = nullptr;
    p.habit = h->habit.data(); p.current_mode = h->mode.data(); p.norm = h->norm.data();
    p.social_rowptr = h->s_rowptr.data(); p.social_col = h->s_col.data();
    p.neighbour_rowptr = h->n_rowptr.data(); p.neighbour_col = h->n_col.data();
    *out = h.release();
    return MVB_OK;
}

}  // extern "C"

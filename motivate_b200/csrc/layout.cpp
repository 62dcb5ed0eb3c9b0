// layout.cpp — builds the device layout described in layout.h from the caller's two CSR link graphs.
#include "layout.h"

#include <algorithm>
#include <numeric>

#include "mvb_internal.h"

namespace mvb {

namespace {
inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
}  // namespace

int build_graph_layout(const mvb_population& pop, uint32_t H, uint32_t capacity_agents, GraphLayout& L) {
    const uint32_t N = pop.n_agents;
    const uint64_t* srp = pop.social_rowptr;
    const uint64_t* nrp = pop.neighbour_rowptr;
    L = GraphLayout();
    L.N = N;
    L.Es = srp[N];
    L.En = nrp[N];
    if (capacity_agents < 64 || capacity_agents % 64) { set_error("internal: bad shared-memory capacity %u", capacity_agents); return MVB_ERR_ARG; }

    // ---- degrees, hubs, neighbourhood groups -------------------------------------------------
    std::vector<uint64_t> d(N);
    for (uint32_t i = 0; i < N; ++i) {
        const uint64_t ds = srp[i + 1] - srp[i], dn = nrp[i + 1] - nrp[i];
        if (ds > 0xFFFFFFFFull || dn > 0xFFFFFFFFull) { set_error("agent %u has more than 2^32 links", i); return MVB_ERR_ARG; }
        d[i] = ds + dn;
    }
    std::vector<uint32_t> hubs;
    std::vector<std::vector<uint32_t>> groups(H);
    for (uint32_t i = 0; i < N; ++i) {
        if (d[i] > kHubDegree) hubs.push_back(i);
        else groups[pop.neighbourhood[i]].push_back(i);
    }
    auto by_degree = [&](uint32_t a, uint32_t b) { return d[a] != d[b] ? d[a] > d[b] : a < b; };
    std::sort(hubs.begin(), hubs.end(), by_degree);
    for (auto& g : groups) std::sort(g.begin(), g.end(), by_degree);

    // ---- internal order ----------------------------------------------------------------------
    const uint32_t hub_pad = round_up((uint32_t)hubs.size(), 64);
    uint64_t total = hub_pad;
    std::vector<uint32_t> gstart(H), gpad(H);
    for (uint32_t h = 0; h < H; ++h) {
        gstart[h] = (uint32_t)total;
        gpad[h] = round_up((uint32_t)groups[h].size(), 64);
        total += gpad[h];
    }
    if (total >= (1ull << 29)) { set_error("too many agents for the 25-bit word index of a link code"); return MVB_ERR_ARG; }
    L.N_pad = (uint32_t)total;
    L.n_hub_slices = hub_pad / 32;
    L.n_sell_slices = (L.N_pad - hub_pad) / 32;
    L.int2ext.assign(L.N_pad, kInvalid);
    L.ext2int.assign(N, kInvalid);
    for (uint32_t k = 0; k < hubs.size(); ++k) L.int2ext[k] = hubs[k];
    for (uint32_t h = 0; h < H; ++h)
        for (uint32_t k = 0; k < groups[h].size(); ++k) L.int2ext[gstart[h] + k] = groups[h][k];
    for (uint32_t s = 0; s < L.N_pad; ++s)
        if (L.int2ext[s] != kInvalid) L.ext2int[L.int2ext[s]] = s;

    // ---- hot set: hub region + a degree-descending prefix of every group ---------------------
    // position of an internal slot in the staged vector, or kInvalid if cold
    std::vector<uint32_t> hot_pos(L.N_pad, kInvalid);
    uint32_t staged = 0;
    auto add_range = [&](uint32_t start, uint32_t count) {       // both multiples of 64
        if (!count) return;
        L.ranges.push_back(StageRange{start / 16, count / 16, staged / 16, 0});
        for (uint32_t k = 0; k < count; ++k) hot_pos[start + k] = staged + k;
        staged += count;
    };
    if (L.N_pad <= capacity_agents) {
        add_range(0, L.N_pad);
    } else {
        // largest degree threshold theta such that {hubs} ∪ {d >= theta} (rounded per group) fits
        uint32_t budget = capacity_agents > hub_pad ? capacity_agents - hub_pad : 0;
        uint32_t hub_take = std::min(hub_pad, capacity_agents);
        std::vector<uint32_t> take(H, 0);
        uint64_t lo = 0, hi = kHubDegree + 1;                    // theta in (lo, hi]: hi always fits
        auto count_for = [&](uint64_t theta, std::vector<uint32_t>& out) {
            uint64_t sum = 0;
            for (uint32_t h = 0; h < H; ++h) {
                const auto& g = groups[h];
                // g sorted by degree descending: first index with d < theta
                uint32_t c = (uint32_t)(std::partition_point(g.begin(), g.end(), [&](uint32_t a) { return d[a] >= theta; }) - g.begin());
                out[h] = std::min(round_up(c, 64), gpad[h]);
                sum += out[h];
            }
            return sum;
        };
        std::vector<uint32_t> tmp(H);
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) / 2;
            if (count_for(mid, tmp) <= budget) hi = mid; else lo = mid;
        }
        uint64_t used = count_for(hi, take);
        // spend what is left on the next-lower degree class, group by group
        for (uint32_t h = 0; h < H && used < budget; ++h) {
            const uint32_t extra = std::min<uint64_t>(gpad[h] - take[h], (budget - used) / 64 * 64);
            take[h] += extra; used += extra;
        }
        add_range(0, hub_take);
        for (uint32_t h = 0; h < H; ++h) add_range(gstart[h], take[h]);
    }
    L.hot_words = staged / 16;

    // ---- link codes ---------------------------------------------------------------------------
    auto code_of = [&](uint32_t ext_target) -> uint32_t {
        const uint32_t t = L.ext2int[ext_target];
        const uint32_t hp = hot_pos[t];
        const uint32_t pos = hp != kInvalid ? hp : t;
        return ((pos & 15u) * 2u) << 27 | (pos >> 4) << 2 | (hp != kInvalid ? 0u : 2u);
    };
    // hub rows: [social codes | neighbour codes | pad to 4]
    const uint32_t n_hub_rows = L.n_hub_slices * 32;
    L.hub_rowptr.assign((size_t)n_hub_rows + 1, 0);
    L.hub_ds.assign(n_hub_rows, 0);
    L.hub_dn.assign(n_hub_rows, 0);
    for (uint32_t k = 0; k < n_hub_rows; ++k) {
        uint64_t len = 0;
        if (k < hubs.size()) {
            const uint32_t e = hubs[k];
            L.hub_ds[k] = (uint32_t)(srp[e + 1] - srp[e]);
            L.hub_dn[k] = (uint32_t)(nrp[e + 1] - nrp[e]);
            len = (d[e] + 3) / 4 * 4;
        }
        L.hub_rowptr[k + 1] = L.hub_rowptr[k] + len;
    }
    L.hub_codes.assign(L.hub_rowptr[n_hub_rows], 0);
    for (uint32_t k = 0; k < hubs.size(); ++k) {
        const uint32_t e = hubs[k];
        uint32_t* o = L.hub_codes.data() + L.hub_rowptr[k];
        for (uint64_t x = srp[e]; x < srp[e + 1]; ++x) *o++ = code_of(pop.social_col[x]);
        for (uint64_t x = nrp[e]; x < nrp[e + 1]; ++x) *o++ = code_of(pop.neighbour_col[x]);
    }
    // SELL slices
    L.deg.assign(L.N_pad, 0);
    L.sell_off.assign(L.n_sell_slices, 0);
    L.sell_K.assign(L.n_sell_slices, 0);
    uint64_t off = 0;
    for (uint32_t s = 0; s < L.n_sell_slices; ++s) {
        const uint32_t base = hub_pad + s * 32;
        uint32_t K = 0;
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t e = L.int2ext[base + l];
            if (e != kInvalid) K = std::max<uint32_t>(K, (uint32_t)d[e]);
        }
        L.sell_off[s] = off;
        L.sell_K[s] = K;
        off += (uint64_t)32 * K;
    }
    L.sell_codes.assign(off, 0);
    std::vector<uint32_t> row;
    for (uint32_t s = 0; s < L.n_sell_slices; ++s) {
        const uint32_t base = hub_pad + s * 32, K = L.sell_K[s], K4 = K / 4;
        uint32_t* body = L.sell_codes.data() + L.sell_off[s];
        uint32_t* tail = body + (size_t)K4 * 128;
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t e = L.int2ext[base + l];
            if (e == kInvalid) continue;
            const uint32_t ds = (uint32_t)(srp[e + 1] - srp[e]), dn = (uint32_t)(nrp[e + 1] - nrp[e]);
            L.deg[base + l] = ds | dn << 16;
            row.clear();
            for (uint64_t x = srp[e]; x < srp[e + 1]; ++x) row.push_back(code_of(pop.social_col[x]));
            for (uint64_t x = nrp[e]; x < nrp[e + 1]; ++x) row.push_back(code_of(pop.neighbour_col[x]));
            for (uint32_t k = 0; k < row.size(); ++k) {
                if (k < 4 * K4) body[(size_t)(k / 4) * 128 + l * 4 + (k & 3)] = row[k];
                else tail[(size_t)(k - 4 * K4) * 32 + l] = row[k];
            }
        }
    }
    L.sell_order.resize(L.n_sell_slices);
    std::iota(L.sell_order.begin(), L.sell_order.end(), 0u);
    std::stable_sort(L.sell_order.begin(), L.sell_order.end(),
                     [&](uint32_t a, uint32_t b) { return L.sell_K[a] > L.sell_K[b]; });
    return MVB_OK;
}

}  // namespace mvb

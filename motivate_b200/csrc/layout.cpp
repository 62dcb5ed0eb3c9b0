// layout.cpp — builds the device layout described in layout.h from the caller's two CSR link graphs.
#include "layout.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "mvb_internal.h"

namespace mvb {

namespace {
inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// Stable counting sort of v[b, e) by key DESCENDING (keys <= max_key): elements with equal keys keep their order.
template <class KeyFn>
void counting_sort_desc(std::vector<uint32_t>& v, size_t b, size_t e, uint32_t max_key, KeyFn key,
                        std::vector<uint32_t>& tmp, std::vector<uint32_t>& cnt) {
    if (e - b < 2) return;
    cnt.assign((size_t)max_key + 2, 0);
    for (size_t i = b; i < e; ++i) cnt[max_key - key(v[i]) + 1]++;
    for (uint32_t k = 0; k <= max_key; ++k) cnt[k + 1] += cnt[k];
    tmp.resize(e - b);
    for (size_t i = b; i < e; ++i) tmp[cnt[max_key - key(v[i])]++] = v[i];
    std::copy(tmp.begin(), tmp.end(), v.begin() + b);
}
}  // namespace

LayoutMode pick_layout_mode(uint32_t N_pad, uint32_t capacity_agents, uint32_t n_replicas) {
    if (N_pad <= capacity_agents) return kLayoutGlobal;          // a population that fits is staged whole
    if (const char* e = getenv("MVB_LAYOUT")) {
        if (!strcmp(e, "local")) return kLayoutLocal;
        if (!strcmp(e, "global")) return kLayoutGlobal;
    }
    // Measured on B200 (S=4, H=10, BA 5+10), % of the HBM roofline, GLOBAL vs LOCAL:
    //   one population:  1M 49.7 / 54.4   2M 29.7 / 41.9   3M 23.4 / 36.9   6M 17.9 / 30.7   8M 16.6 / 27.8   50M (H=64) 12.8 / 24.4
    //   8 x 1M 81 / 64,  64 x 1M 83.5 / 64.4: LOCAL's parts are neighbourhood-pure, so a launch of many replicas has several
    //   times more of them, and at 1M agents only 4 % of the links are cold to begin with.
    if (n_replicas <= 4) return kLayoutLocal;
    return (uint64_t)N_pad * 2 > (uint64_t)capacity_agents * 3 ? kLayoutLocal : kLayoutGlobal;
}

bool pick_push_mode(uint32_t N_pad, uint32_t n_replicas) {
    if (const char* e = getenv("MVB_PUSH")) return atoi(e) != 0;
    // one large population per handle (BASELINE config 4).  Measured on one B200, ms per weekday pull / push (S=4, H=10, BA
    // 5+10, profiles/r02_push_threshold.txt): 2M agents 0.1135 / 0.1108, 3M 0.196 / 0.167, 4M 0.284 / 0.235, 5M 0.369 /
    // 0.282, 8M 0.698 / 0.487, 50M 8.04 / 4.28; at 1M agents only 4 % of the links are cold and the extra launch costs more
    // than it saves.  With several replicas per launch the populations are small enough for the pull path.
    return n_replicas == 1 && N_pad > 1800u * 1000u;
}
uint32_t push_chunk_agents(uint64_t own_slots, uint32_t grid) {
    uint32_t cap = kPushChunkMax;
    if (const char* e = getenv("MVB_PUSH_CHUNK")) cap = std::min<uint32_t>(kPushChunkMax, std::max(32, atoi(e)) / 32 * 32);   // tests: several chunks on small populations
    if (!own_slots) return 32;
    if (own_slots == ~0ull) return cap;
    uint64_t rounds = 1;
    while ((own_slots + rounds * grid - 1) / (rounds * grid) > cap) ++rounds;
    const uint64_t c = (own_slots + rounds * grid - 1) / (rounds * grid);
    return (uint32_t)std::min<uint64_t>(cap, (c + 31) / 32 * 32);
}

int layout_begin(const mvb_population& pop, uint32_t H, uint32_t capacity_agents, GraphLayout& L, LayoutWork& W, uint32_t n_replicas) {
    const uint32_t N = pop.n_agents;
    const uint64_t* srp = pop.social_rowptr;
    const uint64_t* nrp = pop.neighbour_rowptr;
    L = GraphLayout();
    W = LayoutWork();
    L.N = N;
    L.Es = srp[N];
    L.En = nrp[N];
    if (capacity_agents < 64 || capacity_agents % 64) { set_error("internal: bad shared-memory capacity %u", capacity_agents); return MVB_ERR_ARG; }

    // ---- degrees, hubs, neighbourhood groups -------------------------------------------------
    std::vector<uint64_t>& d = W.d;
    d.resize(N);
    // Non-hub agents are bucketed by (neighbourhood, degree): one counting pass, one scatter pass in id order, and every
    // group comes out in (degree desc, id asc) order.  With very many neighbourhoods the bucket table would not stay
    // in cache; then the groups are filled in id order and sorted one by one.
    const bool bucketed = (uint64_t)H * (kHubDegree + 1) <= (1u << 18);
    std::vector<uint32_t> bucket(bucketed ? (size_t)H * (kHubDegree + 1) : 0, 0), gsize(H, 0);
    for (uint32_t i = 0; i < N; ++i) {
        const uint64_t ds = srp[i + 1] - srp[i], dn = nrp[i + 1] - nrp[i];
        if (ds > 0xFFFFFFFFull || dn > 0xFFFFFFFFull) { set_error("agent %u has more than 2^32 links", i); return MVB_ERR_ARG; }
        d[i] = ds + dn;
        if (d[i] <= kHubDegree) {
            gsize[pop.neighbourhood[i]]++;
            if (bucketed) bucket[(size_t)pop.neighbourhood[i] * (kHubDegree + 1) + (kHubDegree - (uint32_t)d[i])]++;
        }
    }
    std::vector<uint32_t>& hubs = W.hubs;
    std::vector<std::vector<uint32_t>>& groups = W.groups;
    groups.resize(H);
    if (bucketed) {
        for (uint32_t h = 0; h < H; ++h) {               // bucket counts -> first position of the bucket in its group
            groups[h].resize(gsize[h]);
            uint32_t at = 0;
            for (uint32_t k = 0; k <= kHubDegree; ++k) { uint32_t& b = bucket[(size_t)h * (kHubDegree + 1) + k]; const uint32_t c = b; b = at; at += c; }
        }
        for (uint32_t i = 0; i < N; ++i) {
            if (d[i] > kHubDegree) hubs.push_back(i);
            else { const uint32_t h = pop.neighbourhood[i]; groups[h][bucket[(size_t)h * (kHubDegree + 1) + (kHubDegree - (uint32_t)d[i])]++] = i; }
        }
    } else {
        for (uint32_t h = 0; h < H; ++h) groups[h].reserve(gsize[h]);
        for (uint32_t i = 0; i < N; ++i) {
            if (d[i] > kHubDegree) hubs.push_back(i);
            else groups[pop.neighbourhood[i]].push_back(i);
        }
        std::vector<uint32_t> tmp, cnt;
        // ids in increasing order, degrees <= kHubDegree: a stable counting sort gives (degree desc, id asc)
        for (auto& g : groups) counting_sort_desc(g, 0, g.size(), kHubDegree, [&](uint32_t a) { return (uint32_t)d[a]; }, tmp, cnt);
    }
    auto by_degree = [&](uint32_t a, uint32_t b) { return d[a] != d[b] ? d[a] > d[b] : a < b; };
    std::sort(hubs.begin(), hubs.end(), by_degree);
    const uint32_t hub_pad = W.hub_pad = round_up((uint32_t)hubs.size(), 64);
    {   // Deal the hubs (biggest first) round the hub slices like cards, so that every slice - the unit in which hub
        // work is handed to the parts of a launch - holds about the same number of links.  With the 32 biggest hubs
        // in one slice, the part that drew it was the straggler of every launch with few replicas.
        const uint32_t n_hs = hub_pad / 32;
        std::vector<uint32_t> dealt(hub_pad, kInvalid);
        for (uint32_t j = 0; j < hubs.size(); ++j) dealt[(j % n_hs) * 32 + j / n_hs] = hubs[j];
        hubs.swap(dealt);                               // from here on: hubs[slot], kInvalid where a slice has a free place
    }

    // ---- hot set: all hubs + the highest-degree agents of every group, as many as fit ---------
    std::vector<uint32_t>& take = W.take;               // hot agents of group h (a prefix in degree order)
    std::vector<uint32_t>& gpad = W.gpad;
    take.assign(H, 0);
    gpad.resize(H);
    uint64_t total = hub_pad;
    for (uint32_t h = 0; h < H; ++h) { gpad[h] = round_up((uint32_t)groups[h].size(), 64); total += gpad[h]; }
    if (total >= (1ull << 29)) { set_error("too many agents for the word index of a link code"); return MVB_ERR_ARG; }
    L.N_pad = (uint32_t)total;
    const bool all_hot = W.all_hot = L.N_pad <= capacity_agents;
    W.capacity_agents = capacity_agents;
    W.mode = L.mode = pick_layout_mode(L.N_pad, capacity_agents, n_replicas);
    L.push = !all_hot && pick_push_mode(L.N_pad, n_replicas);
    W.staged_hub_slots = all_hot ? hub_pad : std::min(hub_pad, capacity_agents);
    W.take_common.assign(H, 0);
    if (W.mode == kLayoutLocal) {
        // hubs get at most half of shared memory
        W.staged_hub_slots = std::min(hub_pad, capacity_agents / 2 / 64 * 64);
        const uint32_t room = capacity_agents - W.staged_hub_slots;
        uint32_t max_gpad = 0;
        for (uint32_t h = 0; h < H; ++h) max_gpad = std::max(max_gpad, gpad[h]);
        if (max_gpad <= room) {
            // every neighbourhood fits: its blocks stage all of it, and every block the longest degree-threshold prefix of
            // all groups such that   sum_h common[h] + max_h (gpad[h] - common[h]) <= room
            auto need_for = [&](uint64_t theta, std::vector<uint32_t>& out) {
                uint64_t sum = 0, tail = 0;
                for (uint32_t h = 0; h < H; ++h) {
                    const auto& g = groups[h];
                    const uint32_t c = (uint32_t)(std::partition_point(g.begin(), g.end(), [&](uint32_t a) { return d[a] >= theta; }) - g.begin());
                    out[h] = std::min(round_up(c, 64), (uint32_t)g.size() / 64 * 64);
                    sum += out[h];
                    tail = std::max<uint64_t>(tail, gpad[h] - out[h]);
                }
                return sum + tail;
            };
            std::vector<uint32_t> tmp(H);
            uint64_t lo = 0, hi = kHubDegree + 1;        // theta in (lo, hi]; hi always fits (no common prefix at all)
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) / 2;
                if (need_for(mid, tmp) <= room) hi = mid; else lo = mid;
            }
            need_for(hi, W.take_common);
            for (uint32_t h = 0; h < H; ++h) take[h] = (uint32_t)groups[h].size();
        } else {
            for (uint32_t h = 0; h < H; ++h) take[h] = std::min((uint32_t)groups[h].size(), room);
        }
    } else if (all_hot) {
        for (uint32_t h = 0; h < H; ++h) take[h] = (uint32_t)groups[h].size();
    } else {
        // largest degree threshold theta such that {hubs} U {d >= theta} (rounded to 64 per group) fits
        const uint32_t budget = capacity_agents > hub_pad ? capacity_agents - hub_pad : 0;
        auto count_for = [&](uint64_t theta, std::vector<uint32_t>& out) {
            uint64_t sum = 0;
            for (uint32_t h = 0; h < H; ++h) {
                const auto& g = groups[h];
                const uint32_t c = (uint32_t)(std::partition_point(g.begin(), g.end(), [&](uint32_t a) { return d[a] >= theta; }) - g.begin());
                out[h] = std::min(round_up(c, 64), (uint32_t)g.size() / 64 * 64);
                sum += out[h];
            }
            return sum;
        };
        std::vector<uint32_t> tmp(H);
        uint64_t lo = 0, hi = kHubDegree + 1;            // theta in (lo, hi]; hi always fits (nothing taken)
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) / 2;
            if (count_for(mid, tmp) <= budget) hi = mid; else lo = mid;
        }
        uint64_t used = count_for(hi, take);
        for (uint32_t h = 0; h < H && used + 64 <= budget; ++h) {   // spend the rest on the next degree class
            const uint32_t room = (uint32_t)groups[h].size() / 64 * 64 - take[h];
            const uint32_t extra = (uint32_t)std::min<uint64_t>(room, (budget - used) / 64 * 64);
            take[h] += extra; used += extra;
        }
    }
    std::vector<uint8_t>& is_hot = W.is_hot;
    is_hot.assign(N, 0);
    for (uint32_t k = 0; k < W.staged_hub_slots; ++k)     // hub slots that get staged
        if (hubs[k] != kInvalid) is_hot[hubs[k]] = 1;
    for (uint32_t h = 0; h < H; ++h)
        for (uint32_t k = 0; k < take[h]; ++k) is_hot[groups[h][k]] = (W.mode == kLayoutLocal && k >= W.take_common[h]) ? 2 : 1;
    return MVB_OK;
}

void layout_count_hot(const mvb_population& pop, LayoutWork& W) {
    const uint32_t N = pop.n_agents;
    W.nsh_own.assign(N, 0);
    W.nnh_own.assign(N, 0);
    for (uint32_t i = 0; i < N; ++i) {
        uint32_t a = 0, b = 0;
        const bool local_ok = W.d[i] <= kHubDegree;            // LOCAL: a hub source only sees the staged hubs
        auto hot = [&](uint32_t t) { const uint8_t c = W.is_hot[t]; return c == 1 || (c == 2 && local_ok && pop.neighbourhood[t] == pop.neighbourhood[i]); };
        for (uint64_t x = pop.social_rowptr[i]; x < pop.social_rowptr[i + 1]; ++x) a += hot(pop.social_col[x]);
        for (uint64_t x = pop.neighbour_rowptr[i]; x < pop.neighbour_rowptr[i + 1]; ++x) b += hot(pop.neighbour_col[x]);
        W.nsh_own[i] = a; W.nnh_own[i] = b;
    }
    W.nsh = W.nsh_own.data();
    W.nnh = W.nnh_own.data();
}

int layout_finish(const mvb_population& pop, LayoutWork& W, GraphLayout& L, bool want_tcode) {
    const uint32_t N = pop.n_agents, H = (uint32_t)W.groups.size();
    const uint64_t* srp = pop.social_rowptr;
    const uint64_t* nrp = pop.neighbour_rowptr;
    const uint32_t* nsh = W.nsh;
    const uint32_t* nnh = W.nnh;
    std::vector<uint64_t>& d = W.d;
    std::vector<uint32_t>& hubs = W.hubs;
    std::vector<std::vector<uint32_t>>& groups = W.groups;
    const std::vector<uint32_t>& take = W.take;
    const uint32_t hub_pad = W.hub_pad;

    // hot | cold << 16 list lengths of the agents that are not hubs (both <= kHubDegree)
    std::vector<uint32_t> hc(N, 0);
    for (uint32_t i = 0; i < N; ++i)
        if (d[i] <= kHubDegree) { const uint32_t hl = nsh[i] + nnh[i]; hc[i] = hl | ((uint32_t)d[i] - hl) << 16; }
    // ---- internal order ----------------------------------------------------------------------
    std::vector<uint32_t> gstart(H);
    {
        uint64_t at = hub_pad;
        for (uint32_t h = 0; h < H; ++h) { gstart[h] = (uint32_t)at; at += W.gpad[h]; }
    }
    L.n_hub_slices = hub_pad / 32;
    L.n_sell_slices = (L.N_pad - hub_pad) / 32;
    L.int2ext.assign(L.N_pad, kInvalid);
    for (uint32_t k = 0; k < hubs.size(); ++k) L.int2ext[k] = hubs[k];            // hub_pad entries, kInvalid = free place
    // Group by group: order (hot length desc, degree desc, id asc) — the groups are in (degree desc, id asc) order
    // already, so one more stable pass by hot length does it, separately for the hot part and the cold part — then
    // the group's SELL slices.  The list lengths travel with the ids through the sort (high half of a u64), so the
    // only random reads are one gather per agent; one group's worth of scratch is reused for all groups.
    L.sell_off.assign(L.n_sell_slices, 0);
    L.sell_K.assign(L.n_sell_slices, 0);
    uint64_t off = 0;
    {
        std::vector<uint64_t> kv, tmp;
        std::vector<uint32_t> cnt;
        uint32_t s = 0;
        for (uint32_t h = 0; h < H; ++h) {
            const std::vector<uint32_t>& g = groups[h];
            kv.resize(g.size());
            for (size_t i = 0; i < g.size(); ++i) kv[i] = (uint64_t)hc[g[i]] << 32 | g[i];
            auto sort_part = [&](size_t b, size_t e) {          // stable counting sort by hot length, descending
                if (e - b < 2) return;
                cnt.assign(kHubDegree + 2, 0);
                for (size_t i = b; i < e; ++i) cnt[kHubDegree - (uint32_t)((kv[i] >> 32) & 0xFFFFu) + 1]++;
                for (uint32_t k = 0; k <= kHubDegree; ++k) cnt[k + 1] += cnt[k];
                tmp.resize(e - b);
                for (size_t i = b; i < e; ++i) tmp[cnt[kHubDegree - (uint32_t)((kv[i] >> 32) & 0xFFFFu)]++] = kv[i];
                std::copy(tmp.begin(), tmp.end(), kv.begin() + b);
            };
            sort_part(0, W.take_common[h]);                    // (empty unless the LOCAL layout has a common prefix)
            sort_part(W.take_common[h], take[h]);
            sort_part(take[h], g.size());
            for (size_t i = 0; i < kv.size(); ++i) L.int2ext[gstart[h] + i] = (uint32_t)kv[i];
            for (size_t b = 0; b < W.gpad[h]; b += 32, ++s) {   // the group's slices: its agents, 32 at a time, then padding
                uint32_t K = 0, Kc = 0;
                for (size_t i = b; i < std::min(b + 32, kv.size()); ++i) {
                    const uint32_t key = (uint32_t)(kv[i] >> 32);
                    K = std::max(K, key & 0xFFFFu);
                    Kc = std::max(Kc, key >> 16);
                }
                L.sell_off[s] = off;
                L.sell_K[s] = K | Kc << 16;
                off += (uint64_t)32 * (K + (L.push ? 0u : Kc));     // push mode: the block holds hot rows only
            }
        }
    }

    // ---- staged ranges: which pieces of the packed vector go to shared memory, and where
    uint32_t staged = 0;
    auto add_range = [&](uint32_t start, uint32_t count) {       // both multiples of 64
        if (!count) return;
        L.ranges.push_back(StageRange{start / 16, count / 16, staged / 16, 0});
        staged += count;
    };
    L.nb_slice_begin.resize((size_t)H + 1);
    for (uint32_t h = 0; h < H; ++h) L.nb_slice_begin[h] = (gstart[h] - hub_pad) / 32;
    L.nb_slice_begin[H] = L.n_sell_slices;
    if (W.mode == kLayoutLocal) {
        // [0] the hub region, [1 + h] the common prefix of group h; then [1 + H + h] the piece of neighbourhood h that only its
        // own blocks stage, every one of them placed right behind the common set
        add_range(0, W.staged_hub_slots);
        if (L.ranges.empty()) L.ranges.push_back(StageRange{0, 0, 0, 0});
        for (uint32_t h = 0; h < H; ++h) {
            L.ranges.push_back(StageRange{gstart[h] / 16, W.take_common[h] / 16, staged / 16, 0});
            staged += W.take_common[h];
        }
        const uint32_t common = staged;
        uint32_t longest = 0;
        for (uint32_t h = 0; h < H; ++h) {
            const uint32_t words = ((take[h] + 63) / 64 * 64 - W.take_common[h]) / 16;
            L.ranges.push_back(StageRange{(gstart[h] + W.take_common[h]) / 16, words, common / 16, 0});
            longest = std::max(longest, words);
        }
        L.n_common_ranges = 1 + H;
        staged = common + longest * 16;
    } else if (W.all_hot) {
        add_range(0, L.N_pad);
    } else {
        add_range(0, std::min(hub_pad, W.capacity_agents));
        for (uint32_t h = 0; h < H; ++h) add_range(gstart[h], take[h]);     // take[h] is a multiple of 64 here
    }
    L.hot_words = staged / 16;
    if (W.mode != kLayoutLocal) L.n_common_ranges = (uint32_t)L.ranges.size();

    // ---- link codes ---------------------------------------------------------------------------
    // code of every possible target; the arrays of codes are filled on the device (fill_kernels.cuh).  With
    // want_tcode == false the caller computes this table on the device as well (tcode_kernel, same rule).
    if (want_tcode && W.mode == kLayoutLocal) {
        L.tcode.resize(N); L.tcold.resize(N);
        for (uint32_t t = 0; t < L.N_pad; ++t) {
            const uint32_t e = L.int2ext[t];
            if (e == kInvalid) continue;
            L.tcold[e] = (t >> 4) << 7 | (30u - 2u * (t & 15u));
            uint32_t pos = t;                                    // a staged hub sits at its own slot in the hub region
            if (W.is_hot[e] && t >= hub_pad) {
                const StageRange& r = L.ranges[(W.is_hot[e] == 2 ? 1 + H : 1) + pop.neighbourhood[e]];
                pos = r.dst_word * 16 + (t - r.src_word * 16);
            }
            L.tcode[e] = W.is_hot[e] ? (((pos >> 4) << 7 | (30u - 2u * (pos & 15u))) | 1u) : L.tcold[e];
        }
    } else if (want_tcode) {
        L.tcode.resize(N);
        for (const StageRange& r : L.ranges)
            for (uint32_t t = r.src_word * 16; t < (r.src_word + r.n_words) * 16; ++t) {
                const uint32_t e = L.int2ext[t];
                if (e == kInvalid || !W.is_hot[e]) continue;
                const uint32_t pos = r.dst_word * 16 + (t - r.src_word * 16);
                L.tcode[e] = ((pos >> 4) << 7 | (30u - 2u * (pos & 15u))) | 1u;
            }
        for (uint32_t t = 0; t < L.N_pad; ++t) {
            const uint32_t e = L.int2ext[t];
            if (e == kInvalid || W.is_hot[e]) continue;
            L.tcode[e] = (t >> 4) << 7 | (30u - 2u * (t & 15u));
        }
    }
    // hub blocks: lane l of a block holds links [l*K, (l+1)*K) of what is left of the list
    const uint32_t n_hub_rows = L.n_hub_slices * 32;
    L.hub_off.assign(n_hub_rows, 0);
    L.hub_deg.assign((size_t)4 * n_hub_rows, 0);
    uint64_t hoff = 0;
    L.hub_units.clear();
    L.hub_unit_begin.assign((size_t)L.n_hub_slices + 1, 0);
    for (uint32_t k = 0; k < n_hub_rows; ++k) {
        L.hub_off[k] = hoff;
        if ((k & 31u) == 0) L.hub_unit_begin[k >> 5] = (uint32_t)L.hub_units.size();
        if (hubs[k] == kInvalid) continue;
        const uint32_t e = hubs[k];
        const uint32_t ds = (uint32_t)(srp[e + 1] - srp[e]), dn = (uint32_t)(nrp[e + 1] - nrp[e]);
        uint32_t* dg = &L.hub_deg[(size_t)4 * k];
        dg[0] = nsh[e]; dg[1] = nnh[e]; dg[2] = ds - nsh[e]; dg[3] = dn - nnh[e];
        const uint32_t nh = nsh[e] + nnh[e], nc = (uint32_t)d[e] - nh;
        // what is left of a list at the start of a unit, as far as one unit can see (32 lanes x 32 rows)
        auto left = [](uint32_t total, uint32_t done) { return total > done ? std::min(total - done, 1024u) : 0u; };
        for (uint32_t dh = 0, dc = 0; dh < nh || dc < nc;) {
            uint32_t K, Kc;
            hub_segment(nh - std::min(dh, nh), nc - std::min(dc, nc), K, Kc);
            L.hub_units.push_back(uint4x{(uint32_t)(hoff / 32), k, K | Kc << 8 | left(dg[0], dh) << 16, left(dg[2], dc) | left(nc, dc) << 16});
            hoff += (uint64_t)32 * (K + Kc);
            dh += 32 * K; dc += 32 * Kc;
        }
    }
    L.hub_unit_begin[L.n_hub_slices] = (uint32_t)L.hub_units.size();
    L.hub_units.push_back(uint4x{0, 0, 0, 0});            // slack
    if (hoff / 32 >= (1ull << 32)) { set_error("hub link lists too long for a 32-bit row offset"); return MVB_ERR_ARG; }
    L.hub_codes_len = hoff + 256;                        // + slack for the row prefetch
    L.sell_codes_len = off + 256;                        // + slack: the kernel's row prefetch may run 2 rows past a slice
    // one 16-byte record per slice; 64 zero records of slack let the kernel prefetch past the end
    L.sell_meta.assign((size_t)L.n_sell_slices + 64, uint4x{0, 0, 0, 0});
    for (uint32_t s = 0; s < L.n_sell_slices; ++s)
        L.sell_meta[s] = uint4x{(uint32_t)L.sell_off[s], (uint32_t)(L.sell_off[s] >> 32), L.push ? L.sell_K[s] & 0xFFFFu : L.sell_K[s], 0};
    return MVB_OK;
}

int build_graph_layout(const mvb_population& pop, uint32_t H, uint32_t capacity_agents, GraphLayout& L, uint32_t n_replicas) {
    LayoutWork W;
    if (int rc = layout_begin(pop, H, capacity_agents, L, W, n_replicas)) return rc;
    layout_count_hot(pop, W);
    return layout_finish(pop, W, L, true);
}

namespace {
// Work model, in units of one shared-memory gather.  A link read from L2 costs about 2 of them when few links are cold
// (GLOBAL layout on populations just above the shared-memory capacity) and about 8 in the LOCAL layout, where a third of
// the links are cold gathers and the L2 gather path (~0.4 per clock per SM against ~3.4 from shared memory) sets the
// pace; + per-agent work; a hub's links are walked in a phase of their own, between two block barriers (x 2.5 GLOBAL,
// calibrated on 8M agents over 2 B200s; x 1.5 LOCAL).  LOCAL values from tools/partition_balance.py (every rank's share
// of 50M agents timed on one GPU for a grid of weights): 8 ranks, H = 10: 1.16-1.20 ms per rank with these values
// (0.83-1.71 with the GLOBAL ones), H = 64: 0.73-0.79 ms (0.66-1.10).  The same model hands every neighbourhood its share
// of a rank's parts (create.cuh): a neighbourhood's last, cold-heavy slices cost several times its first ones.
struct WorkModel { uint64_t cold, per_agent, hub10, hub_cold; };
WorkModel work_model(const GraphLayout& L) {
    WorkModel m = L.mode == kLayoutLocal ? WorkModel{8, 40, 15, 8} : WorkModel{2, 64, 25, 2};
    // push mode: a SELL agent's cold link costs a sorted gather + a shared-memory reduction (~1.2 per clock per SM), a
    // hub's cold links stay on the pull path (hub_cold), and with the cold links cheap the per-agent work weighs more
    // (tools/partition_balance.py, 50M agents over 8 ranks, one round of parts: 0.45-0.50 ms per rank at H = 64 and
    // 0.50-0.59 at H = 10 with these values; 0.45-0.55 and 0.54-0.68 with {4, 40, 15})
    if (L.push) { m.cold = 3; m.per_agent = 80; m.hub10 = 22; }
    if (const char* e = getenv("MVB_PART_HUBCOLD")) m.hub_cold = (uint64_t)atoi(e);
    if (const char* e = getenv("MVB_PART_COLD")) m.cold = (uint64_t)atoi(e);          // calibration experiments
    if (const char* e = getenv("MVB_PART_AGENT")) m.per_agent = (uint64_t)atoi(e);
    if (const char* e = getenv("MVB_PART_HUB10")) m.hub10 = (uint64_t)atoi(e);
    return m;
}
}  // namespace

uint64_t sell_slice_work(const GraphLayout& L, uint32_t s) {
    const WorkModel m = work_model(L);
    return 32ull * ((L.sell_K[s] & 0xFFFFu) + m.cold * (L.sell_K[s] >> 16) + m.per_agent);
}

void partition_slices(const GraphLayout& L, uint32_t world, std::vector<uint32_t>& bounds) {
    const uint32_t n = L.n_hub_slices + L.n_sell_slices;
    std::vector<uint64_t> work(n);
    const WorkModel m = work_model(L);
    for (uint32_t hs = 0; hs < L.n_hub_slices; ++hs) {
        uint64_t w = 0;
        for (uint32_t k = 0; k < 32; ++k) {
            const uint32_t* dg = &L.hub_deg[(size_t)4 * (hs * 32 + k)];
            w += m.hub10 * ((uint64_t)dg[0] + dg[1] + m.hub_cold * (dg[2] + dg[3])) / 10 + m.per_agent;
        }
        work[hs] = w;
    }
    for (uint32_t s = 0; s < L.n_sell_slices; ++s) work[L.n_hub_slices + s] = sell_slice_work(L, s);
    uint64_t total = 0;
    for (uint64_t w : work) total += w;
    bounds.assign(world + 1, n);
    bounds[0] = 0;
    uint64_t acc = 0;
    uint32_t r = 1;
    for (uint32_t i = 0; i < n && r < world; ++i) {
        acc += work[i];
        while (r < world && acc * world >= total * r) bounds[r++] = i + 1;
    }
}

}  // namespace mvb

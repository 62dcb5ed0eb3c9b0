// layout_invariants.cpp — CPU check of the device layout (motivate_b200/csrc/layout.h) that mvb_create* and
// mvb_partition_plan build from a population: compiled and run by tests/test_layout.py against libmotivate_b200.so.
// usage: layout_invariants N S H capacity_agents seed   -> prints "ok ..." or "FAILED: ..." (exit 1)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "layout.h"

using namespace mvb;

#define CHECK(cond, ...) do { if (!(cond)) { printf("FAILED: " __VA_ARGS__); printf("  [%s]\n", #cond); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 6) { printf("usage\n"); return 2; }
    const uint32_t N = (uint32_t)atol(argv[1]), S = (uint32_t)atol(argv[2]), H = (uint32_t)atol(argv[3]);
    const uint32_t cap = (uint32_t)atol(argv[4]);
    mvb_synth_spec sp{};
    sp.n_agents = N; sp.n_subcultures = S; sp.n_neighbourhoods = H; sp.social_links_min = 5; sp.neighbour_links_min = 10;
    sp.n_cars = N / 2; sp.n_bikes = N / 2; sp.seed = (uint64_t)atol(argv[5]); sp.n_gmm = 1;
    sp.gmm[0][0] = 9845.09; sp.gmm[0][1] = 3691.81; sp.gmm[0][2] = 1.0;
    mvb_synth* syn = nullptr;
    CHECK(mvb_synth_create(&sp, &syn) == MVB_OK, "synth: %s", mvb_last_error());
    const mvb_population& pop = *mvb_synth_population(syn);
    GraphLayout L;
    CHECK(build_graph_layout(pop, H, cap, L) == MVB_OK, "layout: %s", mvb_last_error());

    // ---- internal order: a permutation of the agents plus padding, regions padded to 64
    CHECK(L.N == N && L.N_pad % 64 == 0 && L.N_pad >= N, "sizes");
    CHECK(L.int2ext.size() == L.N_pad && (L.n_hub_slices + L.n_sell_slices) * 32 == L.N_pad, "slices");
    std::vector<uint32_t> slot_of(N, kInvalid);
    for (uint32_t s = 0; s < L.N_pad; ++s) {
        const uint32_t e = L.int2ext[s];
        if (e == kInvalid) continue;
        CHECK(e < N && slot_of[e] == kInvalid, "agent %u placed twice", e);
        slot_of[e] = s;
    }
    for (uint32_t e = 0; e < N; ++e) CHECK(slot_of[e] != kInvalid, "agent %u not placed", e);
    auto deg = [&](uint32_t e) { return (uint64_t)(pop.social_rowptr[e + 1] - pop.social_rowptr[e]) + (pop.neighbour_rowptr[e + 1] - pop.neighbour_rowptr[e]); };
    // hubs (more than kHubDegree links) fill the hub slices and nothing else does
    uint32_t n_hubs = 0;
    for (uint32_t e = 0; e < N; ++e) {
        const bool hub = deg(e) > kHubDegree;
        n_hubs += hub;
        CHECK(hub == (slot_of[e] < L.n_hub_slices * 32), "agent %u (degree %llu) on the wrong side of the hub boundary", e, (unsigned long long)deg(e));
    }
    CHECK(L.n_hub_slices * 32 >= n_hubs && L.n_hub_slices * 32 < n_hubs + 64, "hub padding");
    // a SELL slice never mixes neighbourhoods, and neighbourhoods follow each other in order
    uint32_t last_nb = 0;
    for (uint32_t s = 0; s < L.n_sell_slices; ++s) {
        uint32_t nb = kInvalid;
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t e = L.int2ext[(L.n_hub_slices + s) * 32 + l];
            if (e == kInvalid) continue;
            if (nb == kInvalid) nb = pop.neighbourhood[e];
            CHECK(pop.neighbourhood[e] == nb, "slice %u mixes neighbourhoods", s);
        }
        if (nb != kInvalid) { CHECK(nb >= last_nb, "neighbourhood order"); last_nb = nb; }
    }

    // ---- staged ranges.  GLOBAL: disjoint, increasing, inside the capacity; "hot" = inside a range.
    // LOCAL: [0] = the hub region, [1 + h] = the prefix of neighbourhood h placed right behind it; a block stages [0] and one of the others.
    std::vector<uint8_t> hot_slot(L.N_pad, 0);            // 1 = staged for every block, 2 = for blocks of its own neighbourhood
    const bool local = L.mode == kLayoutLocal;
    if (local) {
        // [0] hub region, [1 + h] common prefix of group h (staged by every block), [1 + H + h] the piece only blocks of h stage
        CHECK(L.ranges.size() == 2 * (size_t)H + 1 && L.n_common_ranges == H + 1 && L.ranges[0].src_word == 0 && L.ranges[0].dst_word == 0, "local ranges");
        CHECK(L.ranges[0].n_words * 16 <= L.n_hub_slices * 32, "hub region");
        for (uint32_t t = 0; t < L.ranges[0].n_words * 16; ++t) hot_slot[t] = 1;
        uint32_t common = L.ranges[0].n_words, longest = 0;
        for (uint32_t h = 0; h < H; ++h) {
            const StageRange& r = L.ranges[1 + h];
            CHECK(r.n_words % 4 == 0 && r.dst_word == common, "common range geometry");
            CHECK((r.src_word * 16 - L.n_hub_slices * 32) / 32 == L.nb_slice_begin[h] || r.n_words == 0, "common range start");
            for (uint32_t t = r.src_word * 16; t < (r.src_word + r.n_words) * 16; ++t) { CHECK(L.int2ext[t] != kInvalid, "padding in a common prefix"); hot_slot[t] = 1; }
            common += r.n_words;
        }
        for (uint32_t h = 0; h < H; ++h) {
            const StageRange& r = L.ranges[1 + H + h];
            CHECK(r.n_words % 4 == 0 && r.dst_word == common, "own range geometry");
            CHECK(((uint64_t)common + r.n_words) * 16 <= cap, "neighbourhood %u stages more than fits", h);
            CHECK(r.src_word == L.ranges[1 + h].src_word + L.ranges[1 + h].n_words, "own piece follows the common prefix");
            for (uint32_t t = r.src_word * 16; t < (r.src_word + r.n_words) * 16; ++t) {
                CHECK(t < L.N_pad && (t - L.n_hub_slices * 32) / 32 < L.nb_slice_begin[h + 1], "own range end");
                if (L.int2ext[t] != kInvalid) hot_slot[t] = 2;
            }
            longest = std::max(longest, r.n_words);
        }
        CHECK(L.hot_words == common + longest, "hot_words");
    } else {
        uint32_t staged = 0, prev_end = 0;
        for (const StageRange& r : L.ranges) {
            CHECK(r.n_words % 4 == 0 && r.dst_word == staged && r.src_word >= prev_end, "range geometry");
            for (uint32_t t = r.src_word * 16; t < (r.src_word + r.n_words) * 16; ++t) { CHECK(t < L.N_pad, "range end"); hot_slot[t] = 1; }
            staged += r.n_words; prev_end = r.src_word + r.n_words;
        }
        CHECK(staged == L.hot_words && (uint64_t)staged * 16 <= cap, "staged %u words for a capacity of %u agents", staged, cap);
        if (L.N_pad <= cap) CHECK((uint64_t)staged * 16 == L.N_pad, "everything fits but not everything is staged");
    }
    CHECK(L.nb_slice_begin.size() == (size_t)H + 1 && L.nb_slice_begin[H] == L.n_sell_slices, "nb_slice_begin");

    // ---- link codes of the targets (host table; mvb_create* computes the same on the device)
    CHECK(L.tcode.size() == N && (!local || L.tcold.size() == N), "tcode");
    uint32_t n_hot = 0;
    for (uint32_t e = 0; e < N; ++e) {
        const uint32_t c = L.tcode[e], hot = c & 1u, word = c >> 7, shift = c & 0x1Eu, s = slot_of[e];
        CHECK(hot == (hot_slot[s] != 0), "hot flag of agent %u", e);
        n_hot += hot;
        uint32_t pos = s;
        if (hot && !local) for (const StageRange& r : L.ranges) if (s >= r.src_word * 16 && s < (r.src_word + r.n_words) * 16) pos = r.dst_word * 16 + (s - r.src_word * 16);
        if (local && hot && s >= L.n_hub_slices * 32) { const StageRange& r = L.ranges[(hot_slot[s] == 2 ? 1 + H : 1) + pop.neighbourhood[e]]; pos = r.dst_word * 16 + (s - r.src_word * 16); }
        CHECK(word == pos / 16 && shift == 30 - 2 * (pos % 16), "code of agent %u: word %u shift %u for position %u", e, word, shift, pos);
        if (local) CHECK(L.tcold[e] == ((s >> 4) << 7 | (30u - 2u * (s & 15u))), "cold code of agent %u", e);
    }

    // ---- SELL slices: K / Kc are the longest hot / cold list of the slice, offsets are the running sum
    std::vector<uint32_t> n_hot_links(N, 0);
    for (uint32_t e = 0; e < N; ++e) {
        // a link is hot if its target is staged for every block, or for the blocks of the source's own neighbourhood (not for hub sources)
        auto hot = [&](uint32_t t) { const uint8_t k = hot_slot[slot_of[t]]; return k == 1 || (k == 2 && deg(e) <= kHubDegree && pop.neighbourhood[t] == pop.neighbourhood[e]); };
        for (uint64_t x = pop.social_rowptr[e]; x < pop.social_rowptr[e + 1]; ++x) n_hot_links[e] += hot(pop.social_col[x]);
        for (uint64_t x = pop.neighbour_rowptr[e]; x < pop.neighbour_rowptr[e + 1]; ++x) n_hot_links[e] += hot(pop.neighbour_col[x]);
    }
    uint64_t off = 0, rows = 0, links = 0;
    for (uint32_t s = 0; s < L.n_sell_slices; ++s) {
        uint32_t K = 0, Kc = 0;
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t e = L.int2ext[(L.n_hub_slices + s) * 32 + l];
            if (e == kInvalid) continue;
            K = std::max(K, n_hot_links[e]); Kc = std::max<uint32_t>(Kc, (uint32_t)deg(e) - n_hot_links[e]);
            links += deg(e);
        }
        CHECK(L.sell_K[s] == (K | Kc << 16) && L.sell_off[s] == off, "slice %u: K %u Kc %u off %llu", s, K, Kc, (unsigned long long)off);
        // push mode (layout.h): the block holds hot rows only and the record says so; sell_K keeps Kc for the work model
        CHECK(L.sell_meta[s].x == (uint32_t)off && L.sell_meta[s].y == (uint32_t)(off >> 32) &&
              L.sell_meta[s].z == (L.push ? L.sell_K[s] & 0xFFFFu : L.sell_K[s]), "slice record %u", s);
        off += 32ull * (K + (L.push ? 0u : Kc)); rows += K + Kc;
    }
    CHECK(L.sell_codes_len >= off, "sell_codes_len");
    const double padding = links ? (double)rows * 32 / (double)links : 1.0;

    // ---- hub units: each hub's hot and cold lists are covered exactly once, in order, by its units
    CHECK(L.hub_unit_begin.size() == (size_t)L.n_hub_slices + 1, "hub_unit_begin");
    uint64_t max_units = 0, sum_units = 0;
    for (uint32_t hs = 0; hs < L.n_hub_slices; ++hs) {
        const uint32_t ub = L.hub_unit_begin[hs], ue = L.hub_unit_begin[hs + 1];
        CHECK(ub <= ue && ue < L.hub_units.size(), "unit range");
        max_units = std::max<uint64_t>(max_units, ue - ub); sum_units += ue - ub;
        std::vector<uint32_t> done_h(32, 0), done_c(32, 0);
        uint32_t last_row = hs * 32;
        for (uint32_t u = ub; u < ue; ++u) {
            const uint4x& r = L.hub_units[u];
            const uint32_t row = r.y, K = r.z & 0xFFu, Kc = (r.z >> 8) & 0xFFu;
            CHECK(row / 32 == hs && row >= last_row, "unit %u belongs to hub row %u", u, row);
            last_row = row;
            const uint32_t* dg = &L.hub_deg[(size_t)4 * row];
            const uint32_t nh = dg[0] + dg[1], nc = dg[2] + dg[3], l = row % 32;
            CHECK(K + Kc <= kHubUnitRows && K + Kc > 0, "unit rows");
            CHECK((uint64_t)r.x * 32 >= L.hub_off[row], "unit offset");
            auto left = [](uint32_t total, uint32_t done) { return total > done ? std::min(total - done, 1024u) : 0u; };
            CHECK((r.z >> 16) == left(dg[0], done_h[l]) && (r.w & 0xFFFFu) == left(dg[2], done_c[l]) && (r.w >> 16) == left(nc, done_c[l]), "unit %u: lists left", u);
            done_h[l] += 32 * K; done_c[l] += 32 * Kc;
            (void)nh;
        }
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t e = L.int2ext[hs * 32 + l];
            const uint32_t* dg = &L.hub_deg[(size_t)4 * (hs * 32 + l)];
            if (e == kInvalid) { CHECK(dg[0] + dg[1] + dg[2] + dg[3] == 0 && done_h[l] + done_c[l] == 0, "free hub place has links"); continue; }
            CHECK(dg[0] + dg[1] == n_hot_links[e] && dg[0] + dg[1] + dg[2] + dg[3] == deg(e), "hub %u list lengths", e);
            CHECK(done_h[l] >= dg[0] + dg[1] && done_h[l] < dg[0] + dg[1] + 32 * kHubUnitRows, "hub %u: hot list covered", e);
            CHECK(done_c[l] >= dg[2] + dg[3] && done_c[l] < dg[2] + dg[3] + 32 * kHubUnitRows, "hub %u: cold list covered", e);
        }
    }
    const double hub_balance = sum_units ? (double)max_units * L.n_hub_slices / (double)sum_units : 1.0;

    // ---- agent-partitioned mode: contiguous slice ranges that cover everything
    for (uint32_t world : {1u, 2u, 3u, 8u}) {
        std::vector<uint32_t> b;
        partition_slices(L, world, b);
        CHECK(b.size() == world + 1 && b[0] == 0 && b[world] == L.n_hub_slices + L.n_sell_slices, "bounds");
        for (uint32_t r = 0; r < world; ++r) CHECK(b[r] <= b[r + 1], "bounds order");
    }
    uint64_t all_links = 0, hot_links = 0;
    for (uint32_t e = 0; e < N; ++e) { all_links += deg(e); hot_links += n_hot_links[e]; }
    printf("ok push %d N_pad %u hubs %u hot %u ranges %zu sell_padding %.3f hub_slice_balance %.2f mode %s hot_link_share %.3f\n", (int)L.push, L.N_pad, n_hubs, n_hot, L.ranges.size(), padding, hub_balance,
           local ? "local" : "global", all_links ? (double)hot_links / (double)all_links : 1.0);
    mvb_synth_destroy(syn);
    return 0;
}

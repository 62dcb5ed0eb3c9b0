// synth.cpp — seeded synthetic populations and Barabási–Albert networks (host only).
//
// Follows the distributions of the reference's setup code, which is NOT on the hot
// path and is not reproducible there (thread_rng): src/agent_generation.rs:114-325,
// src/social_network.rs:12-48, src/simulation.rs:155-189.  Own PRNG (SplitMix64).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#ifndef MVB_NO_CUDA            // oracle/Makefile builds this file alone, without CUDA, as the CPU arm's input generator
#include <cuda_runtime_api.h>
#endif

#include "mvb_internal.h"
#include "synth.h"

namespace mvb {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* last_error_cstr() { return g_err.c_str(); }

// Preferential attachment as src/social_network.rs:12-48 does it: the first `min`
// nodes form a clique (node i links to everyone already present while fewer than
// `min` nodes exist); afterwards node i makes `min` draws WITH replacement from the
// nodes present, weighted by their degree as it was BEFORE node i arrived
// (the WeightedChoice is built once per node, :30-34), duplicates kept (:37-41).
// Sampling a uniform element of the endpoint list is exactly degree-weighted.
// Emits directed pairs (a -> b) and (b -> a) for every link.
static void ba_edges(uint32_t min_links, uint32_t n, uint64_t seed, std::vector<uint32_t>& src,
                     std::vector<uint32_t>& dst) {
    SplitMix64 rng(seed);
    std::vector<uint32_t> endpoints;
    endpoints.reserve((size_t)2 * min_links * n);
    src.clear(); dst.clear();
    src.reserve((size_t)min_links * n); dst.reserve((size_t)min_links * n);
    for (uint32_t i = 0; i < n; ++i) {
        if (i < min_links) {
            for (uint32_t k = 0; k < i; ++k) {
                src.push_back(i); dst.push_back(k);
                endpoints.push_back(i); endpoints.push_back(k);
            }
        } else {
            const uint64_t len_before = endpoints.size();
            for (uint32_t d = 0; d < min_links; ++d) {
                const uint32_t id = endpoints[rng.below(len_before)];
                src.push_back(i); dst.push_back(id);
                endpoints.push_back(i); endpoints.push_back(id);
            }
        }
    }
}

}  // namespace mvb

using namespace mvb;

extern "C" {

const char* mvb_last_error(void) { return mvb::last_error_cstr(); }
int mvb_abi_version(void) { return MVB_ABI_VERSION; }
void mvb_free(void* p) { free(p); }

int mvb_synth_ba_network(uint32_t min_links, uint32_t n_nodes, uint64_t seed, uint64_t** rowptr_out,
                         uint32_t** col_out) {
    if (!rowptr_out || !col_out) { set_error("null output pointer"); return MVB_ERR_ARG; }
    if (min_links < 2 && n_nodes > min_links) {
        // number_of_social_network_links: 1 panics in the reference (total weight 0 when node 1
        // arrives, src/social_network.rs:30-34) — SURVEY.md §5 defect 5.
        set_error("min_links must be >= 2 (the reference panics for 1)");
        return MVB_ERR_DOMAIN;
    }
    std::vector<uint32_t> src, dst;
    ba_edges(min_links, n_nodes, seed, src, dst);
    uint64_t* rp = (uint64_t*)calloc((size_t)n_nodes + 1, sizeof(uint64_t));
    uint32_t* cl = (uint32_t*)malloc(sizeof(uint32_t) * std::max<size_t>(1, 2 * src.size()));
    if (!rp || !cl) { free(rp); free(cl); set_error("out of memory"); return MVB_ERR_NOMEM; }
    for (size_t e = 0; e < src.size(); ++e) { rp[(size_t)src[e] + 1]++; rp[(size_t)dst[e] + 1]++; }
    for (uint32_t i = 0; i < n_nodes; ++i) rp[i + 1] += rp[i];
    std::vector<uint64_t> cur(rp, rp + n_nodes);
    for (size_t e = 0; e < src.size(); ++e) {
        cl[cur[src[e]]++] = dst[e];
        cl[cur[dst[e]]++] = src[e];
    }
    *rowptr_out = rp; *col_out = cl;
    return MVB_OK;
}

int mvb_synth_create(const mvb_synth_spec* sp, mvb_synth** out) {
    if (!sp || !out) { set_error("null argument"); return MVB_ERR_ARG; }
    const uint32_t N = sp->n_agents, S = sp->n_subcultures, H = sp->n_neighbourhoods;
    if (N == 0 || S == 0 || S > 256 || H == 0 || H > 65536 || sp->n_gmm == 0 || sp->n_gmm > 8) {
        set_error("bad synth spec (N=%u S=%u H=%u n_gmm=%u)", N, S, H, sp->n_gmm);
        return MVB_ERR_ARG;
    }
    if (sp->social_links_min < 2 || sp->neighbour_links_min < 2) {
        set_error("link minimums must be >= 2 (the reference panics for 1)");
        return MVB_ERR_DOMAIN;
    }
    if (sp->n_cars > N || sp->n_bikes > N) {
        set_error("more cars/bikes than agents");   // sample_slice_ref would panic
        return MVB_ERR_DOMAIN;
    }
    mvb_synth* s = new (std::nothrow) mvb_synth();
    if (!s) { set_error("out of memory"); return MVB_ERR_NOMEM; }
    SplitMix64 rng(sp->seed * 0x9E3779B97F4A7C15ull + 0x1234567ull);
    s->sub.resize(N); s->nbhd.resize(N); s->commute.resize(N); s->car.assign(N, 0); s->bike.assign(N, 0);
    s->mode.resize(N); s->norm.resize(N); s->ws.resize(N); s->habit.assign((size_t)4 * N, 0.0f);

    // create_unlinked_agent: uniform subculture / neighbourhood, U[0,1) sensitivity (:214-256)
    for (uint32_t i = 0; i < N; ++i) {
        s->sub[i]  = (uint8_t)rng.below(S);
        s->nbhd[i] = (uint16_t)rng.below(H);
        s->ws[i]   = rng.f32();
    }
    // cars then bikes: uniform samples without replacement (:149-164)
    {
        std::vector<uint32_t> idx(N);
        for (int pass = 0; pass < 2; ++pass) {
            std::iota(idx.begin(), idx.end(), 0u);
            const uint32_t k = pass == 0 ? sp->n_cars : sp->n_bikes;
            for (uint32_t j = 0; j < k; ++j) {
                const uint32_t r = j + (uint32_t)rng.below(N - j);
                std::swap(idx[j], idx[r]);
                (pass == 0 ? s->car : s->bike)[idx[j]] = 1;
            }
        }
    }
    // commute distance = |GMM sample| (:166-189, gaussian.rs:36-98)
    {
        double total_w = 0;
        for (uint32_t g = 0; g < sp->n_gmm; ++g) total_w += sp->gmm[g][2];
        for (uint32_t i = 0; i < N; ++i) {
            const double r = rng.f64() * total_w;
            double cdf = 0; uint32_t g = sp->n_gmm - 1;
            for (uint32_t k = 0; k < sp->n_gmm; ++k) { cdf += sp->gmm[k][2]; if (r < cdf) { g = k; break; } }
            double u1 = rng.f64(), u2 = rng.f64();
            if (u1 < 1e-300) u1 = 1e-300;
            const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
            const double d = std::fabs(sp->gmm[g][0] + sp->gmm[g][1] * z);
            s->commute[i] = d < 4241.0 ? 0 : (d < 19457.0 ? 1 : 2);
        }
    }
    // initial mode by ownership (:286-325); habit = {mode: 1}, norm = mode (:198-200)
    for (uint32_t i = 0; i < N; ++i) {
        const double r = rng.f64();
        uint8_t m;
        if (s->car[i] && s->bike[i]) m = r < 0.4 ? 0 : (r < 0.7 ? 2 : (r < 0.85 ? 3 : 1));
        else if (s->car[i])          m = r < 0.57 ? 0 : (r < 0.79 ? 3 : 1);
        else if (s->bike[i])         m = r < 0.5 ? 2 : (r < 0.75 ? 3 : 1);
        else                         m = r < 0.5 ? 3 : 1;
        s->mode[i] = m; s->norm[i] = m; s->habit[(size_t)4 * i + m] = 1.0f;
    }
    // social network over all agents
    {
        std::vector<uint32_t> src, dst;
        ba_edges(sp->social_links_min, N, sp->seed ^ 0xA5A5A5A5ull, src, dst);
        s->s_rowptr.assign((size_t)N + 1, 0);
        for (size_t e = 0; e < src.size(); ++e) { s->s_rowptr[(size_t)src[e] + 1]++; s->s_rowptr[(size_t)dst[e] + 1]++; }
        for (uint32_t i = 0; i < N; ++i) s->s_rowptr[i + 1] += s->s_rowptr[i];
        s->s_col.resize(s->s_rowptr[N]);
        std::vector<uint64_t> cur(s->s_rowptr.begin(), s->s_rowptr.end() - 1);
        for (size_t e = 0; e < src.size(); ++e) { s->s_col[cur[src[e]]++] = dst[e]; s->s_col[cur[dst[e]]++] = src[e]; }
    }
    // one BA network per neighbourhood over its residents in agent order (simulation.rs:165-189)
    {
        std::vector<std::vector<uint32_t>> groups(H);
        for (uint32_t i = 0; i < N; ++i) groups[s->nbhd[i]].push_back(i);
        std::vector<std::vector<uint32_t>> gsrc(H), gdst(H);
        s->n_rowptr.assign((size_t)N + 1, 0);
        for (uint32_t h = 0; h < H; ++h) {
            ba_edges(sp->neighbour_links_min, (uint32_t)groups[h].size(), (sp->seed + 1) * 1000003ull + h, gsrc[h], gdst[h]);
            for (size_t e = 0; e < gsrc[h].size(); ++e) {
                s->n_rowptr[(size_t)groups[h][gsrc[h][e]] + 1]++;
                s->n_rowptr[(size_t)groups[h][gdst[h][e]] + 1]++;
            }
        }
        for (uint32_t i = 0; i < N; ++i) s->n_rowptr[i + 1] += s->n_rowptr[i];
        s->n_col.resize(s->n_rowptr[N]);
        std::vector<uint64_t> cur(s->n_rowptr.begin(), s->n_rowptr.end() - 1);
        for (uint32_t h = 0; h < H; ++h)
            for (size_t e = 0; e < gsrc[h].size(); ++e) {
                const uint32_t a = groups[h][gsrc[h][e]], b = groups[h][gdst[h][e]];
                s->n_col[cur[a]++] = b; s->n_col[cur[b]++] = a;
            }
    }
    mvb_population& p = s->pop;
    p.n_agents = N;
    p.subculture = s->sub.data(); p.neighbourhood = s->nbhd.data(); p.commute = s->commute.data();
    p.owns_car = s->car.data(); p.owns_bike = s->bike.data(); p.weather_sensitivity = s->ws.data();
    p.consistency = p.social_connectivity = p.subculture_connectivity = nullptr;
    p.neighbourhood_connectivity = p.average_weight = nullptr;
    p.habit = s->habit.data(); p.current_mode = s->mode.data(); p.norm = s->norm.data();
    p.social_rowptr = s->s_rowptr.data(); p.social_col = s->s_col.data();
    p.neighbour_rowptr = s->n_rowptr.data(); p.neighbour_col = s->n_col.data();
    *out = s;
    return MVB_OK;
}

void mvb_synth_destroy(mvb_synth* s) {
    if (!s) return;
#ifndef MVB_NO_CUDA
    if (s->d_block || s->d_block2) { cudaSetDevice(s->device); cudaFree(s->d_block); cudaFree(s->d_block2); }
#endif
    delete s;
}
const mvb_population* mvb_synth_population(const mvb_synth* s) { return s ? &s->pop : nullptr; }

}  // extern "C"

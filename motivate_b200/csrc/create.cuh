// create.cuh — mvb_create* (included by sim.cu): populations -> device layout (layout.h) -> resident handle.
//
// Everything that touches an agent or a link runs on the GPU (setup_kernels.cuh, fill_kernels.cuh); the host keeps
// the O(H) plan (which neighbourhood group starts where, how much of it is staged in shared memory) and a few
// scalars.  Two streams: the copy stream moves the caller's arrays (host-resident populations) into stream-ordered
// scratch, running several replicas ahead; the handle's stream builds the layouts behind it, tied by events, so for
// host-resident input the whole of set-up costs about what the PCIe copies cost.  Populations that already live in
// device memory (mvb_synth_create_device) are used where they are.
#pragma once

namespace {

struct SetupScalars {                 // device -> host after each stage
    SetupFlag flag;                   // first error of the graph-side checks
    uint32_t n_hubs, n_units;
    uint64_t sell_total, hub_total;
    uint32_t push_total, push_chunks; // push mode: cold links of the owned SELL agents, chunks they are cut into
};

// layout_begin's decisions (layout.cpp) from the degree histogram cnt[h][kHubDegree - d] of the non-hub agents.
struct Plan {
    uint32_t N = 0, H = 0, n_hubs = 0, hub_pad = 0, n_hs = 0, N_pad = 0, n2 = 0, staged_hub_slots = 0, hot_words = 0;
    bool all_hot = false;
    uint32_t mode = kLayoutGlobal, n_common_ranges = 0, common_words = 0;
    std::vector<uint32_t> gsize, gpad, gstart, gofs, take, take_common, nb_slice_begin;
    std::vector<StageRange> ranges;
};
inline uint32_t round_up_u32(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

int make_plan(uint32_t N, uint32_t H, uint32_t capacity_agents, const uint32_t* cnt, uint32_t n_hubs, uint32_t n_replicas, Plan& P) {
    constexpr uint32_t B = kHubDegree + 1;
    if (capacity_agents < 64 || capacity_agents % 64) { set_error("internal: bad shared-memory capacity %u", capacity_agents); return MVB_ERR_ARG; }
    P = Plan();
    P.N = N; P.H = H; P.n_hubs = n_hubs;
    P.gsize.assign(H, 0); P.gpad.resize(H); P.gstart.resize(H); P.gofs.resize(H); P.take.assign(H, 0); P.take_common.assign(H, 0);
    // cum[h][k] = agents of group h with degree >= kHubDegree - k
    std::vector<uint32_t> cum((size_t)H * B);
    for (uint32_t h = 0; h < H; ++h) {
        uint32_t run = 0;
        for (uint32_t k = 0; k < B; ++k) { run += cnt[(size_t)h * B + k]; cum[(size_t)h * B + k] = run; }
        P.gsize[h] = run;
    }
    P.hub_pad = round_up_u32(n_hubs, 64);
    P.n_hs = P.hub_pad / 32;
    uint64_t total = P.hub_pad, seen = 0;
    for (uint32_t h = 0; h < H; ++h) {
        P.gpad[h] = round_up_u32(P.gsize[h], 64);
        P.gstart[h] = (uint32_t)total; P.gofs[h] = (uint32_t)seen;
        total += P.gpad[h]; seen += P.gsize[h];
    }
    if (seen + n_hubs != N) { set_error("internal: degree histogram does not add up"); return MVB_ERR_STATE; }
    if (total >= (1ull << 29)) { set_error("too many agents for the word index of a link code"); return MVB_ERR_ARG; }
    P.N_pad = (uint32_t)total;
    P.n2 = N - n_hubs;
    P.all_hot = P.N_pad <= capacity_agents;
    P.mode = pick_layout_mode(P.N_pad, capacity_agents, n_replicas);
    P.staged_hub_slots = P.all_hot ? P.hub_pad : std::min(P.hub_pad, capacity_agents);
    if (P.mode == kLayoutLocal) {
        // hubs get at most half of shared memory; then as in layout_begin (layout.cpp): whole neighbourhoods behind the longest
        // common prefix that leaves them room, or, if a neighbourhood does not fit, its prefix alone
        P.staged_hub_slots = std::min(P.hub_pad, capacity_agents / 2 / 64 * 64);
        const uint32_t room = capacity_agents - P.staged_hub_slots;
        uint32_t max_gpad = 0;
        for (uint32_t h = 0; h < H; ++h) max_gpad = std::max(max_gpad, P.gpad[h]);
        if (max_gpad <= room) {
            auto need_for = [&](uint64_t theta, std::vector<uint32_t>& out) {
                uint64_t sum = 0, tail = 0;
                for (uint32_t h = 0; h < H; ++h) {
                    const uint32_t c = theta > kHubDegree ? 0u : cum[(size_t)h * B + (kHubDegree - (uint32_t)theta)];
                    out[h] = std::min(round_up_u32(c, 64), P.gsize[h] / 64 * 64);
                    sum += out[h];
                    tail = std::max<uint64_t>(tail, P.gpad[h] - out[h]);
                }
                return sum + tail;
            };
            std::vector<uint32_t> tmp(H);
            uint64_t lo = 0, hi = kHubDegree + 1;
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) / 2;
                if (need_for(mid, tmp) <= room) hi = mid; else lo = mid;
            }
            need_for(hi, P.take_common);
            P.take = P.gsize;
        } else {
            for (uint32_t h = 0; h < H; ++h) P.take[h] = std::min(P.gsize[h], room);
        }
    } else if (P.all_hot) {
        P.take = P.gsize;
    } else {
        // largest degree threshold theta such that {hubs} U {d >= theta} (rounded to 64 per group) fits
        const uint32_t budget = capacity_agents > P.hub_pad ? capacity_agents - P.hub_pad : 0;
        auto count_for = [&](uint64_t theta, std::vector<uint32_t>& out) {
            uint64_t sum = 0;
            for (uint32_t h = 0; h < H; ++h) {
                const uint32_t c = theta > kHubDegree ? 0u : cum[(size_t)h * B + (kHubDegree - (uint32_t)theta)];
                out[h] = std::min(round_up_u32(c, 64), P.gsize[h] / 64 * 64);
                sum += out[h];
            }
            return sum;
        };
        std::vector<uint32_t> tmp(H);
        uint64_t lo = 0, hi = kHubDegree + 1;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) / 2;
            if (count_for(mid, tmp) <= budget) hi = mid; else lo = mid;
        }
        uint64_t used = count_for(hi, P.take);
        for (uint32_t h = 0; h < H && used + 64 <= budget; ++h) {
            const uint32_t room = P.gsize[h] / 64 * 64 - P.take[h];
            const uint32_t extra = (uint32_t)std::min<uint64_t>(room, (budget - used) / 64 * 64);
            P.take[h] += extra; used += extra;
        }
    }
    P.nb_slice_begin.resize((size_t)H + 1);
    for (uint32_t h = 0; h < H; ++h) P.nb_slice_begin[h] = (P.gstart[h] - P.hub_pad) / 32;
    P.nb_slice_begin[H] = (P.N_pad - P.hub_pad) / 32;
    uint32_t staged = 0;
    auto add_range = [&](uint32_t start, uint32_t count) {
        if (!count) return;
        P.ranges.push_back(StageRange{start / 16, count / 16, staged / 16, 0});
        staged += count;
    };
    if (P.mode == kLayoutLocal) {
        add_range(0, P.staged_hub_slots);
        if (P.ranges.empty()) P.ranges.push_back(StageRange{0, 0, 0, 0});
        for (uint32_t h = 0; h < H; ++h) {
            P.ranges.push_back(StageRange{P.gstart[h] / 16, P.take_common[h] / 16, staged / 16, 0});
            staged += P.take_common[h];
        }
        const uint32_t common = staged;
        uint32_t longest = 0;
        for (uint32_t h = 0; h < H; ++h) {
            const uint32_t words = (round_up_u32(P.take[h], 64) - P.take_common[h]) / 16;
            P.ranges.push_back(StageRange{(P.gstart[h] + P.take_common[h]) / 16, words, common / 16, 0});
            longest = std::max(longest, words);
        }
        P.n_common_ranges = 1 + H;
        P.common_words = common / 16;
        staged = common + longest * 16;
    } else if (P.all_hot) add_range(0, P.N_pad);
    else {
        add_range(0, std::min(P.hub_pad, capacity_agents));
        for (uint32_t h = 0; h < H; ++h) add_range(P.gstart[h], P.take[h]);
    }
    P.hot_words = staged / 16;
    if (P.mode != kLayoutLocal) { P.n_common_ranges = (uint32_t)P.ranges.size(); P.common_words = P.hot_words; }
    return MVB_OK;
}

const char* setup_error_text(uint32_t code) {
    switch (code) {
        case kErrSubculture: return "subculture index out of range (\"Subculture not found\")";
        case kErrNeighbourhood: return "neighbourhood index out of range";
        case kErrCode: return "bad commute/mode/norm code";
        case kErrRowptr: return "rowptr is not monotone from 0 to the link count";
        case kErrDegree: return "more than 2^32 links";
        case kErrTarget: return "a link target is out of range";
    }
    return "unknown error";
}

bool is_device_pointer(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice;
}

// Device scratch that may be allocated on one stream and given back on another.
struct UpBuf {
    void* p = nullptr;
    UpBuf() = default;
    UpBuf(const UpBuf&) = delete;
    UpBuf& operator=(const UpBuf&) = delete;
    ~UpBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n, cudaStream_t st) { return cudaMallocAsync(&p, n ? n : 1, st); }
    void release(cudaStream_t st) { if (p) cudaFreeAsync(p, st); p = nullptr; }
    template <class T> T* as() const { return (T*)p; }
};

struct GraphInput {                   // the caller's graph-side arrays, on the device
    const uint64_t* srp = nullptr; const uint64_t* nrp = nullptr;
    const uint32_t* scol = nullptr; const uint32_t* ncol = nullptr;
    const uint16_t* nbhd = nullptr;
    uint64_t Es = 0, En = 0;
    UpBuf b[5];
    cudaEvent_t ev = nullptr;         // uploads done (host-resident input only)
    bool issued = false;
};
struct StateInput {                   // the caller's per-agent state of one replica, on the device
    const uint8_t *sub = nullptr, *commute = nullptr, *car = nullptr, *bike = nullptr, *mode = nullptr, *norm = nullptr;
    const uint16_t* nbhd = nullptr;
    const float* ws = nullptr; const float4* habit = nullptr;
    const float* pa[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    UpBuf b[14];
    cudaEvent_t ev = nullptr;
    bool issued = false;
};

}  // namespace

static int create_impl(uint32_t R, const mvb_population* const* pops, const mvb_tables* const* tabs,
                       const mvb_params* prm, int device, uint32_t rank, uint32_t world, const void* unique_id,
                       mvb_sim** out) {
    if (out) *out = nullptr;
    if (!out || !pops || !tabs || !prm || R == 0) { set_error("null argument or zero replicas"); return MVB_ERR_ARG; }
    int ndev = mvb_device_count();
    if (ndev <= 0) { set_error("no CUDA device available (this library has no CPU path)"); return MVB_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return MVB_ERR_ARG; }
    for (uint32_t r = 0; r < R; ++r) {
        int rc = check_tables(tabs[r]);
        if (rc) return rc;
    }
    const bool timing = getenv("MVB_TIMING") != nullptr;       // MVB_TIMING=1 prints where the set-up time went (stderr)
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tm0 = now();
    double tm_wait = 0;
    MVB_CUDA(cudaSetDevice(device));
    // which populations already live in device memory (all arrays of a population are taken to be of one kind)
    std::vector<char> on_device(R, 0);
    for (uint32_t r = 0; r < R; ++r) {
        const mvb_population* p = pops[r];
        if (!p || p->n_agents == 0 || !p->subculture || !p->neighbourhood || !p->commute || !p->owns_car || !p->owns_bike ||
            !p->weather_sensitivity || !p->habit || !p->current_mode || !p->norm || !p->social_rowptr || !p->neighbour_rowptr) {
            set_error("replica %u: population has null arrays or zero agents", r);
            return MVB_ERR_ARG;
        }
        on_device[r] = is_device_pointer(p->social_rowptr) ? 1 : 0;
    }
    std::unique_ptr<mvb_sim, void (*)(mvb_sim*)> sp(new (std::nothrow) mvb_sim(), mvb_destroy);
    mvb_sim* s = sp.get();
    if (!s) { set_error("out of memory"); return MVB_ERR_NOMEM; }
    s->device = device;
    s->prm = *prm;
    {   // the uniform-weight kernel's division shortcut (decide.cuh FAST_SHARES) needs tame connectivities;
        // anything else goes through the general kernel, which divides with plain IEEE `/`
        auto tame = [](float w) { return w >= 0x1p-60f && w <= 0x1p60f; };
        if (!tame(prm->social_connectivity) || !tame(prm->neighbourhood_connectivity)) s->any_per_agent = true;
    }
    MVB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    MVB_CUDA(cudaEventCreate(&s->ev0));
    MVB_CUDA(cudaEventCreate(&s->ev1));
    struct CopyStream {
        cudaStream_t st = nullptr;
        ~CopyStream() { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } }
    } copy;
    MVB_CUDA(cudaStreamCreateWithFlags(&copy.st, cudaStreamNonBlocking));
    cudaStream_t S = s->stream, CS = copy.st;

    // shared-memory budget: everything the device lets one block opt in to (227 KB on B200); what the
    // tables do not need holds the staged mode vector
    int sms = 0, smem_optin = 0;
    MVB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    MVB_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    uint32_t tail_max = 0, census_words = 0, H_max = 0;
    for (uint32_t r = 0; r < R; ++r) {
        tail_max = std::max(tail_max, tail_words(tabs[r]->n_subcultures, tabs[r]->n_neighbourhoods));
        census_words = std::max(census_words, rec_width(tabs[r]->n_subcultures, tabs[r]->n_neighbourhoods));
        H_max = std::max(H_max, tabs[r]->n_neighbourhoods);
    }
    tail_max = (tail_max + 3u) & ~3u;
    if ((uint64_t)tail_max * 4 + 4096 > (uint64_t)smem_optin) {
        set_error("too many neighbourhoods/subcultures: the per-block tables need %u bytes of shared memory", tail_max * 4);
        return MVB_ERR_ARG;
    }
    size_t vec_piece = 0;                          // agent-partitioned: stride of vec[0], vec[1], normvec inside s->px_vec
    s->smem_bytes = (size_t)smem_optin;
    s->vec_cap_words = ((uint32_t)smem_optin / 4 - tail_max) & ~3u;
    s->census_smem = sizeof(uint32_t) * census_words;
    uint32_t capacity_agents = staged_capacity_agents(s->vec_cap_words);
    if (const char* e = getenv("MVB_STAGE_AGENTS"))               // tests: a smaller staged set exercises the cold-link path on small populations
        capacity_agents = std::min<uint32_t>(capacity_agents, std::max(64, atoi(e)) / 64 * 64);
    s->grid = (uint32_t)sms;

    s->reps.resize(R);
    s->h_reps.resize(R);
    // distinct link graphs: replicas that pass the same graph + neighbourhood arrays share one (scenario sweeps)
    std::vector<std::shared_ptr<Graph>> graphs;
    std::vector<uint32_t> graph_of(R), first_rep;
    for (uint32_t r = 0; r < R; ++r) {
        const mvb_population* p = pops[r];
        const void* key[5] = {p->social_rowptr, p->social_col, p->neighbour_rowptr, p->neighbour_col, p->neighbourhood};
        uint32_t g = 0;
        for (; g < graphs.size(); ++g)
            if (pops[first_rep[g]]->n_agents == p->n_agents && !memcmp(graphs[g]->key, key, sizeof key) &&
                tabs[first_rep[g]]->n_neighbourhoods == tabs[r]->n_neighbourhoods) break;
        if (g == graphs.size()) {
            graphs.push_back(std::make_shared<Graph>());
            memcpy(graphs.back()->key, key, sizeof key);
            first_rep.push_back(r);
        }
        graph_of[r] = g;
    }
    const uint32_t G = (uint32_t)graphs.size();
    // link counts: rowptr[N] (one small read per graph when the population lives on the device)
    std::vector<GraphInput> gin(G);
    for (uint32_t g = 0; g < G; ++g) {
        const mvb_population* p = pops[first_rep[g]];
        if (on_device[first_rep[g]]) {
            MVB_CUDA(cudaMemcpy(&gin[g].Es, p->social_rowptr + p->n_agents, 8, cudaMemcpyDeviceToHost));
            MVB_CUDA(cudaMemcpy(&gin[g].En, p->neighbour_rowptr + p->n_agents, 8, cudaMemcpyDeviceToHost));
        } else {
            gin[g].Es = p->social_rowptr[p->n_agents];
            gin[g].En = p->neighbour_rowptr[p->n_agents];
        }
        if ((gin[g].Es && !p->social_col) || (gin[g].En && !p->neighbour_col)) { set_error("replica %u: null column array", first_rep[g]); return MVB_ERR_ARG; }
        if (gin[g].Es >= (1ull << 40) || gin[g].En >= (1ull << 40)) { set_error("replica %u: rowptr[n_agents] is not a plausible link count", first_rep[g]); return MVB_ERR_ARG; }
    }
    {   // first chunk of device memory: an estimate of everything the graphs and replicas will need (link codes ~ the
        // caller's CSR columns + slice padding, ~16 B per agent of graph arrays, 27-47 B per agent of state)
        size_t est = 64u << 20;
        for (uint32_t g = 0; g < G; ++g) {
            const size_t e = (gin[g].Es + gin[g].En) / (world > 1 ? world : 1);
            est += 4 * e + e / 4 + 24 * (size_t)pops[first_rep[g]]->n_agents + (1u << 20);
        }
        for (uint32_t r = 0; r < R; ++r) est += 48 * (size_t)pops[r]->n_agents + (1u << 16);
        size_t free_b = 0, total_b = 0;
        MVB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        s->mem.chunk_bytes = std::max<size_t>(256u << 20, std::min(est, free_b / 2));
    }
    {   // scratch comes from the device's stream-ordered pool; let it keep what it has between replicas
        cudaMemPool_t pool;
        MVB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        MVB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    struct Events {                                            // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~Events() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
        cudaError_t make(cudaEvent_t* e) { cudaError_t rc = cudaEventCreateWithFlags(e, cudaEventDisableTiming); if (rc == cudaSuccess) v.push_back(*e); return rc; }
    } events;
    std::vector<StateInput> sin(R);
    // ---- copy stream: the caller's arrays, a few replicas ahead of the layout work ------------------------------
    auto up = [&](UpBuf& b, const void* src, size_t bytes, const void** view) -> int {
        MVB_CUDA(b.alloc(bytes, CS));
        if (bytes) MVB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, CS));
        *view = b.p;
        return MVB_OK;
    };
    auto issue_graph = [&](uint32_t g) -> int {
        GraphInput& gi = gin[g];
        if (gi.issued) return MVB_OK;
        gi.issued = true;
        const mvb_population* p = pops[first_rep[g]];
        if (on_device[first_rep[g]]) {
            gi.srp = p->social_rowptr; gi.nrp = p->neighbour_rowptr; gi.scol = p->social_col; gi.ncol = p->neighbour_col; gi.nbhd = p->neighbourhood;
            return MVB_OK;
        }
        const size_t nb_rp = 8 * ((size_t)p->n_agents + 1);
        int rc;
        if ((rc = up(gi.b[0], p->social_rowptr, nb_rp, (const void**)&gi.srp)) || (rc = up(gi.b[1], p->neighbour_rowptr, nb_rp, (const void**)&gi.nrp)) ||
            (rc = up(gi.b[2], p->social_col, 4 * gi.Es, (const void**)&gi.scol)) || (rc = up(gi.b[3], p->neighbour_col, 4 * gi.En, (const void**)&gi.ncol)) ||
            (rc = up(gi.b[4], p->neighbourhood, 2 * (size_t)p->n_agents, (const void**)&gi.nbhd)))
            return rc;
        MVB_CUDA(events.make(&gi.ev));
        MVB_CUDA(cudaEventRecord(gi.ev, CS));
        return MVB_OK;
    };
    auto issue_state = [&](uint32_t r) -> int {
        StateInput& si = sin[r];
        if (si.issued) return MVB_OK;
        si.issued = true;
        const mvb_population* p = pops[r];
        const float* pa_src[5] = {p->consistency, p->social_connectivity, p->subculture_connectivity, p->neighbourhood_connectivity, p->average_weight};
        if (on_device[r]) {
            si.sub = p->subculture; si.nbhd = p->neighbourhood; si.commute = p->commute; si.car = p->owns_car; si.bike = p->owns_bike;
            si.ws = p->weather_sensitivity; si.habit = reinterpret_cast<const float4*>(p->habit); si.mode = p->current_mode; si.norm = p->norm;
            for (int k = 0; k < 5; ++k) si.pa[k] = pa_src[k];
            return MVB_OK;
        }
        const size_t N = p->n_agents;
        int rc;
        if ((rc = up(si.b[0], p->subculture, N, (const void**)&si.sub)) || (rc = up(si.b[1], p->neighbourhood, 2 * N, (const void**)&si.nbhd)) ||
            (rc = up(si.b[2], p->commute, N, (const void**)&si.commute)) || (rc = up(si.b[3], p->owns_car, N, (const void**)&si.car)) ||
            (rc = up(si.b[4], p->owns_bike, N, (const void**)&si.bike)) || (rc = up(si.b[5], p->weather_sensitivity, 4 * N, (const void**)&si.ws)) ||
            (rc = up(si.b[6], p->habit, 16 * N, (const void**)&si.habit)) || (rc = up(si.b[7], p->current_mode, N, (const void**)&si.mode)) ||
            (rc = up(si.b[8], p->norm, N, (const void**)&si.norm)))
            return rc;
        for (int k = 0; k < 5; ++k)
            if (pa_src[k] && (rc = up(si.b[9 + k], pa_src[k], 4 * N, (const void**)&si.pa[k]))) return rc;
        MVB_CUDA(events.make(&si.ev));
        MVB_CUDA(cudaEventRecord(si.ev, CS));
        return MVB_OK;
    };
    constexpr uint32_t kAhead = 4;                             // replicas whose copies are queued ahead of the one being laid out
    uint32_t issued_upto = 0;
    auto issue_upto = [&](uint32_t upto) -> int {
        for (; issued_upto < std::min(upto, R); ++issued_upto) {
            int rc;
            if ((rc = issue_graph(graph_of[issued_upto])) || (rc = issue_state(issued_upto))) return rc;
        }
        return MVB_OK;
    };

    // pinned landing zone of the small device -> host reads (one graph at a time) and source of the plan uploads
    const size_t cnt_bytes = (sizeof(uint32_t) * (size_t)H_max * (kHubDegree + 1) + 255) & ~(size_t)255;
    const size_t plan_bytes = sizeof(uint32_t) * (5 * (size_t)H_max + 4) + sizeof(StageRange) * (2 * (size_t)H_max + 2);
    char* pin = nullptr; size_t pin_cap = 0;
    MVB_CUDA(pinned_cache().get(cnt_bytes + plan_bytes + 256, &pin, &pin_cap));
    struct PinGuard { char* p; size_t cap; ~PinGuard() { pinned_cache().put(p, cap); } } pin_guard{pin, pin_cap};
    uint32_t* h_cnt = reinterpret_cast<uint32_t*>(pin);
    SetupScalars* h_scal = reinterpret_cast<SetupScalars*>(pin + cnt_bytes);
    char* h_plan = pin + cnt_bytes + 128;
    cudaEvent_t ev_host = nullptr;
    MVB_CUDA(events.make(&ev_host));
    auto host_wait = [&]() -> int {                            // the host needs what the handle's stream has produced so far
        const double t = now();
        MVB_CUDA(cudaEventRecord(ev_host, S));
        MVB_CUDA(cudaEventSynchronize(ev_host));
        tm_wait += now() - t;
        return MVB_OK;
    };
    UpBuf rep_flags;                                           // SetupFlag[R]: per-replica state checks, read once at the end
    MVB_CUDA(rep_flags.alloc(sizeof(SetupFlag) * R, S));
    MVB_CUDA(cudaMemsetAsync(rep_flags.p, 0, sizeof(SetupFlag) * R, S));

    std::vector<char> built(G, 0);
    std::list<std::vector<float>> host_f;
    const double tm_first = now();
    for (uint32_t r = 0; r < R; ++r) {
        const mvb_population* p = pops[r];
        const mvb_tables* t = tabs[r];
        Replica& rp = s->reps[r];
        const uint32_t N = p->n_agents, g_idx = graph_of[r];
        rp.N = N; rp.S = t->n_subcultures; rp.H = t->n_neighbourhoods;
        rp.rec_width = rec_width(rp.S, rp.H);
        rp.rec_off = s->rec_stride;
        s->rec_stride += rp.rec_width;
        rp.graph = graphs[g_idx];
        if (int rc = issue_upto(r + 1 + kAhead)) return rc;
        Graph* g = rp.graph.get();
        GraphLayout& L = g->L;
        if (!built[g_idx]) {
            built[g_idx] = 1;
            GraphInput& gi = gin[g_idx];
            const uint32_t H = rp.H, B = kHubDegree + 1;
            if (gi.ev) MVB_CUDA(cudaStreamWaitEvent(S, gi.ev, 0));
            // ---- stage A: degrees, first sort keys, histogram -------------------------------------------------
            UpBuf deg, key[2], val[2], cnt, scal, scratch, is_hot, nsh, nnh;
            MVB_CUDA(deg.alloc(4 * (size_t)N, S)); MVB_CUDA(cnt.alloc(4 * (size_t)H * B, S)); MVB_CUDA(scal.alloc(sizeof(SetupScalars), S));
            for (int k = 0; k < 2; ++k) { MVB_CUDA(key[k].alloc(4 * (size_t)N, S)); MVB_CUDA(val[k].alloc(4 * (size_t)N, S)); }
            MVB_CUDA(scratch.alloc(4 * radix_scratch_u32(N), S));
            MVB_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * (size_t)H * B, S));
            MVB_CUDA(cudaMemsetAsync(scal.p, 0, sizeof(SetupScalars), S));
            SetupScalars* d_scal = scal.as<SetupScalars>();
            degree_kernel<<<(N + 255) / 256, 256, 0, S>>>(N, H, gi.srp, gi.nrp, gi.Es, gi.En, gi.nbhd, deg.as<uint32_t>(), key[0].as<uint32_t>(),
                                                        val[0].as<uint32_t>(), cnt.as<uint32_t>(), &d_scal->n_hubs, &d_scal->flag);
            MVB_CUDA(cudaGetLastError());
            MVB_CUDA(cudaMemcpyAsync(h_cnt, cnt.p, 4 * (size_t)H * B, cudaMemcpyDeviceToHost, S));
            MVB_CUDA(cudaMemcpyAsync(h_scal, scal.p, sizeof(SetupScalars), cudaMemcpyDeviceToHost, S));
            if (int rc = host_wait()) return rc;
            if (h_scal->flag.code) { set_error("replica %u: agent %u: %s", r, h_scal->flag.index, setup_error_text(h_scal->flag.code)); return MVB_ERR_ARG; }
            Plan P;
            if (int rc = make_plan(N, H, capacity_agents, h_cnt, h_scal->n_hubs, R, P)) return rc;
            L = GraphLayout();
            L.N = N; L.N_pad = P.N_pad; L.Es = gi.Es; L.En = gi.En;
            L.n_hub_slices = P.n_hs; L.n_sell_slices = (P.N_pad - P.hub_pad) / 32; L.hot_words = P.hot_words;
            L.ranges = P.ranges; L.mode = P.mode; L.nb_slice_begin = P.nb_slice_begin; L.n_common_ranges = P.n_common_ranges;
            L.push = !P.all_hot && pick_push_mode(P.N_pad, R);      // cold links of SELL agents through push_kernel (layout.h)
            g->common_words = P.common_words;
            if (L.hot_words + 4 > s->vec_cap_words) { set_error("internal: staged vector larger than its shared-memory slot"); return MVB_ERR_STATE; }
            const uint32_t n_hub_rows = P.n_hs * 32, n_sell = L.n_sell_slices, n2 = P.n2, n_hubs = P.n_hubs;
            // ---- stage B: the two orders, hot counts, slices, offsets ------------------------------------------
            uint32_t* hp = reinterpret_cast<uint32_t*>(h_plan);
            memcpy(hp, P.gofs.data(), 4 * (size_t)H); memcpy(hp + H, P.take.data(), 4 * (size_t)H); memcpy(hp + 2 * H, P.gstart.data(), 4 * (size_t)H);
            memcpy(hp + 3 * H, P.take_common.data(), 4 * (size_t)H);
            memcpy(hp + 4 * H, P.nb_slice_begin.data(), 4 * ((size_t)H + 1));
            uint32_t* hr = hp + 5 * H + 4;
            memcpy(hr, P.ranges.data(), sizeof(StageRange) * P.ranges.size());
            UpBuf plan;
            MVB_CUDA(plan.alloc(4 * 4 * (size_t)H, S));
            MVB_CUDA(cudaMemcpyAsync(plan.p, hp, 4 * 4 * (size_t)H, cudaMemcpyHostToDevice, S));
            const uint32_t* d_gofs = plan.as<uint32_t>(); const uint32_t* d_take = d_gofs + H; const uint32_t* d_gstart = d_gofs + 2 * H;
            const uint32_t* d_take_common = d_gofs + 3 * H;
            MVB_CARVE(s->mem, g->ranges, sizeof(StageRange) * P.ranges.size());
            MVB_CUDA(cudaMemcpyAsync(g->ranges.p, hr, sizeof(StageRange) * P.ranges.size(), cudaMemcpyHostToDevice, S));
            MVB_CARVE(s->mem, g->nb_slice_begin, 4 * ((size_t)H + 1));
            MVB_CUDA(cudaMemcpyAsync(g->nb_slice_begin.p, hp + 4 * H, 4 * ((size_t)H + 1), cudaMemcpyHostToDevice, S));
            MVB_CARVE(s->mem, g->int2ext, 4 * (size_t)P.N_pad);
            MVB_CARVE(s->mem, g->deg, 4 * (size_t)P.N_pad);
            MVB_CARVE(s->mem, g->hub_off, 8 * (size_t)n_hub_rows + 8);
            MVB_CARVE(s->mem, g->hub_deg, 16 * (size_t)n_hub_rows + 16);
            MVB_CARVE(s->mem, g->hub_unit_begin, 4 * ((size_t)P.n_hs + 1));
            MVB_CARVE(s->mem, g->sell_meta, 16 * ((size_t)n_sell + 64));
            uint32_t* kk[2] = {key[0].as<uint32_t>(), key[1].as<uint32_t>()};
            uint32_t* vv[2] = {val[0].as<uint32_t>(), val[1].as<uint32_t>()};
            int cur = 0;
            MVB_CUDA(radix_sort_pairs(kk, vv, N, 8 + bits_for(H), scratch.as<uint32_t>(), &cur, S));     // (neighbourhood, degree desc, id asc); hubs last
            uint32_t* order1 = vv[cur]; uint32_t* key1s = kk[cur];
            fill_u32_kernel<<<sms * 4, 256, 0, S>>>(g->int2ext.as<uint32_t>(), P.N_pad, kInvalid);
            MVB_CUDA(is_hot.alloc(N, S)); MVB_CUDA(nsh.alloc(4 * (size_t)N, S)); MVB_CUDA(nnh.alloc(4 * (size_t)N, S));
            MVB_CUDA(cudaMemsetAsync(is_hot.p, 0, N, S));
            UpBuf hk[2], hv[2];
            if (n_hubs) {                                      // hubs: (degree desc, id asc), dealt round the hub slices
                for (int k = 0; k < 2; ++k) { MVB_CUDA(hk[k].alloc(4 * (size_t)n_hubs, S)); MVB_CUDA(hv[k].alloc(4 * (size_t)n_hubs, S)); }
                MVB_CUDA(cudaMemcpyAsync(hv[0].p, order1 + n2, 4 * (size_t)n_hubs, cudaMemcpyDeviceToDevice, S));
                hub_keys_kernel<<<(n_hubs + 255) / 256, 256, 0, S>>>(n_hubs, hv[0].as<uint32_t>(), deg.as<uint32_t>(), hk[0].as<uint32_t>());
                uint32_t* hkk[2] = {hk[0].as<uint32_t>(), hk[1].as<uint32_t>()};
                uint32_t* hvv[2] = {hv[0].as<uint32_t>(), hv[1].as<uint32_t>()};
                int hc = 0;
                MVB_CUDA(radix_sort_pairs(hkk, hvv, n_hubs, 32, scratch.as<uint32_t>(), &hc, S));
                place_hubs_kernel<<<(n_hubs + 255) / 256, 256, 0, S>>>(n_hubs, P.n_hs, P.staged_hub_slots, hvv[hc], g->int2ext.as<uint32_t>(), is_hot.as<uint8_t>());
            }
            if (n2) mark_hot_kernel<<<(n2 + 255) / 256, 256, 0, S>>>(n2, key1s, order1, d_gofs, d_take, P.mode == kLayoutLocal ? d_take_common : nullptr, is_hot.as<uint8_t>());
            uint32_t* d_bad = &d_scal->flag.code;              // count_hot_kernel ORs 1 in: translated below
            MVB_CUDA(cudaMemsetAsync(d_bad, 0, 4, S));
            count_hot_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)N * 32 + 255) / 256, (uint64_t)sms * 32), 256, 0, S>>>(
                N, gi.srp, gi.scol, gi.nrp, gi.ncol, is_hot.as<uint8_t>(), gi.nbhd, nsh.as<uint32_t>(), nnh.as<uint32_t>(), d_bad);
            MVB_CUDA(cudaGetLastError());
            if (n2) {
                key2_kernel<<<(n2 + 255) / 256, 256, 0, S>>>(n2, key1s, order1, d_gofs, d_take, P.mode == kLayoutLocal ? d_take_common : nullptr,
                                                              nsh.as<uint32_t>(), nnh.as<uint32_t>(), kk[cur ^ 1]);
                // second order: (group, hot part first, hot length desc), ties keep the first order
                uint32_t* k2[2] = {kk[cur ^ 1], kk[cur]};
                uint32_t* v2[2] = {vv[cur], vv[cur ^ 1]};
                int c2 = 0;
                MVB_CUDA(radix_sort_pairs(k2, v2, n2, 8 + bits_for(4 * H), scratch.as<uint32_t>(), &c2, S));
                place_agents_kernel<<<(n2 + 255) / 256, 256, 0, S>>>(n2, k2[c2], v2[c2], d_gofs, d_gstart, g->int2ext.as<uint32_t>());
            }
            UpBuf sell_K, sell_len, sell_off, hub_nu, hub_len, unit_base, sums;
            MVB_CUDA(sell_K.alloc(4 * (size_t)n_sell, S)); MVB_CUDA(sell_len.alloc(8 * (size_t)n_sell, S)); MVB_CUDA(sell_off.alloc(8 * (size_t)n_sell + 8, S));
            MVB_CUDA(hub_nu.alloc(4 * (size_t)n_hub_rows, S)); MVB_CUDA(hub_len.alloc(8 * (size_t)n_hub_rows, S)); MVB_CUDA(unit_base.alloc(4 * (size_t)n_hub_rows + 4, S));
            MVB_CUDA(sums.alloc(8 * ((size_t)scan_blocks(std::max<uint64_t>(n_sell, n_hub_rows)) + 8), S));
            if (n_sell) {
                slice_dims_kernel<<<(unsigned)(((uint64_t)n_sell * 32 + 255) / 256), 256, 0, S>>>(P.n_hs, n_sell, g->int2ext.as<uint32_t>(), deg.as<uint32_t>(),
                                                                                                nsh.as<uint32_t>(), nnh.as<uint32_t>(), L.push ? 1u : 0u, sell_K.as<uint32_t>(), sell_len.as<uint64_t>());
                MVB_CUDA(exclusive_scan<uint64_t>(sell_len.as<uint64_t>(), sell_off.as<uint64_t>(), n_sell, sums.as<uint64_t>(), &d_scal->sell_total, S));
            }
            if (n_hub_rows) {
                hub_dims_kernel<<<(n_hub_rows + 255) / 256, 256, 0, S>>>(n_hub_rows, g->int2ext.as<uint32_t>(), gi.srp, gi.nrp, nsh.as<uint32_t>(), nnh.as<uint32_t>(),
                                                                       g->hub_deg.as<uint4>(), hub_nu.as<uint32_t>(), hub_len.as<uint64_t>());
                MVB_CUDA(exclusive_scan<uint32_t>(hub_nu.as<uint32_t>(), unit_base.as<uint32_t>(), n_hub_rows, sums.as<uint32_t>(), &d_scal->n_units, S));
                MVB_CUDA(exclusive_scan<uint64_t>(hub_len.as<uint64_t>(), g->hub_off.as<uint64_t>(), n_hub_rows, sums.as<uint64_t>(), &d_scal->hub_total, S));
            }
            MVB_CUDA(cudaGetLastError());
            MVB_CUDA(cudaMemcpyAsync(h_scal, scal.p, sizeof(SetupScalars), cudaMemcpyDeviceToHost, S));
            if (int rc = host_wait()) return rc;
            if (h_scal->flag.code) { set_error("replica %u: a link target is out of range", r); return MVB_ERR_ARG; }
            const uint64_t sell_total = n_sell ? h_scal->sell_total : 0, hub_total = n_hub_rows ? h_scal->hub_total : 0;
            const uint32_t n_units = n_hub_rows ? h_scal->n_units : 0;
            if (hub_total / 32 >= (1ull << 32)) { set_error("hub link lists too long for a 32-bit row offset"); return MVB_ERR_ARG; }
            // ---- agent-partitioned handles keep the link codes of their own slices only ------------------------
            uint32_t own_hub_b = 0, own_hub_e = P.n_hs, own_sell_b = 0, own_sell_e = n_sell;
            uint64_t sell_rebase = 0, sell_own_len = sell_total, hub_rebase = 0, hub_own_len = hub_total;
            if (world > 1) {
                L.sell_K.resize(n_sell); L.hub_deg.resize((size_t)4 * n_hub_rows);
                std::vector<uint64_t> h_hub_off((size_t)n_hub_rows + 1, 0);
                if (n_sell) MVB_CUDA(cudaMemcpyAsync(L.sell_K.data(), sell_K.p, 4 * (size_t)n_sell, cudaMemcpyDeviceToHost, S));
                if (n_hub_rows) {
                    MVB_CUDA(cudaMemcpyAsync(L.hub_deg.data(), g->hub_deg.p, 16 * (size_t)n_hub_rows, cudaMemcpyDeviceToHost, S));
                    MVB_CUDA(cudaMemcpyAsync(h_hub_off.data(), g->hub_off.p, 8 * (size_t)n_hub_rows, cudaMemcpyDeviceToHost, S));
                }
                if (int rc = host_wait()) return rc;
                h_hub_off[n_hub_rows] = hub_total;
                partition_slices(L, world, s->bounds);
                const uint32_t b = s->bounds[rank], e = s->bounds[rank + 1], nh = P.n_hs;
                own_hub_b = std::min(b, nh); own_hub_e = std::min(e, nh);
                own_sell_b = std::max(b, nh) - nh; own_sell_e = std::max(e, nh) - nh;
                std::vector<uint64_t> off((size_t)n_sell + 1, 0);
                for (uint32_t k = 0; k < n_sell; ++k) off[k + 1] = off[k] + 32ull * ((L.sell_K[k] & 0xFFFFu) + (L.push ? 0u : L.sell_K[k] >> 16));
                sell_rebase = off[own_sell_b]; sell_own_len = off[own_sell_e] - off[own_sell_b];
                hub_rebase = h_hub_off[(size_t)own_hub_b * 32]; hub_own_len = h_hub_off[(size_t)own_hub_e * 32] - hub_rebase;
            }
            // ---- stage C: link codes ---------------------------------------------------------------------------
            L.hub_codes_len = hub_own_len + 256;               // + slack for the row prefetch
            L.sell_codes_len = sell_own_len + 256;             // + slack: the kernel's row prefetch may run 2 rows past a slice
            MVB_CARVE(s->mem, g->hub_units, 16 * ((size_t)n_units + 1));
            MVB_CARVE(s->mem, g->hub_codes, 4 * L.hub_codes_len);
            MVB_CARVE(s->mem, g->sell_codes, 4 * L.sell_codes_len);
            sell_meta_kernel<<<(n_sell + 64 + 255) / 256, 256, 0, S>>>(n_sell, sell_off.as<uint64_t>(), sell_K.as<uint32_t>(), sell_rebase, own_sell_b, own_sell_e, L.push ? 1u : 0u, g->sell_meta.as<uint4>());
            hub_units_kernel<<<(n_hub_rows + 1 + 255) / 256, 256, 0, S>>>(n_hub_rows, g->hub_deg.as<uint4>(), unit_base.as<uint32_t>(), g->hub_off.as<uint64_t>(), hub_rebase,
                                                                        g->hub_units.as<uint4>(), g->hub_unit_begin.as<uint32_t>(), n_units);
            if (hub_rebase && n_hub_rows) rebase_u64_kernel<<<(n_hub_rows + 255) / 256, 256, 0, S>>>(g->hub_off.as<uint64_t>(), n_hub_rows, hub_rebase);
            // unused entries = the code of the always-zero shared-memory word (layout.h)
            fill_u32_kernel<<<sms * 8, 256, 0, S>>>(g->hub_codes.as<uint32_t>(), L.hub_codes_len, pad_code(s->vec_cap_words));
            fill_u32_kernel<<<sms * 8, 256, 0, S>>>(g->sell_codes.as<uint32_t>(), L.sell_codes_len, pad_code(s->vec_cap_words));
            UpBuf tcode, tcold;
            MVB_CUDA(tcode.alloc(4 * (size_t)N, S));
            if (P.mode == kLayoutLocal) MVB_CUDA(tcold.alloc(4 * (size_t)N, S));
            tcode_kernel<<<(P.N_pad + 255) / 256, 256, 0, S>>>(P.N_pad, g->int2ext.as<uint32_t>(), g->ranges.as<StageRange>(), (uint32_t)P.ranges.size(),
                                                             is_hot.as<uint8_t>(), P.mode == kLayoutLocal ? gi.nbhd : nullptr, H, P.hub_pad, tcode.as<uint32_t>(), tcold.as<uint32_t>());
            MVB_CUDA(cudaGetLastError());
            // ---- push mode: where every owned SELL agent's cold links go in the (still unsorted) list ------------------
            UpBuf pcnt, poff, pk[2], pv[2], pscratch, psums, pfirst;
            const uint32_t n_own_slots = (own_sell_e - own_sell_b) * 32;
            uint32_t n_push = 0, n_chunks = 0;
            const uint32_t push_max_agents = push_chunk_agents(~0ull, s->grid);       // (the cap itself: kPushChunkMax unless a test lowers it)
            if (L.push && n_own_slots) {
                const uint32_t chunk_cap = std::min<uint32_t>(65535, n_own_slots / 32 + 1);        // chunk ids travel in 16 bits
                MVB_CUDA(pcnt.alloc(4 * (size_t)n_own_slots, S)); MVB_CUDA(poff.alloc(4 * (size_t)n_own_slots + 4, S));
                MVB_CUDA(psums.alloc(4 * ((size_t)scan_blocks(n_own_slots) + 8), S));
                MVB_CUDA(pfirst.alloc(4 * ((size_t)chunk_cap + 1), S));
                push_count_kernel<<<(n_own_slots + 255) / 256, 256, 0, S>>>(P.n_hs, own_sell_b * 32, n_own_slots, g->int2ext.as<uint32_t>(), deg.as<uint32_t>(),
                                                                           nsh.as<uint32_t>(), nnh.as<uint32_t>(), pcnt.as<uint32_t>());
                MVB_CUDA(exclusive_scan<uint32_t>(pcnt.as<uint32_t>(), poff.as<uint32_t>(), n_own_slots, psums.as<uint32_t>(), &d_scal->push_total, S));
                push_plan_kernel<<<1, 32, 0, S>>>(n_own_slots, poff.as<uint32_t>(), &d_scal->push_total, s->grid, push_max_agents, chunk_cap,
                                                 pfirst.as<uint32_t>(), &d_scal->push_chunks);
                MVB_CUDA(cudaMemcpyAsync(h_scal, scal.p, sizeof(SetupScalars), cudaMemcpyDeviceToHost, S));
                if (int rc = host_wait()) return rc;
                n_push = h_scal->push_total; n_chunks = h_scal->push_chunks;
                if (n_chunks == 0xFFFFFFFFu) { set_error("internal: too many push chunks for %u agents", n_own_slots); return MVB_ERR_STATE; }
                for (int k = 0; k < 2; ++k) { MVB_CUDA(pk[k].alloc(4 * (size_t)n_push + 4, S)); MVB_CUDA(pv[k].alloc(4 * (size_t)n_push + 4, S)); }
                MVB_CUDA(pscratch.alloc(4 * radix_scratch_u32(n_push), S));
            }
            FillArgs fa{gi.srp, gi.scol, gi.nrp, gi.ncol, tcode.as<uint32_t>(), g->int2ext.as<uint32_t>(), P.n_hs, n_sell,
                        g->sell_meta.as<uint4>(), g->sell_codes.as<uint32_t>(), g->deg.as<uint32_t>(),
                        g->hub_off.as<uint64_t>(), g->hub_deg.as<uint4>(), g->hub_codes.as<uint32_t>(),
                        own_hub_b, own_hub_e, own_sell_b, own_sell_e,
                        P.mode == kLayoutLocal ? tcold.as<uint32_t>() : nullptr, is_hot.as<uint8_t>(), gi.nbhd,
                        n_chunks ? pk[0].as<uint32_t>() : nullptr, pv[0].as<uint32_t>(), poff.as<uint32_t>(), pfirst.as<uint32_t>(), n_chunks};
            MVB_CUDA(cudaMemsetAsync(g->deg.p, 0, g->deg.bytes, S));
            if (n_sell) fill_sell_kernel<<<(unsigned)(((uint64_t)n_sell * 32 + 255) / 256), 256, 0, S>>>(fa);
            if (P.n_hs) fill_hub_kernel<<<(unsigned)(((uint64_t)n_hub_rows * 32 + 255) / 256), 256, 0, S>>>(fa, n_hub_rows);
            MVB_CUDA(cudaGetLastError());
            if (world > 1) { s->own[0] = own_hub_b; s->own[1] = own_hub_e; s->own[2] = own_sell_b; s->own[3] = own_sell_e; }
            if (n_chunks) {
                // ---- push mode: sort every chunk's entries by target word (the chunks are contiguous already: a stable sort
                // by word, then by chunk), and keep {code u32, source u16} + the chunk boundaries
                uint32_t* kk2[2] = {pk[0].as<uint32_t>(), pk[1].as<uint32_t>()};
                uint32_t* vv2[2] = {pv[0].as<uint32_t>(), pv[1].as<uint32_t>()};
                int pc = 0;
                MVB_CUDA(radix_sort_pairs(kk2, vv2, n_push, bits_for(P.N_pad / 16), pscratch.as<uint32_t>(), &pc, S, 7));          // code >> 7 = word of the vector
                MVB_CUDA(radix_sort_pairs(vv2, kk2, n_push, bits_for(n_chunks - 1), pscratch.as<uint32_t>(), &pc, S, 16, pc));    // roles swapped: key = chunk << 16 | source
                MVB_CARVE(s->mem, g->push_code, 4 * (size_t)n_push + 64 * 1024);     // + slack: the kernel reads rows of 4 096 entries
                MVB_CARVE(s->mem, g->push_src, 2 * (size_t)n_push + 32 * 1024);
                MVB_CARVE(s->mem, g->push_chunk_begin, 4 * ((size_t)n_chunks + 1));
                MVB_CARVE(s->mem, g->push_chunk_first, 4 * ((size_t)n_chunks + 1));
                MVB_CUDA(cudaMemcpyAsync(g->push_chunk_first.p, pfirst.p, 4 * ((size_t)n_chunks + 1), cudaMemcpyDeviceToDevice, S));
                MVB_CUDA(cudaMemcpyAsync(g->push_code.p, kk2[pc], 4 * (size_t)n_push, cudaMemcpyDeviceToDevice, S));
                if (n_push) push_src_kernel<<<sms * 8, 256, 0, S>>>(n_push, vv2[pc], g->push_src.as<uint16_t>());
                push_chunk_begin_kernel<<<(n_chunks + 1 + 255) / 256, 256, 0, S>>>(n_chunks, n_own_slots, pfirst.as<uint32_t>(), poff.as<uint32_t>(), &d_scal->push_total,
                                                                                  g->push_chunk_begin.as<uint32_t>());
                MVB_CUDA(cudaGetLastError());
                g->push_chunks = n_chunks; g->push_chunk_agents = push_max_agents; g->push_own_slots = n_own_slots; g->push_links = n_push;
                g->push_first_slot = (P.n_hs + own_sell_b) * 32;
            }
            // everything temporary goes back to the pool in stream order
            for (UpBuf* b : {&deg, &key[0], &key[1], &val[0], &val[1], &cnt, &scal, &scratch, &is_hot, &nsh, &nnh, &plan, &hk[0], &hk[1], &hv[0], &hv[1],
                             &sell_K, &sell_len, &sell_off, &hub_nu, &hub_len, &unit_base, &sums, &tcode, &tcold, &pcnt, &poff, &pk[0], &pk[1], &pv[0], &pv[1], &pscratch, &psums, &pfirst})
                b->release(S);
            for (UpBuf& b : gi.b) b.release(S);
        }
        const uint32_t NP = L.N_pad;
        s->max_npad = std::max(s->max_npad, NP);
        // ---- per-agent state of the replica: the caller's arrays, permuted into internal slot order on the device ----
        const float pa_def[5] = {prm->consistency, prm->social_connectivity, prm->subculture_connectivity,
                                 prm->neighbourhood_connectivity, prm->average_weight};
        StateInput& si = sin[r];
        rp.per_agent = false;
        for (int k = 0; k < 5; ++k) rp.per_agent |= (si.pa[k] != nullptr);
        if (rp.per_agent) s->any_per_agent = true;
        {
            const size_t hub_cnt_bytes = sizeof(uint32_t) * 8 * 32 * (size_t)L.n_hub_slices;
            vec_piece = Arena::piece(NP / 4 + 64);
            MVB_CARVE(s->mem, rp.state, (size_t)(NP / 32) * kRecWords * 4);
            // + 64 bytes: the early gather of a padded cold-list entry reads the word a padding code names (at most
            // 16 bytes past a vector that only just exceeds the shared-memory capacity); its value is never used
            if (world > 1) {
                // agent-partitioned: the vectors get an allocation of their own, which the other ranks can map (setup_peer_exchange)
                MVB_CUDA(s->px_vec.alloc(3 * vec_piece));
                rp.vec[0].view(s->px_vec.p, NP / 4 + 64); rp.vec[1].view(s->px_vec.as<char>() + vec_piece, NP / 4 + 64);
                rp.normvec.view(s->px_vec.as<char>() + 2 * vec_piece, NP / 4);
            } else {
                MVB_CARVE(s->mem, rp.vec[0], NP / 4 + 64); MVB_CARVE(s->mem, rp.vec[1], NP / 4 + 64); MVB_CARVE(s->mem, rp.normvec, NP / 4);
            }
            if (rp.per_agent) MVB_CARVE(s->mem, rp.pa, 20 * (size_t)NP);
            MVB_CARVE(s->mem, rp.tables_f, sizeof(float) * 4 * (rp.H + rp.S));
            MVB_CARVE(s->mem, rp.tables_u, sizeof(uint32_t) * 5 * rp.H);
            MVB_CARVE(s->mem, rp.cong, sizeof(float) * 4 * rp.H);
            MVB_CARVE(s->mem, rp.hub_cnt, hub_cnt_bytes);
            if (g->push_chunks) MVB_CARVE(s->mem, rp.push_cnt, 8 * (size_t)g->push_own_slots);
            if (si.ev) MVB_CUDA(cudaStreamWaitEvent(S, si.ev, 0));
            // tables + residents per neighbourhood (counted on the device, neighbourhood.rs:71)
            if (int rc = upload_tables(s, rp, t)) return rc;
            uint32_t* d_nres = rp.tables_u.as<uint32_t>() + 4 * rp.H;
            MVB_CUDA(cudaMemsetAsync(d_nres, 0, 4 * rp.H, S));
            validate_state_kernel<<<(N + 255) / 256, 256, 0, S>>>(N, rp.S, rp.H, si.sub, si.nbhd, si.commute, si.mode, si.norm, d_nres,
                                                                rep_flags.as<SetupFlag>() + r);
            InitArgs ia{};
            ia.int2ext = g->int2ext.as<uint32_t>(); ia.N = N; ia.N_pad = NP;
            ia.sub = si.sub; ia.nbhd = si.nbhd; ia.commute = si.commute; ia.car = si.car; ia.bike = si.bike; ia.ws = si.ws;
            ia.habit = si.habit; ia.mode = si.mode; ia.norm = si.norm;
            for (int k = 0; k < 5; ++k) { ia.pa_src[k] = si.pa[k]; ia.pa_def[k] = pa_def[k]; }
            ia.lens = g->deg.as<uint32_t>(); ia.state = rp.state.as<uint32_t>();
            ia.pa_out = rp.per_agent ? rp.pa.as<float>() : nullptr;
            ia.vec0 = rp.vec[0].as<uint2>(); ia.vec1 = rp.vec[1].as<uint2>(); ia.normvec = rp.normvec.as<uint2>();
            init_state_kernel<<<(NP + 255) / 256, 256, 0, S>>>(ia);
            MVB_CUDA(cudaGetLastError());
            for (UpBuf& b : si.b) b.release(S);
        }
        host_f.emplace_back(4 * rp.H, 1.0f);          // default_congestion_modifier, neighbourhood.rs:34-41; lives until the final synchronize
        MVB_CUDA(cudaMemcpyAsync(rp.cong.p, host_f.back().data(), 16 * rp.H, cudaMemcpyHostToDevice, S));

        RepDev& d = s->h_reps[r];
        d.N = N; d.N_pad = NP; d.S = rp.S; d.H = rp.H; d.rec_off = rp.rec_off;
        d.n_hub_slices = L.n_hub_slices; d.n_sell_slices = L.n_sell_slices;
        d.n_ranges = L.n_common_ranges; d.hot_words = g->common_words;      // what every block stages (LOCAL: + one more range, of its neighbourhood)
        d.hub_begin = 0; d.hub_end = L.n_hub_slices; d.sell_begin = 0; d.sell_end = L.n_sell_slices;
        if (world > 1) { d.hub_begin = s->own[0]; d.hub_end = s->own[1]; d.sell_begin = s->own[2]; d.sell_end = s->own[3]; }
        d.state = rp.state.as<uint32_t>();
        d.pa = rp.per_agent ? rp.pa.as<float>() : nullptr;
        d.vec[0] = rp.vec[0].as<uint2>(); d.vec[1] = rp.vec[1].as<uint2>(); d.normvec = rp.normvec.as<uint2>();
        d.ranges = g->ranges.as<StageRange>();
        if (rp.hub_cnt.bytes) MVB_CUDA(cudaMemsetAsync(rp.hub_cnt.p, 0, rp.hub_cnt.bytes, S));
        d.hub_units = g->hub_units.as<uint4>(); d.hub_unit_begin = g->hub_unit_begin.as<uint32_t>();
        d.hub_off = g->hub_off.as<uint64_t>(); d.hub_deg = g->hub_deg.as<uint4>(); d.hub_cnt = rp.hub_cnt.as<uint32_t>();
        d.hub_codes = g->hub_codes.as<uint32_t>();
        d.sell_meta = g->sell_meta.as<uint4>(); d.sell_codes = g->sell_codes.as<uint32_t>();
        d.supp = rp.tables_f.as<float>(); d.desir = rp.tables_f.as<float>() + 4 * rp.H;
        d.cap = rp.tables_u.as<uint32_t>(); d.nres = rp.tables_u.as<uint32_t>() + 4 * rp.H;
        d.cong = rp.cong.as<float>();
        d.int2ext = g->int2ext.as<uint32_t>();
        if (g->push_chunks) s->solo.push_cnt = rp.push_cnt.as<uint2>() - g->push_first_slot;      // indexed by internal slot (R == 1 in push mode)
        d.local_mode = L.mode == kLayoutLocal ? 1u : 0u;
        // cold gathers bypass L1 when the vector is several times what shared memory + L1 hold (day_kernels.cuh: ld_cold;
        // tools/ab_cold_cg.sh: 8M agents -4.8 %, 2M +0.6 %, 1M +1 % with the bypass)
        d.cold_cg = (size_t)NP / 4 > (1u << 20) ? 1u : 0u;
        if (const char* e = getenv("MVB_COLD_CG")) d.cold_cg = atoi(e) ? 1u : 0u;
        d.nb_slice_begin = g->nb_slice_begin.as<uint32_t>();
    }
    const double tm3 = now();
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    if (R == 1) {      // the SOLO instantiations (push mode, L1-bypassing cold gathers, fused exchange)
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
        MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    }
    MVB_CUDA(cudaFuncSetAttribute(push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kPushChunkMax)));
    if (s->census_smem > 48 * 1024)
        MVB_CUDA(cudaFuncSetAttribute(census_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->census_smem));
    s->rank = rank; s->world = world;
    if (world > 1 && !getenv("MVB_DEBUG_NO_EXCHANGE")) {      // (the debug switch lets ONE process time any rank's share of the work)
        Nccl* n = nccl();
        if (!n) { set_error("agent-partitioned mode needs NCCL (libnccl.so.2 could not be loaded: %s)", dlerror()); return MVB_ERR_COMM; }
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof id);
        MVB_NCCL(n->CommInitRank(&s->comm, (int)world, id, (int)rank));
        if (int rc = setup_peer_exchange(s, vec_piece, S)) return rc;      // sets s->exchange_mode: P2P, or NCCL if any rank could not map the others
        if (s->exchange_mode == MVB_EXCHANGE_P2P) s->solo.px = s->d_px.as<PeerExchange>();
        for (uint32_t k = 0; k < world; ++k) s->xchg_words = std::max(s->xchg_words, s->bounds[k + 1] - s->bounds[k]);
        if (s->exchange_mode == MVB_EXCHANGE_NCCL) MVB_CUDA(s->xchg.alloc((size_t)world * 2 * s->xchg_words * sizeof(uint2)));
        MVB_CUDA(s->d_bounds.alloc(4 * ((size_t)world + 1)));
        MVB_CUDA(cudaMemcpyAsync(s->d_bounds.p, s->bounds.data(), 4 * ((size_t)world + 1), cudaMemcpyHostToDevice, S));
    }
    // LOCAL layout: the parts of a launch are neighbourhood-pure.  Every neighbourhood gets a share of the T parts per
    // replica in proportion to the slices of it this rank updates (all of them unless agent-partitioned), so that parts
    // cost about the same; T = a few rounds of the grid, enough slices per warp to amortise a part's staging.
    {
        const uint32_t n_pieces = (R + kMaxRepsPerLaunch - 1) / kMaxRepsPerLaunch;
        s->local_P.assign(n_pieces, 0);
        std::vector<uint4x> tables;
        std::vector<size_t> table_at(R, 0);
        for (uint32_t c = 0; c < n_pieces; ++c) {
            const uint32_t r0 = c * kMaxRepsPerLaunch, r1 = std::min(R, r0 + kMaxRepsPerLaunch);
            uint32_t first_local = r1;
            for (uint32_t r = r1; r-- > r0;) if (s->h_reps[r].local_mode) first_local = r;
            if (first_local == r1) continue;                     // no replica of this launch uses the LOCAL layout
            // (replicas in the GLOBAL layout that share the launch take class p of the same P = T parts: any P works for them)
            const RepDev& d0 = s->h_reps[first_local];
            const uint32_t own0 = d0.sell_end - d0.sell_begin;
            // ONE round of the grid: every extra round costs each block another staging pass and another ragged end (measured
            // with MVB_LOCAL_ROUNDS = 1, 2, 6 on one B200: 3M agents 0.194 / 0.200 / 0.299 ms per weekday, 6M 0.356 / 0.382 /
            // 0.426, 8M 0.487 / 0.512 / 0.561, 50M 4.28 / 4.41 / 4.64); the work model below balances the parts instead
            uint32_t rounds = 1;
            (void)own0;
            if (const char* e = getenv("MVB_LOCAL_ROUNDS")) rounds = (uint32_t)std::max(1, atoi(e));       // tuning experiments
            uint32_t T = std::max<uint32_t>(d0.H, rounds * s->grid / (r1 - r0));
            for (uint32_t r = r0; r < r1; ++r) if (s->h_reps[r].local_mode) T = std::max(T, s->h_reps[r].H);
            s->local_P[c] = T;
            for (uint32_t r = r0; r < r1; ++r) {
                const RepDev& d = s->h_reps[r];
                if (!d.local_mode) continue;
                const std::vector<uint32_t>& nb = s->reps[r].graph->L.nb_slice_begin;
                // share of a neighbourhood = modelled work of the slices of it this rank updates (the slice dimensions are on the
                // host in the agent-partitioned mode, where ranges cut through neighbourhoods and a neighbourhood's first
                // slices cost far less than its last, cold-heavy ones); whole neighbourhoods: their slice counts
                const GraphLayout& GL = s->reps[r].graph->L;
                std::vector<uint64_t> n_h(d.H);
                std::vector<uint32_t> Q(d.H, 0);
                uint64_t total = 0;
                for (uint32_t h = 0; h < d.H; ++h) {
                    const uint32_t b = std::max(nb[h], d.sell_begin), e = std::min(nb[h + 1], d.sell_end);
                    n_h[h] = 0;
                    if (e > b) {
                        if (GL.sell_K.size() == GL.n_sell_slices) for (uint32_t k = b; k < e; ++k) n_h[h] += sell_slice_work(GL, k);
                        else n_h[h] = e - b;
                    }
                    total += n_h[h];
                }
                uint32_t used = 0;
                for (uint32_t h = 0; h < d.H; ++h)
                    if (n_h[h]) { Q[h] = std::max<uint32_t>(1, (uint32_t)((uint64_t)T * n_h[h] / std::max<uint64_t>(total, 1))); used += Q[h]; }
                // hand out what rounding left (or take back what the minimum of one part per neighbourhood overdrew)
                for (uint32_t guard = 0; used != T && guard < 4 * T + 4 * d.H; ++guard) {
                    uint32_t best = d.H; double best_v = used < T ? -1.0 : 1e300;
                    for (uint32_t h = 0; h < d.H; ++h) {
                        if (!n_h[h]) continue;
                        const double per_part = (double)n_h[h] / Q[h];
                        if (used < T ? per_part > best_v : (Q[h] > 1 && per_part < best_v)) { best_v = per_part; best = h; }
                    }
                    if (best == d.H) break;
                    if (used < T) { ++Q[best]; ++used; } else { --Q[best]; --used; }
                }
                table_at[r] = tables.size();
                for (uint32_t h = 0; h < d.H; ++h)
                    for (uint32_t q = 0; q < Q[h]; ++q) tables.push_back(uint4x{h, q, Q[h], 0});
                while (tables.size() - table_at[r] < T) tables.push_back(uint4x{0, 0xFFFFFFFFu, 1, 0});     // an empty part: class q beyond every slice count
            }
        }
        if (!tables.empty()) {
            MVB_CARVE(s->mem, s->part_tables, 16 * tables.size());
            MVB_CUDA(cudaMemcpyAsync(s->part_tables.p, tables.data(), 16 * tables.size(), cudaMemcpyHostToDevice, S));
            MVB_CUDA(cudaStreamSynchronize(S));                    // `tables` is pageable and goes out of scope
            for (uint32_t r = 0; r < R; ++r)
                if (s->h_reps[r].local_mode) s->h_reps[r].part_table = s->part_tables.as<uint4>() + table_at[r];
        }
    }
    s->rep_args.assign((R + kMaxRepsPerLaunch - 1) / kMaxRepsPerLaunch, RepArray{});
    for (uint32_t r = 0; r < R; ++r) s->rep_args[r / kMaxRepsPerLaunch].r[r % kMaxRepsPerLaunch] = s->h_reps[r];
    MVB_CUDA(s->d_reps.alloc(sizeof(RepDev) * R));
    int rc = upload_reps(s);
    if (rc) return rc;
    std::vector<IvRep> ivreps(R);
    for (uint32_t r = 0; r < R; ++r) {
        const Replica& rp = s->reps[r];
        ivreps[r] = IvRep{rp.graph->int2ext.as<uint32_t>(), rp.state.as<uint32_t>(), rp.tables_f.as<float>(), rp.tables_u.as<uint32_t>(),
                          rp.graph->L.N_pad, rp.S, rp.H};
    }
    MVB_CUDA(s->d_ivreps.alloc(sizeof(IvRep) * R));
    MVB_CUDA(cudaMemcpyAsync(s->d_ivreps.p, ivreps.data(), sizeof(IvRep) * R, cudaMemcpyHostToDevice, S));   // pageable: staged before return
    MVB_CUDA(cudaEventCreateWithFlags(&s->iv_ev, cudaEventDisableTiming));
    rc = ensure_rows(s, 8);
    if (rc) return rc;
    s->cursor = 0;
    s->row_days.assign(1, 0);
    rc = launch_census(s, 0);
    if (rc) return rc;
    std::vector<SetupFlag> flags(R);
    MVB_CUDA(cudaMemcpyAsync(flags.data(), rep_flags.p, sizeof(SetupFlag) * R, cudaMemcpyDeviceToHost, S));
    rep_flags.release(S);
    MVB_CUDA(cudaStreamSynchronize(S));
    for (uint32_t r = 0; r < R; ++r)
        if (flags[r].code) { set_error("replica %u: agent %u: %s", r, flags[r].index, setup_error_text(flags[r].code)); return MVB_ERR_ARG; }
    const double tm4 = now();
    {   // Scratch stays in the device's stream-ordered pool for the next handle of the process, up to 2 GB (what 64
        // populations of 1M agents need; giving it back costs ~6 ms per create and as much again to get it the next time).
        // Anything above that goes back to the driver at the pool's next synchronisation; MVB_TRIM_POOL=1 returns all now.
        cudaMemPool_t pool;
        MVB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = 2ull << 30;
        if (getenv("MVB_TRIM_POOL")) { keep = 0; }
        MVB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        if (!keep) MVB_CUDA(cudaMemPoolTrimTo(pool, 0));
    }
    if (timing)
        fprintf(stderr, "mvb_create: %u replicas, %u graphs: before the first replica %.3f s, enqueue + layout %.3f s (of which waiting for the device %.3f s), "
                "drain %.3f s, pool trim %.3f s\n", R, G, tm_first - tm0, tm3 - tm_first, tm_wait, tm4 - tm3, now() - tm4);
    *out = sp.release();
    return MVB_OK;
}

// sim.cu — C ABI of libmotivate_b200 (include/motivate_b200.h): handle lifecycle, the
// weekday loop of src/simulation.rs:95-136 driven on one CUDA stream, state access.
// Kernels live in day_kernels.cuh.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>          // types and enums only: the functions are resolved with dlsym (no link-time dependency)

#include <algorithm>
#include <cmath>
#include <numeric>
#include <thread>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <new>
#include <vector>

#include "day_kernels.cuh"
#include "fill_kernels.cuh"
#include "setup_kernels.cuh"
#include "layout.h"
#include "mvb_internal.h"

using namespace mvb;

namespace {

struct DevBuf {                       // device memory: its own allocation, or a piece of an Arena
    void* p = nullptr;
    size_t bytes = 0;
    bool owned = true;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), owned(o.owned) { o.p = nullptr; o.bytes = 0; }
    ~DevBuf() { if (p && owned) cudaFree(p); }
    cudaError_t alloc(size_t n) {
        if (p && owned) cudaFree(p);
        p = nullptr; owned = true;
        bytes = n;
        return cudaMalloc(&p, n ? n : 1);
    }
    void view(void* ptr, size_t n) {
        if (p && owned) cudaFree(p);
        p = ptr; bytes = n; owned = false;
    }
    template <class T> T* as() const { return (T*)p; }
};

// Device memory of a handle: a few large allocations cut into 256-byte-aligned pieces.  On a B200 box one 10 GB
// cudaMalloc + cudaFree costs about 5 ms, the same memory as 128 separate buffers 100-300 ms
// (tools/microbench_alloc.cu), and tear-down of many replicas was dominated by it.
struct Arena {
    std::vector<DevBuf> chunks;
    size_t used = 0, chunk_bytes = 256u << 20;                 // size of the next chunk (a request larger than it gets its own)
    static size_t piece(size_t n) { return (n + 255) & ~(size_t)255; }
    cudaError_t carve(DevBuf& b, size_t n) {
        const size_t need = piece(n);
        if (chunks.empty() || used + need > chunks.back().bytes) {
            DevBuf c;
            cudaError_t e = c.alloc(std::max(chunk_bytes, need));
            if (e != cudaSuccess) return e;
            chunks.push_back(std::move(c));
            used = 0;
            if (chunks.size() == 1) chunk_bytes = std::max<size_t>(256u << 20, chunk_bytes / 8);   // an estimate that fell short grows in steps
        }
        b.view((char*)chunks.back().p + used, n);
        used += need;
        return cudaSuccess;
    }
};
#define MVB_CARVE(arena, buf, n) MVB_CUDA((arena).carve((buf), (n)))

// The link graph of one population in device layout (layout.h), shareable between replicas
// (scenario sweeps over one population pass the same graph + neighbourhood arrays).
struct Graph {
    GraphLayout L;                 // host copy: permutation, sizes (code arrays are released after upload)
    DevBuf deg, ranges, hub_off, hub_deg, hub_codes, hub_units, hub_unit_begin, sell_meta, sell_codes, int2ext, nb_slice_begin;
    const void* key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // caller pointers it was built from
    uint32_t common_words = 0;     // words of the staged vector every block copies (layout.h)
    // push mode (layout.h): the cold links of the owned SELL agents, chunk by chunk, sorted by target
    DevBuf push_code, push_src, push_chunk_begin, push_chunk_first;
    uint32_t push_chunks = 0, push_chunk_agents = 0, push_own_slots = 0, push_links = 0, push_first_slot = 0;
};

struct Replica {
    uint32_t N = 0, S = 0, H = 0;
    uint32_t rec_width = 0, rec_off = 0;
    std::shared_ptr<Graph> graph;
    DevBuf state, vec[2], normvec, pa, tables_f, tables_u, cong, hub_cnt;     // state = slice records (layout.h)
    DevBuf push_cnt;                     // push mode: {social, neighbour} counts of every owned SELL agent's cold links, per weekday
    bool per_agent = false;
};

// NCCL, loaded on first use: the replica path (no communication) never needs it.
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl* nccl() {
    static Nccl n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (n.lib) {
#define MVB_SYM(f) n.f = reinterpret_cast<decltype(n.f)>(dlsym(n.lib, "nccl" #f))
            MVB_SYM(GetUniqueId); MVB_SYM(CommInitRank); MVB_SYM(CommDestroy); MVB_SYM(GroupStart); MVB_SYM(GroupEnd);
            MVB_SYM(Broadcast); MVB_SYM(AllReduce); MVB_SYM(AllGather); MVB_SYM(GetErrorString);
#undef MVB_SYM
            if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.GroupStart || !n.GroupEnd || !n.Broadcast ||
                !n.AllReduce || !n.AllGather || !n.GetErrorString) n.lib = nullptr;
        }
    }
    return n.lib ? &n : nullptr;
}
#define MVB_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            mvb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, nccl()->GetErrorString(r_)); \
            return MVB_ERR_COMM;                                                                    \
        }                                                                                           \
    } while (0)

}  // namespace

struct mvb_sim {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    mvb_params prm{};
    Arena mem;                           // device memory of the graphs and replicas
    std::vector<Replica> reps;
    DevBuf d_reps;                       // RepDev[R]
    std::vector<RepDev> h_reps;
    std::vector<RepArray> rep_args;      // h_reps cut into kernel-parameter sized pieces
    std::vector<uint32_t> local_P;       // per piece: parts per replica when its replicas use the LOCAL layout (0 = GLOBAL)
    DevBuf part_tables;
    uint32_t rec_stride = 0;             // u32 per record row (all replicas)
    uint32_t grid = 0;                   // persistent blocks (one per SM)
    uint32_t max_npad = 0;
    uint32_t vec_cap_words = 0;          // shared-memory words reserved for the staged mode vector
    size_t smem_bytes = 0, census_smem = 0;
    bool any_per_agent = false;
    // record ring: rows [0, rec_cap); row `cursor` holds the census of the current state
    DevBuf rec;
    uint32_t rec_cap = 0, cursor = 0;
    std::vector<uint32_t> row_days;      // Day column of rows 0..cursor
    uint32_t parity = 0;                 // which mode buffer holds current_mode
    uint8_t prev_weather = 0;
    bool stepped = false;
    float last_ms = 0.f;
    uint64_t last_launches = 0;
    bool timing_valid = false;
    // agent-partitioned mode (one population over `world` GPUs)
    uint32_t rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    std::vector<uint32_t> bounds;        // [world+1] global slice index where each rank's range starts
    uint32_t own[4] = {0, 0, 0, 0};      // this rank's hub slices [0],[1]) and SELL slices [2],[3])
    DevBuf stamps;                       // debug (MVB_DEBUG_STAMPS): {first block start, last block end} in ns per day-kernel launch
    uint32_t stamp_first = 0, stamp_cap = 0;
    DevBuf xchg, d_bounds;               // exchange buffer [world][2][xchg_words] u64 words (modes, norms); bounds on the device
    uint32_t xchg_words = 0;             // longest rank range, in slices
    // exchange over peer memory (MVB_EXCHANGE_P2P): this rank's exported buffers, the other ranks' buffers mapped here
    int exchange_mode = MVB_EXCHANGE_NONE;
    DevBuf px_vec, px_ctl, d_px;         // {vec[0], vec[1], normvec}; {landing zone for partial records, flags, done, error, scratch}; PeerExchange
    std::vector<void*> px_open;          // cudaIpcOpenMemHandle results, closed by mvb_destroy
    uint32_t epoch = 0;                  // day-kernel launches so far: the flag value a rank raises when it has finished one
    SoloExtra solo{nullptr, nullptr, 0, 0};   // what a handle with ONE population passes to the day kernel besides the replica descriptors
    // interventions of the current run, staged before the day loop (apply_interventions_kernel)
    DevBuf d_ivreps;                     // IvRep[R]
    char* iv_pinned = nullptr;           // library-owned pinned staging: the caller's arrays are only borrowed for the call
    size_t iv_pinned_cap = 0;
    cudaEvent_t iv_ev = nullptr;         // the copy out of iv_pinned has finished
};

namespace {

constexpr uint32_t kBlock = 256;

int upload_reps(mvb_sim* s) {
    MVB_CUDA(cudaMemcpyAsync(s->d_reps.p, s->h_reps.data(), sizeof(RepDev) * s->h_reps.size(),
                             cudaMemcpyHostToDevice, s->stream));
    return MVB_OK;
}

// Make sure rows cursor+1 .. cursor+n exist and are zero.
int ensure_rows(mvb_sim* s, uint32_t n) {
    if ((uint64_t)s->cursor + n < s->rec_cap) return MVB_OK;
    const uint32_t new_cap = std::max<uint64_t>((uint64_t)s->cursor + n + 1, (uint64_t)s->rec_cap * 2);
    DevBuf nb;
    MVB_CUDA(nb.alloc(sizeof(uint32_t) * (size_t)new_cap * s->rec_stride));
    MVB_CUDA(cudaMemsetAsync(nb.p, 0, nb.bytes, s->stream));
    if (s->rec.p)
        MVB_CUDA(cudaMemcpyAsync(nb.p, s->rec.p, sizeof(uint32_t) * (size_t)(s->cursor + 1) * s->rec_stride,
                                 cudaMemcpyDeviceToDevice, s->stream));
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    std::swap(s->rec.p, nb.p);
    std::swap(s->rec.bytes, nb.bytes);
    s->rec_cap = new_cap;
    return MVB_OK;
}

int launch_census(mvb_sim* s, uint32_t row) {
    uint32_t* rec = s->rec.as<uint32_t>() + (size_t)row * s->rec_stride;
    MVB_CUDA(cudaMemsetAsync(rec, 0, sizeof(uint32_t) * s->rec_stride, s->stream));
    dim3 grid((s->max_npad + kBlock - 1) / kBlock, (unsigned)s->reps.size());
    census_kernel<<<grid, kBlock, s->census_smem, s->stream>>>(s->d_reps.as<RepDev>(), s->parity, rec);
    MVB_CUDA(cudaGetLastError());
    return MVB_OK;
}

// How many parts to cut each of the R replicas of one launch into.  A block takes parts b, b+grid, ... and a part
// costs a fixed set-up (staging the mode vector, tables, two barriers) plus its slices, 32 at a time (one per warp).
// Few parts per replica leave blocks idle, many make the set-up and the last, partly filled round of slices dominate
// (small populations); pick the count with the smallest modelled launch time.
uint32_t choose_parts(const mvb_sim* s, uint32_t first_rep, uint32_t R) {
    if (const char* e = getenv("MVB_PARTS")) return (uint32_t)std::max(1, atoi(e));       // tuning experiments
    double slices = 0, vec_bytes = 0;
    for (uint32_t r = 0; r < R; ++r) {
        const RepDev& d = s->h_reps[first_rep + r];
        slices += (d.hub_end - d.hub_begin) * 4.0 + (d.sell_end - d.sell_begin);      // a hub slice is worth a few SELL slices
        vec_bytes += d.hot_words * 4.0;
    }
    slices /= R; vec_bytes /= R;
    const double t_slice = 7.5, t_fixed = 10.0 + vec_bytes / 150e3;                    // microseconds (calibrated on B200)
    uint32_t best = 1;
    double best_t = 1e300;
    for (uint32_t P = 1; P <= s->grid; ++P) {
        const double rounds = std::ceil((double)R * P / s->grid);
        const double t = rounds * (t_fixed + (slices / P / 32.0 + 0.5) * t_slice);        // + half a round: the ragged end
        if (t < best_t * 0.999) { best_t = t; best = P; }
    }
    return best;
}

// with_norms (agent-partitioned mode): also exchange the norm words.  No rank's update reads another agent's norm
// (statistics.rs counts each agent's own), so a run sends them once, with its last weekday, not every day.
int launch_day(mvb_sim* s, uint8_t weather, uint8_t changed, bool with_norms = true) {
    int rc = ensure_rows(s, 1);
    if (rc) return rc;
    const uint32_t* prev = s->rec.as<uint32_t>() + (size_t)s->cursor * s->rec_stride;
    uint32_t* next = s->rec.as<uint32_t>() + (size_t)(s->cursor + 1) * s->rec_stride;
    AgentWeights uw{s->prm.consistency, s->prm.social_connectivity, s->prm.subculture_connectivity,
                    s->prm.neighbourhood_connectivity, s->prm.average_weight};
    // replica descriptors travel as kernel parameters, kMaxRepsPerLaunch at a time
    // Programmatic dependent launch: the next weekday's blocks may start (and run their day-independent prologue) while
    // this one's last blocks finish; the kernel waits with griddepcontrol.wait before it touches anything a previous
    // launch wrote (day_kernels.cuh).  Not in the agent-partitioned mode over NCCL, where NCCL's kernels sit between the days.
    static const bool no_pdl = getenv("MVB_NO_PDL") != nullptr;
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(s->grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = s->smem_bytes; cfg.stream = s->stream;
    cfg.attrs = pdl_attr; cfg.numAttrs = ((s->world == 1 || s->exchange_mode == MVB_EXCHANGE_P2P) && !no_pdl) ? 1 : 0;
    const bool p2p = s->exchange_mode == MVB_EXCHANGE_P2P;
    SoloExtra solo = s->solo;                                    // push mode / agent-partitioned mode: handles with one population
    if (p2p) { solo.px_epoch = ++s->epoch; solo.px_norms = with_norms ? 1u : 0u; }
    // SOLO instantiation (day_kernels.cuh): one population per handle with push mode, L1-bypassing cold gathers or the fused exchange
    const bool solo_kernel = s->reps.size() == 1 && (solo.push_cnt || p2p || s->h_reps[0].cold_cg);
#define MVB_LAUNCH_X(PA, T, G, SO, X) MVB_CUDA(cudaLaunchKernelEx(&cfg, day_kernel<PA, T, G, SO, X>, s->rep_args[c], R, P, s->parity, prev, (uint32_t*)next, \
        (uint32_t)weather, (uint32_t)changed, uw, s->vec_cap_words, solo))
#define MVB_LAUNCH(PA, T, G) do { if (p2p) MVB_LAUNCH_X(PA, T, G, true, true); else if (solo_kernel) MVB_LAUNCH_X(PA, T, G, true, false); \
                                  else MVB_LAUNCH_X(PA, T, G, false, false); } while (0)
    // push mode (one large population, layout.h): the cold links of the SELL agents are counted first, by push_kernel
    for (Replica& rp : s->reps) {
        const Graph& g = *rp.graph;
        if (!g.push_chunks) continue;
        PushDesc pd{g.push_code.as<uint32_t>(), g.push_src.as<uint16_t>(), g.push_chunk_begin.as<uint32_t>(), g.push_chunk_first.as<uint32_t>(),
                    rp.push_cnt.as<uint2>(), g.push_chunks, s->h_reps[&rp - s->reps.data()].cold_cg};
        push_kernel<<<std::min(s->grid, g.push_chunks), 1024, 8 * (size_t)g.push_chunk_agents, s->stream>>>(pd, rp.vec[s->parity].as<uint32_t>());
        MVB_CUDA(cudaGetLastError());
    }
    for (size_t c = 0; c < s->rep_args.size(); ++c) {
        const uint32_t R = std::min<uint32_t>(kMaxRepsPerLaunch, (uint32_t)s->reps.size() - (uint32_t)c * kMaxRepsPerLaunch);
        const uint32_t P = s->local_P[c] ? s->local_P[c] : choose_parts(s, (uint32_t)c * kMaxRepsPerLaunch, R);
        // work items (slices) a warp of a part gets: with many, they are claimed four at a time, else two (day_kernels.cuh)
        const RepDev& r0 = s->h_reps[c * kMaxRepsPerLaunch];
        const uint32_t per_warp = ((r0.hub_end - r0.hub_begin) + (r0.sell_end - r0.sell_begin)) / P / 32;
        const char* force = getenv("MVB_CLAIM_GROUP");          // tests pin the instantiation: "2" or "4"
        const bool many = force ? atoi(force) >= 4 : per_warp >= 32;
        if (s->any_per_agent) { if (many) MVB_LAUNCH(true, 1024, 4); else MVB_LAUNCH(true, 1024, 2); }
        else                  { if (many) MVB_LAUNCH(false, 1024, 4); else MVB_LAUNCH(false, 1024, 2); }
        MVB_CUDA(cudaGetLastError());
        s->last_launches++;
    }
#undef MVB_LAUNCH
#undef MVB_LAUNCH_X
    if (p2p) {
        // the exchange itself happened inside the day kernel (stores into every rank's vectors); what is left is to wait
        // for the other ranks' flags and to add the partial records up (day_kernels.cuh)
        exchange_finish_kernel<<<1, 256, 0, s->stream>>>(s->d_px.as<PeerExchange>(), s->epoch, next);
        MVB_CUDA(cudaGetLastError());
    }
    static const bool skip_exchange = getenv("MVB_DEBUG_NO_EXCHANGE") != nullptr;      // timing experiments only: results are wrong
    if (s->world > 1 && !skip_exchange && !p2p) {
        // agent-partitioned, NCCL path: every rank now holds new modes / norms for its own slices and a partial record.  The
        // social network is global (simulation.rs:160), so tomorrow every rank needs everybody's modes: ONE all-gather
        // of the packed 2-bit words (each rank's piece: its mode words, padded to the longest range, and on the last
        // weekday of a run its norm words behind them) and ONE all-reduce of the day's counters.
        Replica& rp = s->reps[0];
        Nccl* n = nccl();
        uint2* vec_next = rp.vec[s->parity ^ 1u].as<uint2>();
        uint2* normvec = rp.normvec.as<uint2>();
        const uint32_t M = s->xchg_words;
        const uint32_t piece = with_norms ? 2 * M : M;               // u64 words per rank in the all-gather buffer
        uint2* send = s->xchg.as<uint2>() + (size_t)s->rank * piece;
        const uint32_t n_slices = s->bounds[s->world];
        pack_exchange_kernel<<<(s->bounds[s->rank + 1] - s->bounds[s->rank] + 255) / 256, 256, 0, s->stream>>>(
            vec_next, with_norms ? normvec : nullptr, s->bounds[s->rank], s->bounds[s->rank + 1], M, send);
        MVB_NCCL(n->AllGather(send, s->xchg.p, (size_t)piece * sizeof(uint2), ncclUint8, s->comm, s->stream));      // in place
        unpack_exchange_kernel<<<(n_slices + 255) / 256, 256, 0, s->stream>>>(s->xchg.as<uint2>(), s->d_bounds.as<uint32_t>(), s->world, s->rank, M, piece,
                                                                            n_slices, vec_next, with_norms ? normvec : nullptr);
        MVB_CUDA(cudaGetLastError());
        MVB_NCCL(n->AllReduce(next, next, s->rec_stride, ncclUint32, ncclSum, s->comm, s->stream));
    }
    s->cursor++;
    s->parity ^= 1u;
    s->stepped = true;
    return MVB_OK;
}

// Device scratch of the create path: comes from the device's stream-ordered pool and goes back to it in stream
// order, so releasing it never waits for the device.
struct TmpBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    TmpBuf() = default;
    TmpBuf(const TmpBuf&) = delete;
    TmpBuf& operator=(const TmpBuf&) = delete;
    ~TmpBuf() { release(); }
    void release() { if (p) cudaFreeAsync(p, st); p = nullptr; }
    cudaError_t alloc(size_t n, cudaStream_t stream) {
        release();
        st = stream;
        return cudaMallocAsync(&p, n ? n : 1, stream);
    }
    template <class T> T* as() const { return (T*)p; }
};

// Pinned host buffers are slow to make (about 15 ms for 9 MB on the B200 boxes), so the few that set-up needs are
// kept for the next handle of the process (at most 12; never returned to the driver before exit).
struct PinnedCache {
    std::mutex m;
    std::vector<std::pair<char*, size_t>> idle;
    cudaError_t get(size_t bytes, char** out, size_t* cap) {
        {
            std::lock_guard<std::mutex> lk(m);
            for (size_t k = 0; k < idle.size(); ++k)
                if (idle[k].second >= bytes) { *out = idle[k].first; *cap = idle[k].second; idle.erase(idle.begin() + k); return cudaSuccess; }
        }
        *cap = bytes + bytes / 8;
        return cudaHostAlloc((void**)out, *cap, cudaHostAllocDefault);
    }
    void put(char* p, size_t cap) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(m);
        if (idle.size() < 12) idle.emplace_back(p, cap); else cudaFreeHost(p);
    }
};
static PinnedCache& pinned_cache() { static PinnedCache* c = new PinnedCache(); return *c; }

// the u64 of one slice (lanes 0-15 in .x, 16-31 in .y, as finish_slice packs them) -> 32 two-bit values
void unpack_slices(const std::vector<uint32_t>& int2ext, const std::vector<uint2>& in, uint8_t* values) {
    for (size_t i = 0; i < int2ext.size(); ++i) {
        if (int2ext[i] == kInvalid) continue;
        const uint32_t l = (uint32_t)(i & 31);
        const uint2 w = in[i >> 5];
        values[int2ext[i]] = (uint8_t)(((l < 16 ? w.x : w.y) >> ((l & 15u) * 2u)) & 3u);
    }
}

// The internal order lives on the device (create.cuh builds it there); the host copy is fetched when a caller asks
// for state in its own agent order.
int host_int2ext(mvb_sim* s, Graph& g) {
    if (!g.L.int2ext.empty() || g.L.N_pad == 0) return MVB_OK;
    MVB_CUDA(cudaSetDevice(s->device));
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    g.L.int2ext.resize(g.L.N_pad);
    MVB_CUDA(cudaMemcpy(g.L.int2ext.data(), g.int2ext.p, 4 * (size_t)g.L.N_pad, cudaMemcpyDeviceToHost));
    return MVB_OK;
}

int check_tables(const mvb_tables* t) {
    if (!t || !t->supportiveness || !t->capacity || !t->desirability || t->n_subcultures == 0 ||
        t->n_subcultures > 256 || t->n_neighbourhoods == 0 || t->n_neighbourhoods > 65536) {
        set_error("bad tables (need 1..256 subcultures, 1..65536 neighbourhoods, non-null arrays)");
        return MVB_ERR_ARG;
    }
    return MVB_OK;
}

int upload_tables(mvb_sim* s, Replica& r, const mvb_tables* t) {
    // tables_f = supportiveness[H][4] | desirability[S][4];  tables_u = capacity[H][4] | residents[H]
    MVB_CUDA(cudaMemcpyAsync(r.tables_f.p, t->supportiveness, sizeof(float) * 4 * r.H, cudaMemcpyHostToDevice, s->stream));
    MVB_CUDA(cudaMemcpyAsync(r.tables_f.as<float>() + 4 * r.H, t->desirability, sizeof(float) * 4 * r.S,
                             cudaMemcpyHostToDevice, s->stream));
    MVB_CUDA(cudaMemcpyAsync(r.tables_u.p, t->capacity, sizeof(uint32_t) * 4 * r.H, cudaMemcpyHostToDevice, s->stream));
    return MVB_OK;
}

}  // namespace

extern "C" {

int mvb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint32_t mvb_run_rows(uint32_t n_days) {
    uint32_t r = 1;
    for (uint32_t d = 1; d < n_days; ++d) r += (d % 7 < 5);
    return r;
}

void mvb_destroy(mvb_sim* s) {
    if (!s) return;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    const double t1 = now();
    // nobody writes into this rank's memory any more (its last exchange_finish_kernel saw every rank's flag); unmap the
    // other ranks' buffers (also after a set-up that failed half way), and free the exported ones only when every rank has
    for (void* p : s->px_open) cudaIpcCloseMemHandle(p);
    s->px_open.clear();
    if (s->exchange_mode == MVB_EXCHANGE_P2P) {
        uint32_t* scratch = s->px_ctl.as<uint32_t>() + (s->px_ctl.bytes - 256) / 4;
        if (s->comm && nccl() && nccl()->AllReduce(scratch, scratch, 1, ncclUint32, ncclSum, s->comm, s->stream) == ncclSuccess)
            cudaStreamSynchronize(s->stream);
    }
    if (s->comm && nccl()) nccl()->CommDestroy(s->comm);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->iv_ev) cudaEventDestroy(s->iv_ev);
    pinned_cache().put(s->iv_pinned, s->iv_pinned_cap);
    if (s->stream) cudaStreamDestroy(s->stream);
    const double t2 = now();
    s->mem.chunks.clear();
    const double t3 = now();
    delete s;
    if (getenv("MVB_TIMING"))
        fprintf(stderr, "mvb_destroy: stream sync %.3f s, events/stream %.3f s, device memory %.3f s, host memory %.3f s\n",
                t1 - t0, t2 - t1, t3 - t2, now() - t3);
}

}  // extern "C"

namespace {

// Agent-partitioned mode, exchange over peer memory: export this rank's vectors and control block, map everybody else's
// (handles travel through one ncclAllGather), agree that it worked on EVERY rank (else all fall back to the NCCL
// exchange) and describe the result to the kernels (PeerExchange).  Also the barrier after which every rank's vectors are
// initialised and may be written by the others.  s->px_vec holds vec[0], vec[1], normvec at `vec_piece` strides.
int setup_peer_exchange(mvb_sim* s, size_t vec_piece, cudaStream_t S) {
    Nccl* n = nccl();
    const uint32_t world = s->world, rank = s->rank, W = s->rec_stride;
    const char* force = getenv("MVB_EXCHANGE");
    bool ok = world <= kMaxPeers && !(force && !strcmp(force, "nccl"));
    const size_t xrec_bytes = Arena::piece(sizeof(uint32_t) * 2 * (size_t)world * W);
    MVB_CUDA(s->px_ctl.alloc(xrec_bytes + 4 * 256));              // + flags, done, error, scratch (a 256-byte line each)
    MVB_CUDA(cudaMemsetAsync(s->px_ctl.p, 0, s->px_ctl.bytes, S));
    struct Handles { cudaIpcMemHandle_t vec, ctl; };
    Handles mine{};
    if (ok && (cudaIpcGetMemHandle(&mine.vec, s->px_vec.p) != cudaSuccess || cudaIpcGetMemHandle(&mine.ctl, s->px_ctl.p) != cudaSuccess)) {
        ok = false; cudaGetLastError();
    }
    DevBuf d_h;
    MVB_CUDA(d_h.alloc(sizeof(Handles) * world + 256));
    std::vector<Handles> all(world);
    MVB_CUDA(cudaMemcpyAsync(d_h.as<Handles>() + rank, &mine, sizeof mine, cudaMemcpyHostToDevice, S));
    MVB_NCCL(n->AllGather(d_h.as<Handles>() + rank, d_h.p, sizeof(Handles), ncclUint8, s->comm, S));
    MVB_CUDA(cudaMemcpyAsync(all.data(), d_h.p, sizeof(Handles) * world, cudaMemcpyDeviceToHost, S));
    MVB_CUDA(cudaStreamSynchronize(S));
    std::vector<char*> vecs(world, nullptr), ctls(world, nullptr);
    vecs[rank] = s->px_vec.as<char>(); ctls[rank] = s->px_ctl.as<char>();
    for (uint32_t p = 0; ok && p < world; ++p) {
        if (p == rank) continue;
        void *a = nullptr, *b = nullptr;
        if (cudaIpcOpenMemHandle(&a, all[p].vec, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
        s->px_open.push_back(a);
        if (cudaIpcOpenMemHandle(&b, all[p].ctl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
        s->px_open.push_back(b);
        vecs[p] = (char*)a; ctls[p] = (char*)b;
    }
    // every rank or none (also the barrier: all ranks' set-up kernels, enqueued on S before this, have finished)
    uint32_t* d_ok = reinterpret_cast<uint32_t*>(d_h.as<char>() + sizeof(Handles) * world);
    uint32_t h_ok = ok ? 1u : 0u;
    MVB_CUDA(cudaMemcpyAsync(d_ok, &h_ok, 4, cudaMemcpyHostToDevice, S));
    MVB_NCCL(n->AllReduce(d_ok, d_ok, 1, ncclUint32, ncclMin, s->comm, S));
    MVB_CUDA(cudaMemcpyAsync(&h_ok, d_ok, 4, cudaMemcpyDeviceToHost, S));
    MVB_CUDA(cudaStreamSynchronize(S));
    if (!h_ok) {
        for (void* p : s->px_open) cudaIpcCloseMemHandle(p);
        s->px_open.clear();
        s->exchange_mode = MVB_EXCHANGE_NCCL;
        return MVB_OK;
    }
    PeerExchange px{};
    px.world = world; px.rank = rank; px.rec_stride = W;
    for (uint32_t p = 0; p < world; ++p) {
        px.vec[0][p] = reinterpret_cast<uint2*>(vecs[p]);
        px.vec[1][p] = reinterpret_cast<uint2*>(vecs[p] + vec_piece);
        px.normvec[p] = reinterpret_cast<uint2*>(vecs[p] + 2 * vec_piece);
        px.xrec[p] = reinterpret_cast<uint32_t*>(ctls[p]);
        px.flags[p] = reinterpret_cast<uint32_t*>(ctls[p] + xrec_bytes);
    }
    px.done = reinterpret_cast<uint32_t*>(ctls[rank] + xrec_bytes + 256);
    px.error = reinterpret_cast<uint32_t*>(ctls[rank] + xrec_bytes + 512);
    MVB_CUDA(s->d_px.alloc(sizeof px));
    MVB_CUDA(cudaMemcpyAsync(s->d_px.p, &px, sizeof px, cudaMemcpyHostToDevice, S));
    MVB_CUDA(cudaStreamSynchronize(S));                            // `px` is pageable and goes out of scope
    s->exchange_mode = MVB_EXCHANGE_P2P;
    return MVB_OK;
}

// A rank that never raised its flag (exchange_finish_kernel gave up waiting): the results are not valid.
int check_peer_exchange(mvb_sim* s) {
    if (s->exchange_mode != MVB_EXCHANGE_P2P) return MVB_OK;
    uint32_t err = 0;
    MVB_CUDA(cudaMemcpy(&err, s->px_ctl.as<char>() + s->px_ctl.bytes - 512, 4, cudaMemcpyDeviceToHost));
    if (err) { set_error("agent-partitioned run: a rank did not finish a weekday within 20 s (peer-to-peer exchange)"); return MVB_ERR_COMM; }
    return MVB_OK;
}

}  // namespace
#include "create.cuh"
#include "synth_device.cuh"
extern "C" {

int mvb_create_batch(uint32_t R, const mvb_population* const* pops, const mvb_tables* const* tabs,
                     const mvb_params* prm, int device, mvb_sim** out) {
    return create_impl(R, pops, tabs, prm, device, 0, 1, nullptr, out);
}

int mvb_create(const mvb_population* pop, const mvb_tables* tab, const mvb_params* prm, int device, mvb_sim** out) {
    return create_impl(1, &pop, &tab, prm, device, 0, 1, nullptr, out);
}

int mvb_comm_unique_id(void* id_out) {
    if (!id_out) { set_error("null argument"); return MVB_ERR_ARG; }
    Nccl* n = nccl();
    if (!n) { set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror()); return MVB_ERR_COMM; }
    static_assert(sizeof(ncclUniqueId) == MVB_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    MVB_NCCL(n->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return MVB_OK;
}

int mvb_create_partitioned(const mvb_population* pop, const mvb_tables* tab, const mvb_params* prm, int device,
                           uint32_t rank, uint32_t world, const void* unique_id, mvb_sim** out) {
    if (world == 0 || rank >= world) { set_error("rank %u out of range for world %u", rank, world); return MVB_ERR_ARG; }
    if (world > 1 && !unique_id) { set_error("a communicator id is required for world > 1"); return MVB_ERR_ARG; }
    return create_impl(1, &pop, &tab, prm, device, rank, world, unique_id, out);
}

static void owners_from_bounds(const GraphLayout& L, const std::vector<uint32_t>& bounds, uint32_t world, uint32_t* rank_out) {
    for (uint32_t r = 0; r < world; ++r)
        for (size_t slot = (size_t)bounds[r] * 32; slot < (size_t)bounds[r + 1] * 32; ++slot)
            if (L.int2ext[slot] != kInvalid) rank_out[L.int2ext[slot]] = r;
}

int mvb_partition_owner(const mvb_sim* s, uint32_t* rank_out) {
    if (!s || !rank_out) { set_error("null argument"); return MVB_ERR_ARG; }
    if (int rc = host_int2ext(const_cast<mvb_sim*>(s), *s->reps[0].graph)) return rc;
    const GraphLayout& L = s->reps[0].graph->L;
    if (s->world <= 1) { for (uint32_t i = 0; i < L.N; ++i) rank_out[i] = 0; return MVB_OK; }
    owners_from_bounds(L, s->bounds, s->world, rank_out);
    return MVB_OK;
}

int mvb_partition_plan(const mvb_population* pop, uint32_t S, uint32_t H, uint32_t world, uint32_t* rank_out) {
    if (!pop || !rank_out || world == 0 || H == 0 || S == 0) { set_error("bad argument"); return MVB_ERR_ARG; }
    GraphLayout L;
    // same shared-memory capacity as mvb_create* computes on a B200 (227 KB opt-in limit), so the plan is the
    // split a B200 handle really uses
    const uint32_t tail = (tail_words(S, H) + 3u) & ~3u;
    const uint32_t capacity = staged_capacity_agents((232448u / 4u - tail) & ~3u);
    int rc = build_graph_layout(*pop, H, capacity, L);
    if (rc) return rc;
    std::vector<uint32_t> bounds;
    partition_slices(L, world, bounds);
    owners_from_bounds(L, bounds, world, rank_out);
    return MVB_OK;
}

int mvb_exchange_mode(const mvb_sim* s) { return s ? s->exchange_mode : MVB_EXCHANGE_NONE; }
uint32_t mvb_n_replicas(const mvb_sim* s) { return s ? (uint32_t)s->reps.size() : 0; }
uint32_t mvb_stats_width(const mvb_sim* s, uint32_t r) {
    return (s && r < s->reps.size()) ? MVB_STATS_FIXED + s->reps[r].S + s->reps[r].H : 0;
}
void* mvb_stream(mvb_sim* s) { return s ? (void*)s->stream : nullptr; }

int mvb_sync(mvb_sim* s) {
    if (!s) { set_error("null handle"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    return check_peer_exchange(s);
}

int mvb_step(mvb_sim* s, uint8_t weather_today, uint8_t weather_changed) {
    if (!s) { set_error("null handle"); return MVB_ERR_ARG; }
    if (weather_today > 1) { set_error("weather code must be 0 or 1"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    int rc = launch_day(s, weather_today, weather_changed ? 1 : 0);
    if (rc) return rc;
    s->row_days.push_back(s->row_days.back() + 1);
    s->prev_weather = weather_today;
    return MVB_OK;
}

int mvb_set_tables(mvb_sim* s, uint32_t r, const mvb_tables* t) {
    if (!s || r >= s->reps.size()) { set_error("bad handle/replica"); return MVB_ERR_ARG; }
    int rc = check_tables(t);
    if (rc) return rc;
    Replica& rp = s->reps[r];
    if (t->n_subcultures != rp.S || t->n_neighbourhoods != rp.H) { set_error("table shape differs from the replica's"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    return upload_tables(s, rp, t);
}

int mvb_set_ownership(mvb_sim* s, uint32_t r, const uint8_t* car, const uint8_t* bike) {
    if (!s || r >= s->reps.size()) { set_error("bad handle/replica"); return MVB_ERR_ARG; }
    if (!car && !bike) return MVB_OK;
    Replica& rp = s->reps[r];
    const uint32_t NP = rp.graph->L.N_pad;
    MVB_CUDA(cudaSetDevice(s->device));
    TmpBuf tc, tb;                                  // stream-ordered pool: no device-wide synchronisation
    if (car)  { MVB_CUDA(tc.alloc(rp.N, s->stream)); MVB_CUDA(cudaMemcpyAsync(tc.p, car, rp.N, cudaMemcpyHostToDevice, s->stream)); }
    if (bike) { MVB_CUDA(tb.alloc(rp.N, s->stream)); MVB_CUDA(cudaMemcpyAsync(tb.p, bike, rp.N, cudaMemcpyHostToDevice, s->stream)); }
    set_ownership_kernel<<<(NP + 255) / 256, 256, 0, s->stream>>>(rp.graph->int2ext.as<uint32_t>(), NP,
                                                                  car ? tc.as<uint8_t>() : nullptr, bike ? tb.as<uint8_t>() : nullptr,
                                                                  rp.state.as<uint32_t>());
    MVB_CUDA(cudaGetLastError());
    MVB_CUDA(cudaStreamSynchronize(s->stream));   // the caller's buffers and the temporaries are released on return
    return MVB_OK;
}

int mvb_reset_rows(mvb_sim* s) {
    if (!s) { set_error("null handle"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    if (s->cursor != 0) {
        // keep the current census as row 0, zero the rest
        MVB_CUDA(cudaMemcpyAsync(s->rec.p, s->rec.as<uint32_t>() + (size_t)s->cursor * s->rec_stride,
                                 sizeof(uint32_t) * s->rec_stride, cudaMemcpyDeviceToDevice, s->stream));
        MVB_CUDA(cudaMemsetAsync(s->rec.as<uint32_t>() + s->rec_stride, 0,
                                 sizeof(uint32_t) * (size_t)(s->rec_cap - 1) * s->rec_stride, s->stream));
        s->cursor = 0;
    }
    const uint32_t d = s->row_days.back();
    s->row_days.assign(1, d);
    return MVB_OK;
}

int mvb_run_async(mvb_sim* s, const uint8_t* weather, uint32_t first_day, uint32_t end_day,
                  const mvb_intervention* const* ivs) {
    if (!s || !weather) { set_error("null argument"); return MVB_ERR_ARG; }
    if (end_day <= first_day) { set_error("empty day range"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    uint32_t d0 = first_day;
    if (first_day == 0) {
        // simulation.rs:95-98: prev = weather[0]; row for day 0 = census of the current state
        int rc = mvb_reset_rows(s);
        if (rc) return rc;
        s->row_days.assign(1, 0);
        s->prev_weather = weather[0];
        d0 = 1;
    }
    uint32_t n_week = 0;
    for (uint32_t d = d0; d < end_day; ++d) n_week += (d % 7 < 5);
    int rc = ensure_rows(s, n_week);
    if (rc) return rc;
    // Interventions that fall inside [d0, end_day): everything they need goes to the device BEFORE the day loop (one
    // copy out of a library-owned pinned buffer into stream-ordered scratch), so that on the day itself `intervene`
    // is one kernel launch for all replicas concerned: no allocation, no copy, no synchronisation inside the loop.
    struct Fire { uint32_t day, r; const mvb_intervention* iv; size_t car, bike, tf, tu; };
    std::vector<Fire> fires;
    size_t stage_bytes = 0;
    auto take = [&](size_t n) { const size_t at = stage_bytes; stage_bytes += (n + 255) & ~(size_t)255; return at; };
    if (ivs)
        for (uint32_t r = 0; r < s->reps.size(); ++r) {
            const mvb_intervention* iv = ivs[r];
            if (!iv || iv->day < d0 || iv->day >= end_day) continue;
            const Replica& rp = s->reps[r];
            if (iv->tables) {
                if ((rc = check_tables(iv->tables))) return rc;
                if (iv->tables->n_subcultures != rp.S || iv->tables->n_neighbourhoods != rp.H) { set_error("table shape differs from the replica's"); return MVB_ERR_ARG; }
            }
            Fire f{iv->day, r, iv, 0, 0, 0, 0};
            if (iv->owns_car) f.car = take(rp.N);
            if (iv->owns_bike) f.bike = take(rp.N);
            if (iv->tables) { f.tf = take(16 * (size_t)(rp.H + rp.S)); f.tu = take(16 * (size_t)rp.H); }
            fires.push_back(f);
        }
    std::stable_sort(fires.begin(), fires.end(), [](const Fire& a, const Fire& b) { return a.day < b.day; });
    TmpBuf iv_scratch;
    const IvDev* d_ivs = nullptr;
    if (!fires.empty()) {
        const size_t list_at = take(sizeof(IvDev) * fires.size());
        MVB_CUDA(cudaEventSynchronize(s->iv_ev));                  // the previous run is done with the pinned buffer
        if (stage_bytes > s->iv_pinned_cap) {
            pinned_cache().put(s->iv_pinned, s->iv_pinned_cap);
            s->iv_pinned = nullptr; s->iv_pinned_cap = 0;
            MVB_CUDA(pinned_cache().get(stage_bytes, &s->iv_pinned, &s->iv_pinned_cap));
        }
        MVB_CUDA(iv_scratch.alloc(stage_bytes, s->stream));
        char* h = s->iv_pinned;
        char* d = iv_scratch.as<char>();
        IvDev* list = reinterpret_cast<IvDev*>(h + list_at);
        for (size_t k = 0; k < fires.size(); ++k) {
            const Fire& f = fires[k];
            const Replica& rp = s->reps[f.r];
            IvDev e{f.r, f.iv->tables ? 1u : 0u, nullptr, nullptr, nullptr, nullptr};
            if (f.iv->owns_car)  { memcpy(h + f.car, f.iv->owns_car, rp.N);   e.car = reinterpret_cast<const uint8_t*>(d + f.car); }
            if (f.iv->owns_bike) { memcpy(h + f.bike, f.iv->owns_bike, rp.N); e.bike = reinterpret_cast<const uint8_t*>(d + f.bike); }
            if (f.iv->tables) {
                float* tf = reinterpret_cast<float*>(h + f.tf);
                memcpy(tf, f.iv->tables->supportiveness, 16 * (size_t)rp.H);
                memcpy(tf + 4 * rp.H, f.iv->tables->desirability, 16 * (size_t)rp.S);
                memcpy(h + f.tu, f.iv->tables->capacity, 16 * (size_t)rp.H);
                e.tf = reinterpret_cast<const float*>(d + f.tf); e.tu = reinterpret_cast<const uint32_t*>(d + f.tu);
            }
            list[k] = e;
        }
        MVB_CUDA(cudaMemcpyAsync(d, h, stage_bytes, cudaMemcpyHostToDevice, s->stream));
        MVB_CUDA(cudaEventRecord(s->iv_ev, s->stream));
        d_ivs = reinterpret_cast<const IvDev*>(d + list_at);
    }
    s->last_launches = 0;
    uint32_t last_weekday = 0;
    for (uint32_t d = d0; d < end_day; ++d) if (d % 7 < 5) last_weekday = d;
    MVB_CUDA(cudaEventRecord(s->ev0, s->stream));
    size_t next_fire = 0;
    for (uint32_t day = d0; day < end_day; ++day) {
        size_t last = next_fire;
        while (last < fires.size() && fires[last].day == day) ++last;                   // simulation.rs:103-105
        if (last > next_fire) {
            uint32_t np_max = 0;
            for (size_t k = next_fire; k < last; ++k) np_max = std::max(np_max, s->reps[fires[k].r].graph->L.N_pad);
            const dim3 grid(std::min<uint32_t>((np_max + 255) / 256, 4 * s->grid), (unsigned)(last - next_fire));
            apply_interventions_kernel<<<grid, 256, 0, s->stream>>>(d_ivs + next_fire, s->d_ivreps.as<IvRep>());
            MVB_CUDA(cudaGetLastError());
            next_fire = last;
        }
        if (day % 7 < 5) {                                                  // simulation.rs:108, 297-299
            const uint8_t w = weather[day];
            if (w > 1) { set_error("weather[%u] = %u is not a weather code", day, w); return MVB_ERR_ARG; }
            if ((rc = launch_day(s, w, (uint8_t)(s->prev_weather != w), day == last_weekday))) return rc;   // :113-128
            s->prev_weather = w;                                            // :131
            s->row_days.push_back(day);
        }
    }
    MVB_CUDA(cudaEventRecord(s->ev1, s->stream));
    s->timing_valid = true;
    return MVB_OK;
}

static void record_to_row(const Replica& rp, const uint32_t* rec, uint32_t* row) {
    // rec: hist[H][4] | joint[4] (index active*2+active_norm) | commute[3] | sub[S]
    const uint32_t H = rp.H, S = rp.S;
    const uint32_t* j = rec + 4 * H;
    row[0] = j[2] + j[3];                 // ActiveMode
    row[1] = j[1] + j[3];                 // ActiveNorm
    row[2] = j[2];                        // ActiveModeCounterToInactiveNorm
    row[3] = j[1];                        // InactiveModeCounterToActiveNorm
    for (uint32_t c = 0; c < 3; ++c) row[4 + c] = rec[4 * H + 4 + c];
    for (uint32_t k = 0; k < S; ++k) row[7 + k] = rec[4 * H + 7 + k];
    for (uint32_t h = 0; h < H; ++h) row[7 + S + h] = rec[4 * h + 2] + rec[4 * h + 3];
}

int mvb_fetch_rows(mvb_sim* s, uint32_t* n_rows, uint32_t* days_out, uint32_t* rows_out, size_t cap_u32) {
    if (!s) { set_error("null handle"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    if (int rc = check_peer_exchange(s)) return rc;
    const uint32_t rows = s->cursor + 1;
    if (n_rows) *n_rows = rows;
    if (days_out) memcpy(days_out, s->row_days.data(), sizeof(uint32_t) * rows);
    if (rows_out) {
        size_t need = 0;
        for (auto& rp : s->reps) need += (size_t)rows * (MVB_STATS_FIXED + rp.S + rp.H);
        if (cap_u32 < need) { set_error("rows buffer too small: need %zu u32, have %zu", need, cap_u32); return MVB_ERR_ARG; }
        std::vector<uint32_t> rec((size_t)rows * s->rec_stride);
        MVB_CUDA(cudaMemcpy(rec.data(), s->rec.p, sizeof(uint32_t) * rec.size(), cudaMemcpyDeviceToHost));
        uint32_t* o = rows_out;
        for (auto& rp : s->reps) {
            const uint32_t W = MVB_STATS_FIXED + rp.S + rp.H;
            for (uint32_t k = 0; k < rows; ++k, o += W)
                record_to_row(rp, rec.data() + (size_t)k * s->rec_stride + rp.rec_off, o);
        }
    }
    return MVB_OK;
}

int mvb_run(mvb_sim* s, const uint8_t* weather, uint32_t n_days, const mvb_intervention* const* ivs,
            uint32_t* days_out, uint32_t* rows_out) {
    if (!s || !weather || n_days == 0) { set_error("null argument"); return MVB_ERR_ARG; }
    int rc;
    if (n_days == 1) {
        if ((rc = mvb_reset_rows(s))) return rc;
        s->row_days.assign(1, 0);
        s->prev_weather = weather[0];
    } else if ((rc = mvb_run_async(s, weather, 0, n_days, ivs))) return rc;
    size_t cap = 0;
    for (auto& rp : s->reps) cap += (size_t)mvb_run_rows(n_days) * (MVB_STATS_FIXED + rp.S + rp.H);
    return mvb_fetch_rows(s, nullptr, days_out, rows_out, rows_out ? cap : 0);
}

int mvb_last_run_stats(mvb_sim* s, float* ms, uint64_t* launches) {
    if (!s) { set_error("null handle"); return MVB_ERR_ARG; }
    if (!s->timing_valid) { set_error("no mvb_run / mvb_run_async yet"); return MVB_ERR_STATE; }
    MVB_CUDA(cudaSetDevice(s->device));
    MVB_CUDA(cudaEventSynchronize(s->ev1));
    MVB_CUDA(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
    if (ms) *ms = s->last_ms;
    if (launches) *launches = s->last_launches;
    return MVB_OK;
}

int mvb_stats_row(mvb_sim* s, uint32_t r, uint32_t* row) {
    if (!s || r >= s->reps.size() || !row) { set_error("bad argument"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    std::vector<uint32_t> rec(s->reps[r].rec_width);
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    MVB_CUDA(cudaMemcpy(rec.data(), s->rec.as<uint32_t>() + (size_t)s->cursor * s->rec_stride + s->reps[r].rec_off,
                        sizeof(uint32_t) * rec.size(), cudaMemcpyDeviceToHost));
    record_to_row(s->reps[r], rec.data(), row);
    return MVB_OK;
}

int mvb_get_state(mvb_sim* s, uint32_t r, float* habit, uint8_t* cur, uint8_t* last, uint8_t* norm,
                  uint32_t* hist, float* cong) {
    if (!s || r >= s->reps.size()) { set_error("bad handle/replica"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    if (int rc = check_peer_exchange(s)) return rc;
    Replica& rp = s->reps[r];
    if (int rc = host_int2ext(s, *rp.graph)) return rc;
    const std::vector<uint32_t>& int2ext = rp.graph->L.int2ext;
    const size_t NP = int2ext.size();
    if (habit) {
        std::vector<uint32_t> recs(NP / 32 * kRecWords);
        MVB_CUDA(cudaMemcpy(recs.data(), rp.state.p, 4 * recs.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < NP; ++i)
            if (int2ext[i] != kInvalid) memcpy(habit + (size_t)4 * int2ext[i], &recs[rec_habit_word((uint32_t)i)], 16);
    }
    std::vector<uint2> v(NP / 32);
    if (cur) {
        MVB_CUDA(cudaMemcpy(v.data(), rp.vec[s->parity].p, 8 * v.size(), cudaMemcpyDeviceToHost));
        unpack_slices(int2ext, v, cur);
    }
    if (last) {
        MVB_CUDA(cudaMemcpy(v.data(), rp.vec[s->stepped ? (s->parity ^ 1u) : s->parity].p, 8 * v.size(), cudaMemcpyDeviceToHost));
        unpack_slices(int2ext, v, last);
    }
    if (norm) {
        MVB_CUDA(cudaMemcpy(v.data(), rp.normvec.p, 8 * v.size(), cudaMemcpyDeviceToHost));
        unpack_slices(int2ext, v, norm);
    }
    if (hist)  MVB_CUDA(cudaMemcpy(hist, s->rec.as<uint32_t>() + (size_t)s->cursor * s->rec_stride + rp.rec_off,
                                   sizeof(uint32_t) * 4 * rp.H, cudaMemcpyDeviceToHost));
    if (cong)  MVB_CUDA(cudaMemcpy(cong, rp.cong.p, sizeof(float) * 4 * rp.H, cudaMemcpyDeviceToHost));
    return MVB_OK;
}

#ifdef MVB_DEBUG_STAMPS
// Debug hook (not part of the ABI): record {start of the first block, end of the last block} (globaltimer, ns) of the next
// `n` day-kernel launches; a second call with out != nullptr copies what was recorded.  tools/launch_gaps.py.
int mvbx_launch_stamps(mvb_sim* s, uint32_t n, unsigned long long* out, uint32_t* n_out) {
    if (!s) return -1;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (out) {
        const uint32_t done = std::min<uint32_t>(s->stamp_cap, s->cursor > s->stamp_first ? s->cursor - s->stamp_first : 0);
        cudaMemcpy(out, s->stamps.p, 16 * (size_t)done, cudaMemcpyDeviceToHost);
        if (n_out) *n_out = done;
        StampCfg off{nullptr, nullptr, 0, 0, 0};
        cudaMemcpyToSymbol(g_stamps, &off, sizeof off);
        return 0;
    }
    if (ensure_rows(s, n + 1)) return -3;                          // the record buffer must not move while launches are stamped
    if (s->stamps.alloc(16 * (size_t)n) != cudaSuccess) return -2;
    std::vector<unsigned long long> init(2 * (size_t)n);
    for (uint32_t k = 0; k < n; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0; }
    cudaMemcpy(s->stamps.p, init.data(), 16 * (size_t)n, cudaMemcpyHostToDevice);
    s->stamp_cap = n; s->stamp_first = s->cursor;
    StampCfg cfg{s->stamps.as<unsigned long long>(), s->rec.as<uint32_t>(), s->rec_stride, s->cursor + 1, n};
    return cudaMemcpyToSymbol(g_stamps, &cfg, sizeof cfg) == cudaSuccess ? 0 : -4;
}
#endif

// Test hook (not part of the ABI in include/): the layout the GPU built for replica r against the host builder
// (layout.cpp, the specification) on the same population (host pointers).  0 = identical; otherwise msg says what differs.
int mvbx_compare_layout_with_host(mvb_sim* s, uint32_t r, const mvb_population* pop, char* msg, size_t cap) {
    if (!s || r >= s->reps.size() || !pop) return -1;
    auto say = [&](const char* what, size_t at) { if (msg && cap) snprintf(msg, cap, "%s differs at %zu", what, at); return 1; };
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    Replica& rp = s->reps[r];
    Graph& g = *rp.graph;
    GraphLayout H;
    uint32_t capacity = staged_capacity_agents(s->vec_cap_words);
    if (const char* e = getenv("MVB_STAGE_AGENTS")) capacity = std::min<uint32_t>(capacity, std::max(64, atoi(e)) / 64 * 64);
    if (build_graph_layout(*pop, rp.H, capacity, H, (uint32_t)s->reps.size())) return -2;
    const GraphLayout& L = g.L;
    if (H.N_pad != L.N_pad || H.n_hub_slices != L.n_hub_slices || H.n_sell_slices != L.n_sell_slices || H.hot_words != L.hot_words) return say("sizes", 0);
    if (H.ranges.size() != L.ranges.size() || H.n_common_ranges != L.n_common_ranges || H.mode != L.mode) return say("range count / layout mode", 0);
    for (size_t k = 0; k < H.ranges.size(); ++k)
        if (memcmp(&H.ranges[k], &L.ranges[k], 12)) return say("ranges", k);
    auto fetch = [&](const DevBuf& b, size_t bytes) { std::vector<char> v(bytes); cudaMemcpy(v.data(), b.p, bytes, cudaMemcpyDeviceToHost); return v; };
    {
        auto v = fetch(g.int2ext, 4 * (size_t)L.N_pad);
        const uint32_t* d = (const uint32_t*)v.data();
        for (size_t k = 0; k < L.N_pad; ++k) if (d[k] != H.int2ext[k]) return say("int2ext", k);
    }
    if (s->world == 1) {
        if (H.sell_codes_len != L.sell_codes_len || H.hub_codes_len != L.hub_codes_len) return say("code lengths", 0);
        auto v = fetch(g.sell_meta, 16 * (size_t)L.n_sell_slices);
        if (memcmp(v.data(), H.sell_meta.data(), v.size())) return say("sell_meta", 0);
        const size_t rows = (size_t)L.n_hub_slices * 32;
        auto o = fetch(g.hub_off, 8 * rows);
        if (memcmp(o.data(), H.hub_off.data(), o.size())) return say("hub_off", 0);
        auto dg = fetch(g.hub_deg, 16 * rows);
        if (memcmp(dg.data(), H.hub_deg.data(), dg.size())) return say("hub_deg", 0);
        auto ub = fetch(g.hub_unit_begin, 4 * ((size_t)L.n_hub_slices + 1));
        if (memcmp(ub.data(), H.hub_unit_begin.data(), ub.size())) return say("hub_unit_begin", 0);
        auto hu = fetch(g.hub_units, 16 * H.hub_units.size());
        if (memcmp(hu.data(), H.hub_units.data(), hu.size())) return say("hub_units", 0);
    }
    return 0;
}

uint64_t mvb_algorithmic_bytes_per_day(const mvb_sim* s, uint32_t r) {
    if (!s || r >= s->reps.size()) return 0;
    const Replica& rp = s->reps[r];
    return 4ull * (rp.graph->L.Es + rp.graph->L.En) + 51ull * rp.N;
}

}  // extern "C"

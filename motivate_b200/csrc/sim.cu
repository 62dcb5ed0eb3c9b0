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
    DevBuf deg, ranges, hub_off, hub_deg, hub_codes, hub_units, hub_unit_begin, sell_meta, sell_codes, int2ext;
    const void* key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // caller pointers it was built from
};

struct Replica {
    uint32_t N = 0, S = 0, H = 0;
    uint32_t rec_width = 0, rec_off = 0;
    std::shared_ptr<Graph> graph;
    DevBuf state, vec[2], normvec, pa, tables_f, tables_u, cong, hub_cnt;     // state = slice records (layout.h)
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
            MVB_SYM(Broadcast); MVB_SYM(AllReduce); MVB_SYM(GetErrorString);
#undef MVB_SYM
            if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.GroupStart || !n.GroupEnd || !n.Broadcast ||
                !n.AllReduce || !n.GetErrorString) n.lib = nullptr;
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

int launch_day(mvb_sim* s, uint8_t weather, uint8_t changed) {
    int rc = ensure_rows(s, 1);
    if (rc) return rc;
    const uint32_t* prev = s->rec.as<uint32_t>() + (size_t)s->cursor * s->rec_stride;
    uint32_t* next = s->rec.as<uint32_t>() + (size_t)(s->cursor + 1) * s->rec_stride;
    AgentWeights uw{s->prm.consistency, s->prm.social_connectivity, s->prm.subculture_connectivity,
                    s->prm.neighbourhood_connectivity, s->prm.average_weight};
    // replica descriptors travel as kernel parameters, kMaxRepsPerLaunch at a time
#define MVB_LAUNCH(PA, T, G) day_kernel<PA, T, G><<<s->grid, T, s->smem_bytes, s->stream>>>( \
        s->rep_args[c], R, P, s->parity, prev, next, weather, changed, uw, s->vec_cap_words)
    for (size_t c = 0; c < s->rep_args.size(); ++c) {
        const uint32_t R = std::min<uint32_t>(kMaxRepsPerLaunch, (uint32_t)s->reps.size() - (uint32_t)c * kMaxRepsPerLaunch);
        const uint32_t P = choose_parts(s, (uint32_t)c * kMaxRepsPerLaunch, R);
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
    if (s->world > 1) {
        // agent-partitioned: every rank now holds new modes / norms for its own slices and a partial record.
        // One NCCL group: each rank broadcasts the packed words of its slice range, the record is all-reduced.
        Replica& rp = s->reps[0];
        Nccl* n = nccl();
        uint2* vec_next = rp.vec[s->parity ^ 1u].as<uint2>();
        uint2* normvec = rp.normvec.as<uint2>();
        MVB_NCCL(n->GroupStart());
        for (uint32_t r = 0; r < s->world; ++r) {
            const size_t b = s->bounds[r], cnt = (size_t)s->bounds[r + 1] - b;
            if (!cnt) continue;
            MVB_NCCL(n->Broadcast(vec_next + b, vec_next + b, cnt * 8, ncclUint8, (int)r, s->comm, s->stream));
            MVB_NCCL(n->Broadcast(normvec + b, normvec + b, cnt * 8, ncclUint8, (int)r, s->comm, s->stream));
        }
        MVB_NCCL(n->AllReduce(next, next, s->rec_stride, ncclUint32, ncclSum, s->comm, s->stream));
        MVB_NCCL(n->GroupEnd());
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

// Two pinned staging slots for the arrays the layout builder leaves in pageable vectors: a copy out of pageable
// memory would make the host wait for everything queued on the stream; out of a slot it is asynchronous.  A slot is
// reused once the copies issued from it have finished (event).
struct PinnedRing {
    char* base[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0}, used = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int cur = 1;
    ~PinnedRing() {
        for (int k = 0; k < 2; ++k) { if (ev[k]) { cudaEventSynchronize(ev[k]); cudaEventDestroy(ev[k]); } pinned_cache().put(base[k], cap[k]); }
    }
    cudaError_t begin(size_t bytes) {                         // next slot, at least `bytes` large
        cur ^= 1; used = 0;
        cudaError_t e;
        if (!ev[cur]) { if ((e = cudaEventCreateWithFlags(&ev[cur], cudaEventDisableTiming)) != cudaSuccess) return e; }
        else if ((e = cudaEventSynchronize(ev[cur])) != cudaSuccess) return e;
        if (bytes > cap[cur]) {
            pinned_cache().put(base[cur], cap[cur]);
            base[cur] = nullptr; cap[cur] = 0;
            if ((e = pinned_cache().get(bytes, &base[cur], &cap[cur])) != cudaSuccess) { base[cur] = nullptr; cap[cur] = 0; return e; }
        }
        return cudaSuccess;
    }
    template <class T> const T* put(const std::vector<T>& v) {
        char* dst = base[cur] + used;
        memcpy(dst, v.data(), sizeof(T) * v.size());
        used += (sizeof(T) * v.size() + 255) & ~(size_t)255;
        return reinterpret_cast<const T*>(dst);
    }
    template <class T> static size_t need(const std::vector<T>& v) { return (sizeof(T) * v.size() + 255) & ~(size_t)255; }
    cudaError_t end(cudaStream_t st) { return cudaEventRecord(ev[cur], st); }
};

template <class T>
int upload_staged(Arena& mem, DevBuf& b, const std::vector<T>& v, PinnedRing& ring, cudaStream_t st) {
    MVB_CARVE(mem, b, sizeof(T) * v.size());
    if (!v.empty()) MVB_CUDA(cudaMemcpyAsync(b.p, ring.put(v), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    return MVB_OK;
}

// Host threads that run `job(i)` for i = 0..n-1 in index order while the caller consumes the results as they come:
// wait(i) blocks until job i is done.  The destructor stops handing out jobs and joins.
class Crew {
public:
    template <class F>
    Crew(uint32_t n, F job) : n_(n), done_(n, 0) {
        const uint32_t nt = std::max(1u, std::min(n, std::thread::hardware_concurrency()));
        for (uint32_t t = 0; t < nt; ++t)
            th_.emplace_back([this, job] {
                for (uint32_t i; !stop_.load() && (i = next_.fetch_add(1)) < n_;) {
                    job(i);
                    { std::lock_guard<std::mutex> lk(m_); done_[i] = 1; }
                    cv_.notify_all();
                }
            });
    }
    void wait(uint32_t i) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return done_[i] != 0; });
    }
    ~Crew() {
        stop_.store(true);
        for (auto& t : th_) t.join();
    }
private:
    uint32_t n_;
    std::vector<char> done_;
    std::atomic<uint32_t> next_{0};
    std::atomic<bool> stop_{false};
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::thread> th_;
};

// the u64 of one slice (lanes 0-15 in .x, 16-31 in .y, as finish_slice packs them) -> 32 two-bit values
void unpack_slices(const std::vector<uint32_t>& int2ext, const std::vector<uint2>& in, uint8_t* values) {
    for (size_t i = 0; i < int2ext.size(); ++i) {
        if (int2ext[i] == kInvalid) continue;
        const uint32_t l = (uint32_t)(i & 31);
        const uint2 w = in[i >> 5];
        values[int2ext[i]] = (uint8_t)(((l < 16 ? w.x : w.y) >> ((l & 15u) * 2u)) & 3u);
    }
}

// Run f(0..n-1) on up to hardware_concurrency host threads (setup only: validation, layout builds).
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

int check_population_arrays(const mvb_population* p) {          // O(1): what must hold before anything is read
    if (!p || p->n_agents == 0 || !p->subculture || !p->neighbourhood || !p->commute || !p->owns_car ||
        !p->owns_bike || !p->weather_sensitivity || !p->habit || !p->current_mode || !p->norm ||
        !p->social_rowptr || !p->neighbour_rowptr) {
        set_error("population has null arrays or zero agents");
        return MVB_ERR_ARG;
    }
    const uint64_t es = p->social_rowptr[p->n_agents], en = p->neighbour_rowptr[p->n_agents];
    if ((es && !p->social_col) || (en && !p->neighbour_col)) { set_error("null column array"); return MVB_ERR_ARG; }
    if (es >= (1ull << 40) || en >= (1ull << 40)) { set_error("rowptr[n_agents] is not a plausible link count"); return MVB_ERR_ARG; }
    return MVB_OK;
}

// O(N), or O(N + E) with check_cols: every index in range.  mvb_create* leaves the link targets to count_hot_kernel.
int validate_population(const mvb_population* p, const mvb_tables* t, bool check_cols) {
    if (int rc = check_population_arrays(p)) return rc;
    const uint32_t N = p->n_agents;
    for (uint32_t i = 0; i < N; ++i) {
        if (p->subculture[i] >= t->n_subcultures) { set_error("agent %u: subculture index %u out of range (\"Subculture not found\")", i, p->subculture[i]); return MVB_ERR_ARG; }
        if (p->neighbourhood[i] >= t->n_neighbourhoods) { set_error("agent %u: neighbourhood index %u out of range", i, p->neighbourhood[i]); return MVB_ERR_ARG; }
        if (p->commute[i] > 2 || p->current_mode[i] > 3 || p->norm[i] > 3) { set_error("agent %u: bad commute/mode/norm code", i); return MVB_ERR_ARG; }
    }
    for (int g = 0; g < 2; ++g) {
        const uint64_t* rp = g ? p->neighbour_rowptr : p->social_rowptr;
        const uint32_t* cl = g ? p->neighbour_col : p->social_col;
        if (rp[0] != 0) { set_error("rowptr[0] must be 0"); return MVB_ERR_ARG; }
        for (uint32_t i = 0; i < N; ++i)
            if (rp[i + 1] < rp[i]) { set_error("rowptr not monotone at %u", i); return MVB_ERR_ARG; }
        if (rp[N] && !cl) { set_error("null column array"); return MVB_ERR_ARG; }
        for (uint64_t e = 0; check_cols && e < rp[N]; ++e)
            if (cl[e] >= N) { set_error("link target %u out of range", cl[e]); return MVB_ERR_ARG; }
    }
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
    if (s->comm && nccl()) nccl()->CommDestroy(s->comm);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    const double t2 = now();
    s->mem.chunks.clear();
    const double t3 = now();
    delete s;
    if (getenv("MVB_TIMING"))
        fprintf(stderr, "mvb_destroy: stream sync %.3f s, events/stream %.3f s, device memory %.3f s, host memory %.3f s\n",
                t1 - t0, t2 - t1, t3 - t2, now() - t3);
}

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
    // MVB_TIMING=1 prints where the set-up time went (stderr)
    const bool timing = getenv("MVB_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tm0 = now();
    double tm_wait = 0, tm_graph = 0, tm_state = 0;
    for (uint32_t r = 0; r < R; ++r)
        if (int rc = check_population_arrays(pops[r])) return rc;
    const double tm1 = now();
    MVB_CUDA(cudaSetDevice(device));
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

    // shared-memory budget: everything the device lets one block opt in to (227 KB on B200); what the
    // tables do not need holds the staged mode vector
    int sms = 0, smem_optin = 0;
    MVB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    MVB_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    uint32_t tail_max = 0, census_words = 0;
    for (uint32_t r = 0; r < R; ++r) {
        tail_max = std::max(tail_max, tail_words(tabs[r]->n_subcultures, tabs[r]->n_neighbourhoods));
        census_words = std::max(census_words, rec_width(tabs[r]->n_subcultures, tabs[r]->n_neighbourhoods));
    }
    tail_max = (tail_max + 3u) & ~3u;
    if ((uint64_t)tail_max * 4 + 4096 > (uint64_t)smem_optin) {
        set_error("too many neighbourhoods/subcultures: the per-block tables need %u bytes of shared memory", tail_max * 4);
        return MVB_ERR_ARG;
    }
    s->smem_bytes = (size_t)smem_optin;
    s->vec_cap_words = ((uint32_t)smem_optin / 4 - tail_max) & ~3u;
    s->census_smem = sizeof(uint32_t) * census_words;
    const uint32_t capacity_agents = staged_capacity_agents(s->vec_cap_words);
    s->grid = (uint32_t)sms;

    s->reps.resize(R);
    s->h_reps.resize(R);
    // distinct link graphs (replicas that pass the same graph + neighbourhood arrays share one), laid out in parallel
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
    {   // first chunk of device memory: an estimate of everything the graphs and replicas will need (link codes ~ the
        // caller's CSR columns + slice padding, 12 B per agent of graph arrays, 27-47 B per agent of state)
        size_t est = 64u << 20;
        for (uint32_t g = 0; g < graphs.size(); ++g) {
            const mvb_population* p = pops[first_rep[g]];
            const size_t e = p->social_rowptr[p->n_agents] + p->neighbour_rowptr[p->n_agents];
            est += 4 * e + e / 4 + 16 * (size_t)p->n_agents + (1u << 20);
        }
        for (uint32_t r = 0; r < R; ++r) est += 48 * (size_t)pops[r]->n_agents + (1u << 16);
        size_t free_b = 0, total_b = 0;
        MVB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        s->mem.chunk_bytes = std::max<size_t>(256u << 20, std::min(est, free_b / 2));
    }
    // Set-up is a pipeline between host threads (the crew) and the device, driven by this thread:
    //   A(r)  crew: range checks of replica r's arrays, residents per neighbourhood, and, if r is the first replica
    //         of its graph g, layout_begin (degree order, hot set)
    //   S1(g) this thread: the caller's CSR goes up as it is; once A is done, is_hot goes up, count_hot_kernel counts
    //         the hot links of every agent (and range-checks the targets), the counts come back to pinned memory
    //   F(g)  crew: waits for the counts, layout_finish (final order, slices, offsets)
    //   S2(g) this thread: layout arrays up, tcode + fill kernels, then the state of the replicas that use g
    // S1 runs kLook graphs ahead of S2 so that the copies of later graphs hide the host work of earlier ones and
    // several F jobs run side by side.
    constexpr uint32_t kLook = 6, kSlots = kLook + 2;          // F jobs of up to kLook + 1 graphs run side by side
    const uint32_t G = (uint32_t)graphs.size();
    struct GraphWork {
        LayoutWork W;
        TmpBuf srp, scol, nrp, ncol, is_hot, counts, flag;
        uint32_t* h_counts = nullptr;                            // pinned: nsh[N] | nnh[N] | bad flag
    };
    std::vector<GraphWork> work(G);
    struct Slots {                                               // pinned buffers the counts come back into
        char* base[kSlots] = {};
        size_t cap[kSlots] = {};
        cudaEvent_t ev[kSlots] = {};
        ~Slots() { for (uint32_t k = 0; k < kSlots; ++k) { if (ev[k]) { cudaEventSynchronize(ev[k]); cudaEventDestroy(ev[k]); } pinned_cache().put(base[k], cap[k]); } }
    } slots;
    {
        size_t slot_bytes = 0;
        for (uint32_t g = 0; g < G; ++g) slot_bytes = std::max(slot_bytes, 9 * (size_t)pops[first_rep[g]]->n_agents + 64);
        for (uint32_t k = 0; k < std::min(kSlots, G); ++k) {
            MVB_CUDA(pinned_cache().get(slot_bytes, &slots.base[k], &slots.cap[k]));
            MVB_CUDA(cudaEventCreateWithFlags(&slots.ev[k], cudaEventDisableTiming));
        }
    }
    struct Signals {                                             // S1(g) queued -> F(g) may wait for its event
        std::mutex m;
        std::condition_variable cv;
        std::vector<char> queued;
        bool abort = false;
    } sig;
    sig.queued.assign(G, 0);
    // crew job list: A(r) in replica order; F(g) follows the A of graph g + kLook
    std::vector<std::pair<char, uint32_t>> jobs;
    std::vector<uint32_t> job_a(R), job_f(G);
    for (uint32_t r = 0; r < R; ++r) {
        job_a[r] = (uint32_t)jobs.size();
        jobs.emplace_back('A', r);
        const uint32_t g = graph_of[r];
        if (first_rep[g] == r && g >= kLook) { job_f[g - kLook] = (uint32_t)jobs.size(); jobs.emplace_back('F', g - kLook); }
    }
    for (uint32_t g = G > kLook ? G - kLook : 0; g < G; ++g) { job_f[g] = (uint32_t)jobs.size(); jobs.emplace_back('F', g); }
    std::vector<int> job_rc(jobs.size(), MVB_OK);
    std::vector<std::string> job_msg(jobs.size());
    std::vector<std::vector<uint32_t>> nres_all(R);
    Crew crew((uint32_t)jobs.size(), [&](uint32_t j) {
        auto fail = [&](int rc) { job_rc[j] = rc; job_msg[j] = mvb_last_error(); };
        if (jobs[j].first == 'A') {
            const uint32_t r = jobs[j].second, g = graph_of[r];
            if (int rc = validate_population(pops[r], tabs[r], false)) return fail(rc);
            nres_all[r].assign(tabs[r]->n_neighbourhoods, 0u);
            for (uint32_t i = 0; i < pops[r]->n_agents; ++i) nres_all[r][pops[r]->neighbourhood[i]]++;
            if (first_rep[g] != r) return;
            if (int rc = layout_begin(*pops[r], tabs[r]->n_neighbourhoods, capacity_agents, graphs[g]->L, work[g].W)) return fail(rc);
        } else {
            const uint32_t g = jobs[j].second;
            {
                std::unique_lock<std::mutex> lk(sig.m);
                sig.cv.wait(lk, [&] { return sig.queued[g] || sig.abort; });
                if (sig.abort) return;
            }
            cudaSetDevice(device);
            if (cudaEventSynchronize(slots.ev[g % kSlots]) != cudaSuccess) { set_error("waiting for the device failed"); return fail(MVB_ERR_CUDA); }
            GraphWork& w = work[g];
            const uint32_t N = pops[first_rep[g]]->n_agents;
            if (w.h_counts[2 * (size_t)N]) { set_error("a link target is out of range"); return fail(MVB_ERR_ARG); }
            w.W.nsh = w.h_counts;
            w.W.nnh = w.h_counts + N;
            if (int rc = layout_finish(*pops[first_rep[g]], w.W, graphs[g]->L, false)) return fail(rc);
            LayoutWork done;                                     // only is_hot's device copy is still needed
            std::swap(done, w.W);
        }
    });
    struct Abort {                                               // declared after the crew: runs before it joins
        Signals& sig;
        ~Abort() { { std::lock_guard<std::mutex> lk(sig.m); sig.abort = true; } sig.cv.notify_all(); }
    } abort_guard{sig};
    auto wait_job = [&](uint32_t j, const char* what, uint32_t idx) -> int {
        crew.wait(j);
        if (job_rc[j]) { set_error("%s %u: %s", what, idx, job_msg[j].c_str()); return job_rc[j]; }
        return MVB_OK;
    };
    uint32_t csr_issued = 0, s1_done = 0;                        // graphs [0, x) have their CSR copies / their S1 queued
    auto issue_csr = [&](uint32_t upto) -> int {
        for (; csr_issued < std::min(upto, G); ++csr_issued) {
            const mvb_population* p = pops[first_rep[csr_issued]];
            const size_t nb_rp = 8 * ((size_t)p->n_agents + 1), es = p->social_rowptr[p->n_agents], en = p->neighbour_rowptr[p->n_agents];
            GraphWork& c = work[csr_issued];
            MVB_CUDA(c.srp.alloc(nb_rp, s->stream)); MVB_CUDA(c.nrp.alloc(nb_rp, s->stream));
            MVB_CUDA(c.scol.alloc(4 * es, s->stream)); MVB_CUDA(c.ncol.alloc(4 * en, s->stream));
            MVB_CUDA(cudaMemcpyAsync(c.srp.p, p->social_rowptr, nb_rp, cudaMemcpyHostToDevice, s->stream));
            MVB_CUDA(cudaMemcpyAsync(c.nrp.p, p->neighbour_rowptr, nb_rp, cudaMemcpyHostToDevice, s->stream));
            if (es) MVB_CUDA(cudaMemcpyAsync(c.scol.p, p->social_col, 4 * es, cudaMemcpyHostToDevice, s->stream));
            if (en) MVB_CUDA(cudaMemcpyAsync(c.ncol.p, p->neighbour_col, 4 * en, cudaMemcpyHostToDevice, s->stream));
        }
        return MVB_OK;
    };
    auto stage1 = [&](uint32_t upto) -> int {
        for (; s1_done < std::min(upto, G); ++s1_done) {
            const uint32_t g = s1_done, r0 = first_rep[g], N = pops[r0]->n_agents;
            if (int rc = issue_csr(g + 1)) return rc;
            const double tw = now();
            if (int rc = wait_job(job_a[r0], "replica", r0)) return rc;
            tm_wait += now() - tw;
            GraphWork& w = work[g];
            char* slot = slots.base[g % kSlots];
            w.h_counts = reinterpret_cast<uint32_t*>(slot);
            uint8_t* h_hot = reinterpret_cast<uint8_t*>(slot) + 8 * (size_t)N + 64;
            memcpy(h_hot, w.W.is_hot.data(), N);
            MVB_CUDA(w.is_hot.alloc(N, s->stream));
            MVB_CUDA(w.counts.alloc(8 * (size_t)N + 4, s->stream));
            MVB_CUDA(cudaMemcpyAsync(w.is_hot.p, h_hot, N, cudaMemcpyHostToDevice, s->stream));
            uint32_t* d_bad = w.counts.as<uint32_t>() + 2 * (size_t)N;
            MVB_CUDA(cudaMemsetAsync(d_bad, 0, 4, s->stream));
            count_hot_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)N * 32 + 255) / 256, (uint64_t)sms * 32), 256, 0, s->stream>>>(
                N, w.srp.as<uint64_t>(), w.scol.as<uint32_t>(), w.nrp.as<uint64_t>(), w.ncol.as<uint32_t>(), w.is_hot.as<uint8_t>(),
                w.counts.as<uint32_t>(), w.counts.as<uint32_t>() + N, d_bad);
            MVB_CUDA(cudaGetLastError());
            MVB_CUDA(cudaMemcpyAsync(w.h_counts, w.counts.p, 8 * (size_t)N + 4, cudaMemcpyDeviceToHost, s->stream));
            MVB_CUDA(cudaEventRecord(slots.ev[g % kSlots], s->stream));
            { std::lock_guard<std::mutex> lk(sig.m); sig.queued[g] = 1; }
            sig.cv.notify_all();
        }
        return MVB_OK;
    };
    PinnedRing ring;
    {   // scratch buffers come from the device's stream-ordered pool; let it keep what it has between replicas
        cudaMemPool_t pool;
        MVB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        MVB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    std::vector<char> uploaded(graphs.size(), 0);
    std::list<std::vector<float>> host_f;
    const double tm2 = now();
    for (uint32_t r = 0; r < R; ++r) {
        const double tr = now();
        const mvb_population* p = pops[r];
        const mvb_tables* t = tabs[r];
        Replica& rp = s->reps[r];
        const uint32_t N = p->n_agents;
        rp.N = N; rp.S = t->n_subcultures; rp.H = t->n_neighbourhoods;
        rp.rec_width = rec_width(rp.S, rp.H);
        rp.rec_off = s->rec_stride;
        s->rec_stride += rp.rec_width;
        rp.graph = graphs[graph_of[r]];
        if (int rc = stage1(graph_of[r] + 1 + kLook)) return rc;
        if (int rc = issue_csr(graph_of[r] + 2 + kLook)) return rc;
        {
            const double tw = now();
            if (int rc = wait_job(job_a[r], "replica", r)) return rc;
            if (!uploaded[graph_of[r]])
                if (int rc = wait_job(job_f[graph_of[r]], "replica", r)) return rc;
            tm_wait += now() - tw;
        }
        if (!uploaded[graph_of[r]]) {
            uploaded[graph_of[r]] = 1;
            Graph* g = rp.graph.get();
            GraphLayout& L = g->L;
            int rc;
            MVB_CUDA(ring.begin(PinnedRing::need(L.ranges) + PinnedRing::need(L.hub_off) + PinnedRing::need(L.hub_deg) +
                                PinnedRing::need(L.sell_meta) + PinnedRing::need(L.hub_units) + PinnedRing::need(L.hub_unit_begin) +
                                PinnedRing::need(L.int2ext)));
            if ((rc = upload_staged(s->mem, g->ranges, L.ranges, ring, s->stream)) || (rc = upload_staged(s->mem, g->hub_off, L.hub_off, ring, s->stream)) ||
                (rc = upload_staged(s->mem, g->hub_deg, L.hub_deg, ring, s->stream)) || (rc = upload_staged(s->mem, g->sell_meta, L.sell_meta, ring, s->stream)) ||
                (rc = upload_staged(s->mem, g->hub_units, L.hub_units, ring, s->stream)) ||
                (rc = upload_staged(s->mem, g->hub_unit_begin, L.hub_unit_begin, ring, s->stream)) ||
                (rc = upload_staged(s->mem, g->int2ext, L.int2ext, ring, s->stream)))
                return rc;
            MVB_CARVE(s->mem, g->deg, sizeof(uint32_t) * (size_t)L.N_pad);
            MVB_CARVE(s->mem, g->hub_codes, sizeof(uint32_t) * L.hub_codes_len);
            MVB_CARVE(s->mem, g->sell_codes, sizeof(uint32_t) * L.sell_codes_len);
            // unused entries = the code of the always-zero shared-memory word (layout.h)
            fill_u32_kernel<<<sms * 8, 256, 0, s->stream>>>(g->hub_codes.as<uint32_t>(), L.hub_codes_len, pad_code(s->vec_cap_words));
            fill_u32_kernel<<<sms * 8, 256, 0, s->stream>>>(g->sell_codes.as<uint32_t>(), L.sell_codes_len, pad_code(s->vec_cap_words));
            MVB_CUDA(cudaGetLastError());
            // the caller's CSR went to the device as it is (issue_csr); the fill kernels re-encode it there, then it is dropped
            GraphWork& c = work[graph_of[r]];
            TmpBuf t_tcode;
            MVB_CUDA(t_tcode.alloc(4 * (size_t)N, s->stream));
            MVB_CUDA(ring.end(s->stream));
            tcode_kernel<<<(L.N_pad + 255) / 256, 256, 0, s->stream>>>(L.N_pad, g->int2ext.as<uint32_t>(), g->ranges.as<StageRange>(),
                                                                     (uint32_t)L.ranges.size(), c.is_hot.as<uint8_t>(), t_tcode.as<uint32_t>());
            MVB_CUDA(cudaGetLastError());
            FillArgs fa{c.srp.as<uint64_t>(), c.scol.as<uint32_t>(), c.nrp.as<uint64_t>(), c.ncol.as<uint32_t>(),
                        t_tcode.as<uint32_t>(), g->int2ext.as<uint32_t>(), L.n_hub_slices, L.n_sell_slices,
                        g->sell_meta.as<uint4>(), g->sell_codes.as<uint32_t>(), g->deg.as<uint32_t>(),
                        g->hub_off.as<uint64_t>(), g->hub_deg.as<uint4>(), g->hub_codes.as<uint32_t>()};
            MVB_CUDA(cudaMemsetAsync(g->deg.p, 0, g->deg.bytes, s->stream));
            if (L.n_sell_slices) {
                fill_sell_kernel<<<(L.n_sell_slices * 32 + 255) / 256, 256, 0, s->stream>>>(fa);
                MVB_CUDA(cudaGetLastError());
            }
            if (L.n_hub_slices) {
                fill_hub_kernel<<<(L.n_hub_slices * 32 * 32 + 255) / 256, 256, 0, s->stream>>>(fa, L.n_hub_slices * 32);
                MVB_CUDA(cudaGetLastError());
            }
            c.srp.release(); c.scol.release(); c.nrp.release(); c.ncol.release(); c.is_hot.release(); c.counts.release();   // back to the pool, in stream order
            std::vector<uint4x>().swap(L.sell_meta);
        }
        const GraphLayout& L = rp.graph->L;
        const uint32_t NP = L.N_pad;
        s->max_npad = std::max(s->max_npad, NP);
        const double ts = now();
        tm_graph += ts - tr;
        // per-agent state: the caller's arrays go up as they are, a kernel permutes them into internal slot order
        const float* pa_src[5] = {p->consistency, p->social_connectivity, p->subculture_connectivity,
                                  p->neighbourhood_connectivity, p->average_weight};
        const float pa_def[5] = {prm->consistency, prm->social_connectivity, prm->subculture_connectivity,
                                 prm->neighbourhood_connectivity, prm->average_weight};
        rp.per_agent = false;
        for (int k = 0; k < 5; ++k) rp.per_agent |= (pa_src[k] != nullptr);
        if (rp.per_agent) s->any_per_agent = true;
        {
            const size_t hub_cnt_bytes = sizeof(uint32_t) * 8 * 32 * (size_t)L.n_hub_slices;
            MVB_CARVE(s->mem, rp.state, (size_t)(NP / 32) * kRecWords * 4);
            // + 64 bytes: the early gather of a padded cold-list entry reads the word a padding code names (at most
            // 16 bytes past a vector that only just exceeds the shared-memory capacity); its value is never used
            MVB_CARVE(s->mem, rp.vec[0], NP / 4 + 64); MVB_CARVE(s->mem, rp.vec[1], NP / 4 + 64); MVB_CARVE(s->mem, rp.normvec, NP / 4);
            if (rp.per_agent) MVB_CARVE(s->mem, rp.pa, 20 * (size_t)NP);
            MVB_CARVE(s->mem, rp.tables_f, sizeof(float) * 4 * (rp.H + rp.S));
            MVB_CARVE(s->mem, rp.tables_u, sizeof(uint32_t) * 5 * rp.H);
            MVB_CARVE(s->mem, rp.cong, sizeof(float) * 4 * rp.H);
            MVB_CARVE(s->mem, rp.hub_cnt, hub_cnt_bytes);
            TmpBuf t[14];
            auto up = [&](TmpBuf& b, const void* src, size_t bytes) -> int {
                MVB_CUDA(b.alloc(bytes, s->stream));
                MVB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s->stream));
                return MVB_OK;
            };
            int rc;
            if ((rc = up(t[0], p->subculture, N)) || (rc = up(t[1], p->neighbourhood, 2 * (size_t)N)) ||
                (rc = up(t[2], p->commute, N)) || (rc = up(t[3], p->owns_car, N)) || (rc = up(t[4], p->owns_bike, N)) ||
                (rc = up(t[5], p->weather_sensitivity, 4 * (size_t)N)) || (rc = up(t[6], p->habit, 16 * (size_t)N)) ||
                (rc = up(t[7], p->current_mode, N)) || (rc = up(t[8], p->norm, N)))
                return rc;
            InitArgs ia{};
            ia.int2ext = rp.graph->int2ext.as<uint32_t>(); ia.N = N; ia.N_pad = NP;
            ia.sub = t[0].as<uint8_t>(); ia.nbhd = t[1].as<uint16_t>(); ia.commute = t[2].as<uint8_t>();
            ia.car = t[3].as<uint8_t>(); ia.bike = t[4].as<uint8_t>(); ia.ws = t[5].as<float>();
            ia.habit = t[6].as<float4>(); ia.mode = t[7].as<uint8_t>(); ia.norm = t[8].as<uint8_t>();
            for (int k = 0; k < 5; ++k) {
                ia.pa_src[k] = nullptr; ia.pa_def[k] = pa_def[k];
                if (rp.per_agent && pa_src[k]) { if ((rc = up(t[9 + k], pa_src[k], 4 * (size_t)N))) return rc; ia.pa_src[k] = t[9 + k].as<float>(); }
            }
            ia.lens = rp.graph->deg.as<uint32_t>(); ia.state = rp.state.as<uint32_t>();
            ia.pa_out = rp.per_agent ? rp.pa.as<float>() : nullptr;
            ia.vec0 = rp.vec[0].as<uint2>(); ia.vec1 = rp.vec[1].as<uint2>(); ia.normvec = rp.normvec.as<uint2>();
            init_state_kernel<<<(NP + 255) / 256, 256, 0, s->stream>>>(ia);
            MVB_CUDA(cudaGetLastError());                        // the scratch goes back to the pool in stream order
        }
        // tables + residents
        if (int rc = upload_tables(s, rp, t)) return rc;
        MVB_CUDA(cudaMemcpyAsync(rp.tables_u.as<uint32_t>() + 4 * rp.H, nres_all[r].data(), 4 * rp.H, cudaMemcpyHostToDevice, s->stream));
        host_f.emplace_back(4 * rp.H, 1.0f);          // default_congestion_modifier, neighbourhood.rs:34-41; lives until the final synchronize
        MVB_CUDA(cudaMemcpyAsync(rp.cong.p, host_f.back().data(), 16 * rp.H, cudaMemcpyHostToDevice, s->stream));

        Graph& g = *rp.graph;
        RepDev& d = s->h_reps[r];
        d.N = N; d.N_pad = NP; d.S = rp.S; d.H = rp.H; d.rec_off = rp.rec_off;
        d.n_hub_slices = L.n_hub_slices; d.n_sell_slices = L.n_sell_slices;
        d.n_ranges = (uint32_t)L.ranges.size(); d.hot_words = L.hot_words;
        d.hub_begin = 0; d.hub_end = L.n_hub_slices; d.sell_begin = 0; d.sell_end = L.n_sell_slices;
        if (world > 1) {                               // this rank's slice range, split into its hub and SELL parts
            partition_slices(L, world, s->bounds);
            const uint32_t b = s->bounds[rank], e = s->bounds[rank + 1], nh = L.n_hub_slices;
            d.hub_begin = std::min(b, nh); d.hub_end = std::min(e, nh);
            d.sell_begin = std::max(b, nh) - nh; d.sell_end = std::max(e, nh) - nh;
        }
        d.state = rp.state.as<uint32_t>();
        d.pa = rp.per_agent ? rp.pa.as<float>() : nullptr;
        d.vec[0] = rp.vec[0].as<uint2>(); d.vec[1] = rp.vec[1].as<uint2>(); d.normvec = rp.normvec.as<uint2>();
        d.ranges = g.ranges.as<StageRange>();
        if (rp.hub_cnt.bytes) MVB_CUDA(cudaMemsetAsync(rp.hub_cnt.p, 0, rp.hub_cnt.bytes, s->stream));
        d.hub_units = g.hub_units.as<uint4>(); d.hub_unit_begin = g.hub_unit_begin.as<uint32_t>();
        d.hub_off = g.hub_off.as<uint64_t>(); d.hub_deg = g.hub_deg.as<uint4>(); d.hub_cnt = rp.hub_cnt.as<uint32_t>();
        d.hub_codes = g.hub_codes.as<uint32_t>();
        d.sell_meta = g.sell_meta.as<uint4>(); d.sell_codes = g.sell_codes.as<uint32_t>();
        d.supp = rp.tables_f.as<float>(); d.desir = rp.tables_f.as<float>() + 4 * rp.H;
        d.cap = rp.tables_u.as<uint32_t>(); d.nres = rp.tables_u.as<uint32_t>() + 4 * rp.H;
        d.cong = rp.cong.as<float>();
        if (L.hot_words + 4 > s->vec_cap_words) { set_error("internal: staged vector larger than its shared-memory slot"); return MVB_ERR_STATE; }
        tm_state += now() - ts;
    }
    const double tm3 = now();
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<true, 1024, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    MVB_CUDA(cudaFuncSetAttribute(day_kernel<false, 1024, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    if (s->census_smem > 48 * 1024)
        MVB_CUDA(cudaFuncSetAttribute(census_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->census_smem));
    s->rank = rank; s->world = world;
    if (world > 1) {
        Nccl* n = nccl();
        if (!n) { set_error("agent-partitioned mode needs NCCL (libnccl.so.2 could not be loaded: %s)", dlerror()); return MVB_ERR_COMM; }
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof id);
        MVB_NCCL(n->CommInitRank(&s->comm, (int)world, id, (int)rank));
    }
    s->rep_args.assign((R + kMaxRepsPerLaunch - 1) / kMaxRepsPerLaunch, RepArray{});
    for (uint32_t r = 0; r < R; ++r) s->rep_args[r / kMaxRepsPerLaunch].r[r % kMaxRepsPerLaunch] = s->h_reps[r];
    MVB_CUDA(s->d_reps.alloc(sizeof(RepDev) * R));
    int rc = upload_reps(s);
    if (rc) return rc;
    rc = ensure_rows(s, 8);
    if (rc) return rc;
    s->cursor = 0;
    s->row_days.assign(1, 0);
    rc = launch_census(s, 0);
    if (rc) return rc;
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    if (timing)
        fprintf(stderr, "mvb_create: checks %.3f s, set-up %.3f s, replicas %.3f s (waiting for host threads %.3f, graph enqueue %.3f, "
                "state enqueue %.3f), drain %.3f s\n", tm1 - tm0, tm2 - tm1, tm3 - tm2, tm_wait, tm_graph - tm_wait, tm_state, now() - tm3);
    {   // hand the scratch memory back
        cudaMemPool_t pool;
        MVB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = 0;
        MVB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        MVB_CUDA(cudaMemPoolTrimTo(pool, 0));
    }
    *out = sp.release();
    return MVB_OK;
}

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

uint32_t mvb_n_replicas(const mvb_sim* s) { return s ? (uint32_t)s->reps.size() : 0; }
uint32_t mvb_stats_width(const mvb_sim* s, uint32_t r) {
    return (s && r < s->reps.size()) ? MVB_STATS_FIXED + s->reps[r].S + s->reps[r].H : 0;
}
void* mvb_stream(mvb_sim* s) { return s ? (void*)s->stream : nullptr; }

int mvb_sync(mvb_sim* s) {
    if (!s) { set_error("null handle"); return MVB_ERR_ARG; }
    MVB_CUDA(cudaSetDevice(s->device));
    MVB_CUDA(cudaStreamSynchronize(s->stream));
    return MVB_OK;
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
    DevBuf tc, tb;
    if (car)  { MVB_CUDA(tc.alloc(rp.N)); MVB_CUDA(cudaMemcpyAsync(tc.p, car, rp.N, cudaMemcpyHostToDevice, s->stream)); }
    if (bike) { MVB_CUDA(tb.alloc(rp.N)); MVB_CUDA(cudaMemcpyAsync(tb.p, bike, rp.N, cudaMemcpyHostToDevice, s->stream)); }
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
    s->last_launches = 0;
    MVB_CUDA(cudaEventRecord(s->ev0, s->stream));
    for (uint32_t day = d0; day < end_day; ++day) {
        if (ivs)
            for (uint32_t r = 0; r < s->reps.size(); ++r) {
                const mvb_intervention* iv = ivs[r];
                if (!iv || iv->day != day) continue;                        // simulation.rs:103-105
                if (iv->tables && (rc = mvb_set_tables(s, r, iv->tables))) return rc;
                if ((rc = mvb_set_ownership(s, r, iv->owns_car, iv->owns_bike))) return rc;
            }
        if (day % 7 < 5) {                                                  // simulation.rs:108, 297-299
            const uint8_t w = weather[day];
            if (w > 1) { set_error("weather[%u] = %u is not a weather code", day, w); return MVB_ERR_ARG; }
            if ((rc = launch_day(s, w, (uint8_t)(s->prev_weather != w)))) return rc;   // :113-128
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
    Replica& rp = s->reps[r];
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

uint64_t mvb_algorithmic_bytes_per_day(const mvb_sim* s, uint32_t r) {
    if (!s || r >= s->reps.size()) return 0;
    const Replica& rp = s->reps[r];
    return 4ull * (rp.graph->L.Es + rp.graph->L.En) + 51ull * rp.N;
}

}  // extern "C"

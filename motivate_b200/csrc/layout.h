// layout.h — device data layout of one population's link graph (host-side builder).
//
// The reference walks `Vec<Rc<RefCell<Agent>>>` per agent (src/agent.rs:322-332).  Here the two
// link lists of every agent are re-encoded once, at create time, for the day kernel:
//
// Internal agent order ("slots"): [ hub agents | neighbourhood 0 | neighbourhood 1 | ... ],
//   hubs   = agents with more than kHubDegree links, dealt by descending degree round the hub slices (slice j % n,
//            place j / n), so that every hub slice holds about the same number of links
//   groups = the remaining residents of one neighbourhood, by degree descending
//   every region is padded to a multiple of 64 agents with invalid agents, so every
//   32-agent slice outside the hub region holds residents of ONE neighbourhood and every
//   region of the 2-bit mode vector starts on a 16-byte boundary.
//
// Mode vector: 2 bits per agent in internal order (a u64 per 32-agent slice), double
//   buffered by day parity in HBM.  Each block stages the "hot" part of yesterday's
//   vector in shared memory: the hub region plus a degree-descending prefix of every
//   group, as many agents as fit (everything if N <= capacity).  The rest is "cold"
//   and gathered from L2.
//
// Link codes (u32 per link).  Every agent's links are split into a HOT list (targets staged in shared
// memory) and a COLD list (targets read from the global vector); inside each list social links come
// first, then neighbour links, so "which list" is a position test, not a flag:
//     hot  code = (byte offset of the target's word in the staged vector) << 5 | (30 - 2*position in word)
//     cold code = (byte offset of the target's word in the global vector) << 5 | (30 - 2*position in word)
//   The low 5 bits are the left-shift that brings the target's 2 bits to bits 31:30 of the word
//   (a wrapping funnel shift uses them directly); code >> 5 is the byte address.
//
// Hub slices  (32 hubs, one WARP per hub): the hub's hot and cold lists are cut into 32 contiguous chunks, one
//   per lane, and stored as blocks ("units") of the SELL format below (lane = chunk) of at most 1 024 links each,
//   see hub_segment(); any warp of the block that owns the hub's slice can take a unit.
// SELL slices (32 agents, one LANE per agent): K = longest hot list, Kc = longest cold list of the slice;
//   codes stored k-major so that a warp's loads are contiguous: a "body" of floor(K/4) uint4 per lane
//   (lane's hot links 4q..4q+3), a "tail" of K%4 u32 per lane, then Kc u32 per lane of cold links.
//   Lists shorter than K / Kc are padded with code 0.  Agents are ordered by (hot length, cold length)
//   inside their region, so slices are uniform and the padding is ~0.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/motivate_b200.h"

namespace mvb {

constexpr uint32_t kHubDegree = 252;        // total links above which an agent is a hub (must be <= 255: 8-bit counters)
constexpr uint32_t kInvalid = 0xFFFFFFFFu;

constexpr uint32_t kHubUnitRows = 32;       // links per lane in one hub block: a block ("unit") is at most 1 024 links,
                                            // about one SELL slice of work, so the biggest hubs spread over many warps
// Lengths (per lane) of the next block of a hub whose lists still have rem_hot / rem_cold links to go.
// Host (layout.cpp) and device (day_kernels.cuh) walk a hub's blocks with this same rule.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void hub_segment(uint32_t rem_hot, uint32_t rem_cold, uint32_t& K, uint32_t& Kc) {
    K = (rem_hot + 31u) / 32u;
    if (K > kHubUnitRows) K = kHubUnitRows;
    Kc = (rem_cold + 31u) / 32u;
    if (Kc > kHubUnitRows - K) Kc = kHubUnitRows - K;
}

// Shared memory holds `vec_cap_words` words of the staged mode vector; the last four are never staged into and the
// very last one is kept zero by the day kernel.  Unused entries of the link-code arrays ("padding": lanes whose list
// is shorter than their slice's) hold a code that points there, so they gather mode 0 (Car), which the kernel does
// not count but derives from the list lengths: no per-link bound check.
inline uint32_t staged_capacity_agents(uint32_t vec_cap_words) { return (vec_cap_words - 4u) * 16u / 64u * 64u; }
inline uint32_t pad_code(uint32_t vec_cap_words) { return ((vec_cap_words - 1u) * 4u) << 5; }

// Per-slice state record: what the 32 agents of a slice read (and, for the habit, write) every weekday sits in one
// contiguous 896-byte block (7 lines), so a slice needs one base address and one L2 prefetch for all of it:
//   words [0, 32)    packed static word of agent `lane` (decide.cuh) | kStatValid
//   words [32, 64)   list lengths of a SELL agent: ds_hot | dn_hot << 8 | ds_cold << 16 | dn_cold << 24
//   words [64, 96)   weather_sensitivity
//   words [96, 224)  dense habit, float4 per agent, updated in place
constexpr uint32_t kRecWords = 224, kRecStat = 0, kRecLens = 32, kRecWs = 64, kRecHabit = 96;
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t rec_word(uint32_t slot, uint32_t field) { return (size_t)(slot >> 5) * kRecWords + field + (slot & 31u); }
#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t rec_habit_word(uint32_t slot) { return (size_t)(slot >> 5) * kRecWords + kRecHabit + 4u * (slot & 31u); }

// Which part of yesterday's mode vector a block stages in shared memory:
//   GLOBAL  one set for the whole population: all hubs + a degree-descending prefix of every neighbourhood group
//           (everything, if the population fits).  Every block stages the same set; a part is any subset of slices.
//   LOCAL   for populations larger than shared memory: the neighbour network never leaves the neighbourhood
//           (simulation.rs:165-173: two thirds of all links), so a block working on slices of neighbourhood h stages
//           a set common to all blocks (the hubs + a degree-descending prefix of every group) AND a piece of h ITSELF
//           right behind it.  If every neighbourhood fits in what the hubs leave, that piece is the whole rest of h (all
//           neighbour links are shared-memory gathers) and the common prefix is as long as the room left allows;
//           otherwise there is no common prefix and the piece is the degree-descending prefix of h, as much as fits.
//           A link is hot if its target is in the common set, or if source and target live in the same neighbourhood and
//           the target is in that neighbourhood's piece (hub SOURCES are handled in mixed slices and only get the first
//           kind).  Parts are neighbourhood-pure.  With 8M agents in 10 neighbourhoods this turns 71 % of the links into
//           shared-memory gathers where one global set reaches 34 %.
enum LayoutMode : uint32_t { kLayoutGlobal = 0, kLayoutLocal = 1 };
// n_replicas = replicas stepped by the same launches (the LOCAL layout has more, smaller parts: with many replicas their
// fixed cost outweighs the cold links it saves on populations just above the capacity).  MVB_LAYOUT=global|local overrides.
LayoutMode pick_layout_mode(uint32_t N_pad, uint32_t capacity_agents, uint32_t n_replicas);

// PUSH mode for the cold links of SELL agents (large single populations).  A divergent gather from the L2-resident vector
// costs ~2 clocks per link whatever the path (L1TEX wavefronts: tools/microbench_cold.cu), so where most links are cold
// they are turned round: the SELL slices a handle updates are cut into chunks of <= kPushChunkMax agents whose packed
// counters fit in shared memory; a chunk's cold links are stored as (target code u32, source u16 = social/neighbour bit
// << 15 | agent index in the chunk) SORTED BY TARGET, so that the 32 gathers of a warp fall into a few 128-byte lines, and
// push_kernel (day_kernels.cuh) adds every gathered mode to its source's counter with a shared-memory reduction
// (tools/microbench_push.cu: 1.0-1.9 links per clock per SM against 0.5-0.85 for the pull).  The day kernel then finds
// the counts of an agent's cold links in push_cnt[slot] and its SELL block holds hot rows only (Kc rows are not stored;
// sell_K keeps Kc for the work model).  Hub agents keep their cold rows (pull).  MVB_PUSH=0|1 overrides the choice.
constexpr uint32_t kPushChunkMax = 16384;    // agents per chunk: 8 bytes of counters each = 128 KB of shared memory; index < 2^15.  Measured
                                             // 8 192 / 12 288 / 16 384 / 20 480 / 24 576 (one B200, ms per weekday): 8M agents 0.516 / 0.507 / 0.486 / 0.490 /
                                             // 0.501, 50M H=64 3.60 / 3.58 / 3.43 / 3.47 / 3.71, 50M H=10 - / 4.08 / 4.03 / 4.12 / 4.26
bool pick_push_mode(uint32_t N_pad, uint32_t n_replicas);
// agents per chunk for a handle that updates `own_slots` SELL slots on a grid of `grid` blocks: as large as allowed (the
// longer a chunk's list, the closer its sorted targets), a whole number of rounds of the grid, a multiple of 32
uint32_t push_chunk_agents(uint64_t own_slots, uint32_t grid);

struct uint4x { uint32_t x, y, z, w; };   // host-side mirror of CUDA's uint4 (16 bytes)

struct StageRange {                // one contiguous piece of the mode vector staged in shared memory
    uint32_t src_word;             // first u32 word in the global packed vector (16 agents per word)
    uint32_t n_words;              // multiple of 4 (16 bytes)
    uint32_t dst_word;             // first word in the shared-memory vector
    uint32_t pad;
};

struct GraphLayout {
    uint32_t N = 0;                // real agents
    uint32_t N_pad = 0;            // internal slots (multiple of 64)
    uint32_t n_hub_slices = 0;     // slices [0, n_hub_slices) are hub slices
    uint32_t n_sell_slices = 0;    // the rest
    uint32_t hot_words = 0;        // staged vector size in u32 words
    uint64_t Es = 0, En = 0;       // link counts of the source graph
    std::vector<uint32_t> int2ext; // [N_pad] internal slot -> caller's agent id, kInvalid for padding
    std::vector<uint32_t> tcode;   // [N] link code of every possible target (caller's id) | 1 if the target is hot
                                   // (the low bit is free: the shift field of a code is even); input of the fill kernels.
                                   // Only filled by build_graph_layout: mvb_create* makes it on the device (tcode_kernel)
    std::vector<StageRange> ranges;   // GLOBAL: every range is staged by every block.  LOCAL: [0] = the hub region, [1 + h] = the
                                      // common prefix of group h (may be empty), [1 + H + h] = the piece of neighbourhood h that
                                      // only its own blocks stage, placed behind the n_common = 1 + H common ranges
    uint32_t n_common_ranges = 0;     // ranges every block stages (all of them in the GLOBAL layout)
    uint32_t mode = kLayoutGlobal;
    bool push = false;                // cold links of SELL agents go through push_kernel: SELL blocks hold hot rows only (see above)
    std::vector<uint32_t> nb_slice_begin;  // [H + 1] SELL slices of neighbourhood h: [nb_slice_begin[h], nb_slice_begin[h + 1])
    std::vector<uint32_t> tcold;      // LOCAL, with tcode: the cold code of every target (a neighbourhood-prefix target is hot
                                      // only for sources of its own neighbourhood)
    // hubs
    std::vector<uint64_t> hub_off;         // [n_hub_slices*32] offset of the hub's first block in hub_codes (u32 units)
    std::vector<uint32_t> hub_deg;         // [n_hub_slices*32][4] ds_hot, dn_hot, ds_cold, dn_cold
    uint64_t hub_codes_len = 0;            // u32 entries of the device array hub_codes (incl. prefetch slack)
    std::vector<uint4x> hub_units;         // one record per unit: {offset in hub_codes / 32, hub (row of hub_deg),
                                           //  K | Kc << 8 | social hot links left << 16, social cold links left | cold links left << 16},
                                           //  "left" = at the start of the unit, capped at 1 024
    std::vector<uint32_t> hub_unit_begin;  // [n_hub_slices + 1] first unit of every hub slice
    // SELL
    std::vector<uint64_t> sell_off;        // [n_sell_slices] offset of the slice in sell_codes (u32 units)
    std::vector<uint32_t> sell_K;          // [n_sell_slices] K | Kc << 16
    std::vector<uint4x> sell_meta;         // [n_sell_slices + 64] {off_lo, off_hi, K | Kc << 16, 0}: what the kernel reads
    uint64_t sell_codes_len = 0;           // u32 entries of the device array sell_codes (incl. prefetch slack)
};

// Host part of the layout: orders, sizes and offsets.  The link-code arrays themselves (hub_codes, sell_codes)
// and the per-agent length words (deg) are written on the device by fill_kernels.cuh from the caller's CSR.
// capacity_agents: how many agents of the mode vector fit in shared memory (multiple of 64, >= 64).
// Returns MVB_OK or an error code (message via set_error).
//
// The work comes in three steps so that mvb_create* can hand the middle one, the only one that touches every link,
// to the GPU (count_hot_kernel) while host threads do the other two for many graphs at once:
//   layout_begin      O(N): degrees, hubs, neighbourhood groups in degree order, which agents are hot
//   layout_count_hot  O(E): per agent, how many of its social / neighbour links point at hot agents
//   layout_finish     O(N): final order, slices, offsets, hub blocks
struct LayoutWork {                        // what layout_begin leaves for the later steps
    std::vector<uint64_t> d;               // [N] total links
    std::vector<uint32_t> hubs;            // [hub_pad] hub id per hub slot (kInvalid = free place), see layout.h on the order
    std::vector<std::vector<uint32_t>> groups;   // per neighbourhood: ids of the other agents, (degree desc, id asc)
    std::vector<uint32_t> take, gpad;      // per group: staged agents (a prefix), size rounded up to 64
    std::vector<uint32_t> take_common;     // LOCAL: the part of that prefix every block stages (the rest only its own neighbourhood's)
    uint32_t hub_pad = 0, capacity_agents = 0;
    bool all_hot = false;
    uint32_t mode = kLayoutGlobal, staged_hub_slots = 0;
    std::vector<uint8_t> is_hot;           // [N] 0 = never staged, 1 = staged for every block, 2 = staged for blocks of its own neighbourhood (LOCAL)
    const uint32_t* nsh = nullptr;         // [N] social links with a hot target   (set by layout_count_hot, or by the
    const uint32_t* nnh = nullptr;         // [N] neighbour links with a hot target  caller from the device's counts)
    std::vector<uint32_t> nsh_own, nnh_own;
};
int layout_begin(const mvb_population& pop, uint32_t n_neighbourhoods, uint32_t capacity_agents, GraphLayout& out, LayoutWork& work,
                 uint32_t n_replicas = 1);
void layout_count_hot(const mvb_population& pop, LayoutWork& work);
int layout_finish(const mvb_population& pop, LayoutWork& work, GraphLayout& out, bool want_tcode);
// all three on the host (mvb_partition_plan, mvb_create_partitioned)
int build_graph_layout(const mvb_population& pop, uint32_t n_neighbourhoods, uint32_t capacity_agents,
                       GraphLayout& out, uint32_t n_replicas = 1);

// Agent-partitioned mode: cut the slices (hub slices first, then SELL slices, i.e. internal slot order) into `world`
// contiguous ranges of about equal link count.  bounds[r] .. bounds[r+1] are rank r's slices (global slice index:
// hub slice hs -> hs, SELL slice s -> n_hub_slices + s); bounds has world + 1 entries.
void partition_slices(const GraphLayout& L, uint32_t world, std::vector<uint32_t>& bounds);
// The work model behind it: modelled cost of SELL slice s (needs L.sell_K) in units of one shared-memory gather.
uint64_t sell_slice_work(const GraphLayout& L, uint32_t s);

}  // namespace mvb

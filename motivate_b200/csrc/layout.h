// layout.h — device data layout of one population's link graph (host-side builder).
//
// The reference walks `Vec<Rc<RefCell<Agent>>>` per agent (src/agent.rs:322-332).  Here the two
// link lists of every agent are re-encoded once, at create time, for the day kernel:
//
// Internal agent order ("slots"): [ hub agents | neighbourhood 0 | neighbourhood 1 | ... ],
//   hubs   = agents with more than kHubDegree links, by degree descending
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
// Link codes (u32 per link): bit 31 = neighbour link (else social link)
//                            bit 30 = cold target: bits 0-29 = internal id (global vector)
//                            else    bits 0-29 = position in the staged shared-memory vector
//   kDummyCode = a social link to a reserved shared-memory slot that always reads mode 0;
//   used as padding and subtracted from the Car count afterwards.
//
// Hub slices  (32 hubs, one WARP per hub): CSR rows, each padded to a multiple of 4 codes.
// SELL slices (32 agents, one LANE per agent): K = longest list of the slice; codes stored
//   k-major so that a warp's loads are contiguous: a "body" of floor(K/4) uint4 per lane
//   (lane's links 4q..4q+3) followed by a "tail" of K%4 u32 per lane; lists shorter than K
//   are padded with kDummyCode (degree-sorted slices make this padding ~0).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/motivate_b200.h"

namespace mvb {

constexpr uint32_t kHubDegree = 128;        // total links above which an agent is a hub (must be <= 255)
constexpr uint32_t kCodeNbr = 0x80000000u;
constexpr uint32_t kCodeCold = 0x40000000u;
constexpr uint32_t kCodeMask = 0x3FFFFFFFu;
constexpr uint32_t kInvalid = 0xFFFFFFFFu;

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
    uint32_t hot_words = 0;        // staged vector size in u32 words, including the dummy word (last)
    uint32_t dummy_code = 0;
    uint64_t Es = 0, En = 0;       // link counts of the source graph
    std::vector<uint32_t> int2ext; // [N_pad] internal slot -> caller's agent id, kInvalid for padding
    std::vector<uint32_t> ext2int; // [N]
    std::vector<uint32_t> deg;     // [N_pad] SELL agents: ds | dn << 16 (hub entries unused)
    std::vector<StageRange> ranges;
    // hubs
    std::vector<uint64_t> hub_rowptr;      // [n_hub_slices*32 + 1] in u32 units, rows padded to 4
    std::vector<uint32_t> hub_ds, hub_dn;  // [n_hub_slices*32]
    std::vector<uint32_t> hub_codes;
    // SELL
    std::vector<uint64_t> sell_off;        // [n_sell_slices] offset of the slice in sell_codes (u32 units)
    std::vector<uint32_t> sell_K;          // [n_sell_slices]
    std::vector<uint32_t> sell_order;      // [n_sell_slices] processing order, longest lists first
    std::vector<uint32_t> sell_codes;
};

// capacity_agents: how many agents of the mode vector fit in shared memory (multiple of 64, >= 64).
// Returns MVB_OK or an error code (message via set_error).
int build_graph_layout(const mvb_population& pop, uint32_t n_neighbourhoods, uint32_t capacity_agents,
                       GraphLayout& out);

}  // namespace mvb

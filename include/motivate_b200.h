/*
 * motivate_b200.h — C ABI of the B200-native per-day agent update of Motivate.
 *
 * The reference (0xr0bert/Motivate, Rust) has NO plugin / FFI interface.  The
 * seam this header creates is the body of `simulation::run`
 * (src/simulation.rs:40-53), i.e. the weekday loop at src/simulation.rs:101-136
 * and what it calls:
 *     Agent::update_habit                     src/agent.rs:244-264
 *     Neighbourhood::update_congestion_modifier  src/neighbourhood.rs:57-89
 *     Agent::choose (+ budget, cost, count_in_subgroup)  src/agent.rs:93-240,269-332
 *     statistics::count_*  + CSV row assembly src/statistics.rs:14-93, src/simulation.rs:220-268
 *     the table / ownership effects of `intervene`   src/simulation.rs:304-464
 *
 * A Rust host keeps Parameters / Scenario / Intervention / --generate / CSV and
 * replaces src/simulation.rs:101-136 by calls into this library (see
 * INTEGRATION.md for the `extern "C"` block).  Everything here is plain old data:
 * no C++ / torch types cross the boundary, no exceptions, caller owns every
 * buffer it passes, the library copies what it needs during the call.
 *
 * Codes (src/transport_mode.rs:3-8, src/journey_type.rs:5-9, src/weather.rs:7-10):
 *     mode    : 0 Car, 1 PublicTransport, 2 Cycle, 3 Walk
 *     commute : 0 LocalCommute, 1 CityCommute, 2 DistantCommute
 *     weather : 0 Good, 1 Bad
 *
 * Threading: a handle is single-threaded; distinct handles may be used from
 * distinct host threads (mirrors one rayon worker per replica, src/main.rs:89-121).
 * Every function returns MVB_OK (0) or a negative error code; the message is
 * available from mvb_last_error() (thread-local).  These codes replace the
 * reference's panics (`expect`/`unwrap`, e.g. src/simulation.rs:318,368).
 *
 * There is no CPU fallback: if no CUDA device is usable every entry point that
 * needs one fails with MVB_ERR_CUDA.
 */
#ifndef MOTIVATE_B200_H
#define MOTIVATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVB_ABI_VERSION 1

#define MVB_OK            0
#define MVB_ERR_ARG      -1   /* bad argument / inconsistent sizes / id out of range        */
#define MVB_ERR_CUDA     -2   /* CUDA runtime error or no device                             */
#define MVB_ERR_NOMEM    -3   /* host or device allocation failed                            */
#define MVB_ERR_STATE    -4   /* call not valid in this state                                */
#define MVB_ERR_DOMAIN   -5   /* input violates a precondition of the reference (see below)  */
#define MVB_ERR_COMM     -6   /* communicator (NCCL) failure, or NCCL not loadable, in partitioned mode */

#define MVB_N_MODES       4
#define MVB_N_COMMUTES    3
/* Fixed part of a statistics row, column order of src/simulation.rs:207-211:
 * ActiveMode, ActiveNorm, ActiveModeCounterToInactiveNorm,
 * InactiveModeCounterToActiveNorm, LocalCommute, CityCommute, DistantCommute,
 * then S subculture columns, then H neighbourhood columns.  (Day and Rain are
 * host-side columns.) */
#define MVB_STATS_FIXED   7

/* ---- inputs ------------------------------------------------------------- */

/* One population = what `load_unlinked_agents_from_file` + `link_agents`
 * (src/agent_generation.rs:26-55, src/simulation.rs:155-189) leave in memory,
 * as structure-of-arrays.  Agent ids are positions 0..n_agents-1 (the ids used
 * by config/networks/N.yaml, src/simulation.rs:274-291).  Link lists keep
 * duplicates: a friend listed twice counts twice (src/agent.rs:322-332). */
/* The arrays may live in host memory (they are copied to the device during mvb_create*; pinned memory makes those
 * copies asynchronous) or, all of them, in device memory of the GPU the handle is created on (then they are read in
 * place during mvb_create* and must not be written until it returns). */
typedef struct mvb_population {
    uint32_t n_agents;
    const uint8_t*  subculture;            /* [N] index into mvb_tables.desirability rows           */
    const uint16_t* neighbourhood;         /* [N] index into mvb_tables.supportiveness/capacity rows */
    const uint8_t*  commute;               /* [N] commute code                                       */
    const uint8_t*  owns_car;              /* [N] 0/1                                                */
    const uint8_t*  owns_bike;             /* [N] 0/1                                                */
    const float*    weather_sensitivity;   /* [N]                                                    */
    /* Per-agent weights (src/agent.rs:44-57).  Each may be NULL, meaning the
     * uniform value from mvb_params (what every generated population has,
     * src/agent_generation.rs:226,243-246). */
    const float*    consistency;               /* [N] or NULL */
    const float*    social_connectivity;       /* [N] or NULL */
    const float*    subculture_connectivity;   /* [N] or NULL */
    const float*    neighbourhood_connectivity;/* [N] or NULL */
    const float*    average_weight;            /* [N] or NULL */
    const float*    habit;                 /* [N][4] dense; absent key == 0.0f (src/agent.rs:62)      */
    const uint8_t*  current_mode;          /* [N] */
    const uint8_t*  norm;                  /* [N] */
    const uint64_t* social_rowptr;         /* [N+1] CSR of agent.social_network                       */
    const uint32_t* social_col;            /* [social_rowptr[N]]                                      */
    const uint64_t* neighbour_rowptr;      /* [N+1] CSR of agent.neighbours                           */
    const uint32_t* neighbour_col;         /* [neighbour_rowptr[N]]                                   */
} mvb_population;

/* Scenario tables (src/neighbourhood.rs:13-31, src/subculture.rs:9-14).  All four
 * modes must be present for every row (SURVEY.md §8 a-K); the host rejects
 * scenarios that omit one. */
typedef struct mvb_tables {
    uint32_t n_subcultures;                /* S */
    uint32_t n_neighbourhoods;             /* H */
    const float*    supportiveness;        /* [H][4] */
    const uint32_t* capacity;              /* [H][4] */
    const float*    desirability;          /* [S][4] */
} mvb_tables;

/* Uniform per-agent weights used where the population passes NULL
 * (src/main.rs:194-198, src/agent_generation.rs:226,246). */
typedef struct mvb_params {
    float consistency;                     /* reference: 1.0f                                         */
    float social_connectivity;
    float subculture_connectivity;
    float neighbourhood_connectivity;
    float average_weight;                  /* 2.0f / (days_in_habit_average as f32 + 1.0f)            */
    uint32_t flags;                        /* MVB_FLAG_*                                              */
} mvb_params;

#define MVB_FLAG_NONE 0u

/* The effect of `intervene` (src/simulation.rs:304-464) on the hot path, already
 * resolved by the host: the tables and ownership bits that hold FROM `day` on.
 * Applied at the start of day `day`, before the weekday gate
 * (src/simulation.rs:103-108).  owns_* may be NULL (= unchanged). */
typedef struct mvb_intervention {
    uint32_t day;
    const mvb_tables* tables;              /* NULL = unchanged */
    const uint8_t* owns_car;               /* [N] or NULL */
    const uint8_t* owns_bike;              /* [N] or NULL */
} mvb_intervention;

typedef struct mvb_sim mvb_sim;            /* opaque: R independent replicas resident on one device */

/* ---- lifecycle ---------------------------------------------------------- */

int  mvb_abi_version(void);
/* Number of usable CUDA devices (0 if none; never fails). */
int  mvb_device_count(void);
const char* mvb_last_error(void);

/* Create one replica on `device`.  Copies everything it needs. */
int  mvb_create(const mvb_population* pop, const mvb_tables* tab, const mvb_params* prm,
                int device, mvb_sim** out);

/* Create R replicas resident on one device and stepped together (one launch
 * per weekday for all of them).  This is the rayon fan-out of
 * src/main.rs:89-121 mapped onto one GPU.  pops[r] / tabs[r] may alias: when
 * two replicas point at the same graph arrays (same pointers) the graph is
 * stored once on the device (scenario sweeps over one population). */
int  mvb_create_batch(uint32_t n_replicas, const mvb_population* const* pops,
                      const mvb_tables* const* tabs, const mvb_params* prm,
                      int device, mvb_sim** out);

void mvb_destroy(mvb_sim* sim);

uint32_t mvb_n_replicas(const mvb_sim* sim);
/* Row width of replica r: MVB_STATS_FIXED + S + H. */
uint32_t mvb_stats_width(const mvb_sim* sim, uint32_t replica);

/* ---- the day loop -------------------------------------------------------- */

/* One weekday for all replicas: update_habit, update_congestion_modifier,
 * choose (src/simulation.rs:116-128).  `weather_changed` is
 * `weather != new_weather` of src/simulation.rs:127.  Asynchronous on the
 * handle's stream. */
int  mvb_step(mvb_sim* sim, uint8_t weather_today, uint8_t weather_changed);

/* The loop of src/simulation.rs:95-136 for all replicas, on the device:
 *   prev = weather[0]; emit row(day 0);
 *   for day in 1..n_days { if iv && day==iv.day {apply};
 *        if day%7<5 { step(weather[day], prev!=weather[day]); prev=weather[day]; emit row(day) } }
 * `ivs` is NULL or an array of n_replicas pointers (entries may be NULL).
 * rows_out (nullable): [n_replicas][mvb_run_rows(n_days)][stats_width] u32, replica-major,
 * days_out (nullable): [mvb_run_rows(n_days)] the Day column.
 * Synchronous: returns after the results are in the caller's buffers. */
int  mvb_run(mvb_sim* sim, const uint8_t* weather, uint32_t n_days,
             const mvb_intervention* const* ivs,
             uint32_t* days_out, uint32_t* rows_out);
/* Rows mvb_run emits for n_days: 1 (day 0) + #{1<=d<n_days : d%7<5}. */
uint32_t mvb_run_rows(uint32_t n_days);

/* Same loop, nothing copied back; continues from the current state and leaves
 * the rows on the device (fetch with mvb_fetch_rows).  first_day lets a caller
 * run the calendar in pieces: days first_day..end_day-1, `prev_weather` carried
 * in the handle.  When first_day==0 the day-0 row is emitted first. */
int  mvb_run_async(mvb_sim* sim, const uint8_t* weather, uint32_t first_day, uint32_t end_day,
                   const mvb_intervention* const* ivs);
/* Copy the rows produced since the last mvb_run_async(first_day==0,...) /
 * mvb_reset_rows: returns the row count through n_rows; buffers nullable. */
int  mvb_fetch_rows(mvb_sim* sim, uint32_t* n_rows, uint32_t* days_out, uint32_t* rows_out,
                    size_t rows_capacity_u32);
int  mvb_reset_rows(mvb_sim* sim);

/* Block until everything queued on the handle's stream is done. */
int  mvb_sync(mvb_sim* sim);
/* cudaStream_t of the handle, for callers that time with their own events. */
void* mvb_stream(mvb_sim* sim);
/* Device time (ms, CUDA events on the handle's stream) of the day loop of the
 * last mvb_run / mvb_run_async, and how many day-kernels it launched. */
int  mvb_last_run_stats(mvb_sim* sim, float* device_ms, uint64_t* kernel_launches);

/* ---- intervention / state ------------------------------------------------ */

/* Replace a replica's tables (table part of `intervene`). */
int  mvb_set_tables(mvb_sim* sim, uint32_t replica, const mvb_tables* tab);
/* Replace ownership bits (car / bike part of `intervene`); either may be NULL. */
int  mvb_set_ownership(mvb_sim* sim, uint32_t replica, const uint8_t* owns_car, const uint8_t* owns_bike);

/* Statistics row of the current state (src/simulation.rs:220-268 without Day,Rain). */
int  mvb_stats_row(mvb_sim* sim, uint32_t replica, uint32_t* row_out);

/* Full mutable state in the caller's agent order; any pointer may be NULL.
 * hist = residents' current_mode histogram [H][4] (tomorrow's congestion input,
 * src/neighbourhood.rs:68-75); congestion = the modifier used by the last step
 * (1.0f where the reference's map has no key). */
int  mvb_get_state(mvb_sim* sim, uint32_t replica, float* habit, uint8_t* current_mode,
                   uint8_t* last_mode, uint8_t* norm, uint32_t* hist, float* congestion);

/* Bytes of HBM the kernel is expected to move for one weekday of replica r,
 * by the definition of SURVEY.md §8(d): sum over agents of
 * 4*(d_s+d_n) + 51.  Used by bench.py for the roofline. */
uint64_t mvb_algorithmic_bytes_per_day(const mvb_sim* sim, uint32_t replica);

/* ---- one population partitioned over the GPUs of a box ---------------------------------------
 * BASELINE config 4 / SURVEY.md §8e: the reference cannot do this (one replica is single-threaded,
 * src/simulation.rs:64); it is the "number_of_people" axis scaled over GPUs.  One process per GPU.  Every rank
 * passes the SAME population and tables; the library cuts the agents into `world` contiguous ranges of its internal
 * order, balanced by modelled work, and each rank updates its own agents (and stores the link codes of those only).
 * The social network is global (src/simulation.rs:160): tomorrow every rank needs every agent's mode, and the congestion
 * needs the whole population's histogram.  Two exchange paths, same results bit for bit (mvb_exchange_mode says which):
 *   MVB_EXCHANGE_P2P   (default where the ranks' GPUs can map each other's memory: CUDA IPC over NVLink / NVSwitch)
 *       the exchange is fused into the weekday kernel: the warp that packs 32 new modes into a word stores that word
 *       into EVERY rank's vector with one store instruction (one lane per rank); the last block of a rank to finish
 *       copies the rank's partial counters into every rank's landing zone and raises the rank's flag there; a one-block
 *       kernel waits for all flags and adds the partial counters up.  No collective, no pack / unpack pass.
 *   MVB_EXCHANGE_NCCL  (fallback, or MVB_EXCHANGE=nccl) after every weekday ONE ncclAllGather of the packed 2-bit
 *       current_mode words each rank owns and ONE ncclAllReduce of the day's counters (hist + statistics).
 * The norm words, which no other agent's update reads, travel with the last weekday of a mvb_run / mvb_run_async call
 * (and with every mvb_step).  Afterwards every rank holds the full mode / norm vectors and identical statistics rows;
 * habits are current only for the agents a rank owns (mvb_partition_owner).  NCCL always carries the set-up (memory
 * handles, barriers).  All other entry points (mvb_step, mvb_run, mvb_set_*, mvb_get_state, ...) work as for a
 * one-replica handle and must be called by all ranks in the same order. */
#define MVB_UNIQUE_ID_BYTES 128
/* Rank 0 calls this and hands the 128 bytes to the other ranks (file, socket, torch.distributed broadcast ...). */
int  mvb_comm_unique_id(void* id_out);
int  mvb_create_partitioned(const mvb_population* pop, const mvb_tables* tab, const mvb_params* prm,
                            int device, uint32_t rank, uint32_t world, const void* unique_id, mvb_sim** out);
/* rank_out[N]: which rank updates each agent (caller's agent order).  Host-only overload below needs no device. */
int  mvb_partition_owner(const mvb_sim* sim, uint32_t* rank_out);
/* The same plan without creating anything (no device needed): how `world` ranks would split this population. */
int  mvb_partition_plan(const mvb_population* pop, uint32_t n_subcultures, uint32_t n_neighbourhoods,
                        uint32_t world, uint32_t* rank_out);
/* How this handle's ranks exchange modes and counters after a weekday (MVB_EXCHANGE_NONE: not partitioned). */
#define MVB_EXCHANGE_NONE 0
#define MVB_EXCHANGE_NCCL 1
#define MVB_EXCHANGE_P2P  2
int  mvb_exchange_mode(const mvb_sim* sim);

/* ---- synthetic populations (host-side utility, seeded, no device needed) --- */

/* Generates a population following src/agent_generation.rs:114-325 and
 * src/social_network.rs:12-48 / src/simulation.rs:165-189 with a seeded
 * SplitMix64 stream instead of thread_rng / HC-128.  The arrays
 * are owned by the returned object; mvb_synth_population() exposes them. */
typedef struct mvb_synth mvb_synth;

typedef struct mvb_synth_spec {
    uint32_t n_agents;
    uint32_t n_subcultures;
    uint32_t n_neighbourhoods;
    uint32_t social_links_min;         /* number_of_social_network_links (>=2)          */
    uint32_t neighbour_links_min;      /* number_of_neighbour_links (>=2)               */
    uint32_t n_cars;                   /* scenario.number_of_cars                        */
    uint32_t n_bikes;                  /* bikes handed out (reference uses number_of_cars,
                                          src/agent_generation.rs:160)                   */
    uint64_t seed;
    /* commute GMM, (mean, sd, weight) x n_gmm (src/main.rs:209-211); n_gmm<=8 */
    uint32_t n_gmm;
    double   gmm[8][3];
} mvb_synth_spec;

int  mvb_synth_create(const mvb_synth_spec* spec, mvb_synth** out);
/* The same kind of population generated ON THE GPU `device` (every agent and every link is a thread;
 * preferential attachment by chasing the chain of endpoint picks, see csrc/synth_device.cuh): 50M agents with
 * 1.5 G link entries take well under a second where the host generator needs minutes.  The arrays
 * mvb_synth_population() exposes then live in DEVICE memory of that GPU; mvb_create* accepts such a population as
 * it is (no host copy is made), mvb_synth_copy_to_host() brings it to the host.  Its own seeded stream: the same
 * distributions as mvb_synth_create, not the same draws. */
int  mvb_synth_create_device(const mvb_synth_spec* spec, int device, mvb_synth** out);
int  mvb_synth_copy_to_host(const mvb_synth* device_population, mvb_synth** host_out);
void mvb_synth_destroy(mvb_synth* s);
const mvb_population* mvb_synth_population(const mvb_synth* s);
/* Barabási–Albert network exactly as src/social_network.rs:12-48 builds it
 * (clique on the first `min` nodes, then `min` draws with replacement weighted
 * by degree), as CSR.  Caller frees with mvb_free. */
int  mvb_synth_ba_network(uint32_t min_links, uint32_t n_nodes, uint64_t seed,
                          uint64_t** rowptr_out, uint32_t** col_out);
void mvb_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MOTIVATE_B200_H */

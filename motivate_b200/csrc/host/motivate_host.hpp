// motivate_host.hpp — C++ host layer above the C ABI (include/motivate_b200.h) mirroring the reference's own
// interface for this path: same YAML files in, same CSV out, same names and argument meaning, and what panics there
// throws std::runtime_error here with the reference's message.
//
//   Parameters / Parameters::from_file        src/main.rs:185-228
//   Scenario / Scenario::from_file             src/scenario.rs:11-45 (+ neighbourhood.rs:13-31, subculture.rs:9-14)
//   Intervention, NeighbourhoodChange, SubcultureChange      src/intervention.rs:6-52
//   Weather::make_pattern                      src/weather.rs:19-50 (matrix of src/main.rs:73-85; StdRng = HC-128 restated)
//   read_network / write_network               src/main.rs:135-148, 173-179
//   load_unlinked_agents_from_file / save_agents   src/agent_generation.rs:26-63
//   link_agents_to_neighbours                  src/simulation.rs:165-189
//   intervene                                  src/simulation.rs:304-464 (resolved into tables + ownership bits)
//   generate_csv_header / generate_csv_output  src/simulation.rs:194-268
//   simulation::run                            src/simulation.rs:40-149 — the weekday loop itself is mvb_run on the GPU
//   run_replicas                               the rayon fan-out of src/main.rs:89-121 as one batch per GPU
// This layer prepares inputs and formats outputs; it does no per-agent arithmetic.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../../include/motivate_b200.h"

namespace motivate {

enum TransportMode { Car = 0, PublicTransport = 1, Cycle = 2, Walk = 3 };          // src/transport_mode.rs:3-8
enum JourneyType { LocalCommute = 0, CityCommute = 1, DistantCommute = 2 };        // src/journey_type.rs:5-9
extern const char* const kModeNames[4];
extern const char* const kJourneyNames[3];
int mode_from_name(const std::string& s);                                           // throws on an unknown name

struct Parameters {                                                                  // src/main.rs:185-212
    uint32_t total_years = 0, number_of_people = 0, number_of_simulations = 0;
    float social_connectivity = 0, subculture_connectivity = 0, neighbourhood_connectivity = 0;
    uint32_t number_of_social_network_links = 0, number_of_neighbour_links = 0, days_in_habit_average = 0;
    std::vector<std::array<double, 3>> distributions;                                // (mean, sd, weight)
    static Parameters from_file(const std::string& path);
};

struct NeighbourhoodChange {                                                         // src/intervention.rs:24-39
    std::string id;
    std::map<int, float> increase_in_supportiveness;                                 // #[serde(default)]
    std::map<int, long long> increase_in_capacity;                                   // #[serde(default)]
};
struct SubcultureChange { std::string id; std::map<int, float> increase_in_desirability; };
struct Intervention {                                                                // src/intervention.rs:6-21
    uint32_t day = 0;
    std::vector<NeighbourhoodChange> neighbourhood_changes;
    std::vector<SubcultureChange> subculture_changes;
    int change_in_number_of_bikes = 0, change_in_number_of_cars = 0;
};

struct Scenario {                                                                    // src/scenario.rs:11-27
    std::string id;
    std::vector<std::string> subculture_ids, neighbourhood_ids;
    std::vector<float> supportiveness, desirability;                                 // [H][4], [S][4] in TransportMode order
    std::vector<uint32_t> capacity;                                                  // [H][4]
    uint32_t number_of_bikes = 0, number_of_cars = 0;
    Intervention intervention;
    static Scenario from_file(const std::string& path);
    mvb_tables tables() const;
};

// What `load_unlinked_agents_from_file` + `link_agents` leave in memory, as the SoA the C ABI takes.
struct Population {
    std::vector<uint8_t> subculture, commute, owns_car, owns_bike, current_mode, norm;
    std::vector<uint16_t> neighbourhood;
    std::vector<float> weather_sensitivity, consistency, social_connectivity, subculture_connectivity,
        neighbourhood_connectivity, average_weight, habit;
    std::vector<uint64_t> social_rowptr, neighbour_rowptr;
    std::vector<uint32_t> social_col, neighbour_col;
    uint32_t size() const { return (uint32_t)subculture.size(); }
    mvb_population view() const;
};

struct ResolvedIntervention {                      // tables and ownership that hold from `day` on
    uint32_t day = 0;
    bool fires = false;                            // day >= 1 (the loop starts at day 1, src/simulation.rs:101-103)
    std::vector<float> supportiveness, desirability;
    std::vector<uint32_t> capacity;
    std::vector<uint8_t> owns_car, owns_bike;      // empty = unchanged
    uint32_t S = 0, H = 0;
    mvb_tables tables() const { return mvb_tables{S, H, supportiveness.data(), capacity.data(), desirability.data()}; }
};

// `StdRng::from_seed` of rand 0.5.3 (Cargo.lock), which weather.rs:24 draws from: the HC-128 stream cipher, key = seed
// bytes 0-15, IV = bytes 16-31; `next_u64` = two keystream words, low first; `gen::<f64>()` = (u64 >> 11) * 2^-53.
// The crate is not under /root/reference: HC-128 is restated from its specification and pinned on the
// specification's zero key / IV keystream (tests/test_weather.py).
class Hc128 {
public:
    Hc128(const uint32_t key[4], const uint32_t iv[4]);
    uint32_t next_u32() { return step(); }
private:
    uint32_t step();
    uint32_t p_[512], q_[512], i_ = 0;
    friend class StdRng;
};
class StdRng {
public:
    explicit StdRng(const uint8_t seed[32]);
    uint64_t next_u64() { const uint64_t lo = core_.next_u32(); return lo | (uint64_t)core_.next_u32() << 32; }
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
private:
    Hc128 core_;
};
// Weather::make_pattern (src/weather.rs:19-50) with the matrix of src/main.rs:73-85: the rain sequence of a reference run.
std::vector<uint8_t> make_pattern(size_t days, double good_given_good = 0.886, double good_given_bad = 0.699,
                                  double chance_of_rain = 0.14);
// Same chain from this repo's own seeded generator (benchmarks and tests that want several different sequences).
std::vector<uint8_t> make_weather_pattern(size_t days, uint64_t seed = 0);
void read_network(const std::string& path, uint32_t n_agents, std::vector<uint64_t>& rowptr, std::vector<uint32_t>& col);
void write_network(const std::string& path, const std::vector<uint64_t>& rowptr, const std::vector<uint32_t>& col);
Population load_unlinked_agents_from_file(const std::string& path, const Scenario& scenario);
void save_agents(const std::string& path, const Population& pop, const Scenario& scenario);
void link_agents_to_neighbours(Population& pop, uint32_t n_neighbourhoods, uint32_t number_of_neighbour_links, uint64_t seed);
Population generate_unlinked_agents(const Scenario& scenario, float social_connectivity, float subculture_connectivity,
                                    float neighbourhood_connectivity, uint32_t days_in_habit_average,
                                    uint32_t number_of_people, const std::vector<std::array<double, 3>>& distributions,
                                    uint64_t seed);
ResolvedIntervention intervene(const Scenario& scenario, const Population& pop, uint64_t seed);
std::string generate_csv_header(const Scenario& scenario);
std::string generate_csv_output(uint32_t day, int rain, const uint32_t* row, uint32_t width);

struct Prepared {                                  // one replica, ready for mvb_create*
    std::string id;
    Scenario scenario;
    Population population;
    ResolvedIntervention intervention;
    mvb_params params{};
};

namespace simulation {
// Setup part of `run` (src/simulation.rs:62-78).
Prepared prepare(const std::string& id, bool generate, const std::string& agents_file, const std::string& scenario_file,
                 uint32_t number_of_people, float social_connectivity, float subculture_connectivity,
                 float neighbourhood_connectivity, uint32_t number_of_neighbour_links, uint32_t days_in_habit_average,
                 const std::vector<std::array<double, 3>>& distributions, const std::string& network_file, uint64_t seed);
// src/simulation.rs:40-149: writes {output_dir}/output_{id}.csv.
void run(const std::string& id, bool generate, const std::string& agents_file, const std::string& scenario_file,
         uint32_t total_years, uint32_t number_of_people, float social_connectivity, float subculture_connectivity,
         float neighbourhood_connectivity, uint32_t number_of_neighbour_links, uint32_t days_in_habit_average,
         const std::vector<std::array<double, 3>>& distributions, const std::vector<uint8_t>& weather_pattern,
         const std::string& network_file, const std::string& output_dir, int device, uint64_t seed);
// All replicas of one GPU as one batch (src/main.rs:89-121): one launch per weekday steps them all.
void run_replicas(const std::vector<Prepared>& replicas, uint32_t total_years, const std::vector<uint8_t>& weather_pattern,
                  const std::string& output_dir, int device);
}  // namespace simulation

}  // namespace motivate

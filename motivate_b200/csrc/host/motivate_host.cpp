// motivate_host.cpp — see motivate_host.hpp.  Plain C++17 over the C ABI; no CUDA in this file.
#include "motivate_host.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>

#include "yaml_lite.hpp"

namespace motivate {

const char* const kModeNames[4] = {"Car", "PublicTransport", "Cycle", "Walk"};
const char* const kJourneyNames[3] = {"LocalCommute", "CityCommute", "DistantCommute"};

namespace {

struct SplitMix64 {                                  // the seeded stream used instead of thread_rng / HC-128
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double f64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
};

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }
void check(int rc, const char* what) {
    if (rc != MVB_OK) fail(std::string(what) + ": " + mvb_last_error());
}
float f32_of(const yaml_lite::Node& n) { return (float)n.as_double(); }     // decimal -> f64 -> f32 like serde `as f32`

template <class T, class F>
std::map<int, T> mode_map(const yaml_lite::Node* n, const char* what, bool require_all, F conv) {
    std::map<int, T> out;
    if (n && n->is_map())
        for (auto& kv : n->map) out[mode_from_name(kv.first)] = conv(kv.second);
    if (require_all && out.size() != 4)
        fail(std::string(what) + ": all four transport modes are required");      // SURVEY.md §8 a-K
    return out;
}

int index_of(const std::vector<std::string>& v, const std::string& s) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i] == s) return (int)i;
    return -1;
}

}  // namespace

int mode_from_name(const std::string& s) {
    for (int m = 0; m < 4; ++m) if (s == kModeNames[m]) return m;
    fail("There was an error parsing the file: unknown variant `" + s + "`, expected one of `Car`, `PublicTransport`, `Cycle`, `Walk`");
}

Parameters Parameters::from_file(const std::string& path) {
    const yaml_lite::Node d = yaml_lite::parse_file(path);
    Parameters p;
    p.total_years = (uint32_t)d.at("total_years").as_int();
    p.number_of_people = (uint32_t)d.at("number_of_people").as_int();
    p.number_of_simulations = (uint32_t)d.at("number_of_simulations").as_int();
    p.social_connectivity = f32_of(d.at("social_connectivity"));
    p.subculture_connectivity = f32_of(d.at("subculture_connectivity"));
    p.neighbourhood_connectivity = f32_of(d.at("neighbourhood_connectivity"));
    p.number_of_social_network_links = (uint32_t)d.at("number_of_social_network_links").as_int();
    p.number_of_neighbour_links = (uint32_t)d.at("number_of_neighbour_links").as_int();
    p.days_in_habit_average = (uint32_t)d.at("days_in_habit_average").as_int();
    for (auto& t : d.at("distributions").seq) {
        if (t.seq.size() != 3) fail("There was an error parsing the file: distributions entries are (mean, sd, weight)");
        p.distributions.push_back({t.seq[0].as_double(), t.seq[1].as_double(), t.seq[2].as_double()});
    }
    return p;
}

Scenario Scenario::from_file(const std::string& path) {
    const yaml_lite::Node d = yaml_lite::parse_file(path);
    Scenario sc;
    sc.id = d.at("id").str();
    for (auto& s : d.at("subcultures").seq) {
        sc.subculture_ids.push_back(s.at("id").str());
        auto m = mode_map<float>(s.find("desirability"), ("subculture " + sc.subculture_ids.back()).c_str(), true, f32_of);
        for (int k = 0; k < 4; ++k) sc.desirability.push_back(m[k]);
    }
    for (auto& n : d.at("neighbourhoods").seq) {
        sc.neighbourhood_ids.push_back(n.at("id").str());
        const std::string what = "neighbourhood " + sc.neighbourhood_ids.back();
        auto s = mode_map<float>(n.find("supportiveness"), what.c_str(), true, f32_of);
        auto c = mode_map<long long>(n.find("capacity"), what.c_str(), true, [](const yaml_lite::Node& x) { return x.as_int(); });
        for (int k = 0; k < 4; ++k) {
            sc.supportiveness.push_back(s[k]);
            if (c[k] < 0 || c[k] > 0xFFFFFFFFll) fail(what + ": capacity does not fit u32");
            sc.capacity.push_back((uint32_t)c[k]);
        }
    }
    sc.number_of_bikes = (uint32_t)d.at("number_of_bikes").as_int();
    sc.number_of_cars = (uint32_t)d.at("number_of_cars").as_int();
    const yaml_lite::Node& iv = d.at("intervention");
    sc.intervention.day = (uint32_t)iv.at("day").as_int();
    sc.intervention.change_in_number_of_bikes = (int)iv.at("change_in_number_of_bikes").as_int();
    sc.intervention.change_in_number_of_cars = (int)iv.at("change_in_number_of_cars").as_int();
    if (const yaml_lite::Node* l = iv.find("neighbourhood_changes"))
        for (auto& c : l->seq) {
            NeighbourhoodChange nc;
            nc.id = c.at("id").str();
            nc.increase_in_supportiveness = mode_map<float>(c.find("increase_in_supportiveness"), "increase_in_supportiveness", false, f32_of);
            nc.increase_in_capacity = mode_map<long long>(c.find("increase_in_capacity"), "increase_in_capacity", false,
                                                          [](const yaml_lite::Node& x) { return x.as_int(); });
            sc.intervention.neighbourhood_changes.push_back(nc);
        }
    if (const yaml_lite::Node* l = iv.find("subculture_changes"))
        for (auto& c : l->seq) {
            SubcultureChange s;
            s.id = c.at("id").str();
            s.increase_in_desirability = mode_map<float>(&c.at("increase_in_desirability"), "increase_in_desirability", false, f32_of);
            sc.intervention.subculture_changes.push_back(s);
        }
    return sc;
}

mvb_tables Scenario::tables() const {
    return mvb_tables{(uint32_t)subculture_ids.size(), (uint32_t)neighbourhood_ids.size(), supportiveness.data(),
                      capacity.data(), desirability.data()};
}

mvb_population Population::view() const {
    mvb_population p{};
    p.n_agents = size();
    p.subculture = subculture.data(); p.neighbourhood = neighbourhood.data(); p.commute = commute.data();
    p.owns_car = owns_car.data(); p.owns_bike = owns_bike.data(); p.weather_sensitivity = weather_sensitivity.data();
    p.consistency = consistency.empty() ? nullptr : consistency.data();
    p.social_connectivity = social_connectivity.empty() ? nullptr : social_connectivity.data();
    p.subculture_connectivity = subculture_connectivity.empty() ? nullptr : subculture_connectivity.data();
    p.neighbourhood_connectivity = neighbourhood_connectivity.empty() ? nullptr : neighbourhood_connectivity.data();
    p.average_weight = average_weight.empty() ? nullptr : average_weight.data();
    p.habit = habit.data(); p.current_mode = current_mode.data(); p.norm = norm.data();
    p.social_rowptr = social_rowptr.data(); p.social_col = social_col.data();
    p.neighbour_rowptr = neighbour_rowptr.data(); p.neighbour_col = neighbour_col.data();
    return p;
}

// Weather::make_pattern, src/weather.rs:19-50, with the transition matrix and the 0.14 chance of rain on day 0 of
// src/main.rs:73-85, drawn from a SplitMix64 stream: a seeded realisation of the same chain for synthetic runs.
// The reference's own sequence (StdRng::from_seed([0; 32]) = HC-128) is make_pattern below.
std::vector<uint8_t> make_weather_pattern(size_t days, uint64_t seed) {
    SplitMix64 rng(seed * 0x9E3779B97F4A7C15ull + 104729);
    std::vector<uint8_t> w(days, 0);
    if (!days) return w;
    w[0] = rng.f64() > 0.14 ? 0 : 1;
    for (size_t i = 1; i < days; ++i) {
        const double p_good = w[i - 1] == 0 ? 0.886 : 0.699;
        w[i] = rng.f64() < p_good ? 0 : 1;
    }
    return w;
}

// ---- StdRng of rand 0.5.3 = HC-128 (restated from the cipher's specification; see the header) ----------------
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

uint32_t Hc128::step() {
    const uint32_t i = i_, j = i & 511u;
    i_ = (i + 1) & 1023u;
    const bool first = (i & 512u) == 0;
    uint32_t* t = first ? p_ : q_;
    const uint32_t* o = first ? q_ : p_;
    const uint32_t a = t[(j - 3) & 511u], b = t[(j - 10) & 511u], c = t[(j - 511) & 511u];
    const uint32_t g = first ? (rotr32(a, 10) ^ rotr32(c, 23)) + rotr32(b, 8) : (rotl32(a, 10) ^ rotl32(c, 23)) + rotl32(b, 8);
    t[j] += g;
    const uint32_t x = t[(j - 12) & 511u];
    return (o[x & 0xFFu] + o[256 + ((x >> 16) & 0xFFu)]) ^ t[j];
}

Hc128::Hc128(const uint32_t key[4], const uint32_t iv[4]) {
    std::vector<uint32_t> w(1280);
    for (int k = 0; k < 8; ++k) { w[k] = key[k & 3]; w[8 + k] = iv[k & 3]; }
    for (uint32_t i = 16; i < 1280; ++i) {
        const uint32_t x = w[i - 2], y = w[i - 15];
        w[i] = (rotr32(x, 17) ^ rotr32(x, 19) ^ (x >> 10)) + w[i - 7] + (rotr32(y, 7) ^ rotr32(y, 18) ^ (y >> 3)) + w[i - 16] + i;
    }
    for (int k = 0; k < 512; ++k) { p_[k] = w[256 + k]; q_[k] = w[768 + k]; }
    for (uint32_t k = 0; k < 1024; ++k) {                       // outputs of the first 1024 steps replace the table words
        uint32_t* t = (k & 512u) == 0 ? p_ : q_;
        const uint32_t s = step();
        t[k & 511u] = s;
    }
    i_ = 0;
}

static Hc128 hc128_from_seed(const uint8_t seed[32]) {
    uint32_t w[8];
    for (int k = 0; k < 8; ++k)
        w[k] = (uint32_t)seed[4 * k] | (uint32_t)seed[4 * k + 1] << 8 | (uint32_t)seed[4 * k + 2] << 16 | (uint32_t)seed[4 * k + 3] << 24;
    return Hc128(w, w + 4);
}
StdRng::StdRng(const uint8_t seed[32]) : core_(hc128_from_seed(seed)) {}

std::vector<uint8_t> make_pattern(size_t days, double good_given_good, double good_given_bad, double chance_of_rain) {
    static const uint8_t zero_seed[32] = {0};
    StdRng rng(zero_seed);                                       // weather.rs:24
    std::vector<uint8_t> w(days, 0);
    if (!days) return w;
    w[0] = rng.gen_f64() > chance_of_rain ? 0 : 1;               // weather.rs:28-32
    for (size_t i = 1; i < days; ++i)                            // weather.rs:35-48
        w[i] = rng.gen_f64() < (w[i - 1] == 0 ? good_given_good : good_given_bad) ? 0 : 1;
    return w;
}

void read_network(const std::string& path, uint32_t n, std::vector<uint64_t>& rowptr, std::vector<uint32_t>& col) {
    const yaml_lite::Node d = yaml_lite::parse_file(path);                      // HashMap<u32, Vec<u32>>
    std::vector<const yaml_lite::Node*> rows(n, nullptr);
    for (auto& kv : d.map) {
        const long long k = std::strtoll(kv.first.c_str(), nullptr, 10);
        if (k < 0 || k >= (long long)n) fail("network id " + kv.first + " is not an agent");   // agents[k] panics there
        rows[k] = &kv.second;
    }
    rowptr.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) rowptr[i + 1] = rowptr[i] + (rows[i] ? rows[i]->seq.size() : 0);
    col.resize(rowptr[n]);
    for (uint32_t i = 0; i < n; ++i)
        if (rows[i]) {
            uint64_t o = rowptr[i];
            for (auto& x : rows[i]->seq) {
                const long long t = x.as_int();
                if (t < 0 || t >= (long long)n) fail("network of agent " + std::to_string(i) + " names an agent that does not exist");
                col[o++] = (uint32_t)t;
            }
        }
}

void write_network(const std::string& path, const std::vector<uint64_t>& rowptr, const std::vector<uint32_t>& col) {
    std::ofstream f(path);
    if (!f) fail("File cannot be created: " + path);
    f << "---\n";
    for (size_t i = 0; i + 1 < rowptr.size(); ++i) {
        if (rowptr[i] == rowptr[i + 1]) { f << i << ": []\n"; continue; }
        f << i << ":\n";
        for (uint64_t x = rowptr[i]; x < rowptr[i + 1]; ++x) f << "  - " << col[x] << "\n";
    }
}

Population load_unlinked_agents_from_file(const std::string& path, const Scenario& sc) {
    const yaml_lite::Node d = yaml_lite::parse_file(path);
    Population p;
    const size_t n = d.seq.size();
    p.subculture.resize(n); p.neighbourhood.resize(n); p.commute.resize(n); p.owns_car.resize(n); p.owns_bike.resize(n);
    p.weather_sensitivity.resize(n); p.consistency.resize(n); p.social_connectivity.resize(n);
    p.subculture_connectivity.resize(n); p.neighbourhood_connectivity.resize(n); p.average_weight.resize(n);
    p.habit.assign(4 * n, 0.f); p.current_mode.resize(n); p.norm.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const yaml_lite::Node& a = d.seq[i];
        const int s = index_of(sc.subculture_ids, a.at("subculture_id").str());
        if (s < 0) fail("Subculture not found");                               // agent_generation.rs:47
        const int h = index_of(sc.neighbourhood_ids, a.at("neighbourhood_id").str());
        if (h < 0) fail("Agent not found");                                    // agent_generation.rs:49 (sic)
        p.subculture[i] = (uint8_t)s; p.neighbourhood[i] = (uint16_t)h;
        const std::string& cl = a.at("commute_length").str();
        int c = -1;
        for (int k = 0; k < 3; ++k) if (cl == kJourneyNames[k]) c = k;
        if (c < 0) fail("There was an error parsing the file: unknown commute_length " + cl);
        p.commute[i] = (uint8_t)c;
        p.owns_car[i] = a.at("owns_car").as_bool(); p.owns_bike[i] = a.at("owns_bike").as_bool();
        p.weather_sensitivity[i] = f32_of(a.at("weather_sensitivity"));
        p.consistency[i] = f32_of(a.at("consistency"));
        p.social_connectivity[i] = f32_of(a.at("social_connectivity"));
        p.subculture_connectivity[i] = f32_of(a.at("subculture_connectivity"));
        p.neighbourhood_connectivity[i] = f32_of(a.at("neighbourhood_connectivity"));
        p.average_weight[i] = f32_of(a.at("average_weight"));
        if (const yaml_lite::Node* hb = a.find("habit"))
            for (auto& kv : hb->map) p.habit[4 * i + mode_from_name(kv.first)] = f32_of(kv.second);
        p.current_mode[i] = (uint8_t)mode_from_name(a.at("current_mode").str());
        p.norm[i] = (uint8_t)mode_from_name(a.at("norm").str());
    }
    return p;
}

void save_agents(const std::string& path, const Population& p, const Scenario& sc) {
    std::ofstream f(path);
    if (!f) fail("File cannot be created: " + path);
    f.precision(9);                                                            // round-trips every f32
    f << "---\n";
    for (uint32_t i = 0; i < p.size(); ++i) {
        const uint32_t m = p.current_mode[i];
        f << "- subculture_id: \"" << sc.subculture_ids[p.subculture[i]] << "\"\n"
          << "  neighbourhood_id: \"" << sc.neighbourhood_ids[p.neighbourhood[i]] << "\"\n"
          << "  commute_length: " << kJourneyNames[p.commute[i]] << "\n  commute_length_continuous: 0.0\n"
          << "  weather_sensitivity: " << p.weather_sensitivity[i] << "\n  consistency: " << p.consistency[i] << "\n"
          << "  social_connectivity: " << p.social_connectivity[i] << "\n  subculture_connectivity: " << p.subculture_connectivity[i]
          << "\n  neighbourhood_connectivity: " << p.neighbourhood_connectivity[i] << "\n  average_weight: " << p.average_weight[i] << "\n"
          << "  habit:\n";
        for (int k = 0; k < 4; ++k)
            if (p.habit[4 * (size_t)i + k] != 0.f) f << "    " << kModeNames[k] << ": " << p.habit[4 * (size_t)i + k] << "\n";
        f << "  current_mode: " << kModeNames[m] << "\n  last_mode: " << kModeNames[m] << "\n  norm: " << kModeNames[p.norm[i]] << "\n"
          << "  owns_bike: " << (p.owns_bike[i] ? "true" : "false") << "\n  owns_car: " << (p.owns_car[i] ? "true" : "false") << "\n";
    }
}

// One Barabási–Albert network per neighbourhood over its residents in agent order, rebuilt at every run because the
// reference never saves them (src/simulation.rs:165-189); seeded, so a --generate run and a later run agree.
void link_agents_to_neighbours(Population& p, uint32_t H, uint32_t min_links, uint64_t seed) {
    const uint32_t n = p.size();
    std::vector<std::vector<uint32_t>> members(H), lists(n);
    for (uint32_t i = 0; i < n; ++i) members[p.neighbourhood[i]].push_back(i);
    for (uint32_t h = 0; h < H; ++h) {
        const auto& mem = members[h];
        if (mem.size() < 2) continue;
        uint64_t* rp = nullptr; uint32_t* cl = nullptr;
        check(mvb_synth_ba_network(min_links, (uint32_t)mem.size(), seed * 1000003ull + h, &rp, &cl), "neighbour network");
        for (size_t k = 0; k < mem.size(); ++k)
            for (uint64_t x = rp[k]; x < rp[k + 1]; ++x) lists[mem[k]].push_back(mem[cl[x]]);
        mvb_free(rp); mvb_free(cl);
    }
    p.neighbour_rowptr.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) p.neighbour_rowptr[i + 1] = p.neighbour_rowptr[i] + lists[i].size();
    p.neighbour_col.clear();
    p.neighbour_col.reserve(p.neighbour_rowptr[n]);
    for (uint32_t i = 0; i < n; ++i) p.neighbour_col.insert(p.neighbour_col.end(), lists[i].begin(), lists[i].end());
}

// generate_unlinked_agents, src/agent_generation.rs:114-205, through the library's seeded generator.
Population generate_unlinked_agents(const Scenario& sc, float soc, float sub, float nei, uint32_t days_in_habit_average,
                                    uint32_t n, const std::vector<std::array<double, 3>>& distributions, uint64_t seed) {
    mvb_synth_spec sp{};
    sp.n_agents = n; sp.n_subcultures = (uint32_t)sc.subculture_ids.size(); sp.n_neighbourhoods = (uint32_t)sc.neighbourhood_ids.size();
    sp.social_links_min = 2; sp.neighbour_links_min = 2;                        // both graphs are replaced by the caller
    sp.n_cars = sc.number_of_cars; sp.n_bikes = sc.number_of_cars;              // bikes use number_of_cars (:160, sic)
    sp.seed = seed;
    static const double kDefault[3][3] = {{9845.09, 3691.81, 0.66987}, {2205.45, 1414.38, 0.25230}, {22606.84, 11965.02, 0.07782}};
    sp.n_gmm = distributions.empty() ? 3 : (uint32_t)std::min<size_t>(distributions.size(), 8);
    for (uint32_t g = 0; g < sp.n_gmm; ++g)
        for (int k = 0; k < 3; ++k) sp.gmm[g][k] = distributions.empty() ? kDefault[g][k] : distributions[g][k];
    mvb_synth* h = nullptr;
    check(mvb_synth_create(&sp, &h), "generate agents");
    const mvb_population* q = mvb_synth_population(h);
    Population p;
    p.subculture.assign(q->subculture, q->subculture + n); p.neighbourhood.assign(q->neighbourhood, q->neighbourhood + n);
    p.commute.assign(q->commute, q->commute + n); p.owns_car.assign(q->owns_car, q->owns_car + n);
    p.owns_bike.assign(q->owns_bike, q->owns_bike + n);
    p.weather_sensitivity.assign(q->weather_sensitivity, q->weather_sensitivity + n);
    p.habit.assign(q->habit, q->habit + 4 * (size_t)n);
    p.current_mode.assign(q->current_mode, q->current_mode + n); p.norm.assign(q->norm, q->norm + n);
    mvb_synth_destroy(h);
    p.consistency.assign(n, 1.0f);                                              // agent_generation.rs:226
    p.social_connectivity.assign(n, soc); p.subculture_connectivity.assign(n, sub); p.neighbourhood_connectivity.assign(n, nei);
    p.average_weight.assign(n, 2.0f / ((float)days_in_habit_average + 1.0f));   // agent_generation.rs:246
    return p;
}

namespace {
void resample(std::vector<uint8_t>& owned, int change, SplitMix64& rng, const char* what) {
    // give to |change| uniformly sampled non-owners or take from |change| owners (src/simulation.rs:383-463)
    if (change == 0) return;
    std::vector<uint32_t> pool;
    for (uint32_t i = 0; i < owned.size(); ++i) if ((owned[i] != 0) == (change < 0)) pool.push_back(i);
    const size_t k = (size_t)std::abs(change);
    if (k > pool.size()) fail(std::string("change_in_number_of_") + what + " exceeds the " + std::to_string(pool.size()) + " eligible agents");
    for (size_t j = 0; j < k; ++j) {                                           // partial Fisher-Yates
        std::swap(pool[j], pool[j + rng.below(pool.size() - j)]);
        owned[pool[j]] = change > 0 ? 1 : 0;
    }
}
}  // namespace

ResolvedIntervention intervene(const Scenario& sc, const Population& pop, uint64_t seed) {
    const Intervention& iv = sc.intervention;
    ResolvedIntervention r;
    r.day = iv.day; r.fires = iv.day >= 1;
    r.supportiveness = sc.supportiveness; r.capacity = sc.capacity; r.desirability = sc.desirability;
    for (auto& c : iv.neighbourhood_changes) {
        const int h = index_of(sc.neighbourhood_ids, c.id);
        if (h < 0) fail("A neighbourhood in your intervention was not found");  // simulation.rs:318
        for (auto& kv : c.increase_in_supportiveness) r.supportiveness[4 * h + kv.first] = r.supportiveness[4 * h + kv.first] + kv.second;
        for (auto& kv : c.increase_in_capacity) {                                // i64 add clamped to [0, u32::MAX] (:328-351)
            const long long v = (long long)r.capacity[4 * h + kv.first] + kv.second;
            r.capacity[4 * h + kv.first] = v < 0 ? 0u : (v > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)v);
        }
    }
    for (auto& c : iv.subculture_changes) {
        const int s = index_of(sc.subculture_ids, c.id);
        if (s < 0) fail("A subculture in your intervention was not found");      // simulation.rs:368
        for (auto& kv : c.increase_in_desirability) r.desirability[4 * s + kv.first] = r.desirability[4 * s + kv.first] + kv.second;
    }
    SplitMix64 rng(seed * 0x2545F4914F6CDD1Dull + 17);
    if (iv.change_in_number_of_bikes) { r.owns_bike = pop.owns_bike; resample(r.owns_bike, iv.change_in_number_of_bikes, rng, "bikes"); }
    if (iv.change_in_number_of_cars)  { r.owns_car = pop.owns_car;   resample(r.owns_car, iv.change_in_number_of_cars, rng, "cars"); }
    r.S = (uint32_t)sc.subculture_ids.size(); r.H = (uint32_t)sc.neighbourhood_ids.size();
    return r;
}

std::string generate_csv_header(const Scenario& sc) {                           // src/simulation.rs:194-212
    std::string s = "Day,Rain,ActiveMode,ActiveNorm,ActiveModeCounterToInactiveNorm,InactiveModeCounterToActiveNorm,"
                    "LocalCommute,CityCommute,DistantCommute,";
    for (size_t i = 0; i < sc.subculture_ids.size(); ++i) s += (i ? "," : "") + sc.subculture_ids[i];
    s += ",";
    for (size_t i = 0; i < sc.neighbourhood_ids.size(); ++i) s += (i ? "," : "") + sc.neighbourhood_ids[i];
    return s + "\n";
}

std::string generate_csv_output(uint32_t day, int rain, const uint32_t* row, uint32_t width) {   // :220-268
    std::string s = std::to_string(day) + "," + std::to_string(rain);
    for (uint32_t k = 0; k < width; ++k) s += "," + std::to_string(row[k]);
    return s + "\n";
}

namespace simulation {

Prepared prepare(const std::string& id, bool generate, const std::string& agents_file, const std::string& scenario_file,
                 uint32_t number_of_people, float soc, float sub, float nei, uint32_t number_of_neighbour_links,
                 uint32_t days_in_habit_average, const std::vector<std::array<double, 3>>& distributions,
                 const std::string& network_file, uint64_t seed) {
    Prepared pr;
    pr.id = id;
    pr.scenario = Scenario::from_file(scenario_file);                           // simulation.rs:62
    if (generate) {
        pr.population = generate_unlinked_agents(pr.scenario, soc, sub, nei, days_in_habit_average, number_of_people, distributions, seed);
        save_agents(agents_file, pr.population, pr.scenario);                   // agent_generation.rs:95
    } else {
        pr.population = load_unlinked_agents_from_file(agents_file, pr.scenario);
    }
    // link_agents, simulation.rs:155-174
    read_network(network_file, pr.population.size(), pr.population.social_rowptr, pr.population.social_col);
    link_agents_to_neighbours(pr.population, (uint32_t)pr.scenario.neighbourhood_ids.size(), number_of_neighbour_links, seed);
    pr.intervention = intervene(pr.scenario, pr.population, seed);
    pr.params = mvb_params{1.0f, soc, sub, nei, 2.0f / ((float)days_in_habit_average + 1.0f), 0u};
    return pr;
}

void run_replicas(const std::vector<Prepared>& reps, uint32_t total_years, const std::vector<uint8_t>& weather,
                  const std::string& output_dir, int device) {
    if (reps.empty()) return;
    const uint32_t n_days = total_years * 365;                                  // loop bound of simulation.rs:101
    if (weather.size() < n_days) fail("weather pattern shorter than total_years * 365");
    const uint32_t R = (uint32_t)reps.size();
    std::vector<mvb_population> pops(R);
    std::vector<mvb_tables> tabs(R);
    std::vector<const mvb_population*> pp(R);
    std::vector<const mvb_tables*> tp(R);
    std::vector<const mvb_intervention*> ivs(R);
    std::vector<mvb_tables> iv_tabs(R);
    std::vector<mvb_intervention> iv_structs(R);
    for (uint32_t r = 0; r < R; ++r) {
        pops[r] = reps[r].population.view(); tabs[r] = reps[r].scenario.tables();
        pp[r] = &pops[r]; tp[r] = &tabs[r];
        const ResolvedIntervention& iv = reps[r].intervention;
        iv_tabs[r] = iv.tables();
        iv_structs[r] = mvb_intervention{iv.day, &iv_tabs[r], iv.owns_car.empty() ? nullptr : iv.owns_car.data(),
                                         iv.owns_bike.empty() ? nullptr : iv.owns_bike.data()};
        ivs[r] = iv.fires ? &iv_structs[r] : nullptr;
    }
    mvb_sim* sim = nullptr;
    check(mvb_create_batch(R, pp.data(), tp.data(), &reps[0].params, device, &sim), "mvb_create_batch");
    const uint32_t rows = mvb_run_rows(n_days);
    std::vector<uint32_t> days(rows), widths(R);
    size_t total = 0;
    for (uint32_t r = 0; r < R; ++r) { widths[r] = mvb_stats_width(sim, r); total += (size_t)rows * widths[r]; }
    std::vector<uint32_t> out(total);
    const int rc = mvb_run(sim, weather.data(), n_days, ivs.data(), days.data(), out.data());   // simulation.rs:95-136
    if (rc != MVB_OK) { const std::string m = mvb_last_error(); mvb_destroy(sim); fail("mvb_run: " + m); }
    mvb_destroy(sim);
    ::mkdir(output_dir.c_str(), 0777);                                          // fs::create_dir_all("output")
    const uint32_t* o = out.data();
    for (uint32_t r = 0; r < R; ++r) {
        std::ofstream f(output_dir + "/output_" + reps[r].id + ".csv");
        if (!f) fail("cannot create " + output_dir + "/output_" + reps[r].id + ".csv");
        f << generate_csv_header(reps[r].scenario);
        for (uint32_t k = 0; k < rows; ++k, o += widths[r]) f << generate_csv_output(days[k], weather[days[k]], o, widths[r]);
    }
}

void run(const std::string& id, bool generate, const std::string& agents_file, const std::string& scenario_file,
         uint32_t total_years, uint32_t number_of_people, float soc, float sub, float nei, uint32_t number_of_neighbour_links,
         uint32_t days_in_habit_average, const std::vector<std::array<double, 3>>& distributions,
         const std::vector<uint8_t>& weather_pattern, const std::string& network_file, const std::string& output_dir,
         int device, uint64_t seed) {
    std::vector<Prepared> one;
    one.push_back(prepare(id, generate, agents_file, scenario_file, number_of_people, soc, sub, nei, number_of_neighbour_links,
                          days_in_habit_average, distributions, network_file, seed));
    run_replicas(one, total_years, weather_pattern, output_dir, device);
}

}  // namespace simulation
}  // namespace motivate

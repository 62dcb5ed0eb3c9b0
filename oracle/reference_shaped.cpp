// reference_shaped.cpp — a THIRD restatement of the reference's per-day update, in the reference's own shape.
//
// TEST INFRASTRUCTURE ONLY (like motivate_oracle.c and literal_model.py): only tests/, __graft_entry__.smoke() and
// bench.py's CPU-baseline legs may load it; the product never does.  parity unpinned: the reference has no tests or
// golden vectors for this path and cannot be built here (no Rust toolchain); this file is checked against the two
// other restatements and the hand-derived micro cases (tests/test_reference_shaped.py).
//
// Where motivate_oracle.c is dense (4-lane arrays, absent key == 0.0) and literal_model.py is too slow for more
// than a few hundred agents, this one keeps what the Rust code does per agent per day — heap objects that point at
// their friends, node-based maps keyed by mode, `union_of` / `intersection_of` folds that allocate a new map each,
// group maps, stable sorts, `max_by` — and is compiled, so it can (a) be compared with the dense oracle on 100k-agent
// populations and (b) stand in for the cost shape of the reference's CPU path in bench.py ("reference-shaped").
// Maps are std::map<int, float>, iterated in key order = the canonical order Car, PublicTransport, Cycle, Walk
// (SURVEY.md §8 a-T); Rust's HashMap order is random per map, any fixed order is one legitimate realisation.
// Same C entry points as motivate_oracle.h, so oracle/binding.py drives both.  Citations are into /root/reference/src.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <utility>
#include <vector>

#include "motivate_oracle.h"

namespace {

enum { Car = 0, PublicTransport = 1, Cycle = 2, Walk = 3 };          // transport_mode.rs:3-8
enum { LocalCommute = 0, CityCommute = 1, DistantCommute = 2 };      // journey_type.rs:5-9
enum { Good = 0, Bad = 1 };                                           // weather.rs:7-10

using ModeMap = std::map<int, float>;
using RankMap = std::map<int, uint32_t>;

// hashmap_union 0.2.0 (Cargo.lock; not under /root/reference): all keys of a and b, f where both have one
template <class V, class F>
std::map<int, V> union_of(const std::map<int, V>& a, const std::map<int, V>& b, F f) {
    std::map<int, V> out;
    for (const auto& kv : a) {
        auto it = b.find(kv.first);
        out[kv.first] = it != b.end() ? f(kv.second, it->second) : kv.second;
    }
    for (const auto& kv : b)
        if (!a.count(kv.first)) out[kv.first] = kv.second;
    return out;
}
// keys in both
template <class V, class F>
std::map<int, V> intersection_of(const std::map<int, V>& a, const std::map<int, V>& b, F f) {
    std::map<int, V> out;
    for (const auto& kv : a) {
        auto it = b.find(kv.first);
        if (it != b.end()) out[kv.first] = f(kv.second, it->second);
    }
    return out;
}

// `sorted_by` of itertools 0.7.8 = collect + slice::sort_by (stable): plain insertion sort, an element moves in front
// of its predecessor only if it is strictly before it, so unordered (NaN) pairs keep their incoming order.
template <class Before>
void stable_sort_pairs(std::vector<std::pair<int, float>>& v, Before before) {
    for (size_t i = 1; i < v.size(); ++i) {
        const std::pair<int, float> x = v[i];
        size_t j = i;
        while (j > 0 && before(x, v[j - 1])) { v[j] = v[j - 1]; --j; }
        v[j] = x;
    }
}

ModeMap journey_cost(int commute) {                                   // journey_type.rs:16-35
    if (commute == LocalCommute) return {{Car, 0.2f}, {PublicTransport, 0.2f}, {Cycle, 0.2f}, {Walk, 0.2f}};
    if (commute == CityCommute) return {{Car, 0.3f}, {PublicTransport, 0.3f}, {Cycle, 0.6f}, {Walk, 0.9f}};
    return {{Car, 0.1f}, {PublicTransport, 0.1f}};
}

struct Agent;

struct Neighbourhood {                                                // neighbourhood.rs:13-31
    ModeMap supportiveness;
    std::map<int, uint32_t> capacity;
    ModeMap congestion_modifier{{Car, 1.0f}, {PublicTransport, 1.0f}, {Cycle, 1.0f}, {Walk, 1.0f}};   // :34-41
    std::vector<Agent*> residents;
    void update_congestion_modifier();
};

struct Subculture { ModeMap desirability; };                          // subculture.rs:9-14

struct Agent {                                                        // agent.rs:16-87
    Subculture* subculture;
    Neighbourhood* neighbourhood;
    int sub_id, nbhd_id, commute_length;
    float weather_sensitivity, consistency, social_connectivity, subculture_connectivity, neighbourhood_connectivity,
        average_weight;
    ModeMap habit;
    int current_mode, last_mode, norm;
    bool owns_bike, owns_car;
    std::vector<Agent*> social_network, neighbours;

    static ModeMap count_in_subgroup(const std::vector<Agent*>& agents, float weight) {      // agent.rs:322-332
        const float agents_size = (float)agents.size();
        std::map<int, std::vector<int>> groups;                       // into_group_map
        for (const Agent* x : agents) groups[x->last_mode].push_back(1);
        ModeMap out;
        for (const auto& kv : groups) out[kv.first] = (float)kv.second.size() * weight / agents_size;
        return out;
    }

    RankMap calculate_mode_budget() {                                 // agent.rs:93-177
        const ModeMap social_vals = count_in_subgroup(social_network, social_connectivity);
        const ModeMap neighbour_vals = count_in_subgroup(neighbours, neighbourhood_connectivity);
        ModeMap subculture_vals;
        for (const auto& kv : subculture->desirability) subculture_vals[kv.first] = kv.second * subculture_connectivity;
        auto add = [](float a, float b) { return a + b; };
        ModeMap acc;
        const ModeMap* norm_terms[3] = {&social_vals, &neighbour_vals, &subculture_vals};
        for (const ModeMap* x : norm_terms) acc = union_of(acc, *x, add);
        {   // max_by keeps the running maximum only while it compares Greater: ties and NaN go to the later element
            bool have = false;
            std::pair<int, float> best{0, 0.0f};
            for (const auto& kv : acc)
                if (!have || !(best.second > kv.second)) { best = kv; have = true; }
            norm = best.first;
        }
        ModeMap habit_vals;
        for (const auto& kv : habit) habit_vals[kv.first] = kv.second * consistency;
        acc.clear();
        const ModeMap* values_to_average[4] = {&social_vals, &neighbour_vals, &subculture_vals, &habit_vals};
        for (const ModeMap* x : values_to_average) acc = union_of(acc, *x, add);
        ModeMap average;
        for (const auto& kv : acc) average[kv.first] = kv.second / 4.0f;
        const ModeMap ownership_modifier{{Car, owns_car ? 1.0f : 0.0f}, {Cycle, owns_bike ? 1.0f : 0.0f}};
        auto mul = [](float a, float b) { return a * b; };
        acc.clear();
        const ModeMap* budget_terms[3] = {&average, &ownership_modifier, &neighbourhood->congestion_modifier};
        for (const ModeMap* x : budget_terms) acc = union_of(acc, *x, mul);
        std::vector<std::pair<int, float>> ordered(acc.begin(), acc.end());
        stable_sort_pairs(ordered, [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });   // descending
        RankMap ranks;
        for (size_t i = 0; i < ordered.size(); ++i) ranks[ordered[i].first] = (uint32_t)i;
        return ranks;
    }

    RankMap calculate_cost(int weather, bool change_in_weather) const {                       // agent.rs:183-240
        ModeMap neighbourhood_vals;
        for (const auto& kv : neighbourhood->supportiveness) neighbourhood_vals[kv.first] = 1.0f - kv.second;
        ModeMap average = intersection_of(journey_cost(commute_length), neighbourhood_vals,
                                          [](float a, float b) { return (a + b) / 2.0f; });
        if (weather == Bad) {
            float resolve;
            if (!change_in_weather && (last_mode == Walk || last_mode == Cycle)) resolve = -0.1f;
            else if (!change_in_weather) resolve = 0.1f;
            else resolve = 0.0f;
            const ModeMap weather_modifier{{PublicTransport, 1.0f}, {Car, 1.0f},
                                           {Cycle, 1.0f + weather_sensitivity + resolve},
                                           {Walk, 1.0f + weather_sensitivity + resolve}};
            average = intersection_of(average, weather_modifier, [](float a, float b) { return a * b; });
        }
        std::vector<std::pair<int, float>> ordered(average.begin(), average.end());
        stable_sort_pairs(ordered, [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second < b.second; });   // ascending
        RankMap ranks;
        for (size_t i = 0; i < ordered.size(); ++i) ranks[ordered[i].first] = (uint32_t)i;
        return ranks;
    }

    void update_habit() {                                             // agent.rs:244-264
        last_mode = current_mode;
        const ModeMap last_mode_map{{last_mode, average_weight}};
        ModeMap old_habit;
        for (const auto& kv : habit) old_habit[kv.first] = kv.second * (1.0f - average_weight);
        ModeMap acc;
        const ModeMap* habit_terms[2] = {&old_habit, &last_mode_map};
        for (const ModeMap* x : habit_terms) acc = union_of(acc, *x, [](float a, float b) { return a + b; });
        habit = acc;
    }

    void choose(int weather, bool change_in_weather) {               // agent.rs:269-316
        const RankMap budget = calculate_mode_budget();
        const RankMap cost = calculate_cost(weather, change_in_weather);
        RankMap total = intersection_of(budget, cost, [](uint32_t a, uint32_t b) { return a + b; });
        if (!owns_car) total.erase(Car);
        if (!owns_bike) total.erase(Cycle);
        std::map<uint32_t, std::vector<int>> groups;                  // rank sum -> modes
        for (const auto& kv : total) groups[kv.second].push_back(kv.first);
        const std::vector<int>& minimum_values = groups.begin()->second;
        if (minimum_values.size() == 1) { current_mode = minimum_values[0]; return; }
        int pick = minimum_values[0];
        for (int m : minimum_values)
            if (budget.at(m) < budget.at(pick)) pick = m;
        current_mode = pick;
    }
};

void Neighbourhood::update_congestion_modifier() {                    // neighbourhood.rs:57-89
    std::map<int, std::vector<Agent*>> groups;
    for (Agent* a : residents) groups[a->last_mode].push_back(a);
    ModeMap fresh;
    for (const auto& kv : groups) {
        const size_t count = kv.second.size(), cap = capacity.at(kv.first);
        if (count <= cap) fresh[kv.first] = 1.0f;
        else fresh[kv.first] = 1.0f - ((float)(count - cap) / (float)(residents.size() - cap));
    }
    congestion_modifier = fresh;
}

}  // namespace

struct mvo_sim {
    uint32_t N = 0, S = 0, H = 0;
    std::vector<std::unique_ptr<Agent>> agents;
    std::vector<Subculture> subcultures;
    std::vector<Neighbourhood> neighbourhoods;
    uint8_t prev_weather = 0;
};

extern "C" {

uint32_t mvo_run_rows(uint32_t n_days) {
    uint32_t r = 1;
    for (uint32_t d = 1; d < n_days; ++d) r += (d % 7 < 5);
    return r;
}

int mvo_set_tables(mvo_sim* s, const mvb_tables* t) {
    if (!s || !t || t->n_subcultures != s->S || t->n_neighbourhoods != s->H) return MVB_ERR_ARG;
    for (uint32_t h = 0; h < s->H; ++h)
        for (int m = 0; m < 4; ++m) {
            s->neighbourhoods[h].supportiveness[m] = t->supportiveness[4 * h + m];
            s->neighbourhoods[h].capacity[m] = t->capacity[4 * h + m];
        }
    for (uint32_t k = 0; k < s->S; ++k)
        for (int m = 0; m < 4; ++m) s->subcultures[k].desirability[m] = t->desirability[4 * k + m];
    return MVB_OK;
}

int mvo_set_ownership(mvo_sim* s, const uint8_t* car, const uint8_t* bike) {
    if (!s) return MVB_ERR_ARG;
    for (uint32_t i = 0; i < s->N; ++i) {
        if (car) s->agents[i]->owns_car = car[i] != 0;
        if (bike) s->agents[i]->owns_bike = bike[i] != 0;
    }
    return MVB_OK;
}

int mvo_create(const mvb_population* p, const mvb_tables* t, const mvb_params* prm, mvo_sim** out) {
    if (!p || !t || !prm || !out) return MVB_ERR_ARG;
    std::unique_ptr<mvo_sim> s(new (std::nothrow) mvo_sim());
    if (!s) return MVB_ERR_NOMEM;
    const uint32_t N = p->n_agents;
    s->N = N; s->S = t->n_subcultures; s->H = t->n_neighbourhoods;
    s->subcultures.resize(s->S);
    s->neighbourhoods.resize(s->H);
    if (mvo_set_tables(s.get(), t)) return MVB_ERR_ARG;
    for (uint32_t i = 0; i < N; ++i) {
        if (p->subculture[i] >= s->S || p->neighbourhood[i] >= s->H || p->commute[i] > 2 || p->current_mode[i] > 3 ||
            p->norm[i] > 3)
            return MVB_ERR_ARG;
        std::unique_ptr<Agent> a(new Agent());
        a->sub_id = p->subculture[i]; a->nbhd_id = p->neighbourhood[i];
        a->subculture = &s->subcultures[a->sub_id];
        a->neighbourhood = &s->neighbourhoods[a->nbhd_id];
        a->commute_length = p->commute[i];
        a->weather_sensitivity = p->weather_sensitivity[i];
        a->consistency = p->consistency ? p->consistency[i] : prm->consistency;
        a->social_connectivity = p->social_connectivity ? p->social_connectivity[i] : prm->social_connectivity;
        a->subculture_connectivity = p->subculture_connectivity ? p->subculture_connectivity[i] : prm->subculture_connectivity;
        a->neighbourhood_connectivity = p->neighbourhood_connectivity ? p->neighbourhood_connectivity[i] : prm->neighbourhood_connectivity;
        a->average_weight = p->average_weight ? p->average_weight[i] : prm->average_weight;
        for (int m = 0; m < 4; ++m)                                   // the habit map is sparse (agent_generation.rs:199)
            if (p->habit[4 * (size_t)i + m] != 0.0f) a->habit[m] = p->habit[4 * (size_t)i + m];
        a->current_mode = p->current_mode[i]; a->last_mode = a->current_mode; a->norm = p->norm[i];
        a->owns_car = p->owns_car[i] != 0; a->owns_bike = p->owns_bike[i] != 0;
        a->neighbourhood->residents.push_back(a.get());
        s->agents.push_back(std::move(a));
    }
    for (uint32_t i = 0; i < N; ++i) {                                // links WITH duplicates (social_network.rs:37-41)
        Agent& a = *s->agents[i];
        for (uint64_t x = p->social_rowptr[i]; x < p->social_rowptr[i + 1]; ++x) {
            if (p->social_col[x] >= N) return MVB_ERR_ARG;
            a.social_network.push_back(s->agents[p->social_col[x]].get());
        }
        for (uint64_t x = p->neighbour_rowptr[i]; x < p->neighbour_rowptr[i + 1]; ++x) {
            if (p->neighbour_col[x] >= N) return MVB_ERR_ARG;
            a.neighbours.push_back(s->agents[p->neighbour_col[x]].get());
        }
    }
    *out = s.release();
    return MVB_OK;
}

void mvo_destroy(mvo_sim* s) { delete s; }

int mvo_step(mvo_sim* s, uint8_t weather_today, uint8_t weather_changed) {   // simulation.rs:113-131
    if (!s || weather_today > 1) return MVB_ERR_ARG;
    for (auto& a : s->agents) a->update_habit();
    for (auto& n : s->neighbourhoods) n.update_congestion_modifier();
    for (auto& a : s->agents) a->choose(weather_today, weather_changed != 0);
    s->prev_weather = weather_today;
    return MVB_OK;
}

int mvo_stats_row(mvo_sim* s, uint32_t* row) {                        // statistics.rs:14-93, simulation.rs:254-267
    if (!s || !row) return MVB_ERR_ARG;
    const uint32_t W = MVB_STATS_FIXED + s->S + s->H;
    memset(row, 0, sizeof(uint32_t) * W);
    auto active = [](int m) { return m == Walk || m == Cycle; };
    for (const auto& a : s->agents) {
        const bool A = active(a->current_mode), AN = active(a->norm);
        row[0] += A; row[1] += AN; row[2] += A && !AN; row[3] += !A && AN;
        if (A) { row[4 + a->commute_length]++; row[7 + a->sub_id]++; row[7 + s->S + a->nbhd_id]++; }
    }
    return MVB_OK;
}

int mvo_run(mvo_sim* s, const uint8_t* weather, uint32_t n_days, const mvb_intervention* iv, uint32_t* days_out,
            uint32_t* rows_out) {                                     // simulation.rs:95-136
    if (!s || !weather || n_days == 0) return MVB_ERR_ARG;
    const uint32_t W = MVB_STATS_FIXED + s->S + s->H;
    uint32_t r = 0;
    uint8_t prev = weather[0];
    if (days_out) days_out[r] = 0;
    if (rows_out) mvo_stats_row(s, rows_out + (size_t)r * W);
    r++;
    for (uint32_t day = 1; day < n_days; ++day) {
        if (iv && day == iv->day) {
            if (iv->tables) mvo_set_tables(s, iv->tables);
            mvo_set_ownership(s, iv->owns_car, iv->owns_bike);
        }
        if (day % 7 < 5) {
            const uint8_t w = weather[day];
            mvo_step(s, w, (uint8_t)(prev != w));
            prev = w;
            if (days_out) days_out[r] = day;
            if (rows_out) mvo_stats_row(s, rows_out + (size_t)r * W);
            r++;
        }
    }
    s->prev_weather = prev;
    return MVB_OK;
}

int mvo_get_state(mvo_sim* s, float* habit, uint8_t* cur, uint8_t* last, uint8_t* norm, uint32_t* hist, float* cong) {
    if (!s) return MVB_ERR_ARG;
    if (hist) memset(hist, 0, sizeof(uint32_t) * 4 * s->H);
    for (uint32_t i = 0; i < s->N; ++i) {
        const Agent& a = *s->agents[i];
        if (habit) {
            for (int m = 0; m < 4; ++m) habit[4 * (size_t)i + m] = 0.0f;
            for (const auto& kv : a.habit) habit[4 * (size_t)i + kv.first] = kv.second;
        }
        if (cur) cur[i] = (uint8_t)a.current_mode;
        if (last) last[i] = (uint8_t)a.last_mode;
        if (norm) norm[i] = (uint8_t)a.norm;
        if (hist) hist[4 * a.nbhd_id + a.current_mode]++;
    }
    if (cong)
        for (uint32_t h = 0; h < s->H; ++h)
            for (int m = 0; m < 4; ++m) {                             // a mode nobody used yesterday has no entry: 1.0
                auto it = s->neighbourhoods[h].congestion_modifier.find(m);
                cong[4 * h + m] = it != s->neighbourhoods[h].congestion_modifier.end() ? it->second : 1.0f;
            }
    return MVB_OK;
}

}  // extern "C"

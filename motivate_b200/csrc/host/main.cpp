// main.cpp — the reference's `main` (src/main.rs:36-130) on top of the C++ host layer: same working-directory
// layout (config/parameters.yaml, config/scenario.yaml, config/networks/{id}.yaml, config/agents/{id}.yaml,
// output/output_{id}.csv) and the same `--generate` flag.  The replicas 1..=number_of_simulations, which the reference
// spreads over rayon workers, go to the GPU as one batch and are stepped together, one launch per weekday.
//   motivate_b200_run [--generate] [--dir DIR] [--device N] [--seed S]      (--dump prints the parsed configuration)
//   motivate_b200_run --weather DAYS      prints the rain sequence `Weather::make_pattern` gives (0 Good / 1 Bad)
//   motivate_b200_run --keystream N       prints N words of the zero-seed StdRng (HC-128) stream, for the known-answer test
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <vector>
#include <string>

#include "motivate_host.hpp"

using namespace motivate;

int main(int argc, char** argv) {
    bool generate = false, dump = false;
    std::string dir = ".";
    int device = 0;
    uint64_t seed = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--generate")) generate = true;
        else if (!strcmp(argv[i], "--dump")) dump = true;
        else if (!strcmp(argv[i], "--dir") && i + 1 < argc) dir = argv[++i];
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--weather") && i + 1 < argc) {
            for (uint8_t w : make_pattern(strtoull(argv[++i], nullptr, 10))) putchar('0' + w);
            putchar('\n');
            return 0;
        } else if (!strcmp(argv[i], "--keystream") && i + 1 < argc) {
            const uint32_t zero[4] = {0, 0, 0, 0};
            Hc128 rng(zero, zero);
            for (unsigned long long k = strtoull(argv[++i], nullptr, 10); k > 0; --k) printf("%08x\n", rng.next_u32());
            return 0;
        }
        else { fprintf(stderr, "usage: %s [--generate] [--dir DIR] [--device N] [--seed S]\n", argv[0]); return 2; }
    }
    try {
        const Parameters prm = Parameters::from_file(dir + "/config/parameters.yaml");        // main.rs:47-50
        if (dump) {                                    // what the YAML reader understood (tests compare it with PyYAML)
            const Scenario sc = Scenario::from_file(dir + "/config/scenario.yaml");
            printf("parameters %u %u %u %a %a %a %u %u %u %zu\n", prm.total_years, prm.number_of_people, prm.number_of_simulations,
                   prm.social_connectivity, prm.subculture_connectivity, prm.neighbourhood_connectivity,
                   prm.number_of_social_network_links, prm.number_of_neighbour_links, prm.days_in_habit_average, prm.distributions.size());
            for (auto& d : prm.distributions) printf("distribution %a %a %a\n", d[0], d[1], d[2]);
            printf("scenario %s %u %u\n", sc.id.c_str(), sc.number_of_bikes, sc.number_of_cars);
            for (size_t i = 0; i < sc.subculture_ids.size(); ++i)
                printf("subculture|%s|%a %a %a %a\n", sc.subculture_ids[i].c_str(), sc.desirability[4 * i], sc.desirability[4 * i + 1],
                       sc.desirability[4 * i + 2], sc.desirability[4 * i + 3]);
            for (size_t i = 0; i < sc.neighbourhood_ids.size(); ++i)
                printf("neighbourhood|%s|%a %a %a %a|%u %u %u %u\n", sc.neighbourhood_ids[i].c_str(), sc.supportiveness[4 * i],
                       sc.supportiveness[4 * i + 1], sc.supportiveness[4 * i + 2], sc.supportiveness[4 * i + 3], sc.capacity[4 * i],
                       sc.capacity[4 * i + 1], sc.capacity[4 * i + 2], sc.capacity[4 * i + 3]);
            const Intervention& iv = sc.intervention;
            printf("intervention %u %d %d %zu %zu\n", iv.day, iv.change_in_number_of_bikes, iv.change_in_number_of_cars,
                   iv.neighbourhood_changes.size(), iv.subculture_changes.size());
            for (auto& c : iv.neighbourhood_changes) {
                printf("nchange|%s|", c.id.c_str());
                for (auto& kv : c.increase_in_supportiveness) printf("s%d=%a ", kv.first, kv.second);
                for (auto& kv : c.increase_in_capacity) printf("c%d=%lld ", kv.first, kv.second);
                printf("\n");
            }
            for (auto& c : iv.subculture_changes) {
                printf("schange|%s|", c.id.c_str());
                for (auto& kv : c.increase_in_desirability) printf("d%d=%a ", kv.first, kv.second);
                printf("\n");
            }
            return 0;
        }
        if (generate) {                                                                        // main.rs:58-71, 154-182
            ::mkdir((dir + "/config/networks").c_str(), 0777);
            ::mkdir((dir + "/config/agents").c_str(), 0777);
            for (uint32_t id = 1; id <= prm.number_of_simulations; ++id) {
                uint64_t* rp = nullptr; uint32_t* cl = nullptr;
                if (mvb_synth_ba_network(prm.number_of_social_network_links, prm.number_of_people, seed * 7919 + id, &rp, &cl))
                    throw std::runtime_error(mvb_last_error());
                std::vector<uint64_t> rowptr(rp, rp + prm.number_of_people + 1);
                std::vector<uint32_t> col(cl, cl + rowptr.back());
                mvb_free(rp); mvb_free(cl);
                write_network(dir + "/config/networks/" + std::to_string(id) + ".yaml", rowptr, col);
            }
            fprintf(stderr, "Generating networks complete\n");
        }
        const std::vector<uint8_t> weather = make_pattern((size_t)365 * prm.total_years);   // main.rs:73-85: the reference's own rain sequence
        std::vector<Prepared> reps;
        for (uint32_t id = 1; id <= prm.number_of_simulations; ++id) {                          // main.rs:89-121
            const std::string n = std::to_string(id);
            reps.push_back(simulation::prepare(n, generate, dir + "/config/agents/" + n + ".yaml", dir + "/config/scenario.yaml",
                                               prm.number_of_people, prm.social_connectivity, prm.subculture_connectivity,
                                               prm.neighbourhood_connectivity, prm.number_of_neighbour_links,
                                               prm.days_in_habit_average, prm.distributions,
                                               dir + "/config/networks/" + n + ".yaml", seed * 1000 + id));
            fprintf(stderr, "[%s] Agents created\n", n.c_str());
        }
        simulation::run_replicas(reps, prm.total_years, weather, dir + "/output", device);
        fprintf(stderr, "TOTAL RUNNING TIME: done, %zu replicas\n", reps.size());
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());                                               // the reference panics here
        return 1;
    }
    return 0;
}

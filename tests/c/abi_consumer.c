/* abi_consumer.c — a plain C99 program over include/motivate_b200.h: what a cgo / bindgen / JNI consumer relies on.
 * It must compile as C (no C++ in the header), link against libmotivate_b200.so and run WITHOUT a GPU: every call here
 * is host-only (synthetic population, partition plan, argument checks, "no CPU path" error).
 * Prints "ok <n_agents> <links> <rows>" and exits 0, or prints what failed and exits 1. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "motivate_b200.h"

#define CHECK(cond, what) do { if (!(cond)) { printf("FAILED: %s (%s)\n", what, mvb_last_error()); return 1; } } while (0)

int main(void) {
    CHECK(mvb_abi_version() == 1, "abi version");
    CHECK(mvb_run_rows(730) == 522, "2 years = day 0 + 521 weekdays");          /* src/simulation.rs:98-136 */
    CHECK(sizeof(mvb_params) == 24, "mvb_params layout");

    mvb_synth_spec spec;
    memset(&spec, 0, sizeof spec);
    spec.n_agents = 5000; spec.n_subcultures = 3; spec.n_neighbourhoods = 4;
    spec.social_links_min = 5; spec.neighbour_links_min = 10; spec.seed = 42; spec.n_cars = 2920; spec.n_bikes = 2920;
    spec.n_gmm = 1; spec.gmm[0][0] = 9845.09; spec.gmm[0][1] = 3691.81; spec.gmm[0][2] = 1.0;
    mvb_synth* syn = NULL;
    CHECK(mvb_synth_create(&spec, &syn) == MVB_OK && syn, "synthetic population");
    const mvb_population* pop = mvb_synth_population(syn);
    CHECK(pop && pop->n_agents == 5000, "population view");
    const unsigned long long links = pop->social_rowptr[5000] + pop->neighbour_rowptr[5000];
    CHECK(links > 100000ull, "links generated");

    uint32_t* owner = (uint32_t*)malloc(sizeof(uint32_t) * 5000);
    CHECK(mvb_partition_plan(pop, 3, 4, 2, owner) == MVB_OK, "partition plan");
    unsigned n1 = 0;
    for (int i = 0; i < 5000; ++i) { CHECK(owner[i] < 2, "owner rank"); n1 += owner[i]; }
    CHECK(n1 > 500 && n1 < 4500, "both ranks own agents");

    /* argument checks come back as codes, never as a crash */
    mvb_sim* sim = (mvb_sim*)1;
    CHECK(mvb_create(NULL, NULL, NULL, 0, &sim) == MVB_ERR_ARG && sim == NULL, "null arguments rejected");
    CHECK(strlen(mvb_last_error()) > 0, "error text");
    CHECK(mvb_step(NULL, 0, 0) == MVB_ERR_ARG, "null handle rejected");
    mvb_destroy(NULL);                                                            /* no-op */
    if (mvb_device_count() == 0) {                                                /* the product has no CPU path */
        float supp[16] = {0}, des[12] = {0};
        uint32_t cap[16] = {0};
        mvb_tables tab = {3, 4, supp, cap, des};
        mvb_params prm = {1.0f, 0.7f, 0.5f, 0.3f, 2.0f / 11.0f, 0};
        CHECK(mvb_create(pop, &tab, &prm, 0, &sim) == MVB_ERR_CUDA && sim == NULL, "create without a GPU fails loudly");
    }
    free(owner);
    mvb_synth_destroy(syn);
    printf("ok %u %llu %u\n", pop ? 5000u : 0u, links, mvb_run_rows(730));
    return 0;
}

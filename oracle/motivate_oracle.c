/*
 * motivate_oracle.c — CPU restatement of Motivate's per-day agent update.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared
 * against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product library never does.
 *
 * PARITY UNPINNED by the reference: 0xr0bert/Motivate ships no tests, no golden
 * vectors, and cannot be built in this image (no cargo/rustc); its shipped
 * output/ CSVs come from unreproducible thread_rng populations.
 * What pins this file instead: (1) hand-computed micro-cases in
 * tests/test_oracle_micro.py, (2) an independent literal model that keeps the
 * reference's HashMap / union_of structure (oracle/literal_model.py) and the
 * golden vectors generated from it (tests/golden/), (3) a compiled restatement
 * in the reference's own shape (oracle/reference_shaped.cpp) compared with this
 * file on 5k random and 100k BASELINE-shaped agents (tests/test_reference_shaped.py).
 *
 * Dense restatement rules (each proven equivalent to the sparse HashMap code in
 * SURVEY.md §8a):  an absent map key == +0.0f for additive folds and == 1.0f for
 * the congestion product; all arithmetic is f32, evaluated in the reference's
 * order, with FMA contraction disabled (-ffp-contract=off).
 *
 * Canonical tie-break (the reference iterates std HashMaps with RandomState, so
 * its own order is non-deterministic; SURVEY.md §8 a-T):  iteration order is the
 * enum order Car, PublicTransport, Cycle, Walk.  Stable sorts give the lower
 * enum index the better rank among equal values; `max_by` returns the LAST
 * maximal element, i.e. the higher enum index among equal maxima.
 *
 * Third-party semantics restated (not under /root/reference):
 *   hashmap_union 0.2.0 (Cargo.toml:17): union_of(a,b,f) = keys of a∪b, f(va,vb)
 *     where both present else the present value; intersection_of = keys in both.
 *   itertools 0.7.8: sorted_by / sorted_by_key = collect + stable sort_by;
 *     into_group_map = HashMap<K, Vec<V>>.
 *   std Iterator::max_by: last maximum wins.
 */
#include "motivate_oracle.h"

#include <stdlib.h>
#include <string.h>

struct mvo_sim {
    uint32_t N, S, H;
    uint8_t  *sub, *commute, *car, *bike;
    uint16_t *nbhd;
    float    *ws, *consistency, *soc_c, *sub_c, *nei_c, *avgw;
    float    *habit;                 /* [N][4] */
    uint8_t  *cur, *last, *norm;
    uint64_t *s_rowptr, *n_rowptr;
    uint32_t *s_col, *n_col;
    float    *supp, *desir;          /* [H][4], [S][4] */
    uint32_t *cap;                   /* [H][4] */
    uint32_t *nres;                  /* [H] residents per neighbourhood */
    float    *cong;                  /* [H][4] modifier used by the last step */
    uint32_t *hist;                  /* [H][4] scratch */
    uint8_t   prev_weather;
};

static void* dup_mem(const void* src, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (p && src && bytes) memcpy(p, src, bytes);
    return p;
}

static float* dup_or_fill(const float* src, float v, uint32_t n) {
    float* p = (float*)malloc(sizeof(float) * (n ? n : 1));
    if (!p) return NULL;
    if (src) memcpy(p, src, sizeof(float) * n);
    else for (uint32_t i = 0; i < n; ++i) p[i] = v;
    return p;
}

int mvo_set_tables(mvo_sim* s, const mvb_tables* t) {
    if (!s || !t || t->n_subcultures != s->S || t->n_neighbourhoods != s->H) return MVB_ERR_ARG;
    memcpy(s->supp,  t->supportiveness, sizeof(float) * 4 * s->H);
    memcpy(s->cap,   t->capacity,       sizeof(uint32_t) * 4 * s->H);
    memcpy(s->desir, t->desirability,   sizeof(float) * 4 * s->S);
    return MVB_OK;
}

int mvo_set_ownership(mvo_sim* s, const uint8_t* car, const uint8_t* bike) {
    if (!s) return MVB_ERR_ARG;
    if (car)  memcpy(s->car,  car,  s->N);
    if (bike) memcpy(s->bike, bike, s->N);
    return MVB_OK;
}

int mvo_create(const mvb_population* p, const mvb_tables* t, const mvb_params* prm, mvo_sim** out) {
    if (!p || !t || !prm || !out) return MVB_ERR_ARG;
    mvo_sim* s = (mvo_sim*)calloc(1, sizeof(mvo_sim));
    if (!s) return MVB_ERR_NOMEM;
    uint32_t N = p->n_agents;
    s->N = N; s->S = t->n_subcultures; s->H = t->n_neighbourhoods;
    for (uint32_t i = 0; i < N; ++i) {
        if (p->subculture[i] >= s->S || p->neighbourhood[i] >= s->H || p->commute[i] > 2 ||
            p->current_mode[i] > 3 || p->norm[i] > 3) { free(s); return MVB_ERR_ARG; }
    }
    s->sub     = (uint8_t*)dup_mem(p->subculture, N);
    s->nbhd    = (uint16_t*)dup_mem(p->neighbourhood, 2 * (size_t)N);
    s->commute = (uint8_t*)dup_mem(p->commute, N);
    s->car     = (uint8_t*)dup_mem(p->owns_car, N);
    s->bike    = (uint8_t*)dup_mem(p->owns_bike, N);
    s->ws      = (float*)dup_mem(p->weather_sensitivity, 4 * (size_t)N);
    s->consistency = dup_or_fill(p->consistency, prm->consistency, N);
    s->soc_c   = dup_or_fill(p->social_connectivity, prm->social_connectivity, N);
    s->sub_c   = dup_or_fill(p->subculture_connectivity, prm->subculture_connectivity, N);
    s->nei_c   = dup_or_fill(p->neighbourhood_connectivity, prm->neighbourhood_connectivity, N);
    s->avgw    = dup_or_fill(p->average_weight, prm->average_weight, N);
    s->habit   = (float*)dup_mem(p->habit, 16 * (size_t)N);
    s->cur     = (uint8_t*)dup_mem(p->current_mode, N);
    s->last    = (uint8_t*)dup_mem(p->current_mode, N);
    s->norm    = (uint8_t*)dup_mem(p->norm, N);
    s->s_rowptr = (uint64_t*)dup_mem(p->social_rowptr, 8 * ((size_t)N + 1));
    s->n_rowptr = (uint64_t*)dup_mem(p->neighbour_rowptr, 8 * ((size_t)N + 1));
    s->s_col   = (uint32_t*)dup_mem(p->social_col, 4 * (size_t)p->social_rowptr[N]);
    s->n_col   = (uint32_t*)dup_mem(p->neighbour_col, 4 * (size_t)p->neighbour_rowptr[N]);
    s->supp    = (float*)malloc(sizeof(float) * 4 * s->H);
    s->cap     = (uint32_t*)malloc(sizeof(uint32_t) * 4 * s->H);
    s->desir   = (float*)malloc(sizeof(float) * 4 * s->S);
    s->nres    = (uint32_t*)calloc(s->H, sizeof(uint32_t));
    s->cong    = (float*)malloc(sizeof(float) * 4 * s->H);
    s->hist    = (uint32_t*)calloc(4 * (size_t)s->H, sizeof(uint32_t));
    mvo_set_tables(s, t);
    for (uint32_t i = 0; i < N; ++i) s->nres[s->nbhd[i]]++;
    /* default_congestion_modifier: 1.0 everywhere (src/neighbourhood.rs:34-41) */
    for (uint32_t k = 0; k < 4 * s->H; ++k) s->cong[k] = 1.0f;
    s->prev_weather = 0;
    *out = s;
    return MVB_OK;
}

void mvo_destroy(mvo_sim* s) {
    if (!s) return;
    free(s->sub); free(s->nbhd); free(s->commute); free(s->car); free(s->bike); free(s->ws);
    free(s->consistency); free(s->soc_c); free(s->sub_c); free(s->nei_c); free(s->avgw);
    free(s->habit); free(s->cur); free(s->last); free(s->norm);
    free(s->s_rowptr); free(s->n_rowptr); free(s->s_col); free(s->n_col);
    free(s->supp); free(s->cap); free(s->desir); free(s->nres); free(s->cong); free(s->hist);
    free(s);
}

/* JourneyType::cost, src/journey_type.rs:16-35.  Returns the key-set mask. */
static unsigned commute_cost(uint8_t commute, float cc[4]) {
    switch (commute) {
    case 0: cc[0] = 0.2f; cc[1] = 0.2f; cc[2] = 0.2f; cc[3] = 0.2f; return 0xFu;
    case 1: cc[0] = 0.3f; cc[1] = 0.3f; cc[2] = 0.6f; cc[3] = 0.9f; return 0xFu;
    default: cc[0] = 0.1f; cc[1] = 0.1f; cc[2] = 0.0f; cc[3] = 0.0f; return 0x3u;  /* Car, PT only */
    }
}

/* count_in_subgroup, src/agent.rs:322-332: v[m] = (cnt as f32) * weight / (len as f32)
 * for modes with cnt>0; absent (== 0.0f here) otherwise; empty list -> all absent. */
static void subgroup_share(const mvo_sim* s, const uint64_t* rowptr, const uint32_t* col,
                           uint32_t i, float weight, float v[4]) {
    uint64_t b = rowptr[i], e = rowptr[i + 1];
    size_t cnt[4] = {0, 0, 0, 0};
    for (uint64_t k = b; k < e; ++k) cnt[s->last[col[k]]]++;
    float len = (float)(size_t)(e - b);
    for (int m = 0; m < 4; ++m)
        v[m] = cnt[m] ? ((float)cnt[m] * weight) / len : 0.0f;
}

/* Agent::update_habit, src/agent.rs:244-264. */
static void update_habit(mvo_sim* s, uint32_t i) {
    s->last[i] = s->cur[i];
    float w = s->avgw[i];
    float keep = 1.0f - w;
    float* h = s->habit + 4 * (size_t)i;
    for (int m = 0; m < 4; ++m) h[m] = h[m] * keep;
    h[s->last[i]] = h[s->last[i]] + w;
}

/* Neighbourhood::update_congestion_modifier, src/neighbourhood.rs:57-89. */
static void update_congestion(mvo_sim* s) {
    memset(s->hist, 0, sizeof(uint32_t) * 4 * s->H);
    for (uint32_t i = 0; i < s->N; ++i) s->hist[4 * s->nbhd[i] + s->last[i]]++;
    for (uint32_t h = 0; h < s->H; ++h)
        for (int m = 0; m < 4; ++m) {
            size_t count = s->hist[4 * h + m], cap = s->cap[4 * h + m];
            float c = 1.0f;                       /* key absent (count==0) or count<=capacity */
            if (count > cap) {
                size_t max_excess = (size_t)s->nres[h] - cap;
                size_t act_excess = count - cap;
                c = 1.0f - ((float)act_excess / (float)max_excess);
            }
            s->cong[4 * h + m] = c;
        }
}

/* Agent::choose, src/agent.rs:269-316 with calculate_mode_budget (93-177) and
 * calculate_cost (183-240). */
static void choose(mvo_sim* s, uint32_t i, uint8_t weather, uint8_t changed) {
    float soc[4], nei[4], sub[4], t[4], avg[4], b[4], c[4], cc[4];
    subgroup_share(s, s->s_rowptr, s->s_col, i, s->soc_c[i], soc);
    subgroup_share(s, s->n_rowptr, s->n_col, i, s->nei_c[i], nei);
    const float* des = s->desir + 4 * s->sub[i];
    const float* h   = s->habit + 4 * (size_t)i;
    const float* cg  = s->cong + 4 * s->nbhd[i];
    const float* sp  = s->supp + 4 * s->nbhd[i];
    for (int m = 0; m < 4; ++m) sub[m] = des[m] * s->sub_c[i];

    /* norm: fold order ((∅∪soc)∪nei)∪sub, max_by -> last maximum (agent.rs:116-124) */
    int norm = 0;
    for (int m = 0; m < 4; ++m) {
        t[m] = (soc[m] + nei[m]) + sub[m];
        if (m == 0 || !(t[norm] > t[m])) norm = m;
    }
    s->norm[i] = (uint8_t)norm;

    /* average of the four components (agent.rs:127-142), ownership and congestion (150-164) */
    for (int m = 0; m < 4; ++m) {
        float hv = h[m] * s->consistency[i];
        avg[m] = (t[m] + hv) / 4.0f;
    }
    b[0] = (avg[0] * (s->car[i]  ? 1.0f : 0.0f)) * cg[0];
    b[1] = avg[1] * cg[1];
    b[2] = (avg[2] * (s->bike[i] ? 1.0f : 0.0f)) * cg[2];
    b[3] = avg[3] * cg[3];

    /* cost (agent.rs:183-221) */
    unsigned K = commute_cost(s->commute[i], cc);
    for (int m = 0; m < 4; ++m) c[m] = (cc[m] + (1.0f - sp[m])) / 2.0f;
    if (weather == 1) {
        float resolve;
        if (!changed && (s->last[i] == 3 || s->last[i] == 2)) resolve = -0.1f;
        else if (!changed) resolve = 0.1f;
        else resolve = 0.0f;
        float mod = (1.0f + s->ws[i]) + resolve;
        c[0] = c[0] * 1.0f; c[1] = c[1] * 1.0f;
        c[2] = c[2] * mod;  c[3] = c[3] * mod;
    }

    /* ranks: budget descending stable over all 4 (agent.rs:168-176), cost ascending stable over K
     * (agent.rs:224-232).  A stable sort over the canonical order puts m before o (o < m) only if m is strictly
     * better; partial_cmp(..).unwrap_or(Equal) makes an unordered (NaN) pair a tie, so o stays first. */
    unsigned brank[4] = {0, 0, 0, 0}, crank[4] = {0, 0, 0, 0};
    for (int o = 0; o < 3; ++o)
        for (int m = o + 1; m < 4; ++m) {
            if (b[m] > b[o]) brank[o]++; else brank[m]++;
            if (((K >> o) & 1u) && ((K >> m) & 1u)) {
                if (c[m] < c[o]) crank[o]++; else crank[m]++;
            }
        }

    /* allowed set and pick (agent.rs:274-315): min sum, tie -> min budget rank */
    unsigned A = K;
    if (!s->car[i])  A &= ~1u;
    if (!s->bike[i]) A &= ~4u;
    unsigned best_key = 0xFFFFFFFFu; int best = -1;
    for (int m = 0; m < 4; ++m) {
        if (!((A >> m) & 1u)) continue;
        unsigned key = 4u * (brank[m] + crank[m]) + brank[m];
        if (key < best_key) { best_key = key; best = m; }
    }
    s->cur[i] = (uint8_t)best;      /* A always contains PublicTransport */
}

int mvo_step(mvo_sim* s, uint8_t weather_today, uint8_t weather_changed) {
    if (!s) return MVB_ERR_ARG;
    for (uint32_t i = 0; i < s->N; ++i) update_habit(s, i);                /* simulation.rs:116-118 */
    update_congestion(s);                                                  /* simulation.rs:121-123 */
    for (uint32_t i = 0; i < s->N; ++i) choose(s, i, weather_today, weather_changed); /* :126-128 */
    return MVB_OK;
}

/* statistics.rs:14-93 in the column order of simulation.rs:254-267. */
int mvo_stats_row(mvo_sim* s, uint32_t* row) {
    if (!s || !row) return MVB_ERR_ARG;
    uint32_t W = MVB_STATS_FIXED + s->S + s->H;
    memset(row, 0, sizeof(uint32_t) * W);
    for (uint32_t i = 0; i < s->N; ++i) {
        int a  = s->cur[i] >= 2;
        int an = s->norm[i] >= 2;
        row[0] += a; row[1] += an; row[2] += (a && !an); row[3] += (!a && an);
        if (a) { row[4 + s->commute[i]]++; row[7 + s->sub[i]]++; row[7 + s->S + s->nbhd[i]]++; }
    }
    return MVB_OK;
}

uint32_t mvo_run_rows(uint32_t n_days) {
    uint32_t r = 1;
    for (uint32_t d = 1; d < n_days; ++d) r += (d % 7 < 5);
    return r;
}

/* simulation.rs:95-136. */
int mvo_run(mvo_sim* s, const uint8_t* weather, uint32_t n_days, const mvb_intervention* iv,
            uint32_t* days_out, uint32_t* rows_out) {
    if (!s || !weather || n_days == 0) return MVB_ERR_ARG;
    uint32_t W = MVB_STATS_FIXED + s->S + s->H, r = 0;
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
            uint8_t w = weather[day];
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

int mvo_get_state(mvo_sim* s, float* habit, uint8_t* cur, uint8_t* last, uint8_t* norm,
                  uint32_t* hist, float* cong) {
    if (!s) return MVB_ERR_ARG;
    if (habit) memcpy(habit, s->habit, 16 * (size_t)s->N);
    if (cur)   memcpy(cur, s->cur, s->N);
    if (last)  memcpy(last, s->last, s->N);
    if (norm)  memcpy(norm, s->norm, s->N);
    if (hist) {
        memset(hist, 0, sizeof(uint32_t) * 4 * s->H);
        for (uint32_t i = 0; i < s->N; ++i) hist[4 * s->nbhd[i] + s->cur[i]]++;
    }
    if (cong)  memcpy(cong, s->cong, sizeof(float) * 4 * s->H);
    return MVB_OK;
}

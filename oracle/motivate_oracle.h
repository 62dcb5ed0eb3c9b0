/*
 * motivate_oracle.h — interface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see
 * motivate_oracle.c).  One replica, same structs and codes as the product ABI
 * (include/motivate_b200.h) so tests drive both through identical harness code.
 */
#ifndef MOTIVATE_ORACLE_H
#define MOTIVATE_ORACLE_H
#include "../include/motivate_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mvo_sim mvo_sim;
int  mvo_create(const mvb_population* pop, const mvb_tables* tab, const mvb_params* prm, mvo_sim** out);
void mvo_destroy(mvo_sim* s);
int  mvo_step(mvo_sim* s, uint8_t weather_today, uint8_t weather_changed);
int  mvo_run(mvo_sim* s, const uint8_t* weather, uint32_t n_days, const mvb_intervention* iv,
             uint32_t* days_out, uint32_t* rows_out);
uint32_t mvo_run_rows(uint32_t n_days);
int  mvo_set_tables(mvo_sim* s, const mvb_tables* tab);
int  mvo_set_ownership(mvo_sim* s, const uint8_t* owns_car, const uint8_t* owns_bike);
int  mvo_stats_row(mvo_sim* s, uint32_t* row_out);
int  mvo_get_state(mvo_sim* s, float* habit, uint8_t* current_mode, uint8_t* last_mode, uint8_t* norm,
                   uint32_t* hist, float* congestion);
#ifdef __cplusplus
}
#endif
#endif

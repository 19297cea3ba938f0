/*
 * oracle/crowd_oracle.h -- TEST INFRASTRUCTURE (CPU oracle). Not product code.
 * C ABI of oracle/_build/libcrowd_oracle.so; bound by oracle/crowd_oracle.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 */
#ifndef ORACLE_CROWD_ORACLE_H
#define ORACLE_CROWD_ORACLE_H

#include "../include/crowdnav_b200.h"

/* Host-memory mirror of CnStateView / CnObsOut / CnStepOut (same field meaning). */
typedef struct OrStateView {
    float *robot;      /* [N,9] */
    float *humans;     /* [N,H,9] */
    float *belief;     /* [N,H,5] */
    float *extras;     /* [N,4] */
    int32_t *counters; /* [N,4] */
    float *episode_return; /* [N] */
    float *groups;     /* [N,CN_MAX_GROUPS,4] radius, cx, cy, valid (group environment) */
} OrStateView;

int oracle_step(const CnConfig *cfg, int n_envs, const OrStateView *st, const float *action,
                const CnStepOut *out, int auto_reset, int n_threads);
int oracle_reset(const CnConfig *cfg, int n_envs, const OrStateView *st, const uint8_t *mask,
                 const CnObsOut *obs, int n_threads);
int oracle_observe(const CnConfig *cfg, int n_envs, const OrStateView *st, const CnObsOut *obs, int n_threads);
/* the counter-based RNG contract shared with the CUDA kernels (exposed for unit tests) */
void oracle_philox(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
/* one ORCA solve as crowd_nav/policy/orca.py:64-139 sets it up (exposed for known-answer tests) */
int oracle_orca_one(const CnConfig *cfg, const float *self9, int n_others, const float *others5, float *out_v);

#endif

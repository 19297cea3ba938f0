/*
 * oracle/rvo2_port.c -- TEST INFRASTRUCTURE (CPU oracle). Not product code.
 *
 * C-ABI simulator object with the subset of rvo2.PyRVOSimulator that the
 * reference uses at crowd_nav/policy/orca.py:87-136 (constructor, addAgent,
 * getNumAgents, setAgentPosition/Velocity/PrefVelocity, doStep,
 * getAgentVelocity).  oracle/refshim/rvo2.py binds it with ctypes so the
 * reference's own Python can be executed in this container.
 *
 * The arithmetic is oracle/orca_core.h (restated RVO2; PARITY UNPINNED against
 * the real, un-vendored and un-pinned rvo2 wheel).
 */
#include <stdlib.h>
#include <string.h>
#include "orca_core.h"

typedef struct {
    orc_v2 pos, vel, pref, new_vel;
    float radius, max_speed, neighbor_dist, time_horizon;
    size_t max_neighbors;
} rvo_agent;

typedef struct {
    float time_step;
    float global_time;
    size_t n, cap;
    rvo_agent *agents;
} rvo_sim;

rvo_sim *rvo_create(float time_step)
{
    rvo_sim *s = (rvo_sim *)calloc(1, sizeof(rvo_sim));
    s->time_step = time_step;
    return s;
}

void rvo_destroy(rvo_sim *s)
{
    if (!s) return;
    free(s->agents);
    free(s);
}

long rvo_add_agent(rvo_sim *s, float px, float py, float neighbor_dist, size_t max_neighbors,
                   float time_horizon, float radius, float max_speed, float vx, float vy)
{
    if (s->n == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 8;
        s->agents = (rvo_agent *)realloc(s->agents, s->cap * sizeof(rvo_agent));
    }
    rvo_agent *a = &s->agents[s->n];
    memset(a, 0, sizeof(*a));
    a->pos = orc_mk(px, py);
    a->vel = orc_mk(vx, vy);
    a->radius = radius;
    a->max_speed = max_speed;
    a->neighbor_dist = neighbor_dist;
    a->time_horizon = time_horizon;
    a->max_neighbors = max_neighbors;
    return (long)(s->n++);
}

size_t rvo_num_agents(const rvo_sim *s) { return s->n; }
void rvo_set_position(rvo_sim *s, size_t i, float x, float y) { s->agents[i].pos = orc_mk(x, y); }
void rvo_set_velocity(rvo_sim *s, size_t i, float x, float y) { s->agents[i].vel = orc_mk(x, y); }
void rvo_set_pref_velocity(rvo_sim *s, size_t i, float x, float y) { s->agents[i].pref = orc_mk(x, y); }
void rvo_get_velocity(const rvo_sim *s, size_t i, float *out) { out[0] = s->agents[i].vel.x; out[1] = s->agents[i].vel.y; }
void rvo_get_position(const rvo_sim *s, size_t i, float *out) { out[0] = s->agents[i].pos.x; out[1] = s->agents[i].pos.y; }

/* RVOSimulator::doStep(): every agent solves, then every agent is updated. */
int rvo_do_step(rvo_sim *s)
{
    if (s->n > ORC_MAX_LINES) return -1;
    orc_v2 o_pos[ORC_MAX_LINES], o_vel[ORC_MAX_LINES];
    float o_rad[ORC_MAX_LINES];
    for (size_t i = 0; i < s->n; ++i) {
        size_t m = 0;
        for (size_t k = 0; k < s->n; ++k) {
            if (k == i) continue;
            o_pos[m] = s->agents[k].pos;
            o_vel[m] = s->agents[k].vel;
            o_rad[m] = s->agents[k].radius;
            ++m;
        }
        rvo_agent *a = &s->agents[i];
        orc_new_velocity(a->pos, a->vel, a->radius, a->max_speed, a->pref, m, o_pos, o_vel, o_rad,
                         a->neighbor_dist, a->max_neighbors, a->time_horizon, s->time_step, &a->new_vel, NULL);
    }
    for (size_t i = 0; i < s->n; ++i) {
        rvo_agent *a = &s->agents[i];
        a->vel = a->new_vel;
        a->pos = orc_add(a->pos, orc_scale(s->time_step, a->vel));
    }
    s->global_time += s->time_step;
    return 0;
}

/*
 * One-shot entry used by unit tests and by the golden-vector generator:
 * solve agent 0 against `n` others exactly as ORCA.predict sets the sim up.
 * out[0..1] = new velocity, returns lp2's fail index (== n_lines when feasible),
 * *n_lines_out = number of neighbours within range.
 */
long rvo_solve_one(float px, float py, float vx, float vy, float radius, float max_speed,
                   float pref_x, float pref_y, size_t n, const float *o_px, const float *o_py,
                   const float *o_vx, const float *o_vy, const float *o_radius,
                   float neighbor_dist, size_t max_neighbors, float time_horizon, float time_step,
                   float *out, long *n_lines_out)
{
    if (n > ORC_MAX_LINES) return -1;
    orc_v2 o_pos[ORC_MAX_LINES], o_vel[ORC_MAX_LINES];
    for (size_t k = 0; k < n; ++k) { o_pos[k] = orc_mk(o_px[k], o_py[k]); o_vel[k] = orc_mk(o_vx[k], o_vy[k]); }
    orc_v2 nv;
    size_t fail = 0;
    size_t nl = orc_new_velocity(orc_mk(px, py), orc_mk(vx, vy), radius, max_speed, orc_mk(pref_x, pref_y),
                                 n, o_pos, o_vel, o_radius, neighbor_dist, max_neighbors, time_horizon, time_step,
                                 &nv, &fail);
    out[0] = nv.x; out[1] = nv.y;
    if (n_lines_out) *n_lines_out = (long)nl;
    return (long)fail;
}

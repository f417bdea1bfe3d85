/* CPU oracle for TrafficManagementEnv -- a plain-C restatement of the reference algorithm.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Follows /root/reference/traffic_management_env/:
 *   config.py:6-35                      constants
 *   utils.py:71-118   TrafficLight      update / _advance_phase / can_pass / set_phase
 *   utils.py:121-171  Intersection      four FIFO queues, process_vehicles, queue lengths
 *   utils.py:174-248  routes            generate_vehicle_route, neighbours (N,S,W,E on the 5x5 grid), direction
 *   utils.py:251-267  calculate_traffic_metrics
 *   environment.py:141-166 reset, :168-203 step, :205-220 _apply_actions, :222-249 _spawn_vehicles,
 *   :251-269 _update_vehicles, :271-285 _process_intersections / _remove_completed_vehicles,
 *   :287-311 _calculate_reward, :313-363 _get_observation
 * and SURVEY.md section 0 facts 5 and 9: the time limit is `terminated`; vehicles never move (they spawn exactly
 * at an intersection, so _update_vehicles is a no-op), wait in their start intersection's queue, and once popped
 * either vanish (route ends where it started) or stay in `self.vehicles` forever ("zombies").
 *
 * The oracle keeps the reference's OBJECT shape on purpose -- vehicle records, FIFO queues of vehicle indices, a
 * `vehicles` list filtered by destination -- and NOT the (count, wait-sum, loop-back count) reduction the CUDA
 * kernel uses, so that the reduction itself is what the parity tests check.
 *
 * Reward and observation are float64 expressions of integers evaluated in the reference's order (np.var uses
 * NumPy's pairwise sum); the observation is cast to float32 at the end (environment.py:363).
 *
 * RNG draw sites, in order: one randint(5,30) per light that enters a green phase (lights
 * in id order), then -- unless the vehicle list is full -- random(), and on a spawn randint(0,NI-1), randint(2,5)
 * and one choice(neighbours) per hop.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beng_oracle_rng.h"

#define MAX_NI 25
#define MAX_VEH 4096

enum { NS_GREEN = 0, NS_YELLOW = 1, EW_GREEN = 2, EW_YELLOW = 3 };
enum { NORTH = 0, EAST = 1, SOUTH = 2, WEST = 3 };

typedef struct {
    int direction;
    int destination; /* -1 == None (completed) */
    int waiting_time;
    int alive; /* still in self.vehicles */
} vehicle;

typedef struct {
    int phase, timer, duration;
    int queue[4][MAX_VEH]; /* vehicle indices, FIFO */
    int qlen[4];
    int vehicles_passed, total_waiting_time;
} intersection;

typedef struct {
    intersection *x; /* [ni] */
    vehicle *veh;    /* every vehicle ever created this episode */
    int n_created;
    int n_listed; /* len(self.vehicles) */
    int timestep;
    double total_reward;
    int needs_reset;
    orc_stream rng;
} traffic_env;

typedef struct {
    int n_envs, mode, rows, cols, ni, max_vehicles, max_timesteps;
    double spawn_rate;
    traffic_env *envs;
    double stats[3]; /* n_episodes, sum_return, sum_length */
} traffic_oracle;

static int obs_dim(const traffic_oracle *o) { return o->ni * 14 + 4; }

static void env_reset(traffic_oracle *o, traffic_env *e) { /* environment.py:141-166 */
    e->timestep = 0; e->total_reward = 0.0; e->n_created = 0; e->n_listed = 0; e->needs_reset = 0;
    for (int i = 0; i < o->ni; ++i) {
        intersection *x = &e->x[i];
        memset(x->qlen, 0, sizeof(x->qlen));
        x->vehicles_passed = 0; x->total_waiting_time = 0;
        x->phase = NS_GREEN; x->timer = 0; x->duration = 5; /* TrafficLight(), utils.py:75-77 */
    }
}

static int neighbours(const traffic_oracle *o, int id, int *out) { /* utils.py:196-214, order N, S, W, E */
    static const int dr[4] = {-1, 1, 0, 0}, dc[4] = {0, 0, -1, 1};
    int row = id / o->cols, col = id % o->cols, n = 0;
    for (int k = 0; k < 4; ++k) {
        int r = row + dr[k], c = col + dc[k];
        if (r >= 0 && r < o->rows && c >= 0 && c < o->cols) out[n++] = r * o->cols + c;
    }
    return n;
}

static int direction_between(const traffic_oracle *o, int from, int to) { /* utils.py:230-248 */
    int fr = from / o->cols, fc = from % o->cols, tr = to / o->cols, tc = to % o->cols;
    if (tr < fr) return NORTH;
    if (tr > fr) return SOUTH;
    if (tc < fc) return WEST;
    return EAST; /* (the same-intersection random.choice branch is unreachable: a hop always moves) */
}

static void spawn(traffic_oracle *o, traffic_env *e) { /* environment.py:222-249 */
    if (e->n_listed >= o->max_vehicles) return;
    if (!(orc_random(&e->rng) < o->spawn_rate)) return;
    int start = (int)orc_randint(&e->rng, 0, o->ni - 1);
    int route_length = (int)orc_randint(&e->rng, 2, o->ni < 5 ? o->ni : 5); /* utils.py:181 */
    int current = start, second = -1, nb[4];
    for (int h = 1; h < route_length; ++h) {
        int n = neighbours(o, current, nb);
        if (n == 0) break;
        current = nb[orc_randint(&e->rng, 0, n - 1)];
        if (h == 1) second = current;
    }
    if (second < 0) return; /* len(route) == 1 */
    if (e->n_created >= MAX_VEH) return;
    vehicle *v = &e->veh[e->n_created];
    v->direction = direction_between(o, start, second);
    v->destination = current; /* route[-1] */
    v->waiting_time = 0;
    v->alive = 1;
    intersection *x = &e->x[start];
    x->queue[v->direction][x->qlen[v->direction]++] = e->n_created;
    e->n_created++;
    e->n_listed++;
}

static int can_pass(int phase, int d) { /* utils.py:99-106 */
    if (phase == NS_GREEN) return d == NORTH || d == SOUTH;
    if (phase == EW_GREEN) return d == EAST || d == WEST;
    return 0;
}

static double np_sum(const double *a, int n) { /* NumPy pairwise sum, n <= 128 */
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

static int queue_total(const intersection *x) { return x->qlen[0] + x->qlen[1] + x->qlen[2] + x->qlen[3]; }

static double calc_reward(const traffic_oracle *o, const traffic_env *e) { /* environment.py:287-311 */
    long passed = 0, waiting = 0, queued = 0;
    double q[MAX_NI];
    for (int i = 0; i < o->ni; ++i) {
        passed += e->x[i].vehicles_passed; waiting += e->x[i].total_waiting_time;
        queued += queue_total(&e->x[i]); q[i] = (double)queue_total(&e->x[i]);
    }
    double reward = 0.0;
    reward += (double)passed * 1.0;
    reward += (double)waiting * -0.1;
    reward += (double)queued * -0.05;
    if (o->ni > 1) { /* np.var: population variance, float64 */
        double mean = np_sum(q, o->ni) / (double)o->ni, sq[MAX_NI];
        for (int i = 0; i < o->ni; ++i) { double d = q[i] - mean; sq[i] = d * d; }
        double var = np_sum(sq, o->ni) / (double)o->ni;
        reward += 0.5 / (1 + var);
    }
    return reward;
}

static void write_obs(const traffic_oracle *o, const traffic_env *e, float *obs) { /* environment.py:313-363 */
    int k = 0;
    for (int i = 0; i < o->ni; ++i) for (int p = 0; p < 4; ++p) obs[k++] = (e->x[i].phase == p) ? 1.0f : 0.0f;
    for (int i = 0; i < o->ni; ++i) for (int d = 0; d < 4; ++d) obs[k++] = (float)(e->x[i].qlen[d] < 20 ? e->x[i].qlen[d] : 20);
    for (int i = 0; i < o->ni; ++i)
        for (int d = 0; d < 4; ++d) {
            const intersection *x = &e->x[i];
            double avg = 0.0;
            if (x->qlen[d]) {
                long s = 0;
                for (int j = 0; j < x->qlen[d]; ++j) s += e->veh[x->queue[d][j]].waiting_time;
                avg = (double)s / (double)x->qlen[d];
            }
            obs[k++] = (float)(avg < 100 ? avg : 100);
        }
    long passed = 0, waiting = 0, queued = 0;
    for (int i = 0; i < o->ni; ++i) {
        const intersection *x = &e->x[i];
        obs[k++] = (float)x->vehicles_passed;
        obs[k++] = (float)(x->total_waiting_time < 1000 ? x->total_waiting_time : 1000);
        passed += x->vehicles_passed; waiting += x->total_waiting_time; queued += queue_total(x);
    }
    double avg_wait = (double)waiting / (double)(passed > 1 ? passed : 1); /* utils.py:257 */
    double avg_queue = (double)queued / (double)o->ni;
    obs[k++] = (float)e->n_listed;
    obs[k++] = (float)(avg_wait < 100 ? avg_wait : 100);
    obs[k++] = (float)(avg_queue < 50 ? avg_queue : 50);
    obs[k++] = (float)((double)passed / (double)o->ni);
}

traffic_oracle *traffic_oracle_create(int n_envs, int rows, int cols, int num_intersections, int max_vehicles,
                                      double spawn_rate, int max_timesteps, uint64_t seed, uint64_t env_id_base,
                                      int mode) {
    traffic_oracle *o = (traffic_oracle *)calloc(1, sizeof(*o));
    o->n_envs = n_envs; o->mode = mode; o->rows = rows; o->cols = cols;
    o->ni = num_intersections < rows * cols ? num_intersections : rows * cols; /* environment.py:82 */
    if (o->ni > MAX_NI) o->ni = MAX_NI;
    o->max_vehicles = max_vehicles; o->spawn_rate = spawn_rate; o->max_timesteps = max_timesteps;
    o->envs = (traffic_env *)calloc((size_t)n_envs, sizeof(traffic_env));
    for (int i = 0; i < n_envs; ++i) {
        traffic_env *e = &o->envs[i];
        e->x = (intersection *)calloc((size_t)o->ni, sizeof(intersection));
        e->veh = (vehicle *)calloc(MAX_VEH, sizeof(vehicle));
        e->rng.seed = seed; e->rng.env = env_id_base + (uint64_t)i; e->rng.stream = 0; e->rng.counter = 0;
        env_reset(o, e);
    }
    return o;
}

void traffic_oracle_destroy(traffic_oracle *o) {
    if (!o) return;
    for (int i = 0; i < o->n_envs; ++i) { free(o->envs[i].x); free(o->envs[i].veh); }
    free(o->envs); free(o);
}

int traffic_oracle_obs_dim(const traffic_oracle *o) { return obs_dim(o); }

void traffic_oracle_reset(traffic_oracle *o, const uint8_t *mask, float *obs) {
    for (int i = 0; i < o->n_envs; ++i) {
        if (!mask || mask[i]) env_reset(o, &o->envs[i]);
        if (obs) write_obs(o, &o->envs[i], obs + (size_t)obs_dim(o) * i);
    }
}

/* actions: int64 [n][ni] */
void traffic_oracle_step(traffic_oracle *o, const int64_t *actions, float *obs, float *reward, uint8_t *terminated,
                         uint8_t *truncated, double *reward64, double *ep_return, int32_t *ep_length) {
    for (int n = 0; n < o->n_envs; ++n) {
        traffic_env *e = &o->envs[n];
        double rew = 0.0;
        int term = 0;
        if (o->mode == 1 && e->needs_reset) {
            env_reset(o, e);
        } else {
            e->timestep += 1; /* :170 */
            for (int i = 0; i < o->ni; ++i) { /* _apply_actions :205-220 */
                intersection *x = &e->x[i];
                int64_t a = actions[(size_t)n * o->ni + i];
                if (a == 1 && x->phase != NS_GREEN) { x->phase = NS_GREEN; x->duration = 5; x->timer = 5; }
                else if (a == 2 && x->phase != EW_GREEN) { x->phase = EW_GREEN; x->duration = 5; x->timer = 5; }
            }
            for (int i = 0; i < o->ni; ++i) { /* TrafficLight.update, utils.py:79-97 */
                intersection *x = &e->x[i];
                x->timer -= 1;
                if (x->timer <= 0) {
                    x->phase = (x->phase + 1) % 4;
                    x->duration = (x->phase == NS_YELLOW || x->phase == EW_YELLOW) ? 3 : (int)orc_randint(&e->rng, 5, 30);
                    x->timer = x->duration;
                }
            }
            spawn(o, e);
            /* _update_vehicles (:251-269) is a no-op: every vehicle sits exactly on its start intersection */
            for (int i = 0; i < o->ni; ++i) { /* _process_intersections :271-281 + utils.py:141-163 */
                intersection *x = &e->x[i];
                for (int d = 0; d < 4; ++d) {
                    if (!x->qlen[d]) continue;
                    if (can_pass(x->phase, d)) {
                        for (int j = 0; j < x->qlen[d]; ++j) {
                            vehicle *v = &e->veh[x->queue[d][j]];
                            x->vehicles_passed += 1;
                            if (v->destination >= 0 && v->destination == i) v->destination = -1;
                        }
                        x->qlen[d] = 0;
                    } else {
                        for (int j = 0; j < x->qlen[d]; ++j) {
                            e->veh[x->queue[d][j]].waiting_time += 1;
                            x->total_waiting_time += 1;
                        }
                    }
                }
            }
            int listed = 0; /* _remove_completed_vehicles :283-285 */
            for (int j = 0; j < e->n_created; ++j) {
                if (e->veh[j].alive && e->veh[j].destination < 0) e->veh[j].alive = 0;
                listed += e->veh[j].alive;
            }
            e->n_listed = listed;
            rew = calc_reward(o, e);
            e->total_reward += rew;
            term = e->timestep >= o->max_timesteps; /* :196 */
            if (term && o->mode != 0) {
                o->stats[0] += 1; o->stats[1] += e->total_reward; o->stats[2] += e->timestep;
                if (ep_return) ep_return[n] = e->total_reward;
                if (ep_length) ep_length[n] = e->timestep;
                if (o->mode == 2) env_reset(o, e); else e->needs_reset = 1;
            }
        }
        if (obs) write_obs(o, e, obs + (size_t)obs_dim(o) * n);
        reward[n] = (float)rew;
        if (reward64) reward64[n] = rew;
        terminated[n] = (uint8_t)term;
        if (truncated) truncated[n] = 0;
    }
}

/* Per env: timestep, num_vehicles, rng counter, total_reward; per intersection: phase, timer, passed, waiting;
 * per queue: length, sum of waiting times. */
void traffic_oracle_get_state(const traffic_oracle *o, int32_t *timestep, int32_t *num_vehicles, uint32_t *rng_counter,
                              double *total_reward, int32_t *phase, int32_t *timer, int32_t *passed, int32_t *waiting,
                              int32_t *qlen, int32_t *qwait) {
    for (int n = 0; n < o->n_envs; ++n) {
        const traffic_env *e = &o->envs[n];
        timestep[n] = e->timestep; num_vehicles[n] = e->n_listed; rng_counter[n] = e->rng.counter;
        total_reward[n] = e->total_reward;
        for (int i = 0; i < o->ni; ++i) {
            const intersection *x = &e->x[i];
            size_t k = (size_t)n * o->ni + i;
            phase[k] = x->phase; timer[k] = x->timer; passed[k] = x->vehicles_passed; waiting[k] = x->total_waiting_time;
            for (int d = 0; d < 4; ++d) {
                long s = 0;
                for (int j = 0; j < x->qlen[d]; ++j) s += e->veh[x->queue[d][j]].waiting_time;
                qlen[k * 4 + d] = x->qlen[d]; qwait[k * 4 + d] = (int32_t)s;
            }
        }
    }
}

void traffic_oracle_get_stats(const traffic_oracle *o, double *out3) { memcpy(out3, o->stats, sizeof(o->stats)); }

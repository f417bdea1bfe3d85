/* CPU oracle for SmartClimateEnv -- a plain-C, float64 restatement of the reference algorithm.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Follows /root/reference/smartclimate_rl-main/smartclimate/:
 *   env.py:16-35  constructor (max_occupancy 8, episode_minutes 1440)     env.py:48-60 _init_state
 *   env.py:62-70  reset          env.py:72-82 _get_obs          env.py:84-117 step
 *   utils.py:5-13 get_outside_temp   :15-22 update_occupancy   :24-28 room_temp_dynamics   :30-50 calculate_reward
 * The time limit is reported as `terminated` (env.py:107), `truncated` is always False (:108).
 *
 * RNG draw sites in order: _init_state = uniform(22, 26), integers(0, max_occupancy + 1), normal(25, 5);
 * step = normal(base(time_of_day), 5), choice(4 values, p).  `choice` with probabilities is the inverse-CDF rule
 * of numpy's Generator.choice: index = number of cdf entries <= u, cdf = cumsum(p) / cumsum(p)[-1]; the two cdf
 * tables below are those float64 values (oracle/replay.py ReplayGenerator computes them with numpy at run time).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beng_oracle_rng.h"

/* cumsum([0.1,0.3,0.4,0.2]) and cumsum([0.2,0.4,0.3,0.1]) / last, as numpy produces them */
static const double CDF_DAY[4] = {0x1.999999999999ap-4, 0x1.999999999999ap-2, 0x1.999999999999ap-1, 0x1p+0};
static const double CDF_NIGHT[4] = {0x1.9999999999998p-3, 0x1.3333333333333p-1, 0x1.cccccccccccccp-1, 0x1p+0};
static const int CHANGE_DAY[4] = {-1, 0, 1, 2}, CHANGE_NIGHT[4] = {-2, -1, 0, 1};

typedef struct {
    double room_temp, outside_temp, ac_setting, total_reward, energy_usage;
    int num_people, lights[4], current_step, comfort_time, needs_reset;
    orc_stream rng;
} climate_env;

typedef struct {
    int n_envs, mode, max_occupancy, episode_minutes;
    climate_env *envs;
    double stats[3]; /* n_episodes, sum_return, sum_length */
} climate_oracle;

static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static double outside_temp(double tod, orc_stream *rng) { /* utils.py:5-13 */
    double base = (0 <= tod && tod < 8) ? 25 : ((8 <= tod && tod < 16) ? 45 : 35);
    return orc_normal(rng, base, 5);
}

static void init_state(climate_oracle *o, climate_env *e) { /* env.py:48-60 */
    e->room_temp = orc_uniform(&e->rng, 22.0, 26.0);
    e->num_people = (int)orc_randint(&e->rng, 0, o->max_occupancy); /* integers(0, max+1): high exclusive */
    e->outside_temp = outside_temp(0.0, &e->rng);
    e->ac_setting = 24.0;
    memset(e->lights, 0, sizeof(e->lights));
    e->total_reward = 0.0; e->comfort_time = 0; e->energy_usage = 0.0;
    e->current_step = 0; e->needs_reset = 0;
}

static void write_obs(const climate_env *e, float *obs) { /* env.py:72-82 */
    obs[0] = (float)e->room_temp; obs[1] = (float)e->num_people;
    obs[2] = (float)((e->current_step % 1440) / 60.0);
    obs[3] = (float)e->outside_temp; obs[4] = (float)e->ac_setting;
    for (int i = 0; i < 4; ++i) obs[5 + i] = (float)e->lights[i];
}

climate_oracle *climate_oracle_create(int n_envs, int max_occupancy, int episode_minutes, uint64_t seed,
                                      uint64_t env_id_base, int mode) {
    climate_oracle *o = (climate_oracle *)calloc(1, sizeof(*o));
    o->n_envs = n_envs; o->mode = mode; o->max_occupancy = max_occupancy; o->episode_minutes = episode_minutes;
    o->envs = (climate_env *)calloc((size_t)n_envs, sizeof(climate_env));
    for (int i = 0; i < n_envs; ++i) {
        o->envs[i].rng.seed = seed; o->envs[i].rng.env = env_id_base + (uint64_t)i;
        o->envs[i].rng.stream = 0; o->envs[i].rng.counter = 0;
    }
    return o;
}

void climate_oracle_destroy(climate_oracle *o) { if (o) { free(o->envs); free(o); } }

void climate_oracle_reset(climate_oracle *o, const uint8_t *mask, float *obs) {
    for (int i = 0; i < o->n_envs; ++i) {
        if (!mask || mask[i]) init_state(o, &o->envs[i]);
        if (obs) write_obs(&o->envs[i], obs + 9 * (size_t)i);
    }
}

/* ac_temp: float32 [n]; lights: int8 [n][4].  info outputs may be NULL. */
void climate_oracle_step(climate_oracle *o, const float *ac_temp, const int8_t *lights, float *obs, float *reward,
                         uint8_t *terminated, uint8_t *truncated, double *reward64, double *comfort,
                         double *ac_penalty, double *light_penalty, double *ep_return, int32_t *ep_length) {
    for (int n = 0; n < o->n_envs; ++n) {
        climate_env *e = &o->envs[n];
        double rew = 0.0, cf = 0.0, acp = 0.0, lp = 0.0;
        int term = 0;
        if (o->mode == 1 && e->needs_reset) {
            init_state(o, e);
        } else {
            e->ac_setting = clipd((double)ac_temp[n], 16.0, 32.0); /* env.py:85 */
            int lights_on = 0;
            for (int i = 0; i < 4; ++i) { e->lights[i] = lights[4 * n + i]; lights_on += e->lights[i]; }
            e->current_step += 1;
            double tod = (e->current_step % 1440) / 60.0; /* :91 */
            e->outside_temp = outside_temp(tod, &e->rng);
            { /* update_occupancy, utils.py:15-22 */
                int day = (9 <= tod && tod < 18);
                double u = orc_random(&e->rng);
                const double *cdf = day ? CDF_DAY : CDF_NIGHT;
                int idx = 0;
                while (idx < 3 && cdf[idx] <= u) ++idx; /* searchsorted(cdf, u, side='right') */
                int change = (day ? CHANGE_DAY : CHANGE_NIGHT)[idx];
                int v = e->num_people + change;
                e->num_people = v < 0 ? 0 : (v > o->max_occupancy ? o->max_occupancy : v);
            }
            /* room_temp_dynamics, utils.py:24-28 */
            double prev = e->room_temp;
            double temp = prev + 0.1 * (e->outside_temp - prev) + 0.2 * (e->ac_setting - prev) + e->num_people * 1.0;
            e->room_temp = clipd(temp, 10, 50);
            /* calculate_reward, utils.py:30-50 */
            double rt = e->room_temp;
            if (20 <= rt && rt <= 24) cf = 10;
            else if (18 <= rt && rt <= 26) cf = 5;
            else if (16 <= rt && rt <= 28) cf = 0;
            else cf = -15 * fabs(rt - 22);
            acp = -0.5 * fabs(e->ac_setting - e->outside_temp);
            int required = (e->num_people + 1) / 2; /* ceil(num_people / 2) */
            if (required > 4) required = 4;
            int extra = lights_on - required;
            lp = -1.0 * (extra > 0 ? extra : 0);
            rew = cf + acp + lp;
            e->total_reward += rew;
            if (20 <= rt && rt <= 24) e->comfort_time += 1;
            e->energy_usage += fabs(e->ac_setting - e->outside_temp) + lights_on;
            term = e->current_step >= o->episode_minutes; /* :107 */
            if (term && o->mode != 0) {
                o->stats[0] += 1; o->stats[1] += e->total_reward; o->stats[2] += e->current_step;
                if (ep_return) ep_return[n] = e->total_reward;
                if (ep_length) ep_length[n] = e->current_step;
                if (o->mode == 2) init_state(o, e); else e->needs_reset = 1;
            }
        }
        if (obs) write_obs(e, obs + 9 * (size_t)n);
        reward[n] = (float)rew;
        if (reward64) reward64[n] = rew;
        if (comfort) comfort[n] = cf;
        if (ac_penalty) ac_penalty[n] = acp;
        if (light_penalty) light_penalty[n] = lp;
        terminated[n] = (uint8_t)term;
        if (truncated) truncated[n] = 0;
    }
}

void climate_oracle_get_state(const climate_oracle *o, double *room_temp, double *outside, double *total_reward,
                              double *energy_usage, int32_t *num_people, int32_t *current_step,
                              int32_t *comfort_time, uint32_t *rng_counter) {
    for (int i = 0; i < o->n_envs; ++i) {
        const climate_env *e = &o->envs[i];
        room_temp[i] = e->room_temp; outside[i] = e->outside_temp; total_reward[i] = e->total_reward;
        energy_usage[i] = e->energy_usage; num_people[i] = e->num_people; current_step[i] = e->current_step;
        comfort_time[i] = e->comfort_time; rng_counter[i] = e->rng.counter;
    }
}

void climate_oracle_get_stats(const climate_oracle *o, double *out3) { memcpy(out3, o->stats, sizeof(o->stats)); }

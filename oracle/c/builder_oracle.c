/* CPU oracle for WorldBuilderEnv -- a plain-C restatement of the reference algorithm (integer dynamics).
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Follows /root/reference/world_builder_env/src/environment/:
 *   world_builder_env.py:38-97 constructor, :99-123 reset, :125-166 step, :186-217 _get_observation,
 *   :233-247 _check_termination
 *   game_logic.py:33-57 reset, :59-123 execute_action (the "smart reward"), :125-156 _try_build,
 *   :158-170 costs, :172-183 _process_production, :185-192 _process_consumption, :194-203 _process_population_growth
 * There is no time limit in the reference (no gymnasium registration either): an episode ends when the population
 * starves (reward -100) or after 50 steps at population >= 20 (reward +100).  `truncated` is always False.
 *
 * RNG: one draw per SUCCESSFUL build -- `np.random.randint(len(empty_positions[0]))` (game_logic.py:137), the index of
 * the chosen cell among the empty cells in row-major order (np.where order).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beng_oracle_rng.h"

enum { FARM = 1, LUMBERYARD = 2, QUARRY = 3, HOUSE = 4 };

typedef struct {
    int8_t *grid;
    int food, wood, stone, population, capacity;
    int counts[5];
    int steps, win_steps, reached, needs_reset;
    long ep_return;
    orc_stream rng;
} builder_env;

typedef struct {
    int n_envs, mode, G;
    builder_env *envs;
    int64_t stats[4]; /* n_episodes, sum_return, sum_length, wins */
} builder_oracle;

static void env_reset(builder_oracle *o, builder_env *e) { /* world_builder_env.py:99-123 + game_logic.py:33-57 */
    memset(e->grid, 0, (size_t)o->G * o->G);
    e->food = 25; e->wood = 20; e->stone = 10; e->population = 3; e->capacity = 10;
    memset(e->counts, 0, sizeof(e->counts));
    e->steps = 0; e->win_steps = 0; e->reached = 0; e->needs_reset = 0; e->ep_return = 0;
}

static int can_afford(const builder_env *e, int b) { /* game_logic.py:13-18, :158-164 */
    switch (b) {
        case FARM: return e->wood >= 5;
        case LUMBERYARD: return e->stone >= 3;
        case QUARRY: return e->wood >= 5;
        default: return e->wood >= 10 && e->stone >= 5;
    }
}

static int try_build(builder_oracle *o, builder_env *e, int b) { /* game_logic.py:125-156 */
    if (!can_afford(e, b)) return 0;
    int cells = o->G * o->G, n_empty = 0;
    for (int i = 0; i < cells; ++i) n_empty += (e->grid[i] == 0);
    if (n_empty == 0) return 0;
    int idx = (int)orc_randint(&e->rng, 0, n_empty - 1); /* np.random.randint(n_empty) */
    int pos = -1;
    for (int i = 0; i < cells; ++i)
        if (e->grid[i] == 0 && idx-- == 0) { pos = i; break; }
    switch (b) { /* _spend_resources */
        case FARM: e->wood -= 5; break;
        case LUMBERYARD: e->stone -= 3; break;
        case QUARRY: e->wood -= 5; break;
        default: e->wood -= 10; e->stone -= 5; break;
    }
    e->grid[pos] = (int8_t)b;
    e->counts[b] += 1;
    if (b == HOUSE) e->capacity += 5;
    return 1;
}

static int execute_action(builder_oracle *o, builder_env *e, int action) { /* game_logic.py:59-123 */
    int prev_population = e->population, prev_capacity = e->capacity, reward = 0;
    if (action != 0) {
        if (try_build(o, e, action)) {
            reward += (action == FARM) ? 3 : (action == HOUSE) ? 4 : 2;
            if (action == HOUSE && prev_population >= prev_capacity - 1) reward += 10;
        } else {
            reward -= 3;
        }
    }
    /* _process_production (dict order: farm, lumberyard, quarry, house) */
    e->food += 2 * e->counts[FARM];
    e->wood += 3 * e->counts[LUMBERYARD];
    e->stone += 2 * e->counts[QUARRY];
    /* _process_consumption */
    if (e->food < e->population) e->population = 0; else e->food -= e->population;
    /* _process_population_growth */
    if (e->population > 0 && e->food > 2 && e->population < e->capacity) { e->population += 1; e->food -= 1; }
    if (e->population > prev_population) reward += 5;
    if (e->population < prev_population) reward -= 50;
    if (e->food > e->population * 2) reward += 1;
    if (e->food < e->population) reward -= 2;
    if (e->food < (e->population > 2 ? e->population : 2)) reward -= 5;
    if (abs(e->wood - e->stone) < 5) reward += 1;
    if (action == FARM && e->food > e->population * 3) reward -= 1;
    return reward;
}

builder_oracle *builder_oracle_create(int n_envs, int grid_size, uint64_t seed, uint64_t env_id_base, int mode) {
    builder_oracle *o = (builder_oracle *)calloc(1, sizeof(*o));
    o->n_envs = n_envs; o->mode = mode; o->G = grid_size;
    o->envs = (builder_env *)calloc((size_t)n_envs, sizeof(builder_env));
    for (int i = 0; i < n_envs; ++i) {
        builder_env *e = &o->envs[i];
        e->grid = (int8_t *)calloc((size_t)grid_size * grid_size, 1);
        e->rng.seed = seed; e->rng.env = env_id_base + (uint64_t)i; e->rng.stream = 0; e->rng.counter = 0;
        env_reset(o, e);
    }
    return o;
}

void builder_oracle_destroy(builder_oracle *o) {
    if (!o) return;
    for (int i = 0; i < o->n_envs; ++i) free(o->envs[i].grid);
    free(o->envs); free(o);
}

static void write_obs(const builder_oracle *o, const builder_env *e, int8_t *grid, float *resources, float *capacity,
                      int32_t *win_steps, size_t i) { /* world_builder_env.py:186-217 */
    size_t cells = (size_t)o->G * o->G;
    if (grid) memcpy(grid + cells * i, e->grid, cells);
    if (resources) {
        resources[4 * i] = (float)e->food; resources[4 * i + 1] = (float)e->wood;
        resources[4 * i + 2] = (float)e->stone; resources[4 * i + 3] = (float)e->population;
    }
    if (capacity) capacity[i] = (float)e->capacity;
    if (win_steps) win_steps[i] = e->win_steps;
}

void builder_oracle_reset(builder_oracle *o, const uint8_t *mask, int8_t *grid, float *resources, float *capacity,
                          int32_t *win_steps) {
    for (int i = 0; i < o->n_envs; ++i) {
        if (!mask || mask[i]) env_reset(o, &o->envs[i]);
        write_obs(o, &o->envs[i], grid, resources, capacity, win_steps, (size_t)i);
    }
}

/* Returns the number of invalid actions (the reference raises ValueError, world_builder_env.py:135-136). */
int builder_oracle_step(builder_oracle *o, const int64_t *actions, int8_t *grid, float *resources, float *capacity,
                        int32_t *win_steps, float *reward, uint8_t *terminated, uint8_t *truncated,
                        int32_t *ep_return, int32_t *ep_length) {
    int invalid = 0;
    for (int n = 0; n < o->n_envs; ++n) {
        builder_env *e = &o->envs[n];
        int rew = 0, term = 0;
        int64_t a = actions[n];
        if (o->mode == 1 && e->needs_reset) {
            env_reset(o, e);
        } else if (a < 0 || a > 4) {
            ++invalid;
        } else {
            e->steps += 1;
            rew = execute_action(o, e, (int)a);
            if (e->population >= 20 && !e->reached) e->reached = 1; /* MAX_POPULATION, :144-145 */
            if (e->reached) e->win_steps += 1;
            term = (e->population <= 0) || (e->reached && e->win_steps >= 50); /* :233-247 */
            if (term) rew = (e->population <= 0) ? -100 : ((e->win_steps >= 50) ? 100 : -50);
            e->ep_return += rew;
            if (term && o->mode != 0) {
                o->stats[0] += 1; o->stats[1] += e->ep_return; o->stats[2] += e->steps;
                o->stats[3] += (e->population > 0);
                if (ep_return) ep_return[n] = (int32_t)e->ep_return;
                if (ep_length) ep_length[n] = e->steps;
                if (o->mode == 2) env_reset(o, e); else e->needs_reset = 1;
            }
        }
        write_obs(o, e, grid, resources, capacity, win_steps, (size_t)n);
        reward[n] = (float)rew;
        terminated[n] = (uint8_t)term;
        if (truncated) truncated[n] = 0;
    }
    return invalid;
}

/* steps, reached flag, rng counter, building counts [n][4] (farm, lumberyard, quarry, house) */
void builder_oracle_get_state(const builder_oracle *o, int32_t *steps, int32_t *reached, uint32_t *rng_counter,
                              int32_t *counts) {
    for (int i = 0; i < o->n_envs; ++i) {
        const builder_env *e = &o->envs[i];
        steps[i] = e->steps; reached[i] = e->reached; rng_counter[i] = e->rng.counter;
        for (int b = 0; b < 4; ++b) counts[4 * i + b] = e->counts[b + 1];
    }
}

void builder_oracle_get_stats(const builder_oracle *o, int64_t *out4) { memcpy(out4, o->stats, sizeof(o->stats)); }

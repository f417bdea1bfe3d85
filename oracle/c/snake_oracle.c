/* CPU oracle for SnakeEnvClassic -- a plain-C restatement of the reference algorithm.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this.  It is never a product path.
 *
 * Follows /root/reference/snake_env_classic/snake_env.py:
 *   reset            :49-65     (head at (G//2, G//2), direction 1, score = steps = 0, food)
 *   step             :67-119    (reversal guard :73-74, move :77-85, wall :88-90, self :93-94,
 *                                insert :97, eat/grow :100-104, pop tail :105-107, steps/limit :109-114)
 *   _place_food      :121-129   (rejection loop, two randint draws per attempt: row then column)
 *   _get_observation :131-143   (zeros; body = 1; food = 2, written last)
 * and SURVEY.md section 0 facts 4-6 (no auto-reset in the reference, time limit reported as
 * terminated, the death step mutates nothing but `direction`).
 *
 * The body is kept the way the reference keeps it -- an ordered list, head first, searched
 * linearly for `in` -- deliberately NOT the ring + occupancy tile the CUDA kernel uses, so that
 * the two implementations do not share a data-structure bug.
 *
 * Auto-reset is the caller loop of SURVEY.md section 3.5 (`if terminated: env.reset()`), folded in:
 *   mode 0 DISABLED : exactly the reference class, nothing else.
 *   mode 1 NEXT_STEP: the step after a terminal one ignores its action and performs reset().
 *   mode 2 SAME_STEP: reset() runs inside the terminal step; the obs returned is the reset obs.
 *
 * Parity pin: tests/golden/snake_*.npz, generated from the reference itself by oracle/gen_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beng_oracle_rng.h"

typedef struct {
    int *body_r, *body_c; /* head first, like the reference's list */
    int length;
    int food_r, food_c;
    int direction;
    int score;
    int steps;
    int needs_reset; /* NEXT_STEP bookkeeping only */
    orc_stream rng;
} snake_env;

typedef struct {
    int n_envs, grid, max_steps, mode;
    snake_env *envs;
    int64_t stats[5]; /* n_episodes, sum_return, sum_length, sum_score, max_score */
} snake_oracle;

static int in_body(const snake_env *e, int r, int c) {
    for (int i = 0; i < e->length; ++i)
        if (e->body_r[i] == r && e->body_c[i] == c) return 1;
    return 0;
}

/* snake_env.py:121-129 */
static void place_food(snake_oracle *o, snake_env *e) {
    if (e->length >= o->grid * o->grid) {
        /* Board full: the reference's rejection loop never returns (unreachable at G = 20 within
         * 1000 steps).  Engine and oracle agree to leave the board without food instead. */
        e->food_r = e->food_c = -1;
        return;
    }
    for (;;) {
        e->food_r = (int)orc_randint(&e->rng, 0, o->grid - 1);
        e->food_c = (int)orc_randint(&e->rng, 0, o->grid - 1);
        if (!in_body(e, e->food_r, e->food_c)) break;
    }
}

/* snake_env.py:49-65 */
static void env_reset(snake_oracle *o, snake_env *e) {
    int center = o->grid / 2;
    e->body_r[0] = center;
    e->body_c[0] = center;
    e->length = 1;
    e->direction = 1;
    e->score = 0;
    e->steps = 0;
    e->needs_reset = 0;
    place_food(o, e);
}

/* snake_env.py:131-143 */
static void write_obs(const snake_oracle *o, const snake_env *e, int8_t *obs) {
    int G = o->grid;
    memset(obs, 0, (size_t)G * G);
    for (int i = 0; i < e->length; ++i) obs[e->body_r[i] * G + e->body_c[i]] = 1;
    if (e->food_r >= 0) obs[e->food_r * G + e->food_c] = 2;
}

snake_oracle *snake_oracle_create(int n_envs, int grid, int max_steps, uint64_t seed, uint64_t env_id_base, int mode) {
    snake_oracle *o = (snake_oracle *)calloc(1, sizeof(*o));
    o->n_envs = n_envs; o->grid = grid; o->max_steps = max_steps; o->mode = mode;
    o->envs = (snake_env *)calloc((size_t)n_envs, sizeof(snake_env));
    o->stats[4] = INT64_MIN;
    for (int i = 0; i < n_envs; ++i) {
        snake_env *e = &o->envs[i];
        e->body_r = (int *)malloc(sizeof(int) * (size_t)(grid * grid + 1));
        e->body_c = (int *)malloc(sizeof(int) * (size_t)(grid * grid + 1));
        e->rng.seed = seed; e->rng.env = env_id_base + (uint64_t)i; e->rng.stream = 0; e->rng.counter = 0;
    }
    return o;
}

void snake_oracle_destroy(snake_oracle *o) {
    if (!o) return;
    for (int i = 0; i < o->n_envs; ++i) { free(o->envs[i].body_r); free(o->envs[i].body_c); }
    free(o->envs);
    free(o);
}

/* mask == NULL resets every env; obs/score/length may be NULL. */
void snake_oracle_reset(snake_oracle *o, const uint8_t *mask, int8_t *obs, int32_t *score, int32_t *length) {
    size_t cells = (size_t)o->grid * o->grid;
    for (int i = 0; i < o->n_envs; ++i) {
        snake_env *e = &o->envs[i];
        if (!mask || mask[i]) env_reset(o, e);
        if (obs) write_obs(o, e, obs + cells * i);
        if (score) score[i] = e->score;
        if (length) length[i] = e->length;
    }
}

static void episode_end(snake_oracle *o, const snake_env *e, int died, int i, float *ep_return, int32_t *ep_length,
                        int32_t *ep_score) {
    int64_t ret = 10 * (int64_t)e->score - (died ? 10 : 0);
    int64_t len = e->steps + (died ? 1 : 0); /* the death step does not increment `steps` (snake_env.py:88-94) */
    o->stats[0] += 1; o->stats[1] += ret; o->stats[2] += len; o->stats[3] += e->score;
    if (e->score > o->stats[4]) o->stats[4] = e->score;
    if (ep_return) ep_return[i] = (float)ret;
    if (ep_length) ep_length[i] = (int32_t)len;
    if (ep_score) ep_score[i] = e->score;
}

/* Returns the number of invalid actions seen (the reference raises ValueError, snake_env.py:69-70;
 * here an invalid action leaves the env untouched, reward 0, terminated 0, and is counted). */
int snake_oracle_step(snake_oracle *o, const int64_t *actions, int8_t *obs, float *reward, uint8_t *terminated,
                      uint8_t *truncated, int32_t *score, int32_t *length, float *ep_return, int32_t *ep_length,
                      int32_t *ep_score) {
    int G = o->grid, invalid = 0;
    size_t cells = (size_t)G * G;
    for (int i = 0; i < o->n_envs; ++i) {
        snake_env *e = &o->envs[i];
        float rew = 0.0f;
        int term = 0;
        int64_t a = actions[i];
        if (o->mode == 1 && e->needs_reset) {
            env_reset(o, e);
        } else if (a < 0 || a > 3) {
            ++invalid;
        } else {
            if (llabs(a - e->direction) != 2) e->direction = (int)a; /* :73-74 */
            int nr = e->body_r[0], nc = e->body_c[0];
            switch (e->direction) { /* :77-85 */
                case 0: nr -= 1; break;
                case 1: nc += 1; break;
                case 2: nr += 1; break;
                default: nc -= 1; break;
            }
            int died = (nr < 0 || nr >= G || nc < 0 || nc >= G) /* :88-90 */
                       || in_body(e, nr, nc);                   /* :93-94, tail still present */
            if (died) {
                rew = -10.0f;
                term = 1;
            } else {
                memmove(e->body_r + 1, e->body_r, sizeof(int) * (size_t)e->length); /* insert(0, new_head) :97 */
                memmove(e->body_c + 1, e->body_c, sizeof(int) * (size_t)e->length);
                e->body_r[0] = nr; e->body_c[0] = nc;
                e->length += 1;
                if (nr == e->food_r && nc == e->food_c) { /* :101-104 */
                    e->score += 1;
                    rew = 10.0f;
                    place_food(o, e);
                } else {
                    e->length -= 1; /* pop() :107 */
                }
                e->steps += 1;                       /* :109 */
                if (e->steps >= o->max_steps) term = 1; /* :112-114 */
            }
            if (term && o->mode != 0) {
                episode_end(o, e, died, i, ep_return, ep_length, ep_score);
                if (o->mode == 2) env_reset(o, e); else e->needs_reset = 1;
            }
        }
        if (obs) write_obs(o, e, obs + cells * i);
        reward[i] = rew;
        terminated[i] = (uint8_t)term;
        if (truncated) truncated[i] = 0; /* the reference never truncates (fact 5) */
        if (score) score[i] = e->score;
        if (length) length[i] = e->length;
    }
    return invalid;
}

/* State readback for field-by-field comparison with the device SoA state. */
void snake_oracle_get_state(const snake_oracle *o, int32_t *head_r, int32_t *head_c, int32_t *food_r, int32_t *food_c,
                            int32_t *direction, int32_t *steps, int32_t *length, uint32_t *rng_counter) {
    for (int i = 0; i < o->n_envs; ++i) {
        const snake_env *e = &o->envs[i];
        head_r[i] = e->body_r[0]; head_c[i] = e->body_c[0];
        food_r[i] = e->food_r; food_c[i] = e->food_c;
        direction[i] = e->direction; steps[i] = e->steps; length[i] = e->length;
        rng_counter[i] = e->rng.counter;
    }
}

/* Body cells of one env, head first, as r*G+c; returns the length. */
int snake_oracle_get_body(const snake_oracle *o, int env, int32_t *cells_out) {
    const snake_env *e = &o->envs[env];
    for (int i = 0; i < e->length; ++i) cells_out[i] = e->body_r[i] * o->grid + e->body_c[i];
    return e->length;
}

void snake_oracle_get_stats(const snake_oracle *o, int64_t *out5) { memcpy(out5, o->stats, sizeof(o->stats)); }

/* Raw stream access so tests can cross-check the three Philox restatements. */
void beng_oracle_draws_u32(uint64_t seed, uint64_t env, uint32_t stream, uint32_t first, uint32_t count, uint32_t *out) {
    orc_stream s = {seed, env, stream, first};
    for (uint32_t i = 0; i < count; ++i) out[i] = orc_u32(&s);
}

/* Synthetic action tape == csrc beng_fill_random_actions == oracle/philox.py action_tape. */
void beng_oracle_action_tape(uint64_t seed, uint64_t env_id_base, int n_envs, uint32_t step, int n_choices, int n_cols,
                             int64_t *out) {
    for (int i = 0; i < n_envs; ++i) {
        orc_stream s = {seed, env_id_base + (uint64_t)i, 1u, step * (uint32_t)n_cols};
        for (int c = 0; c < n_cols; ++c) out[(size_t)i * n_cols + c] = orc_randint(&s, 0, n_choices - 1);
    }
}

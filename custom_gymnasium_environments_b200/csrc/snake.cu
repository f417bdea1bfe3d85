// Batched SnakeEnvClassic for sm_100a: step + reward + termination + auto-reset + food RNG +
// observation encoding in ONE kernel.
//
// Reference behaviour (paths relative to the reference root):
//   SnakeEnvClassic.reset            snake_env_classic/snake_env.py:49-65
//   SnakeEnvClassic.step             snake_env_classic/snake_env.py:67-119
//   SnakeEnvClassic._place_food      snake_env_classic/snake_env.py:121-129
//   SnakeEnvClassic._get_observation snake_env_classic/snake_env.py:131-143
//
// Design:
//   * one THREAD per env does the integer dynamics out of a 16-byte state record (one LDG.128 /
//     STG.128 per env, perfectly coalesced);
//   * one CTA owns a TILE of T consecutive envs; their T x G*G observation bytes are contiguous in
//     global memory, so the tile is composed in shared memory (zero fill, scatter body cells, food
//     last) and drained with ONE bulk asynchronous copy (cp.async.bulk shared->global, the TMA
//     engine, SASS UBLKCP) issued by a single thread -- no per-thread obs stores at all;
//   * CTAs are persistent over tiles with a STAGES-deep ring of tile buffers, so the copy engine
//     drains tile k while the threads compose tile k+1;
//   * the shared-memory tile doubles as the O(1) occupancy map for self-collision and for the
//     food rejection loop (the reference scans a Python list for both);
//   * finished episodes are counted with warp votes, accumulated per CTA in shared memory and
//     flushed with one set of global atomics per CTA; a warp-aggregated compaction emits the
//     list of finished envs.
//
// The kernel is HBM-bound (450 algorithmic bytes per env-step, 400 of them the observation
// write); there is nothing GEMM-shaped here and no tensor-core path is wanted.
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "beng_common.cuh"
#include "beng_rng.cuh"

namespace beng {
namespace {

constexpr uint32_t FLAG_NEEDS_RESET = 1u;  // NEXT_STEP auto-reset bookkeeping
constexpr int FOOD_NONE = 255;             // board full: no food on the grid (see place_food)

// ---- 16-byte state record ---------------------------------------------------------------------
struct Core {
    int head_r, head_c, food_r, food_c;
    int dir;
    uint32_t flags;
    int length, steps, ring_head;
    uint32_t ctr;
};

__device__ __forceinline__ Core unpack(const uint4 v) {
    Core s;
    s.head_r = v.x & 0xFF;
    s.head_c = (v.x >> 8) & 0xFF;
    s.food_r = (v.x >> 16) & 0xFF;
    s.food_c = (v.x >> 24) & 0xFF;
    s.dir = v.y & 0xFF;
    s.flags = (v.y >> 8) & 0xFF;
    s.length = (v.y >> 16) & 0xFFFF;
    s.steps = v.z & 0xFFFF;
    s.ring_head = (v.z >> 16) & 0xFFFF;
    s.ctr = v.w;
    return s;
}

__device__ __forceinline__ uint4 pack(const Core &s) {
    uint4 v;
    v.x = (uint32_t)s.head_r | ((uint32_t)s.head_c << 8) | ((uint32_t)s.food_r << 16) | ((uint32_t)s.food_c << 24);
    v.y = (uint32_t)s.dir | (s.flags << 8) | ((uint32_t)s.length << 16);
    v.z = (uint32_t)s.steps | ((uint32_t)s.ring_head << 16);
    v.w = s.ctr;
    return v;
}

struct Args {
    uint4 *core;
    uint16_t *ring;
    const long long *actions;  // step only
    const uint8_t *mask;       // reset only (nullable)
    beng_snake_io io;
    long long n_envs;
    uint64_t seed, env_id_base;
    int G, max_steps, mode, first_call, tl_trunc;
};

// _place_food (snake_env.py:121-129): draw (row, col) until the cell is not part of the snake.
// `row` holds 1 for every body cell at this point (the food mark is written afterwards), so the
// membership test is one shared-memory byte load.  The reference never returns when the board is
// full (length == G*G, unreachable at G = 20 within 1000 steps); there the engine and the oracle
// both leave the board without food instead of spinning.
__device__ __forceinline__ void place_food(Core &s, EnvStream &rng, const uint8_t *row, int G) {
    if (s.length >= G * G) {
        s.food_r = s.food_c = FOOD_NONE;
        return;
    }
    for (;;) {
        const int r = rng.randint(0, G - 1);
        const int c = rng.randint(0, G - 1);
        if (row[r * G + c] == 0) {
            s.food_r = r;
            s.food_c = c;
            return;
        }
    }
}

// reset (snake_env.py:49-65).  Expects `row` to be all zero.
__device__ __forceinline__ void reset_env(Core &s, EnvStream &rng, uint8_t *row, int G) {
    const int center = G / 2;
    s.head_r = s.head_c = center;
    s.dir = 1;
    s.length = 1;
    s.steps = 0;
    s.ring_head = 0;
    s.flags = 0;
    row[center * G + center] = 1;  // (the ring is not touched: it only holds bodies of length >= 2)
    place_food(s, rng, row, G);
}

// Walk the body oldest -> newest, writing `mark` into the tile row.  Returns the tail cell and
// reports whether `probe` is one of the body cells (`new_head in self.snake`, snake_env.py:93).
// Only called for length >= 2.  `pre_tail` >= 0 is the tail entry fetched one tile ahead; the newest
// entry always equals the head held in the state record, so a length-2 snake needs no ring load here.
__device__ __forceinline__ int scan_body(const Core &s, const uint16_t *ring, uint8_t *row, int cells, int G,
                                         uint8_t mark, int probe, bool &hit, int pre_tail = -1) {
    int idx = s.ring_head - (s.length - 1);
    if (idx < 0) idx += cells;
    const int tail = (pre_tail >= 0) ? pre_tail : (int)ring[idx];
    row[tail] = mark;
    hit |= (tail == probe);
    for (int i = 1; i < s.length - 1; ++i) {
        if (++idx == cells) idx = 0;
        const int c = ring[idx];
        row[c] = mark;
        hit |= (c == probe);
    }
    const int head = s.head_r * G + s.head_c;
    row[head] = mark;
    hit |= (head == probe);
    return tail;
}

// Index of the tail entry in the ring (valid for length >= 2).
__device__ __forceinline__ int tail_index(const Core &s, int cells) {
    const int idx = s.ring_head - (s.length - 1);
    return idx < 0 ? idx + cells : idx;
}

// Per-env inputs of one tile, loaded one tile ahead of their use (software pipeline).
struct Loaded {
    uint4 raw;
    long long act;
    bool selected;
};

// Draw the current body into a zeroed row (no move).
__device__ __forceinline__ void draw_body(const Core &s, const uint16_t *ring, uint8_t *row, int cells, int G) {
    bool hit = false;
    if (s.length == 1) row[s.head_r * G + s.head_c] = 1;
    else if (s.length > 1) scan_body(s, ring, row, cells, G, 1, -1, hit);
}

template <bool IS_RESET>
__device__ __forceinline__ Loaded load_inputs(const Args &a, long long env) {
    Loaded l;
    l.raw = make_uint4(0, 0, 0, 0);
    l.act = 0;
    l.selected = true;
    if (env < a.n_envs) {
        l.raw = ld_stream_u4(a.core + env);
        if constexpr (!IS_RESET) l.act = ld_stream_s64(a.actions + env);
        else if (a.mask) l.selected = a.mask[env] != 0;
    }
    return l;
}

// OWNROW: every thread zero-fills its own env row with 16-byte stores (needs G*G % 16 == 0, and is
// bank-conflict free when G*G/16 is odd, e.g. G = 20: consecutive rows are 25 x 16 B apart so eight
// threads cover all 32 banks).  Otherwise the tile is zero-filled cooperatively + one more barrier.
template <int T, int STAGES, bool IS_RESET, bool OWNROW>
__global__ void __launch_bounds__(T) snake_kernel(const Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_stats[5];  // per-CTA, per-launch: n_episodes, sum_return, sum_length, sum_score, max_score

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int G = a.G;
    const int cells = G * G;
    const int tile_bytes = T * cells;  // multiple of 16 because T is
    const long long n_tiles = (a.n_envs + T - 1) / T;

    if (tid < 5) s_stats[tid] = (tid == 4) ? INT_MIN : 0;
    // PDL: let the next step's grid start becoming resident as this one drains, then wait for the PREVIOUS step's
    // grid to have completed and flushed before touching anything it wrote (state, counters).
    pdl_launch_dependents();
    pdl_wait();
    if (blockIdx.x == 0 && tid == 0 && a.io.done_count_next) *a.io.done_count_next = 0;  // for the NEXT step
    __syncthreads();

    // deferred finished-env list entries of the previous tile (the global atomic's result is consumed one
    // iteration later so that its latency is off the critical path)
    unsigned pend_mask = 0, pend_base = 0;
    int pend_env = 0;

    Loaded cur = load_inputs<IS_RESET>(a, (long long)blockIdx.x * T + tid);
    // tail ring entry of this thread's env, fetched one tile ahead (-1: not fetched / not needed)
    int pre_tail = -1;
    if constexpr (!IS_RESET) {
        const long long env0 = (long long)blockIdx.x * T + tid;
        const Core c0 = unpack(cur.raw);
        if (env0 < a.n_envs && c0.length > 1) pre_tail = a.ring[env0 * cells + tail_index(c0, cells)];
    }
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        uint8_t *buf = smem + (size_t)(it % STAGES) * tile_bytes;
        const long long env = tile * T + tid;
        const bool active = env < a.n_envs;

        // Prefetch the next tile's state + action into registers; consumed next iteration.
        const long long tile_next = tile + gridDim.x;
        Loaded nxt = cur;
        if (tile_next < n_tiles) nxt = load_inputs<IS_RESET>(a, tile_next * T + tid);

        if constexpr (!IS_RESET) {
            if (pend_mask) {  // warp-uniform
                const unsigned base = __shfl_sync(0xFFFFFFFFu, pend_base, 0);
                if (pend_mask & (1u << lane)) a.io.done_env[base + __popc(pend_mask & ((1u << lane) - 1u))] = pend_env;
                pend_mask = 0;
            }
        }

        // The bulk copy that last used this buffer (tile it - STAGES) must have finished reading it.
        if (it >= STAGES) {
            if (tid == 0) bulk_wait_read<STAGES - 1>();
            __syncthreads();
        }
        uint8_t *row = buf + tid * cells;
        if constexpr (OWNROW) {
            uint4 *r16 = reinterpret_cast<uint4 *>(row);
            for (int i = 0; i < (cells >> 4); ++i) r16[i] = make_uint4(0, 0, 0, 0);
        } else {
            uint4 *buf16 = reinterpret_cast<uint4 *>(buf);
            for (int i = tid; i < (tile_bytes >> 4); i += T) buf16[i] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }

        // ---------------------------------------------------------------- per-env dynamics (smem + ring only)
        bool ended = false;
        int ep_ret = 0, ep_len = 0, ep_score = 0;
        float rew = 0.0f;
        int term = 0, trunc = 0;
        bool invalid = false;
        Core s = unpack(cur.raw);
        if (active) {
            uint16_t *ring = a.ring + env * cells;
            if constexpr (IS_RESET) {
                if (cur.selected && a.first_call) s.ctr = 0;
            }
            EnvStream rng(a.seed, a.env_id_base + (uint64_t)env, BENG_STREAM_ENV, s.ctr);

            if constexpr (IS_RESET) {
                if (cur.selected) reset_env(s, rng, row, G);
                else draw_body(s, ring, row, cells, G);
            } else {
                const long long act = cur.act;
                if (a.mode == BENG_AUTORESET_NEXT_STEP && (s.flags & FLAG_NEEDS_RESET)) {
                    reset_env(s, rng, row, G);  // action ignored, reward 0, not terminated
                } else if (act < 0 || act > 3) {
                    // The reference raises ValueError (snake_env.py:69-70); here the env is left
                    // untouched and the event is counted for the host to raise on.
                    invalid = true;
                    draw_body(s, ring, row, cells, G);
                } else {
                    const int action = (int)act;
                    int d = action - s.dir;
                    if (d < 0) d = -d;
                    if (d != 2) s.dir = action;  // reversal guard, snake_env.py:73-74
                    // 0 up, 1 right, 2 down, 3 left in (row, col), snake_env.py:77-85
                    const int nr = s.head_r + (s.dir == 2) - (s.dir == 0);
                    const int nc = s.head_c + (s.dir == 1) - (s.dir == 3);
                    const bool wall = (unsigned)nr >= (unsigned)G || (unsigned)nc >= (unsigned)G;  // :88-90
                    const int new_cell = wall ? -1 : nr * G + nc;
                    const int head_cell = s.head_r * G + s.head_c;
                    const int food_cell = (s.food_r == FOOD_NONE) ? -2 : s.food_r * G + s.food_c;

                    // A wall death under SAME_STEP returns the reset observation: the old body is
                    // never drawn, so the row stays clean for reset_env below.
                    const bool draw = !(wall && a.mode == BENG_AUTORESET_SAME_STEP);
                    bool self_hit = false;
                    int tail_cell = head_cell;
                    if (draw) {
                        if (s.length == 1) row[head_cell] = 1;  // (a length-1 snake cannot hit itself)
                        else tail_cell = scan_body(s, ring, row, cells, G, 1, new_cell, self_hit, pre_tail);  // :93, tail included
                    }
                    const bool died = wall || self_hit;
                    if (died) {
                        rew = -10.0f;  // :90 / :94 -- nothing but `direction` was mutated
                        term = 1;
                    } else {
                        const bool eat = new_cell == food_cell;
                        // insert(0, new_head), :97.  The ring is only maintained for length >= 2: a length-1
                        // body IS the head held in the state record, which spares 93 % of the envs a scattered
                        // 2-byte store (a partial-sector write costs a DRAM read-modify-write).
                        if (s.length > 1 || eat) {
                            if (s.length == 1) ring[s.ring_head] = (uint16_t)head_cell;
                            if (++s.ring_head == cells) s.ring_head = 0;
                            ring[s.ring_head] = (uint16_t)new_cell;
                        }
                        row[new_cell] = 1;
                        s.head_r = nr;
                        s.head_c = nc;
                        if (eat) {  // :101-104
                            s.length += 1;
                            rew = 10.0f;
                            place_food(s, rng, row, G);
                        } else {
                            row[tail_cell] = 0;  // pop(), :107
                        }
                        s.steps = min(s.steps + 1, 65535);     // :109
                        if (s.steps >= a.max_steps) {
                            term = 1;              // :112-114, reported as terminated
                            trunc = a.tl_trunc;    // gym.make's TimeLimit would add truncated=True here
                        }
                    }
                    if (term && a.mode != BENG_AUTORESET_DISABLED) {
                        ended = true;
                        ep_score = s.length - 1;
                        ep_ret = 10 * ep_score - (died ? 10 : 0);
                        ep_len = s.steps + (died ? 1 : 0);  // the death step does not bump `steps`
                        if (a.mode == BENG_AUTORESET_SAME_STEP) {
                            if (draw) {  // un-draw the body so the row is clean again
                                // Self-hit death: the body is unchanged.  Time limit: the body is the moved one
                                // (new head in the state record, everything older in the ring).
                                bool hit = false;
                                if (s.length == 1) row[s.head_r * G + s.head_c] = 0;
                                else scan_body(s, ring, row, cells, G, 0, -1, hit);
                            }
                            reset_env(s, rng, row, G);
                        } else {
                            s.flags |= FLAG_NEEDS_RESET;
                        }
                    }
                }
            }
            if (s.food_r != FOOD_NONE) row[s.food_r * G + s.food_c] = 2;  // food written last, :139-141
            s.ctr = rng.ctr;
        }

        // ---------------------------------------------------------------- drain the tile
        fence_proxy_async_smem();  // this thread's generic-proxy smem writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            const long long first = tile * T;
            const long long n_here = min((long long)T, a.n_envs - first);
            const uint32_t bytes = (uint32_t)(n_here * cells);
            const uint32_t bulk = bytes & ~15u;
            if (bulk) bulk_store_s2g(a.io.obs + first * cells, buf, bulk);
            bulk_commit();
            for (uint32_t i = bulk; i < bytes; ++i) a.io.obs[first * cells + i] = (int8_t)buf[i];  // ragged last tile
        }

        // ---------------------------------------------------------------- per-env outputs (off the critical path)
        if (active) {
            a.core[env] = pack(s);
            if constexpr (!IS_RESET) {
                a.io.reward[env] = rew;
                a.io.terminated[env] = (uint8_t)term;
                if (a.io.truncated) a.io.truncated[env] = (uint8_t)trunc;  // raw class: never truncates, :119
                if (invalid && a.io.invalid_count) atomicAdd(a.io.invalid_count, 1);
            }
            if (a.io.score) a.io.score[env] = s.length - 1;
            if (a.io.snake_length) a.io.snake_length[env] = s.length;
            if (ended) {
                if (a.io.ep_return) a.io.ep_return[env] = (float)ep_ret;
                if (a.io.ep_length) a.io.ep_length[env] = ep_len;
                if (a.io.ep_score) a.io.ep_score[env] = ep_score;
            }
        }

        // ---------------------------------------------------------------- episode-done compaction + stats
        if constexpr (!IS_RESET) {
            const unsigned done_mask = __ballot_sync(0xFFFFFFFFu, ended);
            if (done_mask) {
                if (a.io.done_env) {
                    // one global atomic per warp reserves the slots; the slots are written next iteration
                    if (lane == 0) pend_base = atomicAdd(a.io.done_count, (unsigned)__popc(done_mask));
                    pend_mask = done_mask;
                    pend_env = (int)env;
                }
                if (a.io.stats) {
                    const int sr = __reduce_add_sync(0xFFFFFFFFu, ended ? ep_ret : 0);
                    const int sl = __reduce_add_sync(0xFFFFFFFFu, ended ? ep_len : 0);
                    const int ss = __reduce_add_sync(0xFFFFFFFFu, ended ? ep_score : 0);
                    const int mx = __reduce_max_sync(0xFFFFFFFFu, ended ? ep_score : INT_MIN);
                    if (lane == 0) {
                        atomicAdd(&s_stats[0], __popc(done_mask));
                        atomicAdd(&s_stats[1], sr);
                        atomicAdd(&s_stats[2], sl);
                        atomicAdd(&s_stats[3], ss);
                        atomicMax(&s_stats[4], mx);
                    }
                }
            }
        }
        // Fetch the next tile's tail ring entry now (its state record, requested at the top of this iteration,
        // has arrived by now) so that the dependent ring load is not exposed at the start of the next tile.
        if constexpr (!IS_RESET) {
            pre_tail = -1;
            if (tile_next < n_tiles) {
                const long long env_n = tile_next * T + tid;
                const Core cn = unpack(nxt.raw);
                if (env_n < a.n_envs && cn.length > 1) pre_tail = a.ring[env_n * cells + tail_index(cn, cells)];
            }
        }
        cur = nxt;
    }

    if constexpr (!IS_RESET) {
        if (pend_mask) {
            const unsigned base = __shfl_sync(0xFFFFFFFFu, pend_base, 0);
            if (pend_mask & (1u << lane)) a.io.done_env[base + __popc(pend_mask & ((1u << lane) - 1u))] = pend_env;
        }
    }
    if (tid == 0) bulk_wait_read<0>();  // the copy engine has read every tile before the CTA's shared memory retires
    if constexpr (!IS_RESET) {
        if (a.io.stats) {
            __syncthreads();
            if (tid < 4) {
                if (s_stats[tid] != 0)
                    atomicAdd((unsigned long long *)&a.io.stats[tid], (unsigned long long)(long long)s_stats[tid]);
            } else if (tid == 4) {
                if (s_stats[4] != INT_MIN) atomicMax((long long *)&a.io.stats[4], (long long)s_stats[4]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) snake_export_kernel(const uint4 *__restrict__ core,
                                                           const uint16_t *__restrict__ ring, long long n_envs, int G,
                                                           int *head_r, int *head_c, int *food_r, int *food_c, int *dir,
                                                           int *steps, int *length, uint32_t *ctr, int *body) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    const Core s = unpack(core[e]);
    if (head_r) head_r[e] = s.head_r;
    if (head_c) head_c[e] = s.head_c;
    if (food_r) food_r[e] = (s.food_r == FOOD_NONE) ? -1 : s.food_r;
    if (food_c) food_c[e] = (s.food_c == FOOD_NONE) ? -1 : s.food_c;
    if (dir) dir[e] = s.dir;
    if (steps) steps[e] = s.steps;
    if (length) length[e] = s.length;
    if (ctr) ctr[e] = s.ctr;
    if (body) {
        const int cells = G * G;
        int idx = s.ring_head;
        for (int i = 0; i < cells; ++i) {
            int c = -1;
            if (i < s.length) c = (s.length == 1) ? s.head_r * G + s.head_c : (int)ring[e * cells + idx];
            body[e * cells + i] = c;
            if (--idx < 0) idx = cells - 1;
        }
    }
}

// ---- launch configuration ---------------------------------------------------------------------
struct Config {
    int tile;      // envs (= threads) per CTA tile
    int stages;    // tile buffers per CTA
    int ctas_per_sm;
};

// Measured on B200 at 1M envs, G = 20 (profiles/snake_r1_sweep.txt): every single-stage shape that keeps
// >= 2 CTAs resident per SM reaches the same ~11.5 G env-steps/s (the kernel is DRAM-bound), two-stage shapes
// with fewer CTAs are ~15 % slower.  So: one tile buffer per CTA, as many CTAs per SM as shared memory
// allows, and the largest tile that still gives every SM >= 2 tiles of work.
// BENG_SNAKE_CFG="tile,stages,ctas_per_sm" overrides (read once; profiling sweeps).
Config pick_config(int G, long long n_envs) {
    static int env_tile = -1, env_stages = 0, env_ctas = 0;
    if (env_tile < 0) {
        env_tile = 0;
        if (const char *e = getenv("BENG_SNAKE_CFG")) {
            int t = 0, s = 0, c = 0;
            if (sscanf(e, "%d,%d,%d", &t, &s, &c) == 3) { env_tile = t; env_stages = s; env_ctas = c; }
        }
    }
    const long long cells = (long long)G * G;
    const long long budget = 200 * 1024;  // of 227 KB: leaves room for static smem + 1 KB/CTA reservation
    if (env_tile > 0) return Config{env_tile, env_stages, env_ctas};
    Config c{256, 1, 2};
    const long long sms = device_sm_count();
    while (c.tile > 32 && (c.tile * cells * 2 > budget || (n_envs + c.tile - 1) / c.tile < 2 * sms)) c.tile >>= 1;
    c.ctas_per_sm = (int)(budget / (c.tile * cells));
    if (c.ctas_per_sm < 1) c.ctas_per_sm = 1;
    if (c.ctas_per_sm > 8) c.ctas_per_sm = 8;
    return c;
}

template <int T, int STAGES, bool IS_RESET, bool OWNROW>
int launch_one(const Args &a, const Config &c, cudaStream_t stream) {
    const size_t smem = (size_t)T * a.G * a.G * STAGES;
    auto kern = snake_kernel<T, STAGES, IS_RESET, OWNROW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long n_tiles = (a.n_envs + T - 1) / T;
    long long grid = (long long)device_sm_count() * c.ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    e = launch_pdl(kern, dim3((unsigned)grid), dim3(T), smem, stream, a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)e;
}

template <bool IS_RESET>
int launch(const Args &a, cudaStream_t stream) {
    const Config c = pick_config(a.G, a.n_envs);
    const int cells = a.G * a.G;
    const bool ownrow = (cells % 16 == 0) && ((cells / 16) % 2 == 1) && !getenv("BENG_SNAKE_COOPZERO");
#define BENG_CASE(TT, SS)                                                                   \
    if (c.tile == TT && c.stages == SS)                                                     \
        return ownrow ? launch_one<TT, SS, IS_RESET, true>(a, c, stream)                    \
                      : launch_one<TT, SS, IS_RESET, false>(a, c, stream);
    BENG_CASE(256, 1) BENG_CASE(256, 2) BENG_CASE(128, 1) BENG_CASE(128, 2) BENG_CASE(128, 3) BENG_CASE(64, 1)
    BENG_CASE(224, 1) BENG_CASE(192, 1) BENG_CASE(160, 1) BENG_CASE(96, 1)  /* non-power-of-two tiles (round-2 sweep) */
    BENG_CASE(64, 2) BENG_CASE(64, 3) BENG_CASE(64, 4) BENG_CASE(32, 1) BENG_CASE(32, 2) BENG_CASE(32, 4)
#undef BENG_CASE
    return BENG_ERR_UNSUPPORTED;
}

int check_common(const beng_snake_params *p, const beng_snake_state *st, const beng_snake_io *io, int64_t n_envs) {
    if (!p || !st || !io || !st->core || !st->ring || !io->obs || n_envs < 0) return BENG_ERR_BAD_ARG;
    if (p->grid_size < 2 || p->grid_size > 64) return BENG_ERR_UNSUPPORTED;
    if (p->max_steps < 1 || p->max_steps > 65535) return BENG_ERR_UNSUPPORTED;
    if (p->autoreset_mode < 0 || p->autoreset_mode > 2) return BENG_ERR_BAD_ARG;
    if (((uintptr_t)st->core & 15) || ((uintptr_t)io->obs & 15)) return BENG_ERR_BAD_ARG;
    if (io->done_env && !io->done_count) return BENG_ERR_BAD_ARG;
    return 0;
}

Args make_args(const beng_snake_params *p, const beng_snake_state *st, const beng_snake_io *io, int64_t n_envs) {
    Args a{};
    a.core = (uint4 *)st->core;
    a.ring = st->ring;
    a.io = *io;
    a.n_envs = n_envs;
    a.seed = p->seed;
    a.env_id_base = p->env_id_base;
    a.G = p->grid_size;
    a.max_steps = p->max_steps;
    a.mode = p->autoreset_mode;
    a.tl_trunc = p->time_limit_truncation;
    return a;
}

}  // namespace
}  // namespace beng

extern "C" {

int beng_snake_launch_config(int32_t grid_size, int64_t n_envs, int32_t *tile, int32_t *stages, int32_t *ctas_per_sm) {
    if (grid_size < 2 || grid_size > 64 || n_envs < 0) return BENG_ERR_BAD_ARG;
    const beng::Config c = beng::pick_config(grid_size, n_envs);
    if (tile) *tile = c.tile;
    if (stages) *stages = c.stages;
    if (ctas_per_sm) *ctas_per_sm = c.ctas_per_sm;
    return 0;
}

size_t beng_snake_core_bytes(int64_t n_envs) { return n_envs < 0 ? 0 : (size_t)n_envs * 16; }

size_t beng_snake_ring_bytes(int64_t n_envs, int32_t grid_size) {
    return (n_envs < 0 || grid_size < 1) ? 0 : (size_t)n_envs * (size_t)grid_size * grid_size * sizeof(uint16_t);
}

int beng_snake_reset(const beng_snake_params *p, const beng_snake_state *st, const beng_snake_io *io,
                     const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream) {
    if (int rc = beng::check_common(p, st, io, n_envs)) return rc;
    if (n_envs == 0) return 0;
    beng::Args a = beng::make_args(p, st, io, n_envs);
    a.mask = mask_dev;
    a.first_call = first_call;
    return beng::launch<true>(a, (cudaStream_t)stream);
}

int beng_snake_step(const beng_snake_params *p, const beng_snake_state *st, const int64_t *actions_dev,
                    const beng_snake_io *io, int64_t n_envs, void *stream) {
    if (int rc = beng::check_common(p, st, io, n_envs)) return rc;
    if (!actions_dev || !io->reward || !io->terminated) return BENG_ERR_BAD_ARG;
    if ((uintptr_t)actions_dev & 7) return BENG_ERR_BAD_ARG;  // 64-bit action loads (int64, or float32 pairs)
    if (n_envs == 0) return 0;
    beng::Args a = beng::make_args(p, st, io, n_envs);
    a.actions = (const long long *)actions_dev;
    return beng::launch<false>(a, (cudaStream_t)stream);
}

int beng_snake_step_host(const beng_snake_params *p, const beng_snake_state *st, int64_t *actions_dev,
                         const beng_snake_io *io, int64_t n_envs, const int64_t *actions_host, int8_t *obs_host,
                         float *reward_host, uint8_t *terminated_host, uint8_t *truncated_host, int32_t *score_host,
                         int32_t *snake_length_host, void *stream) {
    if (!actions_host || !actions_dev) return BENG_ERR_BAD_ARG;
    if (int rc = beng::check_common(p, st, io, n_envs)) return rc;
    if ((truncated_host && !io->truncated) || (score_host && !io->score) || (snake_length_host && !io->snake_length))
        return BENG_ERR_BAD_ARG;
    if (n_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)n_envs, cells = (size_t)p->grid_size * p->grid_size;
    cudaError_t e = cudaMemcpyAsync(actions_dev, actions_host, n * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    if (int rc = beng_snake_step(p, st, actions_dev, io, n_envs, stream)) return rc;
#define BENG_D2H(dst, src, bytes) \
    if (dst) { e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s); if (e != cudaSuccess) return (int)e; }
    BENG_D2H(reward_host, io->reward, n * sizeof(float))
    BENG_D2H(terminated_host, io->terminated, n)
    BENG_D2H(truncated_host, io->truncated, n)
    BENG_D2H(score_host, io->score, n * sizeof(int32_t))
    BENG_D2H(snake_length_host, io->snake_length, n * sizeof(int32_t))
    BENG_D2H(obs_host, io->obs, n * cells)
#undef BENG_D2H
    return 0;
}

int beng_snake_export_state(const beng_snake_params *p, const beng_snake_state *st, int64_t n_envs, int32_t *head_r,
                            int32_t *head_c, int32_t *food_r, int32_t *food_c, int32_t *direction, int32_t *steps,
                            int32_t *length, uint32_t *rng_counter, int32_t *body_cells_dev, void *stream) {
    if (!p || !st || !st->core || !st->ring || n_envs < 0) return BENG_ERR_BAD_ARG;
    if (n_envs == 0) return 0;
    const unsigned blocks = (unsigned)((n_envs + 255) / 256);
    beng::snake_export_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const uint4 *)st->core, st->ring, (long long)n_envs, p->grid_size, head_r, head_c, food_r, food_c, direction,
        steps, length, rng_counter, body_cells_dev);
    return beng::finish_launch();
}

}  // extern "C"

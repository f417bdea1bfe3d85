/* beng.h -- C ABI of libbeng.so, the B200 (sm_100a) batched environment engine.
 *
 * This is the drop-in boundary for the step()/reset() hot path of the reference's Gymnasium
 * environments.  The reference is pure Python and has no FFI of its own;
 * the interface each entry point replaces is therefore a METHOD of the reference class, cited
 * per function below (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - return int = cudaError_t (0 == success) or BENG_ERR_* (negative) for argument errors.
 *   - never allocate device memory, never synchronise, never throw: kernels are launched
 *     asynchronously on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - every `*_dev` / state / io pointer is a DEVICE pointer to caller-owned memory, 16-byte
 *     aligned, alive until the stream has drained.  The *_host entry points additionally take
 *     (pinned) HOST buffers and enqueue the H2D/D2H copies around the kernel on the same stream.
 *   - one host thread per GPU/process; env ids are GLOBAL (env_id_base + local index) so a given
 *     env has the same trajectory however the batch is sharded over GPUs.
 */
#ifndef BENG_H
#define BENG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BENG_VERSION 100 /* major*10000 + minor*100 + patch */

enum {
    BENG_ERR_BAD_ARG = -1,      /* NULL where a pointer is required, n_envs < 0, ... */
    BENG_ERR_UNSUPPORTED = -2,  /* e.g. grid_size outside [2, 64] */
};

/* Auto-reset modes of the batched classes (gymnasium.vector.AutoresetMode names).  The reference
 * has no auto-reset: DISABLED is exactly the reference class;
 * SAME_STEP is the reference's caller loop `if terminated: env.reset()` folded into the step. */
enum {
    BENG_AUTORESET_DISABLED = 0,
    BENG_AUTORESET_NEXT_STEP = 1,
    BENG_AUTORESET_SAME_STEP = 2,
};

/* RNG streams of the counter-based generator (Philox4x32-10, csrc/beng_rng.cuh). */
enum { BENG_STREAM_ENV = 0, BENG_STREAM_ACTION = 1 };

/* Library version (BENG_VERSION of the build). */
int beng_version(void);

/* SM architecture the library was compiled for (100 for sm_100a). */
int beng_compiled_arch(void);

/* Number of kernel launches this process has issued through the library since load
 * (bench.py reports it as `gpu_launches`). */
uint64_t beng_launch_count(void);

/* Synthetic uniform action tape on the device (bench / tests; stands in for a policy).
 *   actions_dev[i * n_cols + c] = randint(0, n_choices - 1) drawn from u32 number
 *   (step_index * n_cols + c) of stream BENG_STREAM_ACTION of env (env_id_base + i). */
int beng_fill_random_actions(int64_t *actions_dev, int64_t n_envs, int32_t n_cols, int32_t n_choices,
                             uint32_t step_index, uint64_t env_id_base, uint64_t seed, void *stream);

/* ------------------------------------------------------------------------------------------
 * snake_env_classic  (reference: snake_env_classic/snake_env.py, class SnakeEnvClassic)
 * ------------------------------------------------------------------------------------------ */

/* Constructor arguments of SnakeEnvClassic (snake_env.py:19-47) plus the batching parameters. */
typedef struct beng_snake_params {
    int32_t grid_size;      /* `grid_size` ctor kwarg, default 20 (snake_env.py:19); 2..64 */
    int32_t max_steps;      /* `self.max_steps = 1000` (snake_env.py:47); 1..65535 */
    int32_t autoreset_mode; /* BENG_AUTORESET_* */
    int32_t time_limit_truncation; /* 0: `truncated` is always 0, like the raw reference class (snake_env.py:119).
                                      1: ALSO set truncated on the step that reaches max_steps -- what gymnasium's
                                      TimeLimit wrapper adds when the env is built with gym.make
                                      (snake_env_classic/__init__.py:3-7, max_episode_steps=1000) */
    uint64_t seed;          /* key of the counter-based stream */
    uint64_t env_id_base;   /* global id of local env 0 */
} beng_snake_params;

/* Per-env state, structure-of-arrays in HBM.
 *   core : n_envs x 16 B record {u8 head_r, head_c, food_r, food_c; u8 direction; u8 flags;
 *          u16 length; u16 steps; u16 ring_head; u32 rng_counter}  -- one 128-bit load per env.
 *          (`score` is not stored: score == length - 1 always, snake_env.py:57,102 vs :97,107.)
 *   ring : n_envs x (grid_size^2) u16 body cells (r * G + c); newest at ring_head, oldest
 *          (the tail) at ring_head - length + 1 (mod G^2).  Replaces the Python list
 *          `self.snake` (snake_env.py:54,97,107). */
typedef struct beng_snake_state {
    void *core;     /* uint4[n_envs] */
    uint16_t *ring; /* uint16_t[n_envs * grid_size * grid_size] */
} beng_snake_state;

/* Outputs of one batched step.  `obs`, `reward`, `terminated` are required; the rest may be NULL. */
typedef struct beng_snake_io {
    int8_t *obs;            /* [n_envs, G, G] 0 empty / 1 snake / 2 food  (_get_observation, snake_env.py:131-143) */
    float *reward;          /* [n_envs] -10 / 0 / +10                       (snake_env.py:90,94,100,103) */
    uint8_t *terminated;    /* [n_envs] death or steps >= max_steps        (snake_env.py:90,94,112-114) */
    uint8_t *truncated;     /* [n_envs] always 0                           (snake_env.py:119) */
    int32_t *score;         /* [n_envs] info["score"] after the step       (snake_env.py:117) */
    int32_t *snake_length;  /* [n_envs] info["snake_length"]               (snake_env.py:117) */
    /* written only for envs whose episode ended in this step (auto-reset modes): */
    float *ep_return;       /* [n_envs] 10 * score - 10 * died */
    int32_t *ep_length;     /* [n_envs] env steps in the episode (death step included) */
    int32_t *ep_score;      /* [n_envs] */
    /* warp-compacted list of the envs that finished this step (order unspecified): */
    uint32_t *done_count;   /* [1]  must be zero when the step starts (see done_count_next) */
    int32_t *done_env;      /* [n_envs] LOCAL env indices */
    uint32_t *done_count_next; /* [1] nullable; zeroed BY this launch so it can serve as `done_count` of the
                                  next step (ping-pong two counters: no memset launch between steps) */
    /* running episode statistics, integer (exact, order independent):
     * {n_episodes, sum_return, sum_length, sum_score, max_score} */
    int64_t *stats;         /* [5] accumulated across steps; caller zeroes / all-reduces */
    int32_t *invalid_count; /* [1] number of out-of-space actions seen (reference raises ValueError,
                                   snake_env.py:69-70; here such an env is left untouched) */
} beng_snake_io;

/* Launch shape the step kernel uses for (grid_size, n_envs) on the current device: envs (= threads) per CTA
 * tile, tile buffers per CTA, resident CTAs per SM.  Informational (bench.py names the kernel with it). */
int beng_snake_launch_config(int32_t grid_size, int64_t n_envs, int32_t *tile, int32_t *stages, int32_t *ctas_per_sm);

/* Bytes the caller must allocate. */
size_t beng_snake_core_bytes(int64_t n_envs);
size_t beng_snake_ring_bytes(int64_t n_envs, int32_t grid_size);

/* SnakeEnvClassic.reset (snake_env.py:49-65) for every env with mask[i] != 0 (mask NULL = all),
 * then the observation / info of EVERY env is written to io (io->obs required; reward etc. untouched).
 * `first_call` != 0 additionally zeroes rng_counter of the selected envs (constructor semantics). */
int beng_snake_reset(const beng_snake_params *p, const beng_snake_state *st, const beng_snake_io *io,
                     const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream);

/* SnakeEnvClassic.step (snake_env.py:67-119) + _place_food (:121-129) + _get_observation (:131-143)
 * + the auto-reset named in p->autoreset_mode, for all n_envs envs, in ONE kernel launch. */
int beng_snake_step(const beng_snake_params *p, const beng_snake_state *st, const int64_t *actions_dev,
                    const beng_snake_io *io, int64_t n_envs, void *stream);

/* Same step with HOST action / result buffers (what a caller of the reference's numpy API holds):
 * enqueues H2D(actions) -> kernel -> D2H(obs, reward, terminated[, truncated, score, snake_length])
 * on `stream`.  Host pointers that are NULL are skipped (obs_host NULL keeps observations in HBM).
 * Does not synchronise; the caller waits on the stream before reading the host buffers. */
int beng_snake_step_host(const beng_snake_params *p, const beng_snake_state *st, int64_t *actions_dev,
                         const beng_snake_io *io, int64_t n_envs, const int64_t *actions_host, int8_t *obs_host,
                         float *reward_host, uint8_t *terminated_host, uint8_t *truncated_host, int32_t *score_host,
                         int32_t *snake_length_host, void *stream);

/* Unpack the SoA state into plain int32 arrays (tests, checkpoints, debugging).  Any output may be NULL.
 * body_cells_dev, when given, is [n_envs, G*G] int32 filled head-first with r*G+c and -1 padding. */
int beng_snake_export_state(const beng_snake_params *p, const beng_snake_state *st, int64_t n_envs, int32_t *head_r,
                            int32_t *head_c, int32_t *food_r, int32_t *food_c, int32_t *direction, int32_t *steps,
                            int32_t *length, uint32_t *rng_counter, int32_t *body_cells_dev, void *stream);

/* ------------------------------------------------------------------------------------------
 * crypto_trading_env  (reference: crypto_trading_env/crypto_trading_env.py, class CryptoTradingEnv)
 * ------------------------------------------------------------------------------------------ */

#define BENG_CRYPTO_HISTORY 50  /* TradingConfig.history_length (crypto_trading_env.py:33); fixed */
#define BENG_CRYPTO_OBS_DIM 261 /* 50*5 + 3 + 8: what _get_observation really returns (:505-561); the declared
                                   space says 260 (:285-286) -- SURVEY.md section 0 fact 7 */

/* TradingConfig (crypto_trading_env.py:28-38) + episode limit (:298) + batching parameters. */
typedef struct beng_crypto_params {
    double initial_balance;          /* 10000.0 */
    double trading_fee_rate;         /* 0.001 */
    double slippage_rate;            /* 0.0005 */
    double min_price;                /* 100.0 */
    double max_price;                /* 100000.0 */
    double volatility_base;          /* 0.02 */
    double market_psychology_factor; /* 0.1 */
    int32_t max_steps;               /* 1000 (:298) */
    int32_t autoreset_mode;          /* BENG_AUTORESET_* */
    int32_t action_type;             /* 0 = Discrete(5) int64[n]; 1 = Box(-1,1,(2,)) float32[n][2]  (:288-296) */
    int32_t window_head;             /* slot (0..49) holding the NEWEST candle of every env before this call.
                                        A step writes the new candle to (window_head+1)%50 -- the caller then
                                        advances window_head by one; a reset leaves it unchanged. */
    int32_t time_limit_truncation;   /* 1: also set `truncated` when current_step reaches max_steps (gym.make's
                                        TimeLimit, crypto_trading_env.py:739-743); 0: never (raw class, :388) */
    int32_t reserved;
    uint64_t seed;
    uint64_t env_id_base;
} beng_crypto_params;

/* Per-env state, structure-of-arrays in HBM, env index fastest (coalesced for one thread per env).
 * All money/price arithmetic is float64 like the reference; open/high/low/volume are observation-only
 * features and are stored as float32. */
typedef struct beng_crypto_state {
    double *scal;      /* [4][n]  cash, holdings, trend_strength, market_psychology */
    uint32_t *meta;    /* [2][n]  {step | regime << 16 | flags << 24}, rng_counter */
    double *ep_return; /* [n]     running episode return */
    double *close;     /* [50][pitch] close prices, ring over slots (see window_head); pitch =
                          beng_crypto_window_pitch(n) = n rounded up to a multiple of 32, so that every 32-env unit of
                          every slot is one aligned box for the tensor-map (TMA) loads; 128-byte aligned */
    float *ohlv;       /* [50][pitch][4] open, high, low, volume: one 16-byte record per (slot, env); 128-byte aligned */
} beng_crypto_state;

typedef struct beng_crypto_io {
    float *obs;              /* [n][261]  _get_observation (:505-561) */
    float *reward;           /* [n]       (:440-445) cast to float32 */
    uint8_t *terminated;     /* [n]       step >= max_steps or value <= 0 or value >= 10 * initial (:382-386) */
    uint8_t *truncated;      /* [n]       always 0 (:388) */
    /* info dict (:390-398); all nullable */
    double *reward64;        /* [n] the float64 reward */
    double *portfolio_value; /* [n] */
    double *current_price;   /* [n] */
    uint8_t *trade_kind;     /* [n] 0 = no trade, 1 = buy, 2 = sell (info["trade_info"]["action"]) */
    /* written only for envs whose episode ended in this step (auto-reset modes); nullable */
    double *ep_return_out;   /* [n] */
    int32_t *ep_length;      /* [n] */
    double *stats;           /* [4] running {n_episodes, sum_return, sum_length, sum_final_value}; nullable */
} beng_crypto_io;

/* Row pitch (in envs) of beng_crypto_state.close / .ohlv for a batch of n_envs. */
int64_t beng_crypto_window_pitch(int64_t n_envs);

/* CryptoTradingEnv.reset (:301-340) for envs with mask[i] != 0 (NULL = all): 50 warm-up candles from 50000.0,
 * cash/holdings/step reset, the MarketSimulator state is NOT reset.  first_call != 0 is the
 * constructor: market state = (SIDEWAYS, 0.0, 0.5) (:125-130) and rng_counter = 0 for the selected envs.
 * The observation of EVERY env is written to io->obs. */
int beng_crypto_reset(const beng_crypto_params *p, const beng_crypto_state *st, const beng_crypto_io *io,
                      const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream);

/* CryptoTradingEnv.step (:342-398) + _execute_action/_execute_buy/_execute_sell (:400-503) +
 * MarketSimulator.generate_next_price (:132-221) + _get_observation with all TechnicalIndicators (:41-119,
 * :505-561) + auto-reset, for all envs, in ONE kernel launch (persistent CTAs: per-env float64 dynamics, then the
 * window is streamed once and the indicators are accumulated on the way).  `actions_dev` is int64[n] or
 * float32[n][2] according to p->action_type.  n_envs must be below 2^26 (BENG_ERR_UNSUPPORTED otherwise). */
int beng_crypto_step(const beng_crypto_params *p, const beng_crypto_state *st, const void *actions_dev,
                     const beng_crypto_io *io, int64_t n_envs, void *stream);

/* Same step with HOST action / result buffers: H2D(actions) -> kernel -> D2H(obs, reward, terminated) on `stream`.
 * NULL host outputs are skipped.  Does not synchronise. */
int beng_crypto_step_host(const beng_crypto_params *p, const beng_crypto_state *st, void *actions_dev,
                          const beng_crypto_io *io, int64_t n_envs, const void *actions_host, float *obs_host,
                          float *reward_host, uint8_t *terminated_host, uint8_t *truncated_host, void *stream);

/* ------------------------------------------------------------------------------------------
 * traffic_management_env  (reference: traffic_management_env/{environment,utils,config}.py)
 * ------------------------------------------------------------------------------------------ */

#define BENG_TRAFFIC_MAX_INTERSECTIONS 25

/* Constructor arguments of TrafficManagementEnv (environment.py:61-66, defaults config.py:6-12) + config.py
 * MAX_TIMESTEPS + batching parameters.  The remaining config.py constants (phase durations 5/30/3, reward weights
 * 1.0/-0.1/-0.05/0.5, observation caps 20/100/1000/50) are compiled in. */
typedef struct beng_traffic_params {
    int32_t grid_rows, grid_cols; /* grid_size, default (5, 5) */
    int32_t num_intersections;    /* default 9; effective value = min(this, rows*cols) (environment.py:82); <= 25 */
    int32_t max_vehicles;         /* default 50; <= 255 */
    double spawn_rate;            /* default 0.3 */
    int32_t max_timesteps;        /* config.py:23 MAX_TIMESTEPS = 1000 */
    int32_t autoreset_mode;       /* BENG_AUTORESET_* */
    int32_t time_limit_truncation; /* 1: also set `truncated` at max_timesteps (gym.make's TimeLimit,
                                      traffic_management_env/__init__.py:7-17); 0: never (raw class, environment.py:197) */
    int32_t reserved;
    uint64_t seed;
    uint64_t env_id_base;
} beng_traffic_params;

/* Per-env state, [field][env] arrays (env index fastest).  Vehicles never move in the reference (SURVEY.md
 * section 0 fact 9), so a queue is fully described by (length, sum of waiting times, number of vehicles whose
 * route ends where it started) and the env by the length of `self.vehicles`. */
typedef struct beng_traffic_state {
    uint16_t *light;      /* [ni][n]    phase (low byte: 0 NS_GREEN 1 NS_YELLOW 2 EW_GREEN 3 EW_YELLOW) | timer << 8 */
    int32_t *passed;      /* [ni][n]    Intersection.vehicles_passed */
    int32_t *waiting;     /* [ni][n]    Intersection.total_waiting_time */
    uint32_t *qmeta;      /* [ni*2][n]  row 2i: queue lengths of intersection i, one byte per direction N, E, S, W (byte 0 = N);
                                        row 2i+1: loop-back vehicles (route ends where it started) per direction, same lanes */
    int32_t *qwait;       /* [ni*4][n]  sum of Vehicle.waiting_time over the queue */
    uint32_t *misc;       /* [3][n]     current_timestep | flags << 16 ; len(self.vehicles) ; rng_counter */
    double *total_reward; /* [n]        running episode return (info["total_reward"]) */
} beng_traffic_state;

typedef struct beng_traffic_io {
    float *obs;            /* [n][ni*14+4]  _get_observation (environment.py:313-363) */
    float *reward;         /* [n]           _calculate_reward (:287-311) cast to float32 */
    uint8_t *terminated;   /* [n]           current_timestep >= MAX_TIMESTEPS (:196) */
    uint8_t *truncated;    /* [n]           always 0 (:197) */
    double *reward64;      /* [n] nullable  the float64 reward */
    double *ep_return;     /* [n] nullable  written when an episode ends (auto-reset modes) */
    int32_t *ep_length;    /* [n] nullable */
    double *stats;         /* [3] nullable  running {n_episodes, sum_return, sum_length} */
    int32_t *timestep;     /* [n] nullable  info["timestep"] (environment.py:200): current_timestep after the call, 0 after a reset */
} beng_traffic_io;

/* TrafficManagementEnv.reset (environment.py:141-166) for envs with mask[i] != 0 (NULL = all); first_call != 0 also
 * rewinds the rng counter.  The observation of EVERY env is written to io->obs. */
int beng_traffic_reset(const beng_traffic_params *p, const beng_traffic_state *st, const beng_traffic_io *io,
                       const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream);

/* TrafficManagementEnv.step (environment.py:168-203): _apply_actions, every TrafficLight.update, _spawn_vehicles
 * (+ generate_vehicle_route), _process_intersections, _remove_completed_vehicles, _calculate_reward,
 * _get_observation and auto-reset, for all envs, in ONE kernel launch.  actions_dev: int64 [n][ni], values
 * 0 keep / 1 NS_GREEN / 2 EW_GREEN (anything else keeps, like the reference).  The kernels address the [field][env]
 * arrays with 32-bit element offsets: n_envs * ni * 4 must be below 2^32 (119 M envs on the default 9-intersection grid),
 * otherwise BENG_ERR_UNSUPPORTED. */
int beng_traffic_step(const beng_traffic_params *p, const beng_traffic_state *st, const int64_t *actions_dev,
                      const beng_traffic_io *io, int64_t n_envs, void *stream);

/* Same step with HOST action / result buffers; NULL host outputs are skipped; does not synchronise. */
int beng_traffic_step_host(const beng_traffic_params *p, const beng_traffic_state *st, int64_t *actions_dev,
                           const beng_traffic_io *io, int64_t n_envs, const int64_t *actions_host, float *obs_host,
                           float *reward_host, uint8_t *terminated_host, uint8_t *truncated_host, void *stream);

/* ------------------------------------------------------------------------------------------
 * smartclimate  (reference: smartclimate_rl-main/smartclimate/{env,utils}.py, class SmartClimateEnv)
 * the simplest float env on the same engine / boundary.
 * ------------------------------------------------------------------------------------------ */

#define BENG_CLIMATE_OBS_DIM 9

/* Constructor arguments (env.py:16-24) + batching parameters. */
typedef struct beng_climate_params {
    int32_t max_occupancy;         /* default 8 */
    int32_t episode_minutes;       /* default 1440; reported as `terminated` (env.py:107); <= 65535 */
    int32_t autoreset_mode;        /* BENG_AUTORESET_* */
    int32_t time_limit_truncation; /* 1: also set `truncated` at episode_minutes (gym.make's TimeLimit,
                                      smartclimate/__init__.py:6-10); 0: never (raw class, env.py:108) */
    uint64_t seed;
    uint64_t env_id_base;
} beng_climate_params;

/* Per-env state, [field][env] arrays (env index fastest).  float64 like the reference's Python floats. */
typedef struct beng_climate_state {
    double *f64;  /* [5][n] room_temp, outside_temp, ac_setting, total_reward, energy_usage */
    int32_t *i32; /* [4][n] num_people | lights << 8 | flags << 16 ; current_step ; comfort_time ; rng_counter */
} beng_climate_state;

typedef struct beng_climate_io {
    float *obs;           /* [n][9] room_temp, num_people, time_of_day, outside_temp, ac_setting, lights x4 (env.py:72-82) */
    float *reward;        /* [n]    calculate_reward (utils.py:30-50) cast to float32 */
    uint8_t *terminated;  /* [n]    current_step >= episode_minutes (env.py:107) */
    uint8_t *truncated;   /* [n]    always 0 (env.py:108) unless time_limit_truncation */
    double *reward64;     /* [n]    nullable */
    double *reward_terms; /* [3][n] nullable: info["comfort"], ["ac_penalty"], ["light_penalty"] (utils.py:46-50) */
    double *ep_return;    /* [n]    nullable, written when an episode ends (auto-reset modes) */
    int32_t *ep_length;   /* [n]    nullable */
    double *stats;        /* [3]    nullable running {n_episodes, sum_return, sum_length} */
} beng_climate_io;

/* SmartClimateEnv.reset / _init_state (env.py:48-70) for envs with mask[i] != 0 (NULL = all); first_call != 0 rewinds
 * the rng counter.  The observation of EVERY env is written to io->obs. */
int beng_climate_reset(const beng_climate_params *p, const beng_climate_state *st, const beng_climate_io *io,
                       const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream);

/* SmartClimateEnv.step (env.py:84-117) + utils.py:5-50 + auto-reset, ONE kernel launch.  The Dict action of the
 * reference is passed as two arrays: ac_temp_dev float32 [n] (action['ac_temp'][0]) and lights_dev int8 [n][4]. */
int beng_climate_step(const beng_climate_params *p, const beng_climate_state *st, const float *ac_temp_dev,
                      const int8_t *lights_dev, const beng_climate_io *io, int64_t n_envs, void *stream);

/* Same step with HOST action / result buffers; NULL host outputs are skipped; does not synchronise. */
int beng_climate_step_host(const beng_climate_params *p, const beng_climate_state *st, float *ac_temp_dev,
                           int8_t *lights_dev, const beng_climate_io *io, int64_t n_envs, const float *ac_temp_host,
                           const int8_t *lights_host, float *obs_host, float *reward_host, uint8_t *terminated_host,
                           uint8_t *truncated_host, void *stream);

/* ------------------------------------------------------------------------------------------
 * world_builder_env  (reference: world_builder_env/src/environment/{world_builder_env,game_logic}.py)
 * integer grid builder on the same engine / boundary.
 * ------------------------------------------------------------------------------------------ */

/* Constructor arguments (world_builder_env.py:38-44) + batching parameters.  The game constants (costs, production,
 * MAX_POPULATION 20, WIN_STEPS 50, initial resources) are compiled in (game_logic.py:13-57). */
typedef struct beng_builder_params {
    int32_t grid_size;      /* default 10; 2..15 */
    int32_t autoreset_mode; /* BENG_AUTORESET_* */
    uint64_t seed;
    uint64_t env_id_base;
} beng_builder_params;

/* Per-env scalar state, [word][env] int32 array (env index fastest):
 *   0 food  1 wood  2 stone  3 population  4 population_capacity
 *   5 building counts: farm | lumberyard << 8 | quarry << 16 | house << 24
 *   6 steps  7 win_steps | reached_win_population << 16 | flags << 24  8 rng_counter  9 running episode return
 * The grid itself lives in io->grid: it is both state and observation (one cell changes per successful build). */
typedef struct beng_builder_state {
    int32_t *words; /* [10][n] */
} beng_builder_state;

typedef struct beng_builder_io {
    int8_t *grid;            /* [n][G][G]  obs['grid'] AND the persistent grid state (0 empty, 1 farm, 2 lumberyard,
                                           3 quarry, 4 house); read and rewritten in place by every step */
    float *resources;        /* [n][4]     obs['resources'] = food, wood, stone, population (world_builder_env.py:205-210) */
    float *capacity;         /* [n][1]     obs['population_capacity'] */
    int32_t *win_steps;      /* [n][1]     obs['win_steps'] */
    float *flat_obs;         /* [n][G*G+6] nullable: the flatten_obs=True observation (:189-200) */
    float *reward;           /* [n]        execute_action's reward, or -100 / +100 on termination (:150-157) */
    uint8_t *terminated;     /* [n]        population <= 0, or 50 steps at population >= 20 (:233-247) */
    uint8_t *truncated;      /* [n]        always 0 (:148) */
    int32_t *ep_return;      /* [n] nullable, written when an episode ends (auto-reset modes) */
    int32_t *ep_length;      /* [n] nullable */
    int64_t *stats;          /* [4] nullable running {n_episodes, sum_return, sum_length, wins} (integer, exact) */
    int32_t *invalid_count;  /* [1] nullable: out-of-space actions seen (the reference raises ValueError, :135-136;
                                    such an env is left untouched) */
} beng_builder_io;

/* WorldBuilderEnv.reset (world_builder_env.py:99-123) for envs with mask[i] != 0 (NULL = all); first_call != 0 rewinds the
 * rng counter.  The observation of EVERY env is (re)written. */
int beng_builder_reset(const beng_builder_params *p, const beng_builder_state *st, const beng_builder_io *io,
                       const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream);

/* WorldBuilderEnv.step (:125-166) + GameLogic.execute_action and helpers (game_logic.py:59-203) + auto-reset, ONE
 * kernel launch.  actions_dev: int64 [n], 0 pass / 1 farm / 2 lumberyard / 3 quarry / 4 house. */
int beng_builder_step(const beng_builder_params *p, const beng_builder_state *st, const int64_t *actions_dev,
                      const beng_builder_io *io, int64_t n_envs, void *stream);

/* Same step with HOST action / result buffers; NULL host outputs are skipped; does not synchronise. */
int beng_builder_step_host(const beng_builder_params *p, const beng_builder_state *st, int64_t *actions_dev,
                           const beng_builder_io *io, int64_t n_envs, const int64_t *actions_host, int8_t *grid_host,
                           float *resources_host, float *capacity_host, int32_t *win_steps_host, float *reward_host,
                           uint8_t *terminated_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BENG_H */

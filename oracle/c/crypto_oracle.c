/* CPU oracle for CryptoTradingEnv -- a plain-C, float64 restatement of the reference algorithm.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Follows /root/reference/crypto_trading_env/crypto_trading_env.py:
 *   TradingConfig                      :28-38
 *   TechnicalIndicators.rsi            :45-61     bollinger_bands :64-77     macd :80-104     _ema :107-119
 *   MarketSimulator.generate_next_price:132-164   _update_market_regime :166-186   volatility :188-198
 *                   trend :200-211     _update_market_psychology :213-221
 *   CryptoTradingEnv.reset             :301-340   step :342-398   _execute_action :400-447
 *                   _execute_buy :449-476   _execute_sell :478-503   _get_observation :505-561
 * and SURVEY.md section 0 facts 5, 7, 8: the time limit is reported as `terminated`; the observation has
 * 261 elements (the declared space says 260); the MarketSimulator is NOT reset by reset().
 *
 * Arithmetic is float64 in the reference's operation order.  np.mean / np.std reduce with NumPy's pairwise
 * summation (8 running partial sums for 8 <= n <= 128), restated in np_sum() so that the float64 results agree
 * bit for bit with the reference on the same libm; the observation is cast to float32 at the end (:561).
 * The MACD signal line's O(n^2) prefix loop (:94-100) is evaluated with one forward scan -- `_ema(prices[:i])`
 * is the running EMA at index i-1, same operations in the same order.
 *
 * RNG draw sites, in order: [uniform(0.5,1.5) if a trade executes], uniform(0.5,2.0),
 * random(), [choice(2), uniform(trend range)], normal(0, vol), uniform(1,1.02), uniform(0.98,1); reset adds
 * uniform(0.99,1.01) per warm-up candle.
 *
 * Parity pin: tests/golden/crypto_golden.npz (oracle/gen_golden_crypto.py, from the reference itself).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beng_oracle_rng.h"

#define HIST 50
#define OBS_DIM 261

enum { BULL_RUN = 0, BEAR_MARKET = 1, SIDEWAYS = 2, CRASH = 3, RECOVERY = 4 };
static const double VOL_MULT[5] = {1.2, 1.5, 0.8, 3.0, 2.0};            /* :190-196 */
static const double BASE_TREND[5] = {0.001, -0.001, 0.0, -0.005, 0.002}; /* :202-208 */
static const int NEXT_REGIME[5][2] = {                                   /* :168-174 */
    {SIDEWAYS, CRASH}, {SIDEWAYS, RECOVERY}, {BULL_RUN, BEAR_MARKET}, {RECOVERY, BEAR_MARKET}, {BULL_RUN, SIDEWAYS}};

typedef struct {
    double initial_balance, trading_fee_rate, slippage_rate, min_price, max_price, volatility_base,
        market_psychology_factor;
    int max_steps;
} crypto_cfg;

typedef struct {
    double cash, holdings;
    double candles[HIST][5]; /* oldest first: open, high, low, close, volume */
    int regime;
    double trend_strength, psychology;
    int step;
    int needs_reset;
    orc_stream rng;
    /* last trade (info["trade_info"]) */
    int trade_kind; /* 0 none, 1 buy, 2 sell */
} crypto_env;

typedef struct {
    int n_envs, mode, continuous;
    crypto_cfg cfg;
    crypto_env *envs;
    double stats[4]; /* n_episodes, sum_return, sum_length, sum_final_value */
    double *ep_return_acc;
} crypto_oracle;

/* NumPy pairwise_sum for n <= 128 (numpy/_core/src/umath/loops_utils.h.src). */
static double np_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* :166-186 */
static void update_regime(crypto_env *e) {
    e->regime = NEXT_REGIME[e->regime][orc_randint(&e->rng, 0, 1)];
    if (e->regime == BULL_RUN || e->regime == RECOVERY) e->trend_strength = orc_uniform(&e->rng, 0.5, 1.0);
    else if (e->regime == BEAR_MARKET || e->regime == CRASH) e->trend_strength = orc_uniform(&e->rng, -1.0, -0.5);
    else e->trend_strength = orc_uniform(&e->rng, -0.2, 0.2);
}

/* :132-164 + :213-221 */
static double next_price(const crypto_cfg *c, crypto_env *e, double current_price, double volume) {
    if (orc_random(&e->rng) < 0.01) update_regime(e);
    double volatility = c->volatility_base * VOL_MULT[e->regime];
    double psychology_drift = (e->psychology - 0.5) * c->market_psychology_factor;
    double trend_component = BASE_TREND[e->regime] * e->trend_strength;
    double random_component = orc_normal(&e->rng, 0.0, volatility);
    double volume_factor = 1.0 / (1.0 + volume * 0.1);
    double pct = (trend_component + psychology_drift + random_component) * volume_factor;
    double new_price = clipd(current_price * (1.0 + pct), c->min_price, c->max_price);
    e->psychology += pct * 10.0;
    e->psychology = clipd(e->psychology, 0.0, 1.0);
    e->psychology += (0.5 - e->psychology) * 0.01;
    return new_price;
}

/* :301-340 (the MarketSimulator state is deliberately left alone) */
static void env_reset(const crypto_cfg *c, crypto_env *e) {
    e->cash = c->initial_balance;
    e->holdings = 0.0;
    e->step = 0;
    e->needs_reset = 0;
    e->trade_kind = 0;
    double price = 50000.0;
    for (int k = 0; k < HIST; ++k) {
        double volume = orc_uniform(&e->rng, 0.5, 2.0);
        price = next_price(c, e, price, volume);
        double high = price * orc_uniform(&e->rng, 1.0, 1.02);
        double low = price * orc_uniform(&e->rng, 0.98, 1.0);
        double open = price * orc_uniform(&e->rng, 0.99, 1.01);
        e->candles[k][0] = open; e->candles[k][1] = high; e->candles[k][2] = low; e->candles[k][3] = price;
        e->candles[k][4] = volume;
    }
}

/* :449-476 */
static int do_buy(const crypto_cfg *c, crypto_env *e, double amount, double price) {
    if (amount <= 0 || e->cash < amount) return 0;
    double slippage = price * c->slippage_rate * orc_uniform(&e->rng, 0.5, 1.5);
    double effective = price + slippage;
    double fee = amount * c->trading_fee_rate;
    double net = amount - fee;
    double bought = net / effective;
    e->cash -= amount;
    e->holdings += bought;
    return 1;
}

/* :478-503 */
static int do_sell(const crypto_cfg *c, crypto_env *e, double crypto_amount, double price) {
    if (crypto_amount <= 0 || e->holdings < crypto_amount) return 0;
    double slippage = price * c->slippage_rate * orc_uniform(&e->rng, 0.5, 1.5);
    double effective = price - slippage;
    double received = crypto_amount * effective;
    double fee = received * c->trading_fee_rate;
    double net_cash = received - fee;
    e->holdings -= crypto_amount;
    e->cash += net_cash;
    return 2;
}

/* :505-561 */
static void write_obs(const crypto_cfg *c, const crypto_env *e, float *obs) {
    double cur = e->candles[HIST - 1][3];
    int o = 0;
    for (int k = 0; k < HIST; ++k)
        for (int f = 0; f < 5; ++f) obs[o++] = (float)(e->candles[k][f] / cur);
    double value = e->cash + e->holdings * cur;
    obs[o++] = (float)(e->cash / c->initial_balance);
    obs[o++] = (float)(e->holdings * cur / c->initial_balance);
    obs[o++] = (float)(value / c->initial_balance);

    double p[HIST];
    for (int k = 0; k < HIST; ++k) p[k] = e->candles[k][3];

    /* RSI(14), :45-61 */
    double gains[14], losses[14];
    for (int i = 0; i < 14; ++i) {
        double d = p[HIST - 14 + i] - p[HIST - 15 + i];
        gains[i] = d > 0 ? d : 0.0;
        losses[i] = d < 0 ? -d : 0.0;
    }
    double avg_gain = np_sum(gains, 14) / 14.0, avg_loss = np_sum(losses, 14) / 14.0;
    double rsi;
    if (avg_loss == 0) rsi = 100.0;
    else { double rs = avg_gain / avg_loss; rsi = 100.0 - (100.0 / (1.0 + rs)); }
    obs[o++] = (float)(rsi / 100.0);

    /* MACD(12, 26, 9), :80-119 */
    const double mf = 2.0 / 13.0, ms = 2.0 / 27.0, mg = 2.0 / 10.0;
    double ef = p[0], es = p[0], sig = 0.0;
    for (int k = 1; k < HIST; ++k) {
        ef = (p[k] * mf) + (ef * (1.0 - mf));
        es = (p[k] * ms) + (es * (1.0 - ms));
        if (k == 25) sig = ef - es;                                   /* macd_values[0] = macd of prices[:26] */
        else if (k > 25) sig = ((ef - es) * mg) + (sig * (1.0 - mg)); /* _ema(macd_values, 9) */
    }
    double macd_line = ef - es, histogram = macd_line - sig;
    double mx = p[0], mn = p[0];
    for (int k = 1; k < HIST; ++k) { if (p[k] > mx) mx = p[k]; if (p[k] < mn) mn = p[k]; }
    double price_range = mx - mn;
    if (price_range > 0) {
        obs[o++] = (float)(macd_line / price_range);
        obs[o++] = (float)(sig / price_range);
        obs[o++] = (float)(histogram / price_range);
    } else { obs[o++] = 0.f; obs[o++] = 0.f; obs[o++] = 0.f; }

    /* Bollinger(20, 2 sigma, population std), :64-77 */
    const double *w = p + HIST - 20;
    double sma = np_sum(w, 20) / 20.0;
    double sq[20];
    for (int i = 0; i < 20; ++i) { double d = w[i] - sma; sq[i] = d * d; }
    double std = sqrt(np_sum(sq, 20) / 20.0);
    double upper = sma + (2 * std), lower = sma - (2 * std);
    obs[o++] = (float)((upper > lower) ? (cur - lower) / (upper - lower) : 0.5);
    obs[o++] = (float)((sma > 0) ? (upper - lower) / sma : 0.0);
    obs[o++] = (float)((sma > 0) ? (cur - sma) / sma : 0.0);

    obs[o++] = (float)e->psychology; /* :559 */
}

crypto_oracle *crypto_oracle_create(int n_envs, uint64_t seed, uint64_t env_id_base, int mode, int continuous,
                                    const double *cfg7, int max_steps) {
    crypto_oracle *o = (crypto_oracle *)calloc(1, sizeof(*o));
    o->n_envs = n_envs; o->mode = mode; o->continuous = continuous;
    o->cfg.initial_balance = cfg7[0]; o->cfg.trading_fee_rate = cfg7[1]; o->cfg.slippage_rate = cfg7[2];
    o->cfg.min_price = cfg7[3]; o->cfg.max_price = cfg7[4]; o->cfg.volatility_base = cfg7[5];
    o->cfg.market_psychology_factor = cfg7[6]; o->cfg.max_steps = max_steps;
    o->envs = (crypto_env *)calloc((size_t)n_envs, sizeof(crypto_env));
    o->ep_return_acc = (double *)calloc((size_t)n_envs, sizeof(double));
    for (int i = 0; i < n_envs; ++i) {
        crypto_env *e = &o->envs[i];
        e->regime = SIDEWAYS; e->trend_strength = 0.0; e->psychology = 0.5; /* MarketSimulator.__init__ :125-130 */
        e->rng.seed = seed; e->rng.env = env_id_base + (uint64_t)i; e->rng.stream = 0; e->rng.counter = 0;
    }
    return o;
}

void crypto_oracle_destroy(crypto_oracle *o) {
    if (!o) return;
    free(o->envs); free(o->ep_return_acc); free(o);
}

void crypto_oracle_reset(crypto_oracle *o, const uint8_t *mask, float *obs) {
    for (int i = 0; i < o->n_envs; ++i) {
        crypto_env *e = &o->envs[i];
        if (!mask || mask[i]) { env_reset(&o->cfg, e); o->ep_return_acc[i] = 0.0; }
        if (obs) write_obs(&o->cfg, e, obs + (size_t)OBS_DIM * i);
    }
}

/* actions: int64[n] (discrete) or float[n*2] (continuous).  info outputs may be NULL. */
void crypto_oracle_step(crypto_oracle *o, const void *actions, float *obs, float *reward, uint8_t *terminated,
                        uint8_t *truncated, double *reward64, double *portfolio_value, double *current_price,
                        uint8_t *trade_kind, double *ep_return, int32_t *ep_length) {
    const crypto_cfg *c = &o->cfg;
    for (int i = 0; i < o->n_envs; ++i) {
        crypto_env *e = &o->envs[i];
        double rew = 0.0, value;
        int term = 0;
        if (o->mode == 1 && e->needs_reset) {
            env_reset(c, e);
            o->ep_return_acc[i] = 0.0;
            value = e->cash + e->holdings * e->candles[HIST - 1][3];
        } else {
            /* _execute_action, :400-447 */
            double price = e->candles[HIST - 1][3];
            double initial_value = e->cash + e->holdings * price;
            int trade = 0;
            if (o->continuous) {
                const float *a = (const float *)actions + 2 * (size_t)i;
                double max_buy = e->cash * 0.1, max_sell = e->holdings * 0.1;
                double buy = clipd((double)a[0], 0.0, 1.0) * max_buy, sell = clipd((double)a[1], 0.0, 1.0) * max_sell;
                if (buy > sell && buy > 0) trade = do_buy(c, e, buy, price);
                else if (sell > 0) trade = do_sell(c, e, sell, price);
            } else {
                int64_t a = ((const int64_t *)actions)[i];
                if (a == 1) trade = do_buy(c, e, e->cash * 0.05, price);
                else if (a == 2) trade = do_buy(c, e, e->cash * 0.2, price);
                else if (a == 3) trade = do_sell(c, e, e->holdings * 0.05, price);
                else if (a == 4) trade = do_sell(c, e, e->holdings * 0.2, price);
            }
            e->trade_kind = trade;
            double final_value = e->cash + e->holdings * price;
            rew = final_value - initial_value;
            if (!trade) rew -= 1.0;
            /* new candle, :348-365 */
            double volume = orc_uniform(&e->rng, 0.5, 2.0);
            double new_price = next_price(c, e, price, volume);
            double high = new_price * orc_uniform(&e->rng, 1.0, 1.02);
            double low = new_price * orc_uniform(&e->rng, 0.98, 1.0);
            memmove(e->candles[0], e->candles[1], sizeof(double) * 5 * (HIST - 1));
            double *nc = e->candles[HIST - 1];
            nc[0] = price; nc[1] = high; nc[2] = low; nc[3] = new_price; nc[4] = volume;
            value = e->cash + e->holdings * new_price;
            e->step += 1;
            term = (e->step >= c->max_steps) || (value <= 0) || (value >= c->initial_balance * 10); /* :382-386 */
            o->ep_return_acc[i] += rew;
        }
        if (reward64) reward64[i] = rew;
        if (portfolio_value) portfolio_value[i] = value;
        if (current_price) current_price[i] = e->candles[HIST - 1][3];
        if (trade_kind) trade_kind[i] = (uint8_t)e->trade_kind;
        if (term && o->mode != 0) {
            o->stats[0] += 1; o->stats[1] += o->ep_return_acc[i]; o->stats[2] += e->step; o->stats[3] += value;
            if (ep_return) ep_return[i] = o->ep_return_acc[i];
            if (ep_length) ep_length[i] = e->step;
            if (o->mode == 2) { env_reset(c, e); o->ep_return_acc[i] = 0.0; } else e->needs_reset = 1;
        }
        if (obs) write_obs(c, e, obs + (size_t)OBS_DIM * i);
        reward[i] = (float)rew;
        terminated[i] = (uint8_t)term;
        if (truncated) truncated[i] = 0;
    }
}

/* cash, holdings, trend_strength, psychology (f64); regime, step (i32); rng counter; closes [n][50] */
void crypto_oracle_get_state(const crypto_oracle *o, double *cash, double *holdings, double *trend_strength,
                             double *psychology, int32_t *regime, int32_t *step, uint32_t *rng_counter,
                             double *candles) {
    for (int i = 0; i < o->n_envs; ++i) {
        const crypto_env *e = &o->envs[i];
        cash[i] = e->cash; holdings[i] = e->holdings; trend_strength[i] = e->trend_strength;
        psychology[i] = e->psychology; regime[i] = e->regime; step[i] = e->step; rng_counter[i] = e->rng.counter;
        if (candles) memcpy(candles + (size_t)i * HIST * 5, e->candles, sizeof(e->candles));
    }
}

/* Teacher forcing: overwrite the float64 state of every env (tests re-sync the oracle from the device). */
void crypto_oracle_set_state(crypto_oracle *o, const double *cash, const double *holdings,
                             const double *trend_strength, const double *psychology, const int32_t *regime,
                             const int32_t *step, const uint32_t *rng_counter, const double *candles) {
    for (int i = 0; i < o->n_envs; ++i) {
        crypto_env *e = &o->envs[i];
        e->cash = cash[i]; e->holdings = holdings[i]; e->trend_strength = trend_strength[i];
        e->psychology = psychology[i]; e->regime = regime[i]; e->step = step[i]; e->rng.counter = rng_counter[i];
        if (candles) memcpy(e->candles, candles + (size_t)i * HIST * 5, sizeof(e->candles));
    }
}

void crypto_oracle_get_stats(const crypto_oracle *o, double *out4) { memcpy(out4, o->stats, sizeof(o->stats)); }

// Batched CryptoTradingEnv for sm_100a: trade execution + regime-switching price walk + new candle +
// termination + auto-reset + the 261-feature observation (OHLCV window, portfolio, RSI, MACD, Bollinger,
// psychology) in ONE kernel.
//
// Reference behaviour (paths relative to the reference root, file crypto_trading_env/crypto_trading_env.py):
//   TradingConfig :28-38                 TechnicalIndicators.{rsi,bollinger_bands,macd,_ema} :41-119
//   MarketSimulator.generate_next_price :132-164, _update_market_regime :166-186, volatility :188-198,
//                   trend :200-211, _update_market_psychology :213-221
//   CryptoTradingEnv.reset :301-340, step :342-398, _execute_action :400-447, _execute_buy :449-476,
//                   _execute_sell :478-503, _get_observation :505-561
//
// Design:
//   * one THREAD per env; every state array is laid out [slot/field][env] so that a warp's accesses are
//     contiguous (the 50-candle window is read as 250 fully coalesced loads per thread, high ILP);
//   * the window is a ring over 50 slots with ONE head shared by all envs (params.window_head): a step
//     overwrites the oldest slot, a reset rewrites all 50 slots in rotation -- heads never diverge between
//     envs, so the layout stays coalesced whatever the episode boundaries are;
//   * money, prices and indicators are float64 in the reference's operation order (the MACD is a difference
//     of two EMAs of ~5e4-magnitude prices: float32 closes would break the 1e-5 tolerance); open/high/low/
//     volume only feed the observation and are stored and normalised in float32;
//   * the MACD signal line (an O(n^2) prefix loop in the reference, :94-100) is one forward scan;
//   * the T x 261 float observation tile is contiguous in global memory: composed in shared memory (row
//     stride 261 words = conflict-free) and drained with one bulk asynchronous copy (cp.async.bulk, UBLKCP).
//
// HBM-bound: ~2.4 KB per env-step (window read 1.2 KB + observation write 1.04 KB); no tensor-core work.
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "beng_common.cuh"
#include "beng_rng.cuh"

namespace beng {
namespace {

constexpr int HIST = BENG_CRYPTO_HISTORY;
constexpr int OBS = BENG_CRYPTO_OBS_DIM;
constexpr uint32_t CFLAG_NEEDS_RESET = 1u;

enum { BULL_RUN = 0, BEAR_MARKET = 1, SIDEWAYS = 2, CRASH = 3, RECOVERY = 4 };

struct CArgs {
    beng_crypto_params p;
    beng_crypto_state st;
    beng_crypto_io io;
    const void *actions;
    const uint8_t *mask;
    long long n;
    int first_call;
};

struct Market {
    int regime;
    double trend, psych;
};

__device__ __forceinline__ double vol_mult(int r) {  // :190-196
    return r == BULL_RUN ? 1.2 : r == BEAR_MARKET ? 1.5 : r == SIDEWAYS ? 0.8 : r == CRASH ? 3.0 : 2.0;
}
__device__ __forceinline__ double base_trend(int r) {  // :202-208
    return r == BULL_RUN ? 0.001 : r == BEAR_MARKET ? -0.001 : r == SIDEWAYS ? 0.0 : r == CRASH ? -0.005 : 0.002;
}
__device__ __forceinline__ double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// _update_market_regime, :166-186
__device__ __forceinline__ void update_regime(Market &m, EnvStream &rng) {
    const int pick = rng.randint(0, 1);  // random.choice of the two successors
    int nr;
    switch (m.regime) {
        case BULL_RUN: nr = pick ? CRASH : SIDEWAYS; break;
        case BEAR_MARKET: nr = pick ? RECOVERY : SIDEWAYS; break;
        case SIDEWAYS: nr = pick ? BEAR_MARKET : BULL_RUN; break;
        case CRASH: nr = pick ? BEAR_MARKET : RECOVERY; break;
        default: nr = pick ? SIDEWAYS : BULL_RUN; break;
    }
    m.regime = nr;
    if (nr == BULL_RUN || nr == RECOVERY) m.trend = rng.uniform(0.5, 1.0);
    else if (nr == BEAR_MARKET || nr == CRASH) m.trend = rng.uniform(-1.0, -0.5);
    else m.trend = rng.uniform(-0.2, 0.2);
}

// generate_next_price, :132-164, and _update_market_psychology, :213-221
__device__ __forceinline__ double next_price(const beng_crypto_params &p, Market &m, EnvStream &rng, double cur,
                                             double volume) {
    if (rng.random53() < 0.01) update_regime(m, rng);
    const double volatility = p.volatility_base * vol_mult(m.regime);
    const double drift = (m.psych - 0.5) * p.market_psychology_factor;
    const double trend = base_trend(m.regime) * m.trend;
    const double eps = rng.normal(0.0, volatility);
    const double volume_factor = 1.0 / (1.0 + volume * 0.1);
    const double pct = (trend + drift + eps) * volume_factor;
    const double np_ = clipd(cur * (1.0 + pct), p.min_price, p.max_price);
    m.psych += pct * 10.0;
    m.psych = clipd(m.psych, 0.0, 1.0);
    m.psych += (0.5 - m.psych) * 0.01;
    return np_;
}

__device__ __forceinline__ void store_candle(double *close_arr, float *ohlv_arr, long long n, long long env, int slot,
                                             double open, double high, double low, double close, double volume) {
    close_arr[(long long)slot * n + env] = close;
    float *o = ohlv_arr + ((long long)slot * 4) * n + env;
    o[0] = (float)open;
    o[n] = (float)high;
    o[2 * n] = (float)low;
    o[3 * n] = (float)volume;
}

// reset, :301-340: 50 warm-up candles from 50000.0; the newest lands in slot `head`.  Market state carries over.
// Deliberately NOT inlined (it runs once per 1000 steps) and deliberately BY VALUE: taking the address of the
// caller's market / RNG state or of the kernel parameter block would force them into local memory for every
// thread on the hot path (measured: 135 us of a 300 us step).
struct WarmupResult {
    Market m;
    uint32_t ctr;
};
__device__ __noinline__ WarmupResult warmup_window(double *close_arr, float *ohlv_arr, long long n, long long env,
                                                   int head, beng_crypto_params p, Market m, uint64_t gid,
                                                   uint32_t ctr) {
    EnvStream rng(p.seed, gid, BENG_STREAM_ENV, ctr);
    double price = 50000.0;
    int slot = head + 1 == HIST ? 0 : head + 1;  // oldest
    for (int k = 0; k < HIST; ++k) {
        const double volume = rng.uniform(0.5, 2.0);
        price = next_price(p, m, rng, price, volume);
        const double high = price * rng.uniform(1.0, 1.02);
        const double low = price * rng.uniform(0.98, 1.0);
        const double open = price * rng.uniform(0.99, 1.01);
        store_candle(close_arr, ohlv_arr, n, env, slot, open, high, low, price, volume);
        slot = slot + 1 == HIST ? 0 : slot + 1;
    }
    return WarmupResult{m, rng.ctr};
}

// NumPy's pairwise summation order for 8 <= n <= 128 (np.mean / np.std in the reference), n static.
template <int N>
__device__ __forceinline__ double np_sum(const double (&v)[N]) {
    static_assert(N >= 8 && N < 24, "restated for the two sizes the reference uses (14, 20)");
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = v[j];
    if (N >= 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += v[8 + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int i = (N >= 16 ? 16 : 8); i < N; ++i) res += v[i];
    return res;
}

// Indicator part of _get_observation (:519-559) for one env, plus the normalised close column.
// `closes` is the CTA's staging area [HIST-1][T] (float64, oldest first) holding the 49 older closes of every env
// in the tile; `cur` is the newest close.  Writes row[k*5+3] for k < 49 and row[250..260].
// `close_at(k)` returns the k-th oldest close (k < 49); WRITE_CLOSE_COL selects whether the normalised close column
// row[k*5+3] is written here.  `row` receives 250.. as row[250+i] when TAIL_ONLY is false, else tail[i] = feature 250+i.
template <bool WRITE_CLOSE_COL, typename CloseAt>
__device__ __forceinline__ void compose_indicators(const beng_crypto_params &p, CloseAt close_at, double cur,
                                                   double cash, double holdings, double psych, float *row,
                                                   float *tail) {
    const double inv = 1.0 / cur;
    const double mf = 2.0 / 13.0, ms = 2.0 / 27.0, mg = 2.0 / 10.0;  // _ema multipliers, :113
    double ef = 0.0, es = 0.0, sig = 0.0, mx = 0.0, mn = 0.0;
    double w[20];  // the last 20 closes (Bollinger window; its last 15 give the 14 RSI deltas)
#pragma unroll
    for (int k = 0; k < HIST; ++k) {
        const double c = (k == HIST - 1) ? cur : close_at(k);
        if (WRITE_CLOSE_COL && k < HIST - 1) row[k * 5 + 3] = (float)(c * inv);  // close / current_price, :513-515
        if (k == 0) {
            ef = es = mx = mn = c;  // _ema seeds at prices[0], :114
        } else {
            ef = (c * mf) + (ef * (1.0 - mf));  // :116-117
            es = (c * ms) + (es * (1.0 - ms));
            if (k == 25) sig = ef - es;                                    // macd_values[0] (prices[:26]), :94-98
            else if (k > 25) sig = ((ef - es) * mg) + (sig * (1.0 - mg));  // _ema(macd_values, 9), :100
            mx = c > mx ? c : mx;
            mn = c < mn ? c : mn;
        }
        if (k >= HIST - 20) w[k - (HIST - 20)] = c;
    }

    const double value = cash + holdings * cur;  // :519-527
    tail[0] = (float)(cash / p.initial_balance);
    tail[1] = (float)(holdings * cur / p.initial_balance);
    tail[2] = (float)(value / p.initial_balance);

    // RSI(14) over the last 14 deltas, :45-61
    double g[14], l[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        const double d = w[6 + i] - w[5 + i];
        g[i] = d > 0 ? d : 0.0;
        l[i] = d < 0 ? -d : 0.0;
    }
    const double avg_gain = np_sum(g) / 14.0, avg_loss = np_sum(l) / 14.0;
    double rsi = 100.0;
    if (avg_loss != 0) {
        const double rs = avg_gain / avg_loss;
        rsi = 100.0 - (100.0 / (1.0 + rs));
    }
    tail[3] = (float)(rsi / 100.0);

    // MACD(12, 26, 9) normalised by the close range, :538-547
    const double macd_line = ef - es, hist = macd_line - sig, range = mx - mn;
    if (range > 0) {
        tail[4] = (float)(macd_line / range);
        tail[5] = (float)(sig / range);
        tail[6] = (float)(hist / range);
    } else {
        tail[4] = tail[5] = tail[6] = 0.0f;
    }

    // Bollinger(20, 2 sigma, population std), :64-77 and :550-554
    const double sma = np_sum(w) / 20.0;
    double sq[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        const double d = w[i] - sma;
        sq[i] = d * d;
    }
    const double sd = sqrt(np_sum(sq) / 20.0);
    const double upper = sma + (2 * sd), lower = sma - (2 * sd);
    tail[7] = (float)((upper > lower) ? (cur - lower) / (upper - lower) : 0.5);
    tail[8] = (float)((sma > 0) ? (upper - lower) / sma : 0.0);
    tail[9] = (float)((sma > 0) ? (cur - sma) / sma : 0.0);
    tail[10] = (float)psych;  // :559
}

// L2-coherent loads (bypass L1) for data another thread of this CTA may have just rewritten.
__device__ __forceinline__ double ld_cg_f64(const double *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_cg_f32(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// _execute_buy, :449-476.  Returns 1 when the order executed.
__device__ __forceinline__ int do_buy(const beng_crypto_params &p, EnvStream &rng, double &cash, double &holdings,
                                      double amount, double price) {
    if (amount <= 0 || cash < amount) return 0;
    const double slippage = price * p.slippage_rate * rng.uniform(0.5, 1.5);
    const double effective = price + slippage;
    const double fee = amount * p.trading_fee_rate;
    const double net = amount - fee;
    cash -= amount;
    holdings += net / effective;
    return 1;
}

// _execute_sell, :478-503.  Returns 2 when the order executed.
__device__ __forceinline__ int do_sell(const beng_crypto_params &p, EnvStream &rng, double &cash, double &holdings,
                                       double crypto_amount, double price) {
    if (crypto_amount <= 0 || holdings < crypto_amount) return 0;
    const double slippage = price * p.slippage_rate * rng.uniform(0.5, 1.5);
    const double effective = price - slippage;
    const double received = crypto_amount * effective;
    const double fee = received * p.trading_fee_rate;
    holdings -= crypto_amount;
    cash += received - fee;
    return 2;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Warp-specialised CTA over a tile of T consecutive envs, 4*T threads:
//   threads [0, T)        "env" threads, one per env: trade, price walk, new candle, termination, auto-reset, then the
//                         sequential indicator scans over the closes staged in shared memory;
//   threads [T, 4T)       "window" threads: at kernel entry they fetch the 49 older candles of the whole tile with many
//                         independent, fully coalesced loads (the memory-level parallelism of the kernel), stage the
//                         closes in shared memory and, once the env threads have published 1/close_now, normalise
//                         open/high/low/volume straight into the observation tile.
// Two CTA barriers: (A) closes staged + 1/close published, (B) tile complete -> one bulk asynchronous store.
// An env that auto-resets rewrites its whole window in this launch; it raises a flag and its elements are re-read
// (L2-coherent) after barrier A.
template <int T, bool IS_RESET>
__global__ void __launch_bounds__(4 * T) crypto_kernel(const CArgs a) {
    constexpr int OLD = HIST - 1;          // candles that exist before this step
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);                                    // [T][261]
    double *s_close = reinterpret_cast<double *>(smem_raw + (size_t)T * OBS * sizeof(float));  // [49][T]
    float *s_inv = reinterpret_cast<float *>(s_close + OLD * T);                          // [T]
    uint8_t *s_reload = reinterpret_cast<uint8_t *>(s_inv + T);                            // [T]

    const int tid = threadIdx.x;
    const long long n = a.n;
    const long long first = (long long)blockIdx.x * T;
    // slot that holds the newest candle once this call is done
    const int head = IS_RESET ? a.p.window_head : (a.p.window_head + 1 == HIST ? 0 : a.p.window_head + 1);
    const int oldest = head + 1 == HIST ? 0 : head + 1;  // slots oldest .. oldest+48 (mod 50) are the older candles

    bool ended = false;
    double st_ret = 0.0, st_len = 0.0, st_val = 0.0;

    if (tid >= T) {
        // ======================================================================== window threads
        // Thread (q, e): env column e = j % T of the tile, slots k = q, q+3, q+6, ... (q = j / T in 0..2).  Per slot it
        // fetches open/high/low/volume (held in registers) and the close (staged in shared memory); slot and field
        // offsets are compile-time, so the address arithmetic is one pointer bump per value.
        const int j = tid - T;
        const int e = j % T, q = j / T;
        constexpr int PER = (OLD + 2) / 3;  // 17 slots per thread
        const bool col_ok = first + e < n;
        const float *obase = a.st.ohlv + first + e;
        const double *cbase = a.st.close + first + e;
        float v[PER][4];
        if constexpr (!IS_RESET) {
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int k = q + 3 * i;
                if (k < OLD && col_ok) {
                    int slot = oldest + k;
                    slot = slot >= HIST ? slot - HIST : slot;
                    const float *o = obase + (long long)slot * 4 * n;
                    v[i][0] = o[0];
                    v[i][1] = o[n];
                    v[i][2] = o[2 * n];
                    v[i][3] = o[3 * n];
                    s_close[k * T + e] = cbase[(long long)slot * n];
                }
            }
        }
        __syncthreads();  // (A)
        // (envs that rewrote their window re-stage their own closes after the barrier; see the env branch)
        if (col_ok) {
            const float inv_f = s_inv[e];
            const bool reload = IS_RESET || s_reload[e];
            float *dst = tile + e * OBS + q * 5;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int k = q + 3 * i;
                if (k < OLD) {
                    float x0 = v[i][0], x1 = v[i][1], x2 = v[i][2], x3 = v[i][3];
                    if (reload) {
                        int slot = oldest + k;
                        slot = slot >= HIST ? slot - HIST : slot;
                        const float *o = obase + (long long)slot * 4 * n;
                        x0 = ld_cg_f32(o);
                        x1 = ld_cg_f32(o + n);
                        x2 = ld_cg_f32(o + 2 * n);
                        x3 = ld_cg_f32(o + 3 * n);
                    }
                    dst[i * 15 + 0] = x0 * inv_f;  // price_data / current_price, :513-515
                    dst[i * 15 + 1] = x1 * inv_f;
                    dst[i * 15 + 2] = x2 * inv_f;
                    dst[i * 15 + 4] = x3 * inv_f;  // (index 3 is the close, written by the env thread)
                }
            }
        }
    } else {
        // ======================================================================== env threads
        const long long env = first + tid;
        const bool active = env < n;
        float *row = tile + tid * OBS;
        double cash = 0.0, holdings = 0.0, cur = 1.0, rew = 0.0, value = 0.0, price_out = 0.0, ep_ret = 0.0;
        float nw_o = 0.f, nw_h = 0.f, nw_l = 0.f, nw_v = 0.f;  // newest candle, float32 like the stored window
        Market m{SIDEWAYS, 0.0, 0.5};
        int step = 0, term = 0, trade = 0;
        bool step_at_limit = false;
        uint32_t flags = 0, ctr = 0;
        bool reloaded = IS_RESET;
        if (active) {
            cash = a.st.scal[env];
            holdings = a.st.scal[n + env];
            m.trend = a.st.scal[2 * n + env];
            m.psych = a.st.scal[3 * n + env];
            const uint32_t meta = a.st.meta[env];
            step = meta & 0xFFFF;
            m.regime = (meta >> 16) & 0xFF;
            flags = meta >> 24;
            ctr = a.st.meta[n + env];
            ep_ret = a.st.ep_return[env];

            bool selected = true;
            if constexpr (IS_RESET) {
                if (a.mask) selected = a.mask[env] != 0;
                if (selected && a.first_call) {  // constructor: MarketSimulator.__init__, :125-130
                    m.regime = SIDEWAYS;
                    m.trend = 0.0;
                    m.psych = 0.5;
                    ctr = 0;
                }
            }
            EnvStream rng(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, ctr);

            auto do_reset = [&]() {
                cash = a.p.initial_balance;
                holdings = 0.0;
                step = 0;
                flags = 0;
                ep_ret = 0.0;
                const WarmupResult wr = warmup_window(a.st.close, a.st.ohlv, n, env, head, a.p, m,
                                                      a.p.env_id_base + (uint64_t)env, rng.ctr);
                m = wr.m;
                rng = EnvStream(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, wr.ctr);
                reloaded = true;
            };

            if constexpr (IS_RESET) {
                if (selected) do_reset();
            } else {
                if (a.p.autoreset_mode == BENG_AUTORESET_NEXT_STEP && (flags & CFLAG_NEEDS_RESET)) {
                    // The ring head moved by one slot with this call: rewrite the whole window at the new rotation.
                    do_reset();
                    price_out = a.st.close[(long long)head * n + env];
                    value = cash + holdings * price_out;
                } else {
                    // _execute_action, :400-447
                    const double price = a.st.close[(long long)a.p.window_head * n + env];
                    const double initial_value = cash + holdings * price;
                    if (a.p.action_type == 1) {
                        const float2 act = reinterpret_cast<const float2 *>(a.actions)[env];
                        const double buy = clipd((double)act.x, 0.0, 1.0) * (cash * 0.1);
                        const double sell = clipd((double)act.y, 0.0, 1.0) * (holdings * 0.1);
                        if (buy > sell && buy > 0) trade = do_buy(a.p, rng, cash, holdings, buy, price);
                        else if (sell > 0) trade = do_sell(a.p, rng, cash, holdings, sell, price);
                    } else {
                        const long long act = reinterpret_cast<const long long *>(a.actions)[env];
                        if (act == 1) trade = do_buy(a.p, rng, cash, holdings, cash * 0.05, price);
                        else if (act == 2) trade = do_buy(a.p, rng, cash, holdings, cash * 0.2, price);
                        else if (act == 3) trade = do_sell(a.p, rng, cash, holdings, holdings * 0.05, price);
                        else if (act == 4) trade = do_sell(a.p, rng, cash, holdings, holdings * 0.2, price);
                        // anything else is a hold: the reference does not validate (:424-436)
                    }
                    const double final_value = cash + holdings * price;
                    rew = final_value - initial_value;  // valued at the OLD price, :440-441
                    if (!trade) rew -= 1.0;             // :444-445
                    // next candle, :348-365
                    const double volume = rng.uniform(0.5, 2.0);
                    const double new_price = next_price(a.p, m, rng, price, volume);
                    const double high = new_price * rng.uniform(1.0, 1.02);
                    const double low = new_price * rng.uniform(0.98, 1.0);
                    store_candle(a.st.close, a.st.ohlv, n, env, head, price, high, low, new_price, volume);
                    cur = new_price;
                    nw_o = (float)price; nw_h = (float)high; nw_l = (float)low; nw_v = (float)volume;
                    value = cash + holdings * new_price;
                    price_out = new_price;
                    step = min(step + 1, 65535);
                    step_at_limit = step >= a.p.max_steps;
                    term = step_at_limit || (value <= 0) || (value >= a.p.initial_balance * 10);  // :382-386
                    ep_ret += rew;
                    if (term && a.p.autoreset_mode != BENG_AUTORESET_DISABLED) {
                        ended = true;
                        st_ret = ep_ret;
                        st_len = (double)step;
                        st_val = value;
                        if (a.io.ep_return_out) a.io.ep_return_out[env] = ep_ret;
                        if (a.io.ep_length) a.io.ep_length[env] = step;
                        if (a.p.autoreset_mode == BENG_AUTORESET_SAME_STEP) do_reset();
                        else flags |= CFLAG_NEEDS_RESET;
                    }
                }
            }
            if (reloaded) {  // newest candle from the (re)written window; same-thread read-after-write
                cur = a.st.close[(long long)head * n + env];
                const float *o = a.st.ohlv + ((long long)head * 4) * n + env;
                nw_o = o[0]; nw_h = o[n]; nw_l = o[2 * n]; nw_v = o[3 * n];
            }
            ctr = rng.ctr;
            s_inv[tid] = (float)(1.0 / cur);
        }
        s_reload[tid] = (uint8_t)(active && reloaded && !IS_RESET);
        __threadfence_block();
        __syncthreads();  // (A)
        if (active) {
            if (reloaded) {  // stage this env's 49 older closes again (its window changed in this launch)
                for (int k = 0; k < OLD; ++k)
                    s_close[k * T + tid] = ld_cg_f64(a.st.close + (long long)((oldest + k) % HIST) * n + env);
            }
            const float inv_f = s_inv[tid];
            row[(HIST - 1) * 5 + 0] = nw_o * inv_f;
            row[(HIST - 1) * 5 + 1] = nw_h * inv_f;
            row[(HIST - 1) * 5 + 2] = nw_l * inv_f;
            row[(HIST - 1) * 5 + 3] = (float)(cur * (1.0 / cur));
            row[(HIST - 1) * 5 + 4] = nw_v * inv_f;
            compose_indicators<true>(a.p, [&](int k) { return s_close[k * T + tid]; }, cur, cash, holdings, m.psych, row,
                                     row + 250);

            a.st.scal[env] = cash;
            a.st.scal[n + env] = holdings;
            a.st.scal[2 * n + env] = m.trend;
            a.st.scal[3 * n + env] = m.psych;
            a.st.meta[env] = (uint32_t)step | ((uint32_t)m.regime << 16) | (flags << 24);
            a.st.meta[n + env] = ctr;
            a.st.ep_return[env] = ep_ret;
            if constexpr (!IS_RESET) {
                a.io.reward[env] = (float)rew;
                a.io.terminated[env] = (uint8_t)term;
                if (a.io.truncated) a.io.truncated[env] = (uint8_t)(a.p.time_limit_truncation && term && step_at_limit);
                if (a.io.reward64) a.io.reward64[env] = rew;
                if (a.io.portfolio_value) a.io.portfolio_value[env] = value;
                if (a.io.current_price) a.io.current_price[env] = price_out;
                if (a.io.trade_kind) a.io.trade_kind[env] = (uint8_t)trade;
            }
        }
    }

    // drain the observation tile with one bulk asynchronous copy
    fence_proxy_async_smem();
    __syncthreads();  // (B)
    if (tid == 0) {
        const long long n_here = min((long long)T, n - first);
        const uint32_t bytes = (uint32_t)(n_here * OBS * sizeof(float));
        const uint32_t bulk = bytes & ~15u;
        if (bulk) bulk_store_s2g(a.io.obs + first * OBS, tile, bulk);
        bulk_commit();
        for (uint32_t i = bulk / 4; i < bytes / 4; ++i) a.io.obs[first * OBS + i] = tile[i];  // ragged last tile
    }

    if constexpr (!IS_RESET) {
        if (a.io.stats && tid < T) {
            const unsigned done_mask = __ballot_sync(0xFFFFFFFFu, ended);
            if (done_mask) {  // rare: ~1 step in 1000
                const double r = warp_sum(st_ret), l = warp_sum(st_len), v2 = warp_sum(st_val);
                if ((tid & 31) == 0) {
                    atomicAdd(&a.io.stats[0], (double)__popc(done_mask));
                    atomicAdd(&a.io.stats[1], r);
                    atomicAdd(&a.io.stats[2], l);
                    atomicAdd(&a.io.stats[3], v2);
                }
            }
        }
    }
    if (tid == 0) bulk_wait_read<0>();  // shared memory must outlive the copy engine's reads
}

// ---------------------------------------------------------------------------------------------------------------
// Two-phase CTA (the default): 256 threads own 256 consecutive envs.
//   phase 1  one thread per env: trade, price walk, candle, termination, auto-reset and the indicator scans, closes
//            read straight from global memory (coalesced over envs).  No observation staging is needed here, so 512
//            env threads are resident per SM (4x the warp-specialised kernel above) to hide the long serial float64
//            chain.  Each thread leaves 1/close and its 11 indicator features in shared memory.
//   phase 2  all 256 threads stream the window of eight 32-env sub-tiles: thread (e = tid % 32, g = tid / 32) handles
//            slots k = g, g+8, ... of env e, normalises open/high/low/close/volume into a double-buffered 33 KB
//            observation tile, which one thread drains with a bulk asynchronous copy while the next sub-tile is built.
// Window reads in phase 2 are L2-coherent (ld.global.cg): an env that auto-reset rewrote its window in phase 1.
constexpr int C2_ENVS = 256, C2_SUB = 32, C2_GROUPS = C2_ENVS / C2_SUB;

template <bool IS_RESET>
__global__ void __launch_bounds__(C2_ENVS, 2) crypto2_kernel(const CArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);                     // [2][32][261]
    float *s_tail = tiles + 2 * C2_SUB * OBS;                               // [256][11]
    float *s_inv = s_tail + C2_ENVS * 11;                                   // [256]
    double *s_invd = reinterpret_cast<double *>(s_inv + C2_ENVS);           // [256]

    const int tid = threadIdx.x;
    const long long n = a.n;
    const long long first = (long long)blockIdx.x * C2_ENVS;
    const int head = IS_RESET ? a.p.window_head : (a.p.window_head + 1 == HIST ? 0 : a.p.window_head + 1);
    const int oldest = head + 1 == HIST ? 0 : head + 1;
    pdl_launch_dependents();  // the next step's grid may become resident as this one drains ...
    pdl_wait();               // ... and this one touches nothing before the previous step's grid has flushed

    // ------------------------------------------------------------------------------------------- phase 1
    bool ended = false;
    double st_ret = 0.0, st_len = 0.0, st_val = 0.0;
    {
        const long long env = first + tid;
        if (env < n) {
            // requested together with the state (they are only used on the ordinary step path, but waiting for the flags
            // to arrive before asking for them would put a second memory round trip at the head of the chain)
            double price_pre = 0.0;
            long long act_pre = 0;
            float2 actf_pre = make_float2(0.0f, 0.0f);
            if constexpr (!IS_RESET) {
                price_pre = a.st.close[(long long)a.p.window_head * n + env];
                if (a.p.action_type == 1) actf_pre = reinterpret_cast<const float2 *>(a.actions)[env];
                else act_pre = reinterpret_cast<const long long *>(a.actions)[env];
            }
            double cash = a.st.scal[env], holdings = a.st.scal[n + env];
            Market m;
            m.trend = a.st.scal[2 * n + env];
            m.psych = a.st.scal[3 * n + env];
            const uint32_t meta = a.st.meta[env];
            int step = meta & 0xFFFF;
            m.regime = (meta >> 16) & 0xFF;
            uint32_t flags = meta >> 24;
            uint32_t ctr = a.st.meta[n + env];
            double ep_ret = a.st.ep_return[env];
            bool selected = true;
            if constexpr (IS_RESET) {
                if (a.mask) selected = a.mask[env] != 0;
                if (selected && a.first_call) {  // constructor: MarketSimulator.__init__, :125-130
                    m.regime = SIDEWAYS;
                    m.trend = 0.0;
                    m.psych = 0.5;
                    ctr = 0;
                }
            }
            EnvStream rng(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, ctr);
            double rew = 0.0, value = 0.0, price_out = 0.0, cur = 0.0;
            int term = 0, trade = 0;
            bool step_at_limit = false;
            bool rewrote = false;

            auto do_reset = [&]() {
                cash = a.p.initial_balance;
                holdings = 0.0;
                step = 0;
                flags = 0;
                ep_ret = 0.0;
                const WarmupResult wr = warmup_window(a.st.close, a.st.ohlv, n, env, head, a.p, m,
                                                      a.p.env_id_base + (uint64_t)env, rng.ctr);
                m = wr.m;
                rng = EnvStream(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, wr.ctr);
                rewrote = true;
            };

            if constexpr (IS_RESET) {
                if (selected) do_reset();
                else rewrote = true;  // (just re-read the newest close below)
            } else {
                if (a.p.autoreset_mode == BENG_AUTORESET_NEXT_STEP && (flags & CFLAG_NEEDS_RESET)) {
                    do_reset();  // the ring head moved by one slot with this call: whole window at the new rotation
                    price_out = a.st.close[(long long)head * n + env];
                    value = cash + holdings * price_out;
                } else {
                    // _execute_action, :400-447
                    const double price = price_pre;
                    const double initial_value = cash + holdings * price;
                    if (a.p.action_type == 1) {
                        const float2 act = actf_pre;
                        const double buy = clipd((double)act.x, 0.0, 1.0) * (cash * 0.1);
                        const double sell = clipd((double)act.y, 0.0, 1.0) * (holdings * 0.1);
                        if (buy > sell && buy > 0) trade = do_buy(a.p, rng, cash, holdings, buy, price);
                        else if (sell > 0) trade = do_sell(a.p, rng, cash, holdings, sell, price);
                    } else {
                        const long long act = act_pre;
                        if (act == 1) trade = do_buy(a.p, rng, cash, holdings, cash * 0.05, price);
                        else if (act == 2) trade = do_buy(a.p, rng, cash, holdings, cash * 0.2, price);
                        else if (act == 3) trade = do_sell(a.p, rng, cash, holdings, holdings * 0.05, price);
                        else if (act == 4) trade = do_sell(a.p, rng, cash, holdings, holdings * 0.2, price);
                        // anything else is a hold: the reference does not validate (:424-436)
                    }
                    const double final_value = cash + holdings * price;
                    rew = final_value - initial_value;  // valued at the OLD price, :440-441
                    if (!trade) rew -= 1.0;             // :444-445
                    // next candle, :348-365
                    const double volume = rng.uniform(0.5, 2.0);
                    const double new_price = next_price(a.p, m, rng, price, volume);
                    const double high = new_price * rng.uniform(1.0, 1.02);
                    const double low = new_price * rng.uniform(0.98, 1.0);
                    store_candle(a.st.close, a.st.ohlv, n, env, head, price, high, low, new_price, volume);
                    cur = new_price;
                    value = cash + holdings * new_price;
                    price_out = new_price;
                    step = min(step + 1, 65535);
                    step_at_limit = step >= a.p.max_steps;
                    term = step_at_limit || (value <= 0) || (value >= a.p.initial_balance * 10);  // :382-386
                    ep_ret += rew;
                    if (term && a.p.autoreset_mode != BENG_AUTORESET_DISABLED) {
                        ended = true;
                        st_ret = ep_ret;
                        st_len = (double)step;
                        st_val = value;
                        if (a.io.ep_return_out) a.io.ep_return_out[env] = ep_ret;
                        if (a.io.ep_length) a.io.ep_length[env] = step;
                        if (a.p.autoreset_mode == BENG_AUTORESET_SAME_STEP) do_reset();
                        else flags |= CFLAG_NEEDS_RESET;
                    }
                }
            }
            if (rewrote) cur = a.st.close[(long long)head * n + env];  // same-thread read-after-write

            // indicator features 250..260; the 49 older closes come straight from global memory
            const double *cbase = a.st.close + env;
            compose_indicators<false>(
                a.p,
                [&](int k) {
                    int slot = oldest + k;
                    slot = slot >= HIST ? slot - HIST : slot;
                    return cbase[(long long)slot * n];
                },
                cur, cash, holdings, m.psych, nullptr, s_tail + tid * 11);
            const double inv = 1.0 / cur;
            s_invd[tid] = inv;
            s_inv[tid] = (float)inv;

            a.st.scal[env] = cash;
            a.st.scal[n + env] = holdings;
            a.st.scal[2 * n + env] = m.trend;
            a.st.scal[3 * n + env] = m.psych;
            a.st.meta[env] = (uint32_t)step | ((uint32_t)m.regime << 16) | (flags << 24);
            a.st.meta[n + env] = rng.ctr;
            a.st.ep_return[env] = ep_ret;
            if constexpr (!IS_RESET) {
                a.io.reward[env] = (float)rew;
                a.io.terminated[env] = (uint8_t)term;
                if (a.io.truncated) a.io.truncated[env] = (uint8_t)(a.p.time_limit_truncation && term && step_at_limit);
                if (a.io.reward64) a.io.reward64[env] = rew;
                if (a.io.portfolio_value) a.io.portfolio_value[env] = value;
                if (a.io.current_price) a.io.current_price[env] = price_out;
                if (a.io.trade_kind) a.io.trade_kind[env] = (uint8_t)trade;
            }
        }
        if constexpr (!IS_RESET) {
            if (a.io.stats) {
                const unsigned done_mask = __ballot_sync(0xFFFFFFFFu, ended);
                if (done_mask) {  // rare: ~1 step in 1000
                    const double r = warp_sum(st_ret), l = warp_sum(st_len), v2 = warp_sum(st_val);
                    if ((tid & 31) == 0) {
                        atomicAdd(&a.io.stats[0], (double)__popc(done_mask));
                        atomicAdd(&a.io.stats[1], r);
                        atomicAdd(&a.io.stats[2], l);
                        atomicAdd(&a.io.stats[3], v2);
                    }
                }
            }
        }
    }
    __threadfence();  // this step's candle (and a reset's whole window) must be in L2 before phase 2 reads it
    __syncthreads();

    // ------------------------------------------------------------------------------------------- phase 2
    const int e = tid % C2_SUB, g = tid / C2_SUB;
    constexpr int PER = (HIST + C2_GROUPS - 1) / C2_GROUPS;  // 7 slots per thread
    // Element offsets of this thread's slots (they do not depend on the sub-tile; only the env column moves).
    long long ooff[PER], coff[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int slot = oldest + g + C2_GROUPS * i;
        slot = slot >= HIST ? slot - HIST : slot;
        ooff[i] = (long long)slot * 4 * n;
        coff[i] = (long long)slot * n;
    }
    // Software pipeline over the sub-tiles: the 35 window values of sub-tile s+1 are requested before sub-tile s is
    // fenced, barriered and handed to the copy engine, so their latency overlaps that hand-over.
    float x[PER][4], xn[PER][4];
    double c[PER], cn[PER];
    auto fetch = [&](int sub, float (&xo)[PER][4], double (&co)[PER]) {
        const long long env = first + (long long)sub * C2_SUB + e;
        if (sub < C2_GROUPS && env < n) {
            const float *obase = a.st.ohlv + env;
            const double *cbase = a.st.close + env;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                if (g + C2_GROUPS * i < HIST) {
                    const float *o = obase + ooff[i];
                    xo[i][0] = ld_cg_f32(o);
                    xo[i][1] = ld_cg_f32(o + n);
                    xo[i][2] = ld_cg_f32(o + 2 * n);
                    xo[i][3] = ld_cg_f32(o + 3 * n);
                    co[i] = ld_cg_f64(cbase + coff[i]);
                }
            }
        }
    };
    fetch(0, x, c);
#pragma unroll 1
    for (int sub = 0; sub < C2_GROUPS; ++sub) {
        const long long sub_first = first + (long long)sub * C2_SUB;
        if (sub_first >= n) break;  // CTA-uniform
        float *tile = tiles + (sub & 1) * (C2_SUB * OBS);
        const long long env = sub_first + e;
        const int le = sub * C2_SUB + e;  // env index within the CTA
        if (sub >= 2) {  // the bulk copy that used this buffer two sub-tiles ago must have read it
            if (tid == 0) bulk_wait_read<1>();
            __syncthreads();
        }
        if (env < n) {
            const float inv_f = s_inv[le];
            const double inv_d = s_invd[le];
            float *dst = tile + e * OBS;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int k = g + C2_GROUPS * i;
                if (k < HIST) {
                    dst[k * 5 + 0] = x[i][0] * inv_f;  // price_data / current_price, :513-515
                    dst[k * 5 + 1] = x[i][1] * inv_f;
                    dst[k * 5 + 2] = x[i][2] * inv_f;
                    dst[k * 5 + 3] = (float)(c[i] * inv_d);
                    dst[k * 5 + 4] = x[i][3] * inv_f;
                }
            }
            // the 11 indicator features: groups 0..7 copy them (11 values over 8 groups)
            for (int j = g; j < 11; j += C2_GROUPS) dst[250 + j] = s_tail[le * 11 + j];
        }
        fetch(sub + 1, xn, cn);  // next sub-tile's loads go out before the hand-over below
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const long long n_here = min((long long)C2_SUB, n - sub_first);
            const uint32_t bytes = (uint32_t)(n_here * OBS * sizeof(float));
            const uint32_t bulk = bytes & ~15u;
            if (bulk) bulk_store_s2g(a.io.obs + sub_first * OBS, tile, bulk);
            bulk_commit();
            for (uint32_t i = bulk / 4; i < bytes / 4; ++i) a.io.obs[sub_first * OBS + i] = tile[i];  // ragged tail
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            x[i][0] = xn[i][0]; x[i][1] = xn[i][1]; x[i][2] = xn[i][2]; x[i][3] = xn[i][3];
            c[i] = cn[i];
        }
    }
    if (tid == 0) bulk_wait_read<0>();
}

constexpr size_t crypto2_smem_bytes() {
    return (size_t)2 * C2_SUB * OBS * sizeof(float) + (size_t)C2_ENVS * 11 * sizeof(float) + C2_ENVS * sizeof(float) +
           C2_ENVS * sizeof(double);
}

// ---------------------------------------------------------------------------------------------------------------
// Split step (BENG_CRYPTO_MODE=split; not the default, see launch()): two kernels, each shaped for what bounds it.
//   crypto_dyn_kernel  one thread per env, no shared memory: trade, price walk, candle, termination, auto-reset and
//                      the indicator scans (closes straight from global memory, coalesced over envs).  Leaves the 11
//                      indicator features and 1/close in st.scratch [12][n].
//   crypto_obs_kernel  persistent streaming kernel (like the snake step): per 32-env tile each thread fetches 35
//                      window values one tile ahead, normalises them into one of three 33 KB tile buffers, and one
//                      thread drains the tile with a bulk asynchronous copy.  Launched with PDL behind the first.
template <int N, typename F>
__device__ __forceinline__ double np_sum_stream(F val) {  // NumPy pairwise order for N in {14, 20}, values on demand
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = val(j);
    if (N >= 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += val(8 + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int i = (N >= 16 ? 16 : 8); i < N; ++i) res += val(i);
    return res;
}

template <bool IS_RESET>
__global__ void __launch_bounds__(256, 2) crypto_dyn_kernel(const CArgs a) {
    const long long n = a.n;
    const long long env = (long long)blockIdx.x * 256 + threadIdx.x;
    const int head = IS_RESET ? a.p.window_head : (a.p.window_head + 1 == HIST ? 0 : a.p.window_head + 1);
    const int oldest = head + 1 == HIST ? 0 : head + 1;
    pdl_launch_dependents();
    pdl_wait();  // the previous step's observation kernel still reads the window slot this step overwrites

    bool ended = false;
    double st_ret = 0.0, st_len = 0.0, st_val = 0.0;
    if (env < n) {
        double cash = a.st.scal[env], holdings = a.st.scal[n + env];
        Market m;
        m.trend = a.st.scal[2 * n + env];
        m.psych = a.st.scal[3 * n + env];
        const uint32_t meta = a.st.meta[env];
        int step = meta & 0xFFFF;
        m.regime = (meta >> 16) & 0xFF;
        uint32_t flags = meta >> 24;
        uint32_t ctr = a.st.meta[n + env];
        double ep_ret = a.st.ep_return[env];
        bool selected = true;
        if constexpr (IS_RESET) {
            if (a.mask) selected = a.mask[env] != 0;
            if (selected && a.first_call) {  // constructor: MarketSimulator.__init__, :125-130
                m.regime = SIDEWAYS;
                m.trend = 0.0;
                m.psych = 0.5;
                ctr = 0;
            }
        }
        EnvStream rng(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, ctr);
        double rew = 0.0, value = 0.0, price_out = 0.0, cur = 0.0;
        int term = 0, trade = 0;
        bool step_at_limit = false, rewrote = false;

        auto do_reset = [&]() {
            cash = a.p.initial_balance;
            holdings = 0.0;
            step = 0;
            flags = 0;
            ep_ret = 0.0;
            const WarmupResult wr = warmup_window(a.st.close, a.st.ohlv, n, env, head, a.p, m,
                                                  a.p.env_id_base + (uint64_t)env, rng.ctr);
            m = wr.m;
            rng = EnvStream(a.p.seed, a.p.env_id_base + (uint64_t)env, BENG_STREAM_ENV, wr.ctr);
            rewrote = true;
        };

        if constexpr (IS_RESET) {
            if (selected) do_reset();
            else rewrote = true;  // (just re-read the newest close below)
        } else {
            if (a.p.autoreset_mode == BENG_AUTORESET_NEXT_STEP && (flags & CFLAG_NEEDS_RESET)) {
                do_reset();  // the ring head moved by one slot with this call: whole window at the new rotation
                price_out = a.st.close[(long long)head * n + env];
                value = cash + holdings * price_out;
            } else {
                // _execute_action, :400-447
                const double price = a.st.close[(long long)a.p.window_head * n + env];
                const double initial_value = cash + holdings * price;
                if (a.p.action_type == 1) {
                    const float2 act = reinterpret_cast<const float2 *>(a.actions)[env];
                    const double buy = clipd((double)act.x, 0.0, 1.0) * (cash * 0.1);
                    const double sell = clipd((double)act.y, 0.0, 1.0) * (holdings * 0.1);
                    if (buy > sell && buy > 0) trade = do_buy(a.p, rng, cash, holdings, buy, price);
                    else if (sell > 0) trade = do_sell(a.p, rng, cash, holdings, sell, price);
                } else {
                    const long long act = reinterpret_cast<const long long *>(a.actions)[env];
                    if (act == 1) trade = do_buy(a.p, rng, cash, holdings, cash * 0.05, price);
                    else if (act == 2) trade = do_buy(a.p, rng, cash, holdings, cash * 0.2, price);
                    else if (act == 3) trade = do_sell(a.p, rng, cash, holdings, holdings * 0.05, price);
                    else if (act == 4) trade = do_sell(a.p, rng, cash, holdings, holdings * 0.2, price);
                    // anything else is a hold: the reference does not validate (:424-436)
                }
                const double final_value = cash + holdings * price;
                rew = final_value - initial_value;  // valued at the OLD price, :440-441
                if (!trade) rew -= 1.0;             // :444-445
                // next candle, :348-365
                const double volume = rng.uniform(0.5, 2.0);
                const double new_price = next_price(a.p, m, rng, price, volume);
                const double high = new_price * rng.uniform(1.0, 1.02);
                const double low = new_price * rng.uniform(0.98, 1.0);
                store_candle(a.st.close, a.st.ohlv, n, env, head, price, high, low, new_price, volume);
                cur = new_price;
                value = cash + holdings * new_price;
                price_out = new_price;
                step = min(step + 1, 65535);
                step_at_limit = step >= a.p.max_steps;
                term = step_at_limit || (value <= 0) || (value >= a.p.initial_balance * 10);  // :382-386
                ep_ret += rew;
                if (term && a.p.autoreset_mode != BENG_AUTORESET_DISABLED) {
                    ended = true;
                    st_ret = ep_ret;
                    st_len = (double)step;
                    st_val = value;
                    if (a.io.ep_return_out) a.io.ep_return_out[env] = ep_ret;
                    if (a.io.ep_length) a.io.ep_length[env] = step;
                    if (a.p.autoreset_mode == BENG_AUTORESET_SAME_STEP) do_reset();
                    else flags |= CFLAG_NEEDS_RESET;
                }
            }
        }
        if (rewrote) cur = a.st.close[(long long)head * n + env];  // same-thread read-after-write

        // ---- state, per-step outputs and the cheap features first: frees their registers for the scans below
        float *sc = a.st.scratch + env;
        {
            const double cur_value = cash + holdings * cur;  // :519-527
            sc[0] = (float)(cash / a.p.initial_balance);
            sc[n] = (float)(holdings * cur / a.p.initial_balance);
            sc[2 * n] = (float)(cur_value / a.p.initial_balance);
            sc[10 * n] = (float)m.psych;        // :559
            sc[11 * n] = (float)(1.0 / cur);    // hand-over to the observation kernel
        }
        a.st.scal[env] = cash;
        a.st.scal[n + env] = holdings;
        a.st.scal[2 * n + env] = m.trend;
        a.st.scal[3 * n + env] = m.psych;
        a.st.meta[env] = (uint32_t)step | ((uint32_t)m.regime << 16) | (flags << 24);
        a.st.meta[n + env] = rng.ctr;
        a.st.ep_return[env] = ep_ret;
        if constexpr (!IS_RESET) {
            a.io.reward[env] = (float)rew;
            a.io.terminated[env] = (uint8_t)term;
            if (a.io.truncated) a.io.truncated[env] = (uint8_t)(a.p.time_limit_truncation && term && step_at_limit);
            if (a.io.reward64) a.io.reward64[env] = rew;
            if (a.io.portfolio_value) a.io.portfolio_value[env] = value;
            if (a.io.current_price) a.io.current_price[env] = price_out;
            if (a.io.trade_kind) a.io.trade_kind[env] = (uint8_t)trade;
        }

        // ---- indicator scans over the 50 closes (oldest first); close k < 49 lives in slot (oldest + k) % 50.
        // The last 20 closes stay in registers for the Bollinger / RSI windows: re-reading them through L1 with 1024
        // threads per SM was measured 2x slower (L1 hit rate 41 %) than 512 threads with the window in registers.
        const double *cbase = a.st.close + env;
        float tail[11];
        compose_indicators<false>(
            a.p,
            [&](int k) {
                int slot = oldest + k;
                slot = slot >= HIST ? slot - HIST : slot;
                return cbase[(long long)slot * n];
            },
            cur, 0.0, 0.0, 0.0, nullptr, tail);
#pragma unroll
        for (int j = 3; j < 10; ++j) sc[(long long)j * n] = tail[j];  // RSI, MACD x3, Bollinger x3
    }
    if constexpr (!IS_RESET) {
        if (a.io.stats) {
            const unsigned done_mask = __ballot_sync(0xFFFFFFFFu, ended);
            if (done_mask) {  // rare: ~1 step in 1000
                const double r = warp_sum(st_ret), l = warp_sum(st_len), v2 = warp_sum(st_val);
                if ((threadIdx.x & 31) == 0) {
                    atomicAdd(&a.io.stats[0], (double)__popc(done_mask));
                    atomicAdd(&a.io.stats[1], r);
                    atomicAdd(&a.io.stats[2], l);
                    atomicAdd(&a.io.stats[3], v2);
                }
            }
        }
    }
}

constexpr int OBS_SUB = 32, OBS_THREADS = 256, OBS_GROUPS = OBS_THREADS / OBS_SUB, OBS_STAGES = 3;

__global__ void __launch_bounds__(OBS_THREADS, 2) crypto_obs_kernel(const CArgs a, int head) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);  // [OBS_STAGES][32][261]
    const int tid = threadIdx.x;
    const int e = tid % OBS_SUB, g = tid / OBS_SUB;
    const long long n = a.n;
    const long long n_tiles = (n + OBS_SUB - 1) / OBS_SUB;
    const int oldest = head + 1 == HIST ? 0 : head + 1;
    constexpr int PER = (HIST + OBS_GROUPS - 1) / OBS_GROUPS;  // 7 slots per thread
    long long ooff[PER], coff[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int slot = oldest + g + OBS_GROUPS * i;
        slot = slot >= HIST ? slot - HIST : slot;
        ooff[i] = (long long)slot * 4 * n;
        coff[i] = (long long)slot * n;
    }
    pdl_launch_dependents();
    pdl_wait();  // everything below reads what the dynamics kernel of this step wrote

    float x[PER][4], xn[PER][4], t0 = 0.f, t1 = 0.f, t0n = 0.f, t1n = 0.f, inv = 0.f, invn = 0.f;
    double c[PER], cn[PER];
    auto fetch = [&](long long tile, float (&xo)[PER][4], double (&co)[PER], float &ta, float &tb, float &iv) {
        const long long env = tile * OBS_SUB + e;
        if (tile < n_tiles && env < n) {
            const float *obase = a.st.ohlv + env;
            const double *cbase = a.st.close + env;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                if (g + OBS_GROUPS * i < HIST) {
                    const float *o = obase + ooff[i];
                    xo[i][0] = __ldg(o);
                    xo[i][1] = __ldg(o + n);
                    xo[i][2] = __ldg(o + 2 * n);
                    xo[i][3] = __ldg(o + 3 * n);
                    co[i] = __ldg(cbase + coff[i]);
                }
            }
            const float *sc = a.st.scratch + env;
            ta = __ldg(sc + (long long)g * n);                    // features 250 + g        (g = 0..7)
            tb = (g < 3) ? __ldg(sc + (long long)(g + 8) * n) : 0.f;  // features 258, 259, 260  (g = 0..2)
            iv = __ldg(sc + 11 * n);
        }
    };
    fetch(blockIdx.x, x, c, t0, t1, inv);
    int it = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        float *buf = tiles + (size_t)(it % OBS_STAGES) * (OBS_SUB * OBS);
        const long long first = tile * OBS_SUB;
        const long long env = first + e;
        fetch(tile + gridDim.x, xn, cn, t0n, t1n, invn);  // next tile's values, consumed next iteration
        if (it >= OBS_STAGES) {  // the bulk copy that last used this buffer must have read it
            if (tid == 0) bulk_wait_read<OBS_STAGES - 1>();
            __syncthreads();
        }
        if (env < n) {
            float *dst = buf + e * OBS;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int k = g + OBS_GROUPS * i;
                if (k < HIST) {
                    dst[k * 5 + 0] = x[i][0] * inv;  // price_data / current_price, :513-515 (volume is divided too)
                    dst[k * 5 + 1] = x[i][1] * inv;
                    dst[k * 5 + 2] = x[i][2] * inv;
                    dst[k * 5 + 3] = (float)(c[i] * (double)inv);
                    dst[k * 5 + 4] = x[i][3] * inv;
                }
            }
            dst[250 + g] = t0;
            if (g < 3) dst[258 + g] = t1;
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const long long n_here = min((long long)OBS_SUB, n - first);
            const uint32_t bytes = (uint32_t)(n_here * OBS * sizeof(float));
            const uint32_t bulk = bytes & ~15u;
            if (bulk) bulk_store_s2g(a.io.obs + first * OBS, buf, bulk);
            bulk_commit();
            for (uint32_t i = bulk / 4; i < bytes / 4; ++i) a.io.obs[first * OBS + i] = buf[i];  // ragged last tile
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            x[i][0] = xn[i][0]; x[i][1] = xn[i][1]; x[i][2] = xn[i][2]; x[i][3] = xn[i][3];
            c[i] = cn[i];
        }
        t0 = t0n; t1 = t1n; inv = invn;
    }
    if (tid == 0) bulk_wait_read<0>();
}

template <bool IS_RESET>
int launch_split(const CArgs &a, cudaStream_t stream) {
    const int head = IS_RESET ? a.p.window_head : (a.p.window_head + 1 == HIST ? 0 : a.p.window_head + 1);
    cudaError_t e = launch_pdl(crypto_dyn_kernel<IS_RESET>, dim3((unsigned)((a.n + 255) / 256)), dim3(256), 0, stream, a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)OBS_STAGES * OBS_SUB * OBS * sizeof(float);
    e = cudaFuncSetAttribute(crypto_obs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long n_tiles = (a.n + OBS_SUB - 1) / OBS_SUB;
    long long grid = 2LL * device_sm_count();
    if (grid > n_tiles) grid = n_tiles;
    // (two-argument kernel: launch through the runtime's variadic form with the PDL attribute)
    {
        static const bool use_pdl = getenv("BENG_NO_PDL") == nullptr;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(OBS_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = use_pdl ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, crypto_obs_kernel, a, head);
    }
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)e;
}


template <int T>
constexpr size_t crypto_smem_bytes() {
    return (size_t)T * OBS * sizeof(float) + (size_t)(HIST - 1) * T * sizeof(double) + T * sizeof(float) + T;
}

template <bool IS_RESET>
int launch(const CArgs &a, cudaStream_t stream) {
    static int tile_env = -1;
    if (tile_env < 0) {
        tile_env = 0;
        if (const char *e = getenv("BENG_CRYPTO_TILE")) tile_env = atoi(e);
    }
    // Default: the fused two-phase kernel (237 us/step at 262,144 envs).  BENG_CRYPTO_MODE=split selects the two-kernel
    // step (277 us: its streaming observation kernel runs at 5.1 TB/s, but the stand-alone dynamics kernel loses the
    // overlap it enjoys inside the fused kernel); BENG_CRYPTO_TILE=32|64|128 the warp-specialised kernel (381 us).
    static int split = -1;
    if (split < 0) {
        const char *m = getenv("BENG_CRYPTO_MODE");
        split = (m && m[0] == 's') ? 1 : 0;
    }
    if (split && a.st.scratch) return launch_split<IS_RESET>(a, stream);
    if (tile_env <= 0) {
        const size_t smem = crypto2_smem_bytes();
        auto kern = crypto2_kernel<IS_RESET>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = launch_pdl(kern, dim3((unsigned)((a.n + C2_ENVS - 1) / C2_ENVS)), dim3(C2_ENVS), smem, stream, a);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return (int)e;
    }
    const int T = tile_env;
#define BENG_CCASE(TT)                                                                                          \
    if (T == TT) {                                                                                              \
        const size_t smem = crypto_smem_bytes<TT>();                                                            \
        auto kern = crypto_kernel<TT, IS_RESET>;                                                                \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return (int)e;                                                                    \
        kern<<<(unsigned)((a.n + TT - 1) / TT), 4 * TT, smem, stream>>>(a);                                     \
        return finish_launch();                                                                                 \
    }
    BENG_CCASE(32) BENG_CCASE(64) BENG_CCASE(128)
#undef BENG_CCASE
    return BENG_ERR_UNSUPPORTED;
}

int check(const beng_crypto_params *p, const beng_crypto_state *st, const beng_crypto_io *io, int64_t n) {
    if (!p || !st || !io || n < 0) return BENG_ERR_BAD_ARG;
    if (!st->scal || !st->meta || !st->ep_return || !st->close || !st->ohlv || !io->obs) return BENG_ERR_BAD_ARG;
    if (((uintptr_t)io->obs & 15) || ((uintptr_t)st->close & 7)) return BENG_ERR_BAD_ARG;
    if (p->window_head < 0 || p->window_head >= HIST) return BENG_ERR_BAD_ARG;
    if (p->autoreset_mode < 0 || p->autoreset_mode > 2 || p->action_type < 0 || p->action_type > 1) return BENG_ERR_BAD_ARG;
    if (p->max_steps < 1 || p->max_steps > 65535) return BENG_ERR_UNSUPPORTED;
    return 0;
}

}  // namespace
}  // namespace beng

extern "C" {

int beng_crypto_reset(const beng_crypto_params *p, const beng_crypto_state *st, const beng_crypto_io *io,
                      const uint8_t *mask_dev, int64_t n_envs, int32_t first_call, void *stream) {
    if (int rc = beng::check(p, st, io, n_envs)) return rc;
    if (n_envs == 0) return 0;
    beng::CArgs a{*p, *st, *io, nullptr, mask_dev, (long long)n_envs, first_call};
    return beng::launch<true>(a, (cudaStream_t)stream);
}

int beng_crypto_step(const beng_crypto_params *p, const beng_crypto_state *st, const void *actions_dev,
                     const beng_crypto_io *io, int64_t n_envs, void *stream) {
    if (int rc = beng::check(p, st, io, n_envs)) return rc;
    if (!actions_dev || !io->reward || !io->terminated) return BENG_ERR_BAD_ARG;
    if (n_envs == 0) return 0;
    beng::CArgs a{*p, *st, *io, actions_dev, nullptr, (long long)n_envs, 0};
    return beng::launch<false>(a, (cudaStream_t)stream);
}

int beng_crypto_step_host(const beng_crypto_params *p, const beng_crypto_state *st, void *actions_dev,
                          const beng_crypto_io *io, int64_t n_envs, const void *actions_host, float *obs_host,
                          float *reward_host, uint8_t *terminated_host, uint8_t *truncated_host, void *stream) {
    if (!actions_host || !actions_dev) return BENG_ERR_BAD_ARG;
    if (int rc = beng::check(p, st, io, n_envs)) return rc;
    if (truncated_host && !io->truncated) return BENG_ERR_BAD_ARG;
    if (n_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)n_envs;
    const size_t abytes = p->action_type == 1 ? n * 2 * sizeof(float) : n * sizeof(int64_t);
    cudaError_t e = cudaMemcpyAsync(actions_dev, actions_host, abytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    if (int rc = beng_crypto_step(p, st, actions_dev, io, n_envs, stream)) return rc;
#define BENG_D2H(dst, src, bytes) \
    if (dst) { e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s); if (e != cudaSuccess) return (int)e; }
    BENG_D2H(reward_host, io->reward, n * sizeof(float))
    BENG_D2H(terminated_host, io->terminated, n)
    BENG_D2H(truncated_host, io->truncated, n)
    BENG_D2H(obs_host, io->obs, n * BENG_CRYPTO_OBS_DIM * sizeof(float))
#undef BENG_D2H
    return 0;
}

}  // extern "C"

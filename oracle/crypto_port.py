"""Pure-Python/NumPy restatement of the reference CryptoTradingEnv -- the CPU-baseline "port".

TEST / BASELINE INFRASTRUCTURE.  `/root/reference` cannot travel to the GPU box, so the "reference's
pure-Python per-env step loop" timed beside the GPU number is this port.  It does the same per-step work in the
same interpreter and with the same NumPy calls as crypto_trading_env/crypto_trading_env.py (a list of candle
lists, np.array conversions in the observation, np.mean/np.std/np.diff/np.where for the indicators, and the
reference's O(n^2) MACD signal loop that re-runs both EMAs on every prefix, :94-100), so its speed is
representative; it is validated EXACTLY against tests/golden/crypto_golden.npz (tests/test_crypto_oracle.py).
"""
from __future__ import annotations

import random as _py_random

import numpy as np

_NEXT = {0: (2, 3), 1: (2, 4), 2: (0, 1), 3: (4, 1), 4: (0, 2)}   # regime successors, :168-174
_VOL = (1.2, 1.5, 0.8, 3.0, 2.0)                                   # :190-196
_TREND = (0.001, -0.001, 0.0, -0.005, 0.002)                       # :202-208


class _StdRng:
    """Default randomness: Python `random` + np.random.normal, like the reference."""
    random = staticmethod(_py_random.random)
    uniform = staticmethod(_py_random.uniform)
    choice = staticmethod(_py_random.choice)

    @staticmethod
    def normal(loc, scale):
        return np.random.normal(loc, scale)


def _ema(values, period):  # :107-119
    if len(values) == 0:
        return 0.0
    if len(values) < period:
        return np.mean(values)
    k = 2 / (period + 1)
    acc = values[0]
    for v in values[1:]:
        acc = (v * k) + (acc * (1 - k))
    return acc


def _rsi(p, period=14):  # :45-61
    if len(p) < period + 1:
        return 50.0
    d = np.diff(p)
    up = np.where(d > 0, d, 0)
    dn = np.where(d < 0, -d, 0)
    g, l = np.mean(up[-period:]), np.mean(dn[-period:])
    if l == 0:
        return 100.0
    return 100 - (100 / (1 + g / l))


def _macd(p, fast=12, slow=26, signal=9):  # :80-104
    if len(p) < slow:
        return 0.0, 0.0, 0.0
    line = _ema(p, fast) - _ema(p, slow)
    if len(p) < slow + signal:
        sig = 0.0
    else:
        hist = [_ema(p[:i], fast) - _ema(p[:i], slow) for i in range(slow, len(p) + 1)]
        sig = _ema(np.array(hist), signal)
    return line, sig, line - sig


def _bollinger(p, period=20, k=2):  # :64-77
    if len(p) < period:
        c = p[-1] if len(p) > 0 else 1000.0
        return c * 1.02, c, c * 0.98
    mid = np.mean(p[-period:])
    sd = np.std(p[-period:])
    return mid + (k * sd), mid, mid - (k * sd)


class CryptoPort:
    def __init__(self, rng=None, action_type="discrete", initial_balance=10000.0, fee=0.001, slippage=0.0005,
                 history=50, min_price=100.0, max_price=100000.0, vol_base=0.02, psych_factor=0.1, max_steps=1000):
        self.rng = rng if rng is not None else _StdRng
        self.action_type = action_type
        self.c = (initial_balance, fee, slippage, history, min_price, max_price, vol_base, psych_factor)
        self.max_steps = max_steps
        self.regime, self.trend_strength, self.psychology = 2, 0.0, 0.5   # MarketSimulator.__init__, :125-130
        self.cash, self.holdings, self.candles, self.t = initial_balance, 0.0, [], 0

    # MarketSimulator.generate_next_price + helpers, :132-221
    def _next_price(self, price, volume):
        r = self.rng
        if r.random() < 0.01:
            self.regime = r.choice(_NEXT[self.regime])
            if self.regime in (0, 4):
                self.trend_strength = r.uniform(0.5, 1.0)
            elif self.regime in (1, 3):
                self.trend_strength = r.uniform(-1.0, -0.5)
            else:
                self.trend_strength = r.uniform(-0.2, 0.2)
        vol = self.c[6] * _VOL[self.regime]
        drift = (self.psychology - 0.5) * self.c[7]
        trend = _TREND[self.regime] * self.trend_strength
        noise = r.normal(0, vol)
        pct = (trend + drift + noise) * (1.0 / (1.0 + volume * 0.1))
        new = np.clip(price * (1 + pct), self.c[4], self.c[5])
        self.psychology += pct * 10
        self.psychology = np.clip(self.psychology, 0.0, 1.0)
        self.psychology += (0.5 - self.psychology) * 0.01
        return new

    def reset(self, seed=None, options=None):  # :301-340
        self.cash, self.holdings, self.t, self.candles = self.c[0], 0.0, 0, []
        price, r = 50000.0, self.rng
        for _ in range(self.c[3]):
            volume = r.uniform(0.5, 2.0)
            price = self._next_price(price, volume)
            hi = price * r.uniform(1.0, 1.02)
            lo = price * r.uniform(0.98, 1.0)
            op = price * r.uniform(0.99, 1.01)
            self.candles.append([op, hi, lo, price, volume])
        return self._obs(), {}

    def _buy(self, amount, price):  # :449-476
        if amount <= 0 or self.cash < amount:
            return None
        eff = price + price * self.c[2] * self.rng.uniform(0.5, 1.5)
        fee = amount * self.c[1]
        self.cash -= amount
        self.holdings += (amount - fee) / eff
        return "buy"

    def _sell(self, qty, price):  # :478-503
        if qty <= 0 or self.holdings < qty:
            return None
        eff = price - price * self.c[2] * self.rng.uniform(0.5, 1.5)
        got = qty * eff
        self.holdings -= qty
        self.cash += got - got * self.c[1]
        return "sell"

    def step(self, action):  # :342-447
        price = self.candles[-1][3]
        before = self.cash + self.holdings * price
        trade = None
        if self.action_type == "continuous":
            b, s = action
            b = np.clip(b, 0, 1) * (self.cash * 0.1)
            s = np.clip(s, 0, 1) * (self.holdings * 0.1)
            if b > s and b > 0:
                trade = self._buy(b, price)
            elif s > 0:
                trade = self._sell(s, price)
        elif action == 1:
            trade = self._buy(self.cash * 0.05, price)
        elif action == 2:
            trade = self._buy(self.cash * 0.2, price)
        elif action == 3:
            trade = self._sell(self.holdings * 0.05, price)
        elif action == 4:
            trade = self._sell(self.holdings * 0.2, price)
        reward = (self.cash + self.holdings * price) - before
        if trade is None:
            reward -= 1.0
        r = self.rng
        volume = r.uniform(0.5, 2.0)
        new = self._next_price(price, volume)
        hi = new * r.uniform(1.0, 1.02)
        lo = new * r.uniform(0.98, 1.0)
        self.candles.append([price, hi, lo, new, volume])
        if len(self.candles) > self.c[3]:
            self.candles.pop(0)
        value = self.cash + self.holdings * new
        self.t += 1
        done = self.t >= self.max_steps or value <= 0 or value >= self.c[0] * 10
        return self._obs(), reward, done, False, {"portfolio_value": value, "current_price": new, "trade": trade}

    def _obs(self):  # :505-561
        data = np.array(self.candles)
        cur = data[-1, 3]
        out = list((data / cur).flatten())
        out += [self.cash / self.c[0], self.holdings * cur / self.c[0], (self.cash + self.holdings * cur) / self.c[0]]
        closes = np.array([row[3] for row in self.candles])
        out.append(_rsi(closes) / 100.0)
        line, sig, hist = _macd(closes)
        span = max(closes) - min(closes)
        out += [line / span, sig / span, hist / span] if span > 0 else [0.0, 0.0, 0.0]
        up, mid, lo = _bollinger(closes)
        out.append((cur - lo) / (up - lo) if up > lo else 0.5)
        out.append((up - lo) / mid if mid > 0 else 0.0)
        out.append((cur - mid) / mid if mid > 0 else 0.0)
        out.append(self.psychology)
        return np.array(out, dtype=np.float32)

/* Philox4x32-10 + draw contract, plain-C restatement for the CPU oracle.
 *
 * TEST INFRASTRUCTURE.  Written independently of csrc/beng_rng.cuh (the device version)
 * and of oracle/philox.py (numpy); the three are cross-checked in tests/test_rng.py against
 * the Random123 known-answer vectors.  Contract text: oracle/philox.py docstring.
 */
#ifndef BENG_ORACLE_RNG_H
#define BENG_ORACLE_RNG_H
#include <math.h>
#include <stdint.h>

static inline void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* A positioned per-env stream: (seed, env, stream, counter). */
typedef struct {
    uint64_t seed;
    uint64_t env;
    uint32_t stream;
    uint32_t counter; /* index of the next u32 draw */
} orc_stream;

static inline uint32_t orc_u32(orc_stream *s) {
    uint32_t ctr[4] = {s->counter >> 2, (uint32_t)s->env, (uint32_t)(s->env >> 32), s->stream};
    uint32_t key[2] = {(uint32_t)s->seed, (uint32_t)(s->seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[(s->counter++) & 3u];
}

static inline int64_t orc_randint(orc_stream *s, int64_t a, int64_t b) {
    return a + (int64_t)(((uint64_t)orc_u32(s) * (uint64_t)(b - a + 1)) >> 32);
}

static inline double orc_random(orc_stream *s) {
    uint32_t a = orc_u32(s) >> 5, b = orc_u32(s) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

static inline double orc_uniform(orc_stream *s, double a, double b) { return a + (b - a) * orc_random(s); }

static inline double orc_normal(orc_stream *s, double mu, double sd) {
    double u1 = orc_random(s), u2 = orc_random(s);
    return mu + sd * (sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2));
}

#endif

"""The RNG contract: three independent Philox4x32-10 restatements against the Random123 known answers."""
import numpy as np

from oracle import c_oracle, philox, replay

# Random123 kat_vectors, "philox4x32 10": counter, key -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_numpy_known_answers():
    for ctr, key, out in KAT:
        got = tuple(int(x) for x in philox.philox4x32_10(*ctr, *key))
        assert got == out


def test_philox_c_known_answers():
    # draws j=0..3 of env (c1 | c2<<32), stream c3, seed (k0 | k1<<32) with block index c0
    for ctr, key, out in KAT:
        env = ctr[1] | (ctr[2] << 32)
        seed = key[0] | (key[1] << 32)
        if ctr[0] > 0x3FFFFFFF:
            continue  # block index c0 = first >> 2 cannot reach 2^32-1 through a u32 draw index
        got = c_oracle.draws_u32(seed, env, ctr[3], ctr[0] * 4, 4)
        assert tuple(int(x) for x in got) == out


def test_numpy_c_replay_streams_agree():
    for seed, env, stream in [(0, 0, 0), (7, 12345, 1), (0x123456789ABC, (1 << 33) + 5, 0)]:
        a = philox.draws_u32(seed, [env], 3, 301, stream)[0]
        b = c_oracle.draws_u32(seed, env, stream, 3, 301)
        assert np.array_equal(a, b)
        rr = replay.ReplayRandom(seed, env, stream, start=3)
        c = [rr._u32() for _ in range(301)]
        assert np.array_equal(a, np.array(c, dtype=np.uint32))


def test_draw_maps():
    rr = replay.ReplayRandom(1, 2)
    u = philox.draws_u32(1, [2], 0, 64)[0]
    assert rr.randint(0, 19) == int(philox.randint_from_u32(u[0], 0, 19))
    assert rr.randint(5, 30) == int(philox.randint_from_u32(u[1], 5, 30))
    x = rr.random()
    assert x == float(philox.random53_from_u32(u[2], u[3])) and 0.0 <= x < 1.0
    y = rr.uniform(0.5, 2.0)
    assert 0.5 <= y < 2.0 and rr.counter == 6
    vals = [replay.ReplayRandom(3, e).randint(0, 3) for e in range(2000)]
    assert set(vals) == {0, 1, 2, 3}


def test_action_tape_numpy_vs_c():
    for n_choices, n_cols in [(4, 1), (3, 9), (5, 1)]:
        a = philox.action_tape(9, np.arange(100, 140, dtype=np.uint64), 17, 1, n_choices, n_cols)
        b = c_oracle.action_tape(9, 100, 40, 17, n_choices, n_cols)
        assert np.array_equal(a[:, 0] if n_cols == 1 else a[:, 0, :], b)
        assert a.min() >= 0 and a.max() < n_choices

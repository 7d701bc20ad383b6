"""Synthetic problem definitions for the BASELINE.json configs.

Host-side only: default weights, initial-state / goal samplers.  No solver arithmetic lives here.
The reference publishes no problem data, so these are this repo's frozen definitions.
"""
from __future__ import annotations

import numpy as np

MODEL_QUAD12, MODEL_DUAL24 = 0, 1
DT = 0.02


def dims(model: int):
    return (12, 4) if model == MODEL_QUAD12 else (24, 8)


def default_weights(model: int = MODEL_QUAD12):
    """Q = diag(10,10,10, 1,1,1, 5,5,5, .1,.1,.1) per vehicle, R = 0.1 I, Qf = 10 Q."""
    q = np.array([10.0] * 3 + [1.0] * 3 + [5.0] * 3 + [0.1] * 3)
    n, m = dims(model)
    Q = np.diag(np.tile(q, n // 12))
    R = 0.1 * np.eye(m)
    Qf = 10.0 * Q
    return Q, R, Qf


def pack_qrqf(Q, R, Qf):
    return np.concatenate([np.asarray(Q, float).ravel(), np.asarray(R, float).ravel(), np.asarray(Qf, float).ravel()])


def hover_input(model: int = MODEL_QUAD12):
    n, m = dims(model)
    return np.full(m, 1.0 * 9.81 / 4.0)


def sample_x0(batch: int, seed: int = 0x4E4F43, model: int = MODEL_QUAD12):
    """Config 2 / 4 initial states. pos U(-2,2), vel U(-1,1), roll/pitch U(-0.3,0.3), yaw U(-pi,pi),
    body rates U(-0.2,0.2) (narrower than SURVEY §8d's 0.5 so the 2 s open-loop nominal stays clear of
    the Euler-angle singularity)."""
    rng = np.random.default_rng(seed)
    n, m = dims(model)
    x = np.empty((batch, n))
    for v in range(n // 12):
        o = 12 * v
        x[:, o + 0:o + 3] = rng.uniform(-2, 2, (batch, 3))
        x[:, o + 3:o + 6] = rng.uniform(-1, 1, (batch, 3))
        x[:, o + 6:o + 8] = rng.uniform(-0.3, 0.3, (batch, 2))
        x[:, o + 8] = rng.uniform(-np.pi, np.pi, batch)
        x[:, o + 9:o + 12] = rng.uniform(-0.2, 0.2, (batch, 3))
    if model == MODEL_DUAL24:
        x[:, 12] += 1.0  # second vehicle near the spring rest offset
    return x


# ---- SURVEY.md §8d's generator, literally: std::mt19937_64 seeded 0x4E4F43 + instance_id, twelve draws of
# std::uniform_real_distribution<double> per instance in state order, body rates U(-0.5, 0.5).
_MT_F = np.uint64(6364136223846793005)
_MT_A = np.uint64(0xB5026F5AA96619E9)
_MT_UP, _MT_LO = np.uint64(0xFFFFFFFF80000000), np.uint64(0x7FFFFFFF)


def mt19937_64_first_draws(seeds, count: int = 12):
    """First `count` (<= 156) outputs of std::mt19937_64 for every seed in `seeds` (vectorised over the seeds).
    Only state words 0 .. count+156 are ever touched by the first `count` outputs, so the 312-word state is not
    built in full."""
    assert 1 <= count <= 156
    seeds = np.asarray(seeds, dtype=np.uint64)
    L = count + 157
    st = np.empty((seeds.size, L), dtype=np.uint64)
    st[:, 0] = seeds
    with np.errstate(over="ignore"):
        for i in range(1, L):
            prev = st[:, i - 1]
            st[:, i] = _MT_F * (prev ^ (prev >> np.uint64(62))) + np.uint64(i)
        y = (st[:, :count] & _MT_UP) | (st[:, 1:count + 1] & _MT_LO)      # the twist reads the OLD word i+1
        x = st[:, 156:156 + count] ^ (y >> np.uint64(1)) ^ np.where(y & np.uint64(1), _MT_A, np.uint64(0))
    x = x ^ ((x >> np.uint64(29)) & np.uint64(0x5555555555555555))
    x = x ^ ((x << np.uint64(17)) & np.uint64(0x71D67FFFEDA60000))
    x = x ^ ((x << np.uint64(37)) & np.uint64(0xFFF7EEE000000000))
    return x ^ (x >> np.uint64(43))


def uniform_real_from_u64(x, lo, hi):
    """libstdc++'s uniform_real_distribution<double>(lo, hi) on a 64-bit engine: canonical = double(x) / 2^64
    (one draw; a result of 1.0 is replaced by the largest double below 1), value = canonical * (hi - lo) + lo."""
    c = x.astype(np.float64) * (1.0 / 18446744073709551616.0)
    c = np.where(c >= 1.0, np.nextafter(1.0, 0.0), c)
    return c * (hi - lo) + lo


def sample_x0_survey(batch: int, first_instance: int = 0, seed0: int = 0x4E4F43):
    """SURVEY.md §8d config 2 / 4 initial states: instance b draws x0 from std::mt19937_64(seed0 + b) in state
    order: pos U(-2,2), vel U(-1,1), roll/pitch U(-0.3,0.3), yaw U(-pi,pi), body rates U(-0.5,0.5).
    (tests/test_problems.py cross-checks the generator against a g++ program using <random>.)"""
    lo = np.array([-2.0] * 3 + [-1.0] * 3 + [-0.3, -0.3, -np.pi] + [-0.5] * 3)
    out = np.empty((batch, 12))
    for b0 in range(0, batch, 16384):
        b1 = min(batch, b0 + 16384)
        seeds = np.arange(first_instance + b0, first_instance + b1, dtype=np.uint64) + np.uint64(seed0)
        out[b0:b1] = uniform_real_from_u64(mt19937_64_first_draws(seeds, 12), lo, -lo)
    return out


def sample_goals(batch: int, seed: int = 0x534C51, model: int = MODEL_QUAD12):
    """Config 3 / 5 waypoint goals: pos U(-3,3)^3, yaw U(-pi/2,pi/2), everything else zero."""
    rng = np.random.default_rng(seed)
    n, m = dims(model)
    g = np.zeros((batch, n))
    for v in range(n // 12):
        o = 12 * v
        g[:, o + 0:o + 3] = rng.uniform(-3, 3, (batch, 3))
        g[:, o + 8] = rng.uniform(-np.pi / 2, np.pi / 2, batch)
    if model == MODEL_DUAL24:
        g[:, 12:15] = g[:, 0:3] + np.array([1.0, 0.0, 0.0])
        g[:, 20] = g[:, 8]
    return g


def hover_state(batch: int, model: int = MODEL_QUAD12):
    n, m = dims(model)
    x = np.zeros((batch, n))
    if model == MODEL_DUAL24:
        x[:, 12] = 1.0
    return x

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

"""Shared helpers for the lattice tests: synthetic inputs and a brute-force
path enumerator that is independent of any recursion (used to harden the authored spec)."""
import itertools

import numpy as np


def make_inputs(B, T, U, seed=1234, K=None):
    """z ~ N(0,1); log_emit = log sigmoid(z), log_shift = log sigmoid(-z)  (a proper 2-way
    distribution, like the reference's ln[0.8, 0.2] rows, tests/test_decoding.rs:25-29)."""
    rng = np.random.default_rng(seed)
    shape = (B, T, U) if K is None else (B, T, U, K)
    z = rng.standard_normal(shape)
    le = -np.logaddexp(0.0, -z)
    ls = -np.logaddexp(0.0, z)
    if K is None:
        return le.astype(np.float32), ls.astype(np.float32)
    lt = rng.standard_normal((B, U, K))
    lt = lt - np.logaddexp.reduce(lt, axis=-1, keepdims=True)
    return le.astype(np.float32), ls.astype(np.float32), lt.astype(np.float32)


def ragged_lengths(B, T, U, seed=99):
    """T_b ~ U{ceil(.6T)..T}, U_b ~ U{ceil(.6U)..min(U,T_b)}."""
    rng = np.random.default_rng(seed)
    t_len = rng.integers(int(np.ceil(0.6 * T)), T + 1, size=B)
    u_len = np.array([rng.integers(min(int(np.ceil(0.6 * U)), min(U, tb)), min(U, tb) + 1)
                      for tb in t_len])
    return t_len.astype(np.int32), u_len.astype(np.int32)


def brute_force_ll(le, ls, T, U):
    """Sum over all monotonic paths: choose the U-1 frames (out of the first T-1) that shift."""
    if U > T or T <= 0 or U <= 0:
        return -np.inf
    total = -np.inf
    for shifts in itertools.combinations(range(T - 1), U - 1):
        sset = set(shifts)
        u, lp = 0, 0.0
        for t in range(T - 1):
            if t in sset:
                lp += float(ls[t, u])
                u += 1
            else:
                lp += float(le[t, u])
        lp += float(le[T - 1, U - 1])
        total = np.logaddexp(total, lp)
    return total


def brute_force_tone_ll(le, ls, lt, T, U, K):
    total = -np.inf
    for shifts in itertools.combinations(range(T - 1), U - 1):
        sset = set(shifts)
        for tones in itertools.product(range(K), repeat=U):
            u, lp = 0, float(lt[0, tones[0]])
            for t in range(T - 1):
                if t in sset:
                    lp += float(ls[t, u, tones[u]])
                    u += 1
                    lp += float(lt[u, tones[u]])
                else:
                    lp += float(le[t, u, tones[u]])
            lp += float(le[T - 1, U - 1, tones[U - 1]])
            total = np.logaddexp(total, lp)
    return total

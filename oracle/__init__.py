"""ctypes/numpy front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``ssnt-tts-rust_b200``) never does.  See the header of ``oracle/ssnt_oracle.cpp`` for what
is pinned by the reference's own golden vectors and what is "parity unpinned".

Every wrapper mirrors the argument order of the C symbol of the same name in the reference's
``ssnt_tts_c/src/lib.rs`` and returns freshly allocated numpy arrays.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_bool, c_float, c_int

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssnt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (seconds). Building the checker is not using it."""
    src = os.path.join(_HERE, "ssnt_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ssnt_oracle_get_threads.restype = c_int
        _lib.oracle_ssnt_tts_v2_beam_search_decode_checked.restype = c_int
        _lib.oracle_ssnt_upsample_source_indexes_checked.restype = c_int
    return _lib


def set_threads(n: int) -> None:
    lib().ssnt_oracle_set_threads(c_int(n))


def get_threads() -> int:
    return int(lib().ssnt_oracle_get_threads())


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _b(a):
    return np.ascontiguousarray(a, dtype=np.bool_)


def _p(a, ty):
    return a.ctypes.data_as(POINTER(ty)) if a is not None else None


def beam_search_decode(h, log_prob_history, is_finished, t, u, max_t, beam_width):
    """v1 step, single batch. ssnt_tts_c/src/lib.rs:10-83."""
    h, lph, fin, t, u = _f(h), _f(log_prob_history), _b(is_finished), _i(t), _i(u)
    W = int(beam_width)
    pred, lp = np.full(W, -1, np.int32), np.zeros(W, np.float32)
    nt, nu, nf, bb = (np.zeros(W, np.int32), np.zeros(W, np.int32), np.zeros(W, np.bool_),
                      np.zeros(W, np.int32))
    lib().oracle_ssnt_tts_beam_search_decode(
        _p(h, c_float), _p(lph, c_float), _p(fin, c_bool), _p(t, c_int), _p(u, c_int),
        c_int(max_t), c_int(W), _p(pred, c_int), _p(lp, c_float), _p(nt, c_int), _p(nu, c_int),
        _p(nf, c_bool), _p(bb, c_int))
    return pred, lp, nt, nu, nf, bb


def extract_best_beam_branch(best_final_branch, beam_branch, t_history, beam_width):
    """ssnt_tts_c/src/lib.rs:86-116."""
    bb, th = _i(beam_branch), _i(t_history)
    max_u = bb.shape[0]
    ob, ot = np.zeros(max_u, np.int32), np.zeros(max_u, np.int32)
    lib().oracle_ssnt_extract_best_beam_branch(
        c_int(int(best_final_branch)), _p(bb, c_int), _p(th, c_int), c_int(beam_width),
        c_int(max_u), _p(ob, c_int), _p(ot, c_int))
    return ob, ot


def ssnt_tts_v2_beam_search_decode(h, log_prob_history, is_finished, total_duration,
                                   duration_table, t, u, input_length, output_length, beam_width,
                                   duration_class_size, zero_duration_id, allow_skip, test_mode,
                                   checked=True):
    """v2 step. ssnt_tts_c/src/lib.rs:118-218.  checked=True returns the number of batch
    entries that hit the src/v2.rs:292 panic as an extra trailing value instead of aborting."""
    h, lph, fin = _f(h), _f(log_prob_history), _b(is_finished)
    td, dt, t, u = _i(total_duration), _i(duration_table), _i(t), _i(u)
    il, ol = _i(input_length), _i(output_length)
    B, W = lph.shape
    pred = np.full((B, W), zero_duration_id, np.int32)
    lp = np.zeros((B, W), np.float32)
    nt, nu, ntd, bb = (np.zeros((B, W), np.int32) for _ in range(4))
    nf = np.zeros((B, W), np.bool_)
    args = (_p(h, c_float), _p(lph, c_float), _p(fin, c_bool), _p(td, c_int), _p(dt, c_int),
            _p(t, c_int), _p(u, c_int), _p(il, c_int), _p(ol, c_int), c_int(B), c_int(W),
            c_int(duration_class_size), c_int(zero_duration_id), c_bool(allow_skip),
            c_bool(test_mode), _p(pred, c_int), _p(lp, c_float), _p(nt, c_int), _p(nu, c_int),
            _p(nf, c_bool), _p(ntd, c_int), _p(bb, c_int))
    if checked:
        bad = lib().oracle_ssnt_tts_v2_beam_search_decode_checked(*args)
        return pred, lp, nt, nu, nf, ntd, bb, int(bad)
    lib().oracle_ssnt_tts_v2_beam_search_decode(*args)
    return pred, lp, nt, nu, nf, ntd, bb


def order_beam_branch(final_branch, beam_branch, beam_width):
    """ssnt_tts_c/src/lib.rs:220-241. beam_branch (B,T,W) → (B,W,T)."""
    fb, bb = _i(final_branch), _i(beam_branch)
    B, T, W = bb.shape
    out = np.zeros((B, W, T), np.int32)
    lib().oracle_ssnt_order_beam_branch(_p(fb, c_int), _p(bb, c_int), c_int(B), c_int(W),
                                        c_int(T), _p(out, c_int))
    return out


def upsample_source_indexes(duration, output_length, out_of_range_source_index, beam_width,
                            max_u=None, checked=True):
    """ssnt_tts_c/src/lib.rs:244-265; pre-fill as upsample_source_indexes_op.cc:75 does."""
    d, ol = _i(duration), _i(output_length)
    B, W, T = d.shape
    if max_u is None:
        max_u = int(ol.max()) if ol.size else 0
    out = np.full((B, W, max_u), out_of_range_source_index, np.int32)
    if checked:
        bad = lib().oracle_ssnt_upsample_source_indexes_checked(
            _p(d, c_int), _p(ol, c_int), c_int(B), c_int(W), c_int(T), c_int(max_u), _p(out, c_int))
        return out, int(bad)
    lib().oracle_ssnt_upsample_source_indexes(
        _p(d, c_int), _p(ol, c_int), c_int(B), c_int(W), c_int(T), c_int(max_u), _p(out, c_int))
    return out


def tone_latent_beam_search_decode(h, log_prob_history, is_finished, t, u, input_length,
                                   beam_width, tone_class_size, empty_tone_id):
    """ssnt_tts_c/src/lib.rs:267-343."""
    h, lph, fin, t, u, il = _f(h), _f(log_prob_history), _b(is_finished), _i(t), _i(u), _i(input_length)
    B, W = lph.shape
    pred = np.full((B, W), empty_tone_id, np.int32)
    lp = np.zeros((B, W), np.float32)
    nt, nu, bb = (np.zeros((B, W), np.int32) for _ in range(3))
    nf = np.zeros((B, W), np.bool_)
    lib().oracle_tone_latent_beam_search_decode(
        _p(h, c_float), _p(lph, c_float), _p(fin, c_bool), _p(t, c_int), _p(u, c_int),
        _p(il, c_int), c_int(B), c_int(W), c_int(tone_class_size), c_int(empty_tone_id),
        _p(pred, c_int), _p(lp, c_float), _p(nt, c_int), _p(nu, c_int), _p(nf, c_bool),
        _p(bb, c_int))
    return pred, lp, nt, nu, nf, bb


def levenshtein_edit_distance(a, b, a_lengths, b_lengths):
    """ssnt_tts_c/src/lib.rs:346-381."""
    a, b, al, bl = _i(a), _i(b), _i(a_lengths), _i(b_lengths)
    B, L = a.shape
    out = np.zeros(B, np.int32)
    lib().oracle_tone_latent_levenshtein_edit_distance(
        _p(a, c_int), _p(b, c_int), _p(al, c_int), _p(bl, c_int), c_int(B), c_int(L), _p(out, c_int))
    return out


def forward_backward(log_emit, log_shift, t_len=None, u_len=None, precision="f64", grads=True):
    """Authored lattice spec (SURVEY.md §8 a-FB). Returns ll[B], loss, grad_emit, grad_shift."""
    le, ls = _f(log_emit), _f(log_shift)
    B, T, U = le.shape
    tl = _i(t_len) if t_len is not None else None
    ul = _i(u_len) if u_len is not None else None
    ll, loss = np.zeros(B, np.float32), np.zeros(1, np.float32)
    ge = np.empty_like(le) if grads else None
    gs = np.empty_like(le) if grads else None
    lib().oracle_ssnt_tts_forward_backward(
        _p(le, c_float), _p(ls, c_float), _p(tl, c_int), _p(ul, c_int), c_int(B), c_int(T),
        c_int(U), c_int(1 if precision == "f64" else 0), _p(ll, c_float), _p(loss, c_float),
        _p(ge, c_float), _p(gs, c_float))
    return ll, float(loss[0]), ge, gs


def tone_latent_forward_backward(log_emit, log_shift, log_tone, t_len=None, u_len=None,
                                 precision="f64", grads=True):
    """Authored tone-latent lattice (SURVEY.md §8 a-TL)."""
    le, ls, lt = _f(log_emit), _f(log_shift), _f(log_tone)
    B, T, U, K = le.shape
    tl = _i(t_len) if t_len is not None else None
    ul = _i(u_len) if u_len is not None else None
    ll, loss = np.zeros(B, np.float32), np.zeros(1, np.float32)
    ge = np.empty_like(le) if grads else None
    gs = np.empty_like(le) if grads else None
    gt = np.empty_like(lt) if grads else None
    lib().oracle_tone_latent_forward_backward(
        _p(le, c_float), _p(ls, c_float), _p(lt, c_float), _p(tl, c_int), _p(ul, c_int),
        c_int(B), c_int(T), c_int(U), c_int(K), c_int(1 if precision == "f64" else 0),
        _p(ll, c_float), _p(loss, c_float), _p(ge, c_float), _p(gs, c_float), _p(gt, c_float))
    return ll, float(loss[0]), ge, gs, gt

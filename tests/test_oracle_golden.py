"""Pins the CPU oracle against every golden vector the reference's own tests hold
(SURVEY.md §8c).  CPU only."""
import numpy as np

import golden_vectors as G


def test_edit_distance_scalar_cases(oracle_mod):
    # tests/test_edit_distance.rs:9-69 (test_edit_distance0..3)
    for a, b, want in G.EDIT_CASES:
        L = max(len(a), len(b), 1)
        A = np.full((1, L), -7, np.int32)
        B = np.full((1, L), -9, np.int32)
        A[0, :len(a)] = a
        B[0, :len(b)] = b
        got = oracle_mod.levenshtein_edit_distance(A, B, [len(a)], [len(b)])
        assert got.tolist() == [want], (a, b)


def test_edit_distance_batched(oracle_mod):
    # tests/test_edit_distance.rs:71-107
    got = oracle_mod.levenshtein_edit_distance(G.EDIT_BATCH_A, G.EDIT_BATCH_B,
                                               G.EDIT_BATCH_A_LEN, G.EDIT_BATCH_B_LEN)
    np.testing.assert_array_equal(got, G.EDIT_BATCH_EXPECTED)


def test_extract_best_beam_branch(oracle_mod):
    # tests/test_decoding.rs:53-131 (t_history aliased to the same table, as there)
    branch, t_hist = oracle_mod.extract_best_beam_branch(
        G.BACKTRACE_FINAL, G.BACKTRACE_TABLE, G.BACKTRACE_TABLE, G.BACKTRACE_BEAM_WIDTH)
    np.testing.assert_array_equal(branch, G.BACKTRACE_EXPECTED)
    # t_history[u][branch[u]] is by construction the parent of row u, i.e. branch[u-1]
    np.testing.assert_array_equal(t_hist[1:], G.BACKTRACE_EXPECTED[:-1])


def test_order_beam_branch_matches_single_walk(oracle_mod):
    # src/v2_util.rs:6-36 is the same walk for every final w; pin it on the 60x10 table.
    table = G.BACKTRACE_TABLE[None]  # (B=1, T=60, W=10)
    final = np.arange(10, dtype=np.int32)[None]
    out = oracle_mod.order_beam_branch(final, table, 10)
    np.testing.assert_array_equal(out[0, 9], G.BACKTRACE_EXPECTED)
    for w in range(10):
        b, _ = oracle_mod.extract_best_beam_branch(w, G.BACKTRACE_TABLE, G.BACKTRACE_TABLE, 10)
        np.testing.assert_array_equal(out[0, w], b)


def test_upsample_source_indexes(oracle_mod):
    # ssnt-tts-tensorflow/tests/test_upsample_source_indexes.py:13-53
    out, bad = oracle_mod.upsample_source_indexes(G.UPSAMPLE_DURATION, G.UPSAMPLE_OUTPUT_LENGTH,
                                                  G.UPSAMPLE_FILL, 2)
    assert bad == 0
    np.testing.assert_array_equal(out, G.UPSAMPLE_EXPECTED)


def test_upsample_flags_length_mismatch(oracle_mod):
    # src/v2_util.rs:58 assert_eq!
    ol = G.UPSAMPLE_OUTPUT_LENGTH.copy()
    ol[1, 0] = 9
    _, bad = oracle_mod.upsample_source_indexes(G.UPSAMPLE_DURATION, ol, -1, 2, max_u=11)
    assert bad == 1


def test_v1_smoke_hand_derived(oracle_mod):
    # tests/test_decoding.rs:13-51 prints only; expectations hand-derived.
    S = G.V1_SMOKE
    W = S["beam_width"]
    z = np.zeros(W, np.int32)
    fin = np.zeros(W, np.bool_)
    pred, lp, nt, nu, nf, bb = oracle_mod.beam_search_decode(S["h"], np.zeros(W, np.float32), fin,
                                                             z, z, S["max_t"], W)
    e = S["step1"]
    assert pred.tolist() == e["prediction"] and nt.tolist() == e["next_t"]
    assert nu.tolist() == e["next_u"] and bb.tolist() == e["parent"]
    assert nf.tolist() == e["finished"]
    np.testing.assert_array_equal(lp, np.array(e["log_prob"], np.float32))
    pred, lp2, nt, nu, nf, bb = oracle_mod.beam_search_decode(S["h"], lp, fin, z, z, S["max_t"], W)
    e = S["step2"]
    assert pred.tolist() == e["prediction"] and nt.tolist() == e["next_t"]
    assert nu.tolist() == e["next_u"] and bb.tolist() == e["parent"]
    np.testing.assert_array_equal(lp2, np.array(e["log_prob"], np.float32))


def test_v1_last_position_rules(oracle_mod):
    # src/lib.rs:187-205: at t == max_t-1 Emit finishes, Shift is rewritten to a no-add Emit.
    h = np.log(np.array([[0.6, 0.4], [0.3, 0.7]], np.float32))
    hist = np.array([-1.0, -2.0], np.float32)
    t = np.array([3, 3], np.int32)
    u = np.array([5, 6], np.int32)
    pred, lp, nt, nu, nf, bb = oracle_mod.beam_search_decode(h, hist, [False, False], t, u, 4, 2)
    # candidates: w0:E(-1+ln.6,fin) w0:S->E(-1,fin) w1:E(-2+ln.3) w1:S->E(-2) ; sort desc
    assert pred.tolist() == [0, 0]
    np.testing.assert_array_equal(lp, np.array([-1.0, np.float32(-1.0) + h[0, 0]], np.float32))
    assert nt.tolist() == [3, 3] and nu.tolist() == [5, 5] and nf.tolist() == [True, True]
    assert bb.tolist() == [0, 0]


def test_v1_finished_and_out_of_range_beams(oracle_mod):
    # src/lib.rs:175-184 filler: prediction Emit, log-prob carried, (t,u) unchanged, finished.
    h = np.log(np.array([[0.5, 0.5], [0.9, 0.1], [0.2, 0.8]], np.float32))
    hist = np.array([-0.5, -3.0, -0.25], np.float32)
    pred, lp, nt, nu, nf, bb = oracle_mod.beam_search_decode(
        h, hist, [True, False, False], [1, 7, 2], [4, 9, 3], 5, 3)
    # w0 finished → filler(-0.5); w1: t=7 >= max_t → filler(-3.0); w2 live: E(-0.25+ln.2), S(-0.25+ln.8)
    assert bb.tolist() == [2, 0, 2]
    assert pred.tolist() == [1, 0, 0]
    assert nf.tolist() == [False, True, False]
    assert nt.tolist() == [3, 1, 2] and nu.tolist() == [4, 4, 4]

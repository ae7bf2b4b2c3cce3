"""GPU parity for the reference's seven operators: every output array must equal the CPU
oracle's exactly (ints, bools AND the float32 log-probs bit for bit), on the golden vectors
the reference's tests hold and on seeded random cases it does not cover (v2 band pruning,
diagonal injection, zero-duration skip, tone-latent, B>1, padding, ties).  Everything goes
through the C-ABI, with host pointers and with device pointers."""
import numpy as np
import pytest

import golden_vectors as G

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _eq(got, want, what=""):
    got, want = _np(got), np.asarray(want)
    assert got.dtype == want.dtype or got.dtype.kind == want.dtype.kind, what
    if want.dtype.kind == "f":
        np.testing.assert_array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32), err_msg=what)
    else:
        np.testing.assert_array_equal(got, want, err_msg=what)


@pytest.fixture(params=["host", "device"])
def space(request):
    return request.param


def _conv(space, *arrs):
    return tuple(_dev(a) for a in arrs) if space == "device" else arrs


# ---------------------------------------------------------------- edit distance
def test_edit_distance_golden(product, space):
    for a, b, want in G.EDIT_CASES:  # tests/test_edit_distance.rs:9-69
        L = max(len(a), len(b), 1)
        A = np.full((1, L), -7, np.int32)
        B = np.full((1, L), -9, np.int32)
        A[0, :len(a)] = a
        B[0, :len(b)] = b
        al, bl = np.array([len(a)], np.int32), np.array([len(b)], np.int32)
        got = product.levenshtein_edit_distance(*_conv(space, A, B, al, bl))
        assert _np(got).tolist() == [want], (a, b)
    got = product.levenshtein_edit_distance(*_conv(space, G.EDIT_BATCH_A, G.EDIT_BATCH_B,
                                                   G.EDIT_BATCH_A_LEN, G.EDIT_BATCH_B_LEN))
    _eq(got, G.EDIT_BATCH_EXPECTED)  # tests/test_edit_distance.rs:71-107


@pytest.mark.parametrize("B,L,vocab", [(64, 150, 50), (7, 33, 3), (5, 1000, 8), (3, 1500, 50), (4, 1, 2)])
def test_edit_distance_random(product, oracle_mod, space, B, L, vocab):
    rng = np.random.default_rng(B * 1000 + L)
    a = rng.integers(0, vocab, (B, L)).astype(np.int32)
    b = a.copy()
    mask = rng.random((B, L)) < 0.3
    b[mask] = rng.integers(0, vocab, mask.sum())
    al = rng.integers(0, L + 1, B).astype(np.int32)
    bl = rng.integers(0, L + 1, B).astype(np.int32)
    al[0], bl[0] = L, L
    if B > 1:
        al[1], bl[1] = 0, L
    want = oracle_mod.levenshtein_edit_distance(a, b, al, bl)
    got = product.levenshtein_edit_distance(*_conv(space, a, b, al, bl))
    _eq(got, want)


@pytest.mark.parametrize("L", [31, 32, 33, 64, 65, 96, 257, 1023, 1024, 1025])
def test_edit_distance_block_edges(product, oracle_mod, space, L):
    """The bit-parallel kernel cuts sequence a into 32-row blocks, one per lane (max_length <= 1024; 1025 takes the
    wavefront kernel): lengths on either side of every block edge, identical / disjoint / shifted pairs, a two-symbol
    alphabet and arbitrary int32 symbols (negative, large) — bit-exact against src/edit_distance.rs:6-60."""
    rng = np.random.default_rng(L)
    lens = sorted({0, 1, 2, 31, 32, 33, 63, 64, 65, L // 2, L - 33, L - 32, L - 31, L - 1, L} & set(range(L + 1)))
    pairs = [(m, n) for m in lens for n in lens]
    pairs = [pairs[i] for i in rng.permutation(len(pairs))[:48]] + [(L, L), (L, 1), (1, L), (L, L), (L, L), (L, L)]
    B = len(pairs)
    a = rng.integers(0, 2, (B, L)).astype(np.int32)
    b = rng.integers(0, 2, (B, L)).astype(np.int32)
    wide = rng.integers(-2**31, 2**31 - 1, (B, L), dtype=np.int64).astype(np.int32)
    a[8:16] = wide[8:16]
    b[8:16] = np.where(rng.random((8, L)) < 0.2, wide[16:24], wide[8:16])
    a[-3], b[-3] = 7, 7                       # identical: 0
    a[-2], b[-2] = np.arange(L), np.arange(L) + L   # disjoint: L
    a[-1] = rng.integers(0, 5, L)
    b[-1] = np.roll(a[-1], 3)                 # shifted by three
    al = np.array([m for m, _ in pairs], np.int32)
    bl = np.array([n for _, n in pairs], np.int32)
    want = oracle_mod.levenshtein_edit_distance(a, b, al, bl)
    assert want[-3] == 0 and want[-2] == L
    got = product.levenshtein_edit_distance(*_conv(space, a, b, al, bl))
    _eq(got, want)


# ---------------------------------------------------------------- back-trace / upsample
def test_extract_best_beam_branch_golden(product, space):
    bb, th = _conv(space, G.BACKTRACE_TABLE, G.BACKTRACE_TABLE)
    branch, t_hist = product.extract_best_beam_branch(G.BACKTRACE_FINAL, bb, th, G.BACKTRACE_BEAM_WIDTH)
    _eq(branch, G.BACKTRACE_EXPECTED)  # tests/test_decoding.rs:125-130
    _eq(_np(t_hist)[1:], G.BACKTRACE_EXPECTED[:-1])


@pytest.mark.parametrize("max_u,W", [(1, 1), (5, 3), (60, 10), (1000, 8), (333, 32), (40, 300)])
def test_extract_best_beam_branch_random(product, oracle_mod, space, max_u, W):
    rng = np.random.default_rng(max_u * 7 + W)
    bb = rng.integers(0, W, (max_u, W)).astype(np.int32)
    th = rng.integers(0, 500, (max_u, W)).astype(np.int32)
    final = int(rng.integers(0, W))
    want_b, want_t = oracle_mod.extract_best_beam_branch(final, bb, th, W)
    got_b, got_t = product.extract_best_beam_branch(final, *_conv(space, bb, th), W)
    _eq(got_b, want_b)
    _eq(got_t, want_t)


@pytest.mark.parametrize("B,T,W", [(1, 60, 10), (64, 150, 8), (3, 1, 4), (5, 1000, 8), (2, 17, 40)])
def test_order_beam_branch(product, oracle_mod, space, B, T, W):
    rng = np.random.default_rng(B + T + W)
    bb = rng.integers(0, W, (B, T, W)).astype(np.int32)
    final = np.stack([rng.permutation(W) for _ in range(B)]).astype(np.int32)
    want = oracle_mod.order_beam_branch(final, bb, W)
    got = product.order_beam_branch(*_conv(space, final, bb), W)
    _eq(got, want)


def test_upsample_golden(product, space):
    got = product.upsample_source_indexes(*_conv(space, G.UPSAMPLE_DURATION, G.UPSAMPLE_OUTPUT_LENGTH),
                                          G.UPSAMPLE_FILL, 2)
    _eq(got, G.UPSAMPLE_EXPECTED)  # test_upsample_source_indexes.py:40-53


@pytest.mark.parametrize("B,W,T,dmax", [(64, 8, 150, 12), (2, 3, 700, 3), (1, 1, 1, 5), (3, 2, 40, 0)])
def test_upsample_random(product, oracle_mod, space, B, W, T, dmax):
    rng = np.random.default_rng(B * 31 + T)
    d = rng.integers(0, dmax + 1, (B, W, T)).astype(np.int32)
    ol = d.sum(axis=2).astype(np.int32)
    max_u = int(ol.max()) + 3
    want, bad = oracle_mod.upsample_source_indexes(d, ol, -5, W, max_u=max_u)
    assert bad == 0
    got = product.upsample_source_indexes(*_conv(space, d, ol), -5, W, max_u=max_u)
    _eq(got, want)
    # max_u smaller than some rows: writes are clipped to the row (zip with the max_u chunk)
    if max_u > 6:
        want2, _ = oracle_mod.upsample_source_indexes(d, ol, -5, W, max_u=max_u // 2)
        got2 = product.upsample_source_indexes(*_conv(space, d, ol), -5, W, max_u=max_u // 2)
        _eq(got2, want2)


def test_upsample_length_mismatch_raises_flag(product):
    # src/v2_util.rs:58 assert_eq! — device-pointer calls raise the flag instead of aborting
    d, ol = _dev(G.UPSAMPLE_DURATION), G.UPSAMPLE_OUTPUT_LENGTH.copy()
    ol[2, 1] = 12
    product.last_error()
    out = product.upsample_source_indexes(d, _dev(ol), -1, 2, max_u=12)
    assert product.last_error() & product.ERR_UPSAMPLE_LENGTH
    out = _np(out)
    assert (out[2, 1] == -1).all()                  # the failing row is left untouched
    np.testing.assert_array_equal(out[0, 0, :6], G.UPSAMPLE_EXPECTED[0, 0, :6])
    assert product.last_error() == 0


# ---------------------------------------------------------------- beam steps
def test_v1_smoke_hand_derived(product, space):
    S = G.V1_SMOKE  # tests/test_decoding.rs:13-51 (prints only there)
    W = S["beam_width"]
    z, fin, lph = np.zeros(W, np.int32), np.zeros(W, np.bool_), np.zeros(W, np.float32)
    h = S["h"]
    pred, lp, nt, nu, nf, bb = product.beam_search_decode(*_conv(space, h, lph, fin, z, z), S["max_t"], W)
    e = S["step1"]
    assert _np(pred).tolist() == e["prediction"] and _np(nt).tolist() == e["next_t"]
    assert _np(nu).tolist() == e["next_u"] and _np(bb).tolist() == e["parent"]
    assert _np(nf).tolist() == e["finished"]
    _eq(lp, np.array(e["log_prob"], np.float32))
    pred, lp2, nt, nu, nf, bb = product.beam_search_decode(*_conv(space, h, _np(lp), fin, z, z), S["max_t"], W)
    e = S["step2"]
    assert _np(pred).tolist() == e["prediction"] and _np(bb).tolist() == e["parent"]
    assert _np(nt).tolist() == e["next_t"] and _np(nu).tolist() == e["next_u"]
    _eq(lp2, np.array(e["log_prob"], np.float32))


def _log_softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    return (z - np.log(np.exp(z).sum(axis=-1, keepdims=True))).astype(np.float32)


@pytest.mark.parametrize("W,max_t,quant", [(8, 12, False), (3, 5, True), (1, 4, False), (40, 30, True)])
def test_v1_decode_loop_matches_oracle(product, oracle_mod, space, W, max_t, quant):
    """Full v1 decode loop; `quant` draws probabilities from a tiny set so that exact log-prob
    ties (stable-sort order, consecutive-only dedup) occur at every step."""
    rng = np.random.default_rng(W * 100 + max_t)
    lph = np.zeros(W, np.float32)
    fin = np.zeros(W, np.bool_)
    t = np.zeros(W, np.int32)
    u = np.zeros(W, np.int32)
    for step in range(3 * max_t):
        if quant:
            pe = rng.choice([0.25, 0.5, 0.75], size=(W, 1)).astype(np.float32)
            h = np.log(np.concatenate([pe, 1 - pe], axis=1)).astype(np.float32)
        else:
            h = _log_softmax(rng.standard_normal((W, 2)))
        want = oracle_mod.beam_search_decode(h, lph, fin, t, u, max_t, W)
        got = product.beam_search_decode(*_conv(space, h, lph, fin, t, u), max_t, W)
        for g, w, name in zip(got, want, ("pred", "lp", "nt", "nu", "fin", "parent")):
            _eq(g, w, f"step {step} {name}")
        _, lph, t, u, fin, _ = want
        if fin.all():
            break
    assert fin.all()


def _v2_case(rng, B, W, D, in_hi, quant):
    in_len = rng.integers(max(2, in_hi // 2), in_hi + 1, B).astype(np.int32)
    table = np.arange(D, dtype=np.int32)  # class i ↔ i frames, class 0 = zero duration
    # the reference prunes any step that leaves fewer than 3 frames per remaining token
    # (src/v2.rs:106-111), so a decodable case needs output_length >= 3 * input_length
    out_len = np.array([int(n * rng.uniform(3.2, min(D - 2, 5.0))) for n in in_len], np.int32)
    return in_len, out_len, table


@pytest.mark.parametrize("B,W,D,in_hi,allow_skip,test_mode,quant", [
    (64, 8, 16, 24, False, False, False),
    (5, 8, 32, 40, True, False, False),
    (4, 4, 8, 10, False, True, False),
    (3, 6, 8, 12, True, False, True),
    (2, 33, 12, 9, False, False, False),
])
def test_v2_decode_loop_matches_oracle(product, oracle_mod, space, B, W, D, in_hi, allow_skip, test_mode, quant):
    rng = np.random.default_rng(B * 131 + W * 17 + D)
    in_len, out_len, table = _v2_case(rng, B, W, D, in_hi, quant)
    lph = np.zeros((B, W), np.float32)
    fin = np.zeros((B, W), np.bool_)
    tot = np.zeros((B, W), np.int32)
    t = np.zeros((B, W), np.int32)
    u = np.zeros((B, W), np.int32)
    ol_arg = out_len
    if space == "device":
        product.last_error()
    for step in range(int(in_len.max()) + 2):
        if quant:
            h = np.log(rng.choice([0.05, 0.1, 0.2], size=(B, W, D))).astype(np.float32)
        else:
            h = _log_softmax(rng.standard_normal((B, W, D)))
        ol_or = np.zeros_like(out_len) if test_mode else out_len  # wrapper zeroes it (`__init__.py:47`)
        *want, bad = oracle_mod.ssnt_tts_v2_beam_search_decode(h, lph, fin, tot, table, t, u, in_len, ol_or,
                                                             W, D, 0, allow_skip, test_mode)
        if bad:
            # src/v2.rs:292 would panic for those entries; the device-pointer path raises the flag
            if space == "device":
                product.ssnt_tts_v2_beam_search_decode(*_conv(space, h, lph, fin, tot, table, t, u, in_len, ol_arg),
                                                       W, D, 0, allow_skip, test_mode)
                assert product.last_error() & product.ERR_V2_EMPTY_BEAM
            return
        got = product.ssnt_tts_v2_beam_search_decode(*_conv(space, h, lph, fin, tot, table, t, u, in_len, ol_arg),
                                                     W, D, 0, allow_skip, test_mode)
        for g, w, name in zip(got, want, ("pred", "lp", "nt", "nu", "fin", "total", "parent")):
            _eq(g, w, f"step {step} {name}")
        _, lph, t, u, fin, tot, _ = want
    if space == "device":
        assert product.last_error() == 0
    assert fin.all()


def test_v2_empty_beam_flag(product, oracle_mod):
    # every class pruned by the band → src/v2.rs:292 assert_ne!
    B, W, D = 2, 4, 4
    h = np.zeros((B, W, D), np.float32)
    z = np.zeros((B, W), np.int32)
    table = np.array([0, 50, 60, 70], np.int32)
    in_len, out_len = np.array([10, 10], np.int32), np.array([20, 20], np.int32)
    args = (h, np.zeros((B, W), np.float32), np.zeros((B, W), np.bool_), z, table, z, z, in_len, out_len)
    *_, bad = oracle_mod.ssnt_tts_v2_beam_search_decode(*args, W, D, 0, False, False)
    assert bad == B
    product.last_error()
    product.ssnt_tts_v2_beam_search_decode(*(_dev(a) for a in args), W, D, 0, False, False)
    assert product.last_error() & product.ERR_V2_EMPTY_BEAM


@pytest.mark.parametrize("B,W,K,in_hi", [(64, 8, 4, 20), (3, 5, 1, 6), (2, 40, 7, 9)])
def test_tone_decode_loop_matches_oracle(product, oracle_mod, space, B, W, K, in_hi):
    rng = np.random.default_rng(B * 7 + W + K)
    in_len = rng.integers(1, in_hi + 1, B).astype(np.int32)
    lph = np.zeros((B, W), np.float32)
    fin = np.zeros((B, W), np.bool_)
    t = np.zeros((B, W), np.int32)
    u = np.zeros((B, W), np.int32)
    for step in range(in_hi + 2):
        h = _log_softmax(rng.standard_normal((B, W, K)))
        if step % 3 == 2:
            h = np.round(h * 2) / 2  # provoke ties
        want = oracle_mod.tone_latent_beam_search_decode(h, lph, fin, t, u, in_len, W, K, K)
        got = product.tone_latent_beam_search_decode(*_conv(space, h, lph, fin, t, u, in_len), W, K, K)
        for g, w, name in zip(got, want, ("pred", "lp", "nt", "nu", "fin", "parent")):
            _eq(g, w, f"step {step} {name}")
        _, lph, t, u, fin, _ = want
    assert fin.all()


def test_config4_pipeline_end_to_end(product, oracle_mod):
    """BASELINE config 4 in miniature: v2 beam=8 B=64 decode loop on device buffers, back-trace of
    the recorded parents, upsampling of the decoded durations, edit distance of the decoded
    duration-class sequence against a reference sequence — every stage equal to the oracle."""
    import torch
    rng = np.random.default_rng(4)
    B, W, D, Tin = 64, 8, 16, 20
    in_len = np.full(B, Tin, np.int32)
    out_len = rng.integers(64, 100, B).astype(np.int32)
    table = np.arange(D, dtype=np.int32)
    st = dict(lph=np.zeros((B, W), np.float32), fin=np.zeros((B, W), np.bool_), tot=np.zeros((B, W), np.int32),
              t=np.zeros((B, W), np.int32), u=np.zeros((B, W), np.int32))
    d = {k: _dev(v) for k, v in st.items()}
    d_tab, d_il, d_ol = _dev(table), _dev(in_len), _dev(out_len)
    parents_o, preds_o, parents_d, preds_d = [], [], [], []
    product.last_error()
    for step in range(Tin):
        h = _log_softmax(rng.standard_normal((B, W, D)))
        *o, bad = oracle_mod.ssnt_tts_v2_beam_search_decode(h, st["lph"], st["fin"], st["tot"], table, st["t"],
                                                          st["u"], in_len, out_len, W, D, 0, False, False)
        assert bad == 0
        g = product.ssnt_tts_v2_beam_search_decode(_dev(h), d["lph"], d["fin"], d["tot"], d_tab, d["t"], d["u"],
                                                   d_il, d_ol, W, D, 0, False, False)
        st = dict(lph=o[1], t=o[2], u=o[3], fin=o[4], tot=o[5])
        d = dict(lph=g[1], t=g[2], u=g[3], fin=g[4], tot=g[5])
        parents_o.append(o[6]); preds_o.append(o[0]); parents_d.append(g[6]); preds_d.append(g[0])
    assert product.last_error() == 0
    bb_o = np.stack(parents_o, axis=1)                       # (B, T, W)
    bb_d = torch.stack(parents_d, dim=1).contiguous()
    _eq(bb_d, bb_o)
    final = np.tile(np.arange(W, dtype=np.int32), (B, 1))
    ordered_o = oracle_mod.order_beam_branch(final, bb_o, W)  # (B, W, T)
    ordered_d = product.order_beam_branch(_dev(final), bb_d, W)
    _eq(ordered_d, ordered_o)
    # decoded duration class of beam w at step s = prediction[s][ordered[w][s]]
    pred_o = np.stack(preds_o, axis=1)                       # (B, T, W)
    dur_o = np.take_along_axis(pred_o.transpose(0, 2, 1), ordered_o, axis=1) if False else \
        np.stack([[pred_o[b, np.arange(Tin), ordered_o[b, w]] for w in range(W)] for b in range(B)]).astype(np.int32)
    pred_d = torch.stack(preds_d, dim=1)
    dur_d = torch.gather(pred_d, 2, ordered_d.permute(0, 2, 1).long()).permute(0, 2, 1).contiguous().int()
    _eq(dur_d, dur_o)
    total = dur_o.sum(axis=2).astype(np.int32)
    up_o, bad = oracle_mod.upsample_source_indexes(dur_o, total, -1, W)
    assert bad == 0
    up_d = product.upsample_source_indexes(dur_d, _dev(total), -1, W)
    _eq(up_d, up_o)
    ref = rng.integers(0, D, (B, Tin)).astype(np.int32)
    lens = np.full(B, Tin, np.int32)
    want = oracle_mod.levenshtein_edit_distance(dur_o[:, 0], ref, lens, lens)
    got = product.levenshtein_edit_distance(dur_d[:, 0].contiguous(), _dev(ref), _dev(lens), _dev(lens))
    _eq(got, want)


# ---------------------------------------------------------------- whole-loop decoding
def _oracle_v2_loop(oracle_mod, h, table, in_len, out_len, W, D, zero_id, allow_skip, test_mode, max_u, fill):
    """The reference's usage restated with the oracle: one v2 step per output frame (`__init__.py:33-73`), then
    order_beam_branch with final_branch = 0..W-1, the durations along each branch, upsample_source_indexes."""
    B, S = h.shape[0], h.shape[1]
    lph = np.zeros((B, W), np.float32); fin = np.zeros((B, W), np.bool_); tot = np.zeros((B, W), np.int32)
    t = np.zeros((B, W), np.int32); u = np.zeros((B, W), np.int32)
    ol = np.zeros_like(out_len) if test_mode else out_len
    preds, parents = [], []
    for s in range(S):
        *o, bad = oracle_mod.ssnt_tts_v2_beam_search_decode(np.ascontiguousarray(h[:, s]), lph, fin, tot, table, t, u,
                                                          in_len, ol, W, D, zero_id, allow_skip, test_mode)
        if bad:
            return None
        pred, lph, t, u, fin, tot, parent = o
        preds.append(pred); parents.append(parent)
    ph, bh = np.stack(preds, axis=1), np.stack(parents, axis=1)          # (B, S, W)
    final = np.tile(np.arange(W, dtype=np.int32), (B, 1))
    ordered = oracle_mod.order_beam_branch(final, bh, W)                  # (B, W, S)
    dur = np.stack([[table[ph[b, np.arange(S), ordered[b, w]]] for w in range(W)] for b in range(B)]).astype(np.int32)
    up, bad = oracle_mod.upsample_source_indexes(dur, tot, fill, W, max_u=max_u)
    assert bad == 0
    return dict(prediction_history=ph, beam_branch_history=bh, log_probs=lph, t=t, u=u, is_finished=fin,
                total_duration=tot, ordered_beam_branch=ordered, duration=dur, upsampled_source_indexes=up)


@pytest.mark.parametrize("B,W,D,in_hi,allow_skip,test_mode,quant", [
    (64, 8, 16, 150, False, False, False),    # BASELINE configs[3]: beam 8, B=64, 150 tokens, <= 1000 frames
    (5, 8, 32, 40, True, False, False),
    (4, 4, 8, 10, False, True, False),
    (3, 6, 8, 12, True, False, True),
    (2, 33, 12, 9, False, False, False),
])
def test_v2_whole_loop_equals_per_step_calls(product, oracle_mod, space, B, W, D, in_hi, allow_skip, test_mode, quant):
    rng = np.random.default_rng(B * 131 + W * 17 + D + 1)
    in_len, out_len, table = _v2_case(rng, B, W, D, in_hi, quant)
    S = int(in_len.max()) + 2
    if quant:
        h = np.log(rng.choice([0.05, 0.1, 0.2], size=(B, S, W, D))).astype(np.float32)
    else:
        h = _log_softmax(rng.standard_normal((B, S, W, D)))
    max_u = 1000 if in_hi == 150 else int(out_len.max()) + 8
    want = _oracle_v2_loop(oracle_mod, h, table, in_len, out_len, W, D, 0, allow_skip, test_mode, max_u, -7)
    assert want is not None, "case must be decodable"
    if space == "device":
        product.last_error()
    got = product.ssnt_tts_v2_decode_loop(*_conv(space, h, table, in_len, out_len), W, D, 0, allow_skip, test_mode,
                                          max_u, -7)
    if space == "device":
        assert product.last_error() == 0
    for k, w in want.items():
        _eq(got[k], w, k)
    assert want["is_finished"].all()


def test_v2_whole_loop_takes_an_initial_state(product, oracle_mod):
    """Resuming: the loop over steps [s0, S) from the state after s0 per-step calls equals the tail of the full loop."""
    rng = np.random.default_rng(11)
    B, W, D, S, s0 = 6, 8, 16, 30, 11
    in_len, out_len, table = _v2_case(rng, B, W, D, 28, False)
    h = _log_softmax(rng.standard_normal((B, S, W, D)))
    product.last_error()
    run = lambda hh, **state: product.ssnt_tts_v2_decode_loop(*(_dev(a) for a in (hh, table, in_len, out_len)), W, D, 0,
                                                              True, True, None, -1, **state)
    full = run(h)
    head = run(h[:, :s0])
    tail = run(h[:, s0:], log_prob_history=head["log_probs"], is_finished=head["is_finished"],
               total_duration=head["total_duration"], t=head["t"], u=head["u"])
    assert product.last_error() == 0
    for k in ("log_probs", "t", "u", "is_finished", "total_duration"):
        _eq(tail[k], _np(full[k]), k)
    _eq(tail["prediction_history"], _np(full["prediction_history"])[:, s0:])
    _eq(tail["beam_branch_history"], _np(full["beam_branch_history"])[:, s0:])


@pytest.mark.parametrize("B,W,K,in_hi", [(64, 8, 4, 150), (3, 5, 1, 6), (2, 40, 7, 9)])
def test_tone_whole_loop_equals_per_step_calls(product, oracle_mod, space, B, W, K, in_hi):
    rng = np.random.default_rng(B * 7 + W + K + 1)
    in_len = rng.integers(1, in_hi + 1, B).astype(np.int32)
    S = in_hi + 2
    h = _log_softmax(rng.standard_normal((B, S, W, K)))
    h[:, 2::3] = np.round(h[:, 2::3] * 2) / 2   # provoke ties
    lph = np.zeros((B, W), np.float32); fin = np.zeros((B, W), np.bool_)
    t = np.zeros((B, W), np.int32); u = np.zeros((B, W), np.int32)
    preds, parents = [], []
    for s in range(S):
        pred, lph, t, u, fin, parent = oracle_mod.tone_latent_beam_search_decode(np.ascontiguousarray(h[:, s]), lph, fin, t, u,
                                                                                 in_len, W, K, K)
        preds.append(pred); parents.append(parent)
    ph, bh = np.stack(preds, axis=1), np.stack(parents, axis=1)
    ordered = oracle_mod.order_beam_branch(np.tile(np.arange(W, dtype=np.int32), (B, 1)), bh, W)
    tones = np.stack([[ph[b, np.arange(S), ordered[b, w]] for w in range(W)] for b in range(B)]).astype(np.int32)
    got = product.tone_latent_decode_loop(*_conv(space, h, in_len), W, K, K)
    for k, w in dict(prediction_history=ph, beam_branch_history=bh, log_probs=lph, t=t, u=u, is_finished=fin,
                     ordered_beam_branch=ordered, ordered_tone=tones).items():
        _eq(got[k], w, k)


def test_v2_whole_loop_empty_beam_raises_the_flag(product):
    B, W, D, S = 2, 4, 4, 3
    h = np.zeros((B, S, W, D), np.float32)
    table = np.array([0, 50, 60, 70], np.int32)     # every class pruned by the band → src/v2.rs:292 assert_ne!
    in_len, out_len = np.array([10, 10], np.int32), np.array([20, 20], np.int32)
    product.last_error()
    product.ssnt_tts_v2_decode_loop(*(_dev(a) for a in (h, table, in_len, out_len)), W, D, 0, False, False, 16, -1)
    assert product.last_error() & product.ERR_V2_EMPTY_BEAM


# ---------------------------------------------------------------- robustness
def test_beam_step_with_nan_and_inf_scores_does_not_corrupt(product, oracle_mod):
    """inf - inf / NaN logits: the reference's comparator treats NaN as Equal (src/lib.rs:161) and leaves their final
    position to its sort algorithm, but never crashes.  Here NaN ranks below everything: every output slot is a real
    candidate (valid parent, valid class), and rows without NaN still equal the oracle exactly."""
    rng = np.random.default_rng(3)
    B, W, K = 6, 8, 5
    h = _log_softmax(rng.standard_normal((B, W, K)))
    h[0, 2, 1] = np.nan; h[0, 5, :] = np.nan; h[1, :, 3] = -np.inf; h[2, 0, 0] = np.inf; h[3] = np.nan
    lph = np.zeros((B, W), np.float32); lph[1, 3] = -np.inf; lph[2, 1] = np.inf
    fin = np.zeros((B, W), np.bool_); z = np.zeros((B, W), np.int32); in_len = np.full(B, 9, np.int32)
    product.last_error()
    got = product.tone_latent_beam_search_decode(*(_dev(a) for a in (h, lph, fin, z, z, in_len)), W, K, K)
    assert product.last_error() == 0
    pred, lp, nt, nu, nf, parent = (_np(g) for g in got)
    assert ((parent >= 0) & (parent < W)).all() and ((pred >= 0) & (pred < K)).all()
    assert (nt == 1).all() and (nu == 1).all()
    want = oracle_mod.tone_latent_beam_search_decode(h, lph, fin, z, z, in_len, W, K, K)
    for b in (4, 5):   # rows without NaN
        for g, w in zip((pred, lp, nt, nu, nf, parent), want):
            _eq(g[b], w[b])
    # rows 1 and 2 hold infinities but no NaN: a total order exists, so they must match too
    for b in (1,):
        for g, w in zip((pred, lp, nt, nu, nf, parent), want):
            _eq(g[b], w[b])


def test_wide_candidate_tables_run_with_fewer_warps_per_block(product, oracle_mod):
    """beam_width * classes beyond four tables' worth of shared memory (W=32, D=64: 2048 candidates) used to abort."""
    rng = np.random.default_rng(8)
    B, W, D = 3, 32, 64
    h = _log_softmax(rng.standard_normal((B, W, D)))
    lph = np.zeros((B, W), np.float32); fin = np.zeros((B, W), np.bool_); z = np.zeros((B, W), np.int32)
    table = np.arange(D, dtype=np.int32)
    in_len, out_len = np.full(B, 12, np.int32), np.full(B, 60, np.int32)
    *want, bad = oracle_mod.ssnt_tts_v2_beam_search_decode(h, lph, fin, z, table, z, z, in_len, out_len, W, D, 0, False, True)
    assert bad == 0
    got = product.ssnt_tts_v2_beam_search_decode(*(_dev(a) for a in (h, lph, fin, z, table, z, z, in_len, out_len)),
                                                 W, D, 0, False, True)
    for g, w in zip(got, want):
        _eq(g, w)

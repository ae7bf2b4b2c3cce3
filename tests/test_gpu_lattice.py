"""GPU parity of the lattice forward-backward (and its tone-latent variant) against the fp64
CPU oracle, through the C-ABI.

The reference has no forward-backward at all, so the oracle for this part is
an authored specification — "parity unpinned" by the reference; the oracle itself is hardened in
test_oracle_lattice.py.  Tolerances (BASELINE.json north_star): log-likelihood within 1e-5
relative; gradients within 1e-4 relative ELEMENT-WISE for every occupancy >= 1e-6 and within 1e-10
absolute below that (gradients are posterior probabilities in [0, 1]; the relative error of an
entry like 1e-30 is not meaningful in fp32).  Full-size runs are additionally checked through
size-independent properties: every frame's occupancies sum to 1, expected shifts = U-1."""
import numpy as np
import pytest

from lattice_util import make_inputs, ragged_lengths

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _dev(x):
    import torch
    return None if x is None else torch.as_tensor(np.ascontiguousarray(x)).cuda()


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


GRAD_FLOOR = 1e-6   # occupancies below this are compared absolutely (GRAD_RTOL * GRAD_FLOOR = 1e-10)


def _check(got, want, t_len=None, u_len=None, grad_rtol=GRAD_RTOL):
    """log-likelihood: 1e-5 relative (absolute 1e-5 for |LL| < 1).  Gradients: ELEMENT-WISE 1e-4 relative for every
    occupancy >= 1e-6, absolute 1e-10 below (north_star: "gradients within 1e-4 relative")."""
    ll, loss, ge, gs = (_np(g) for g in got)
    ll64, loss64, ge64, gs64 = want
    finite = np.isfinite(ll64)
    assert np.array_equal(np.isfinite(ll), finite)
    assert np.all(np.abs(ll[finite] - ll64[finite]) <= LL_RTOL * np.maximum(np.abs(ll64[finite]), 1.0))
    if finite.all():
        assert abs(float(loss[0]) - loss64) <= LL_RTOL * max(abs(loss64), 1.0)
    for g, g64, name in ((ge, ge64, "grad_emit"), (gs, gs64, "grad_shift")):
        err = np.abs(g.astype(np.float64) - g64)
        bound = grad_rtol * np.maximum(np.abs(g64), GRAD_FLOOR)
        bad = err > bound
        assert not bad.any(), (name, int(bad.sum()), float((err / np.maximum(np.abs(g64), GRAD_FLOOR)).max()))
        # (cells whose fp64 occupancy rounds to 0 in fp32 fall under the absolute bound above; padded cells are
        # checked to be exactly zero below)
    if t_len is not None:
        for b in range(ll.shape[0]):
            assert not ge[b, t_len[b]:].any() and not gs[b, t_len[b]:].any()
            assert not ge[b, :, u_len[b]:].any() and not gs[b, :, u_len[b]:].any()


def _run(product, le, ls, t_len, u_len, space, kind=-1):
    product.set_fb_kernel(kind)
    try:
        if space == "device":
            out = product.forward_backward(_dev(le), _dev(ls), _dev(t_len), _dev(u_len))
        else:
            out = product.forward_backward(le, ls, t_len, u_len)
        used = product.fb_kernel_used()
    finally:
        product.set_fb_kernel(-1)
    return out, used


@pytest.mark.parametrize("space", ["host", "device"])
def test_config1_B1_U32_T120(product, oracle_mod, space):
    # BASELINE configs[0]: the reference's own CPU-runnable shape
    le, ls = make_inputs(1, 120, 32, seed=1234)
    want = oracle_mod.forward_backward(le, ls)
    got, used = _run(product, le, ls, None, None, space)
    assert used == 6  # the time-parallel block-float kernels, not a fallback
    _check(got, want)


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 6, 7, 8, 9])  # generic, log-warp, block-float (+ forced log re-run), time-parallel (+ forced re-run), warp-serial (+ forced re-run)
@pytest.mark.parametrize("B,T,U", [(3, 1, 4), (2, 2, 4), (4, 3, 4), (5, 9, 8), (3, 17, 16), (2, 40, 36),
                                   (3, 64, 64), (2, 100, 128), (2, 70, 200), (1, 90, 260), (1, 600, 520)])
def test_shapes_and_ragged_lengths(product, oracle_mod, kind, B, T, U):
    le, ls = make_inputs(B, T, U, seed=T * 1000 + U)
    rng = np.random.default_rng(T + U)
    t_len = rng.integers(1, T + 1, B).astype(np.int32)
    u_len = np.array([rng.integers(1, min(U, tb) + 1) for tb in t_len], np.int32)
    t_len[0] = T
    u_len[0] = min(U, T)
    if kind >= 2 and U > 256:
        pytest.skip("block-float kernel covers max_u <= 256; larger lattices take the log-warp kernel")
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got, used = _run(product, le, ls, t_len, u_len, "device", kind)
    assert used == kind
    _check(got, want, t_len, u_len)


@pytest.mark.parametrize("U", [1, 5, 30, 33, 150])
def test_unaligned_max_u_takes_generic_kernel(product, oracle_mod, U):
    le, ls = make_inputs(3, 47, U, seed=U)
    t_len, u_len = ragged_lengths(3, 47, U, seed=U)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got, used = _run(product, le, ls, t_len, u_len, "device")
    assert used == (6 if U % 4 == 0 else 0)
    _check(got, want, t_len, u_len)


def test_infeasible_empty_and_masked(product, oracle_mod):
    le, ls = make_inputs(6, 12, 8, seed=5)
    t_len = np.array([12, 0, 3, 12, 1, 12], np.int32)
    u_len = np.array([8, 2, 5, 0, 1, 8], np.int32)   # 1: T=0, 2: U>T, 3: U=0 → ll=-inf, grads 0
    le[5, 4, :] = -np.inf                            # frame 4 cannot emit anywhere …
    ls[5, 4, :] = -np.inf                            # … nor shift: no path at all → -inf
    le[0, 3, 2] = -np.inf                            # a single forbidden cell is fine
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    assert np.isneginf(want[0][[1, 2, 3, 5]]).all() and np.isfinite(want[0][[0, 4]]).all()
    for kind in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9):
        if kind in (4, 5):
            continue  # max_u = 8 here; the split-role kernel needs max_u in {64, 128, 256} (covered below)
        got, _ = _run(product, le, ls, t_len, u_len, "device", kind)
        _check(got, want, t_len, u_len)
        assert np.isposinf(_np(got[1])[0])


@pytest.mark.parametrize("space", ["host", "device"])
def test_config2_B32_U128_T800(product, oracle_mod, space):
    # BASELINE configs[1]: the headline shape, full lengths
    le, ls = make_inputs(32, 800, 128, seed=1234)
    want = oracle_mod.forward_backward(le, ls)
    got, used = _run(product, le, ls, None, None, space)
    assert used == 6  # the time-parallel kernels (chunk operators / boundary vectors / chunk interiors)
    _check(got, want)
    ll, loss, ge, gs = (_np(g) for g in got)
    rows = (ge + gs).sum(axis=2)
    np.testing.assert_allclose(rows, 1.0, atol=2e-4)          # one transition per frame
    np.testing.assert_allclose(gs.sum(axis=(1, 2)), 127.0, rtol=1e-4)  # exactly U-1 shifts


@pytest.mark.parametrize("kind", [1, 2, 4, 6, 8])
def test_config2_ragged(product, oracle_mod, kind):
    le, ls = make_inputs(32, 800, 128, seed=77)
    t_len, u_len = ragged_lengths(32, 800, 128)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got, _ = _run(product, le, ls, t_len, u_len, "device", kind)
    _check(got, want, t_len, u_len)


def test_block_float_accuracy_budget(product, oracle_mod):
    """The hot-path kernel works on probabilities with power-of-two block exponents, so its error
    is fp32-relative: an order of magnitude inside the 1e-4 budget at the headline shape."""
    le, ls = make_inputs(4, 800, 128, seed=21)
    ll64, _, ge64, gs64 = oracle_mod.forward_backward(le, ls)
    got, used = _run(product, le, ls, None, None, "device", 2)
    assert used == 2
    ll, loss, ge, gs = (_np(g) for g in got)
    assert np.all(np.abs(ll - ll64) <= 2e-6 * np.abs(ll64))
    assert np.abs(ge - ge64).max() <= 1e-5 and np.abs(gs - gs64).max() <= 1e-5


def test_time_parallel_kernels_config2_and_no_fallback(product, oracle_mod):
    """Kind 6 (chunk transfer operators built concurrently, fb_tp.cuh) at the headline shape: parity with the
    fp64 oracle at the block-float accuracy budget, per-frame invariants, and no log-domain re-run on typical inputs."""
    le, ls = make_inputs(32, 800, 128, seed=4321)
    want = oracle_mod.forward_backward(le, ls)
    before = product.fb_fallback_count()
    got, used = _run(product, le, ls, None, None, "device", 6)
    assert used == 6
    assert product.fb_fallback_count() == before
    _check(got, want)
    ll, loss, ge, gs = (_np(g) for g in got)
    assert np.all(np.abs(ll - want[0]) <= 2e-6 * np.abs(want[0]))
    assert np.abs(ge - want[2]).max() <= 1e-5 and np.abs(gs - want[3]).max() <= 1e-5
    np.testing.assert_allclose((ge + gs).sum(axis=2), 1.0, atol=2e-4)
    np.testing.assert_allclose(gs.sum(axis=(1, 2)), 127.0, rtol=1e-4)


def test_peaked_and_uniform_inputs(product, oracle_mod):
    """Inputs that stress the dynamic range: exactly uniform probabilities (binomial spread of
    alpha over hundreds of bits) and a sharply peaked, mostly-wrong model (log-probs of -30)."""
    B, T, U = 4, 400, 64
    le = np.full((B, T, U), np.log(0.5), np.float32)
    ls = np.full((B, T, U), np.log(0.5), np.float32)
    rng = np.random.default_rng(3)
    le[1] = -30.0 * (rng.random((T, U)) < 0.5) - 1e-3
    ls[1] = -30.0 * (rng.random((T, U)) < 0.5) - 1e-3
    le[2], ls[2] = np.log(1e-10), np.log1p(-1e-10)        # model insists on shifting every frame
    le[3], ls[3] = np.log1p(-1e-6), np.log(1e-6)          # model never wants to shift
    want = oracle_mod.forward_backward(le, ls)
    for kind in (0, 1, 2, 4, 6, 8):   # kinds 2, 4, 6 and 8 must notice what they cannot hold and re-run it in the log domain
        got, _ = _run(product, le, ls, None, None, "device", kind)
        # |LL| reaches several thousand here: one fp32 ulp of a log-domain quantity of that size is a few 1e-4 of its
        # exponential, so the element-wise bound is 3e-4 for this adversarial case (1e-4 everywhere else)
        _check(got, want, grad_rtol=3e-4)


def test_large_batch_properties(product):
    """Size-independent invariants at a batch far beyond what the CPU oracle covers in seconds:
    B=1024 U=128 T=400 (more CTAs than the GPU holds at once, multi-wave), ragged."""
    import torch
    B, T, U = 1024, 400, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.randn(B, T, U, device="cuda", generator=g)
    le = torch.nn.functional.logsigmoid(z)
    ls = torch.nn.functional.logsigmoid(-z)
    t_len = torch.randint(240, T + 1, (B,), device="cuda", generator=g, dtype=torch.int32)
    u_len = torch.randint(77, U + 1, (B,), device="cuda", generator=g, dtype=torch.int32)
    ll, loss, ge, gs = product.forward_backward(le, ls, t_len, u_len)
    torch.cuda.synchronize()
    rows = (ge + gs).sum(dim=2)
    valid = torch.arange(T, device="cuda")[None, :] < t_len[:, None]
    assert torch.allclose(rows[valid], torch.ones_like(rows[valid]), atol=2e-4)
    assert not rows[~valid].any()
    assert torch.allclose(gs.sum(dim=(1, 2)), (u_len - 1).float(), rtol=1e-4)
    assert torch.allclose(loss[0], -ll.double().sum().float(), rtol=1e-6)
    # linearity of the likelihood in a per-utterance constant: adding c to every log_emit of the
    # last frame adds exactly c to ll and leaves the occupancies unchanged
    le2 = le.clone()
    idx = (t_len - 1).long()
    le2[torch.arange(B), idx] += 0.75
    ll2, _, ge2, _ = product.forward_backward(le2, ls, t_len, u_len)
    assert torch.allclose(ll2, ll + 0.75, rtol=1e-5, atol=1e-3)
    assert torch.allclose(ge2, ge, atol=2e-5)


def test_repeatable_bitwise(product):
    le, ls = make_inputs(8, 200, 64, seed=2)
    a = product.forward_backward(_dev(le), _dev(ls))
    b = product.forward_backward(_dev(le), _dev(ls))
    for x, y in zip(a, b):
        assert np.array_equal(_np(x).view(np.uint32), _np(y).view(np.uint32))


def test_caller_workspace_and_preallocated_outputs(product, oracle_mod):
    import torch
    le, ls = make_inputs(4, 64, 32, seed=6)
    want = oracle_mod.forward_backward(le, ls)
    n = product.forward_backward_workspace_bytes(4, 64, 32)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = (torch.empty(4, device="cuda"), torch.empty(1, device="cuda"),
           torch.full((4, 64, 32), 7.0, device="cuda"), torch.full((4, 64, 32), 7.0, device="cuda"))
    got = product.forward_backward(_dev(le), _dev(ls), workspace=ws, out=out)
    _check(got, want)


# ---------------------------------------------------------------- tone-latent lattice
def _check_tone(got, want):
    ll, loss, ge, gs, gt = (_np(g) for g in got)
    ll64, loss64, ge64, gs64, gt64 = want
    finite = np.isfinite(ll64)
    assert np.array_equal(np.isfinite(ll), finite)
    assert np.all(np.abs(ll[finite] - ll64[finite]) <= LL_RTOL * np.abs(ll64[finite]) + 1e-6)
    for b in range(ll.shape[0]):
        scale = max(float(np.abs(ge64[b]).max()), 1e-30)
        for g, g64 in ((ge[b], ge64[b]), (gs[b], gs64[b]), (gt[b], gt64[b])):
            assert float(np.abs(g - g64).max()) <= GRAD_RTOL * max(scale, float(np.abs(g64).max()))


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("B,T,U,K", [(2, 1, 1, 3), (3, 12, 5, 2), (4, 60, 24, 4), (2, 33, 9, 1), (2, 50, 40, 7)])
def test_tone_latent_small(product, oracle_mod, space, B, T, U, K):
    le, ls, lt = make_inputs(B, T, U, seed=T + U + K, K=K)
    t_len, u_len = ragged_lengths(B, T, U, seed=K)
    t_len[0], u_len[0] = T, min(T, U)
    want = oracle_mod.tone_latent_forward_backward(le, ls, lt, t_len, u_len)
    if space == "device":
        got = product.tone_latent_forward_backward(_dev(le), _dev(ls), _dev(lt), _dev(t_len), _dev(u_len))
    else:
        got = product.tone_latent_forward_backward(le, ls, lt, t_len, u_len)
    _check_tone(got, want)


def test_tone_latent_config3_B32_U128_T800_K4(product, oracle_mod):
    # BASELINE configs[2]
    le, ls, lt = make_inputs(32, 800, 128, seed=1234, K=4)
    want = oracle_mod.tone_latent_forward_backward(le, ls, lt)
    got = product.tone_latent_forward_backward(_dev(le), _dev(ls), _dev(lt))
    _check_tone(got, want)
    ll, loss, ge, gs, gt = (_np(g) for g in got)
    np.testing.assert_allclose((ge + gs).sum(axis=(2, 3)), 1.0, atol=3e-4)
    np.testing.assert_allclose(gt.sum(axis=2), 1.0, atol=3e-4)


@pytest.mark.parametrize("B", [9, 10, 13])
def test_tone_latent_host_buffers_chunked(product, oracle_mod, B):
    """Host-pointer call large enough for the chunked H2D / kernel / D2H pipeline (c_api.cu), with a
    ragged last chunk (B=10 -> 3,3,3,1; B=13 -> six chunks 3,3,3,3,1) and ragged lengths; loss = -sum of the finite likelihoods' total."""
    T, U, K = 300, 64, 4
    le, ls, lt = make_inputs(B, T, U, seed=B, K=K)
    t_len, u_len = ragged_lengths(B, T, U, seed=B + 1)
    t_len[0], u_len[0] = T, U
    want = oracle_mod.tone_latent_forward_backward(le, ls, lt, t_len, u_len)
    got = product.tone_latent_forward_backward(le, ls, lt, t_len, u_len)
    _check_tone(got, want)
    got2 = product.tone_latent_forward_backward(_dev(le), _dev(ls), _dev(lt), _dev(t_len), _dev(u_len))
    for i, (a, b) in enumerate(zip(got, got2)):  # per-utterance results do not depend on the chunking
        if i == 1:
            np.testing.assert_allclose(_np(a), _np(b), rtol=1e-6)
        else:
            np.testing.assert_array_equal(_np(a), _np(b))


def test_tone_latent_infeasible(product, oracle_mod):
    le, ls, lt = make_inputs(3, 6, 8, seed=9, K=2)
    t_len, u_len = np.array([6, 0, 4], np.int32), np.array([7, 3, 4], np.int32)
    want = oracle_mod.tone_latent_forward_backward(le, ls, lt, t_len, u_len)
    got = product.tone_latent_forward_backward(_dev(le), _dev(ls), _dev(lt), _dev(t_len), _dev(u_len))
    _check_tone(got, want)
    assert np.isneginf(_np(got[0])[:2]).all()


# ---- split-role kernel (kind 4; kind 5 = 4 with the log-domain re-run forced) and wide lattices ----------------
@pytest.mark.timeout(120)
@pytest.mark.parametrize("kind", [2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("B,T,U", [(3, 64, 64), (2, 9, 64), (5, 333, 128), (2, 801, 128), (3, 130, 128),
                                   (2, 700, 256), (1, 300, 256), (35, 200, 128)])
def test_full_width_lattices_all_block_float_kernels(product, oracle_mod, kind, B, T, U):
    """max_u in {64, 128, 256}: the shapes the split-role kernel takes, ragged lengths, including
    U=256 (shallow shared-memory ring — a producer batch deeper than the ring once dead-locked the
    fused kernel here) and U close to T (steep fronts: the block-float kernels must fall back)."""
    le, ls = make_inputs(B, T, U, seed=T * 7 + U + kind)
    rng = np.random.default_rng(T + U)
    t_len = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
    u_len = np.array([rng.integers(1, min(U, tb) + 1) for tb in t_len], np.int32)
    t_len[0] = T
    u_len[0] = min(U, T)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got, used = _run(product, le, ls, t_len, u_len, "device", kind)
    assert used == kind
    _check(got, want, t_len, u_len)


@pytest.mark.timeout(120)
def test_split_kernel_infeasible_and_masked(product, oracle_mod):
    le, ls = make_inputs(6, 80, 64, seed=15)
    t_len = np.array([80, 0, 30, 80, 1, 80], np.int32)
    u_len = np.array([64, 2, 50, 0, 1, 64], np.int32)   # 1: T=0, 2: U>T, 3: U=0 → ll=-inf, grads 0
    le[5, 40, :] = -np.inf
    ls[5, 40, :] = -np.inf                                 # no path at all → -inf
    le[0, 3, 2] = -np.inf
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    for kind in (2, 4, 5, 6, 7, 8, 9):
        got, used = _run(product, le, ls, t_len, u_len, "device", kind)
        assert used == kind
        _check(got, want, t_len, u_len)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("B,T,U", [(2048, 40, 128), (2500, 24, 256)])
def test_warp_serial_even_waves(product, oracle_mod, B, T, U):
    """Batch sizes at which the warp-serial kernels cap their resident warps per SM (one full wave plus a sliver
    otherwise; csrc/fb_kernels.cu::ws_resident_cap): auto-dispatch takes kind 8, results against the fp64 oracle."""
    le, ls = make_inputs(B, T, U, seed=B + T)
    t_len, u_len = ragged_lengths(B, T, U, seed=B)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got, used = _run(product, le, ls, t_len, u_len, "device")
    assert used == 8
    _check(got, want, t_len, u_len)


def test_auto_dispatch_small_and_large_batches(product):
    """Every aligned shape up to max_u = 256 takes the time-parallel kernels (kind 6), whatever the batch size."""
    import torch
    for B, want_kind in ((4, 6), (32, 6), (64, 6), (300, 6)):
        z = torch.randn(B, 200, 128, device="cuda")
        le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
        ll, loss, ge, gs = product.forward_backward(le, ls)
        torch.cuda.synchronize()
        assert product.fb_kernel_used() == want_kind
        rows = (ge + gs).sum(dim=2)
        assert torch.allclose(rows, torch.ones_like(rows), atol=2e-4)


# ---- tone-latent lattice: block-float split-role kernel (K = 4, max_u in {32, 64, 128}) ------------------------
@pytest.mark.timeout(180)
@pytest.mark.parametrize("B,T,U", [(2, 40, 32), (3, 100, 64), (2, 300, 128), (1, 801, 128), (35, 64, 32), (2, 5, 32)])
@pytest.mark.parametrize("kind", [1, 0, 2, 3])
def test_tone_latent_block_float_and_log_kernels(product, oracle_mod, B, T, U, kind):
    """Shapes the split-role block-float tone kernel takes (kind 1) against the fp64 oracle, ragged lengths; kind 0 runs
    the log-domain kernel on the same inputs, kind 2 the warp-serial block-float kernels, kind 3 those with every
    utterance re-run in the log domain, so all paths are pinned to the same vectors."""
    K = 4
    _tone_case(product, oracle_mod, B, T, U, K, kind)


# ---- tone-latent lattice: warp-serial block-float kernels (K in {2,4,8}, max_u in {32,64,128,256}) -------------
@pytest.mark.timeout(180)
@pytest.mark.parametrize("B,T,U,K", [(3, 90, 32, 8), (3, 70, 64, 2), (2, 130, 64, 8), (4, 150, 128, 2), (2, 200, 128, 8),
                                     (2, 400, 256, 2), (2, 500, 256, 4), (40, 64, 64, 4), (2, 1100, 128, 4)])
@pytest.mark.parametrize("kind", [2, 3])
def test_tone_latent_warp_serial_kernels(product, oracle_mod, B, T, U, K, kind):
    """Every shape family of the warp-serial tone kernels (csrc/tone_ws.cu) against the fp64 oracle, ragged lengths."""
    _tone_case(product, oracle_mod, B, T, U, K, kind)


def _tone_case(product, oracle_mod, B, T, U, K, kind):
    le, ls, lt = make_inputs(B, T, U, seed=T + U, K=K)
    rng = np.random.default_rng(T * 3 + U)
    t_len = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
    u_len = np.array([rng.integers(1, min(U, tb) + 1) for tb in t_len], np.int32)
    t_len[0] = T
    u_len[0] = min(U, T)
    want = oracle_mod.tone_latent_forward_backward(le, ls, lt, t_len, u_len)
    product.set_tone_kernel(kind)
    try:
        got = product.tone_latent_forward_backward(_dev(le), _dev(ls), _dev(lt), _dev(t_len), _dev(u_len))
        ll, loss, ge, gs, gt = (_np(g) for g in got)
        assert product.tone_kernel_used() == kind
    finally:
        product.set_tone_kernel(-1)
    finite = np.isfinite(want[0])
    assert np.array_equal(np.isfinite(ll), finite)
    assert np.all(np.abs(ll[finite] - want[0][finite]) <= LL_RTOL * np.abs(want[0][finite]) + 1e-6)
    if finite.all():
        assert abs(float(loss[0]) - want[1]) <= LL_RTOL * max(abs(want[1]), 1.0)
    for b in range(B):
        scale = max(float(np.abs(want[2][b]).max()), 1e-30)
        assert np.abs(ge[b] - want[2][b]).max() <= GRAD_RTOL * scale
        assert np.abs(gs[b] - want[3][b]).max() <= GRAD_RTOL * scale
        # tone gradients are sums of occupancies over frames: scale by their own maximum
        assert np.abs(gt[b] - want[4][b]).max() <= GRAD_RTOL * max(1.0, float(np.abs(want[4][b]).max()))
        assert not ge[b, t_len[b]:].any() and not gs[b, t_len[b]:].any()
        assert not ge[b, :, u_len[b]:].any() and not gt[b, u_len[b]:].any()


@pytest.mark.timeout(180)
def test_tone_latent_config3_properties(product):
    """BASELINE configs[2] at full size (B=32 U=128 T=800 K=4): per-frame occupancies sum to 1, every
    token draws exactly one tone (sum_k grad_tone[u,k] = 1), expected shifts = U-1."""
    import torch
    B, T, U, K = 32, 800, 128, 4
    g = torch.Generator(device="cuda").manual_seed(7)
    z = torch.randn(B, T, U, K, device="cuda", generator=g)
    le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
    lt = torch.log_softmax(torch.randn(B, U, K, device="cuda", generator=g), dim=-1)
    ll, loss, ge, gs, gt = product.tone_latent_forward_backward(le, ls, lt)
    torch.cuda.synchronize()
    rows = (ge + gs).sum(dim=(2, 3))
    assert torch.allclose(rows, torch.ones_like(rows), atol=3e-4)
    assert torch.allclose(gt.sum(dim=2), torch.ones(B, U, device="cuda"), atol=3e-4)
    assert torch.allclose(gs.sum(dim=(1, 2, 3)), torch.full((B,), float(U - 1), device="cuda"), rtol=1e-4)
    assert torch.allclose(loss[0], -ll.double().sum().float(), rtol=1e-6)


def test_tone_latent_block_float_does_not_fall_back_on_typical_inputs(product):
    import torch
    B, T, U, K = 8, 400, 128, 4
    z = torch.randn(B, T, U, K, device="cuda")
    le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
    lt = torch.log_softmax(torch.randn(B, U, K, device="cuda"), dim=-1)
    before = product.fb_fallback_count()
    product.tone_latent_forward_backward(le, ls, lt)
    torch.cuda.synchronize()
    assert product.fb_fallback_count() == before


def test_device_pointer_call_is_capturable_in_a_cuda_graph(product, oracle_mod):
    """bench.py replays the step's C-ABI call from CUDA graphs: a device-pointer call with a caller-owned
    workspace enqueues its launch on the stream set through ssnt_tts_set_stream and does nothing a capture
    forbids (no allocation, no synchronisation); replays give the eager call's results bit for bit."""
    import torch
    B, T, U = 6, 200, 128
    le, ls = make_inputs(B, T, U, seed=77)
    want = oracle_mod.forward_backward(le, ls)
    dle, dls = _dev(le), _dev(ls)
    ws = torch.empty(product.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device="cuda")
    out = (torch.empty(B, device="cuda"), torch.empty(1, device="cuda"),
           torch.empty(B, T, U, device="cuda"), torch.empty(B, T, U, device="cuda"))
    product.forward_backward(dle, dls, workspace=ws, out=out)
    torch.cuda.synchronize()
    eager = [o.clone() for o in out]
    for o in out:
        o.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        product.forward_backward(dle, dls, workspace=ws, out=out)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager, out):
        assert torch.equal(a, b)
    _check(out, want)


# ---- host-pointer calls on ordinary (pageable) numpy arrays, and the loss exchange ------------------------------
def test_host_pointer_call_on_pageable_arrays_config2(product, oracle_mod):
    """What the reference's DEVICE_CPU ops hand over (ssnt_tts_v2_beam_search_decode_op.cc:146-177): plain np.empty /
    np.array buffers, not pinned.  The library page-locks them in place for the call; results must not depend on it."""
    le, ls = make_inputs(32, 800, 128, seed=99)
    le, ls = np.array(le, copy=True), np.array(ls, copy=True)
    want = oracle_mod.forward_backward(le, ls)
    out = (np.empty(32, np.float32), np.empty(1, np.float32), np.empty((32, 800, 128), np.float32),
           np.empty((32, 800, 128), np.float32))
    got = product.forward_backward(le, ls, out=out)
    _check(got, want)
    # an unaligned, odd-sized view (page-rounding of the registered range must not matter)
    le2, ls2 = np.array(le[1:8, :301], copy=True), np.array(ls[1:8, :301], copy=True)
    want2 = oracle_mod.forward_backward(le2, ls2)
    _check(product.forward_backward(le2, ls2), want2)


def test_loss_exchange_single_rank(product, oracle_mod):
    """World size 1: the kernel that reduces the loss stores it into this rank's own slot buffer; the all-reduce
    kernel returns it bit for bit, also when the call is replayed from a CUDA graph."""
    import ctypes
    import torch
    L = product.lib()
    buf = (ctypes.c_ubyte * 64)()
    L.ssnt_tts_loss_exchange_export(1, buf)
    L.ssnt_tts_loss_exchange_connect(0, 1, buf)
    try:
        le, ls = make_inputs(5, 96, 64, seed=3)
        ll, loss, ge, gs = product.forward_backward(_dev(le), _dev(ls))
        red = product.loss_allreduce()
        torch.cuda.synchronize()
        assert torch.equal(red, loss)
        # tone lattice: the log-domain kernel's reduction is the one that is exchanged
        le4, ls4, lt4 = make_inputs(3, 40, 32, seed=4, K=4)
        out = product.tone_latent_forward_backward(_dev(le4), _dev(ls4), _dev(lt4))
        red = product.loss_allreduce()
        torch.cuda.synchronize()
        assert torch.equal(red, out[1])
    finally:
        product.disconnect_loss_exchange()


# ---- raw-logit entry point ------------------------------------------------------------------
def _logit_case(B, T, U, seed, ragged=True):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((B, T, U)).astype(np.float32)
    z64 = z.astype(np.float64)
    le = (-np.logaddexp(0.0, -z64)).astype(np.float32)
    ls = (-np.logaddexp(0.0, z64)).astype(np.float32)
    t_len, u_len = ragged_lengths(B, T, U, seed=seed) if ragged else (None, None)
    return z, le, ls, t_len, u_len


def _check_logits(got, z, want, t_len):
    """want = oracle on (log sigmoid(z), log sigmoid(-z)); grad_logits = grad_emit sigmoid(-z) - grad_shift sigmoid(z)."""
    ll, loss, gz = (_np(g) for g in got)
    ll64, loss64, ge64, gs64 = want
    z64 = z.astype(np.float64)
    gz64 = ge64 / (1.0 + np.exp(z64)) - gs64 / (1.0 + np.exp(-z64))
    finite = np.isfinite(ll64)
    assert np.array_equal(np.isfinite(ll), finite)
    # the oracle saw the fp32 roundings of the log-sigmoids, the kernels form them from z: 3e-5 on |LL| >= 1
    assert np.all(np.abs(ll[finite] - ll64[finite]) <= 3e-5 * np.maximum(np.abs(ll64[finite]), 1.0))
    # a difference of two occupancy terms: relative to the larger term, floor 1e-5
    scale = np.maximum(np.maximum(ge64, gs64), 1e-5)
    err = np.abs(gz - gz64)
    assert (err <= 2e-4 * scale).all(), float((err / scale).max())
    if t_len is not None:
        for b in range(ll.shape[0]):
            assert not gz[b, t_len[b]:].any()


@pytest.mark.parametrize("B,T,U,want_kind", [(1, 120, 32, 6), (32, 800, 128, 6), (3, 333, 256, 6), (2, 64, 64, 6),
                                             (2, 300, 512, 1), (3, 47, 30, 6), (2, 40, 33, 0)])
def test_logit_entry_matches_oracle_on_log_sigmoids(product, oracle_mod, B, T, U, want_kind):
    z, le, ls, t_len, u_len = _logit_case(B, T, U, seed=B * 1000 + T + U)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got = product.forward_backward_logits(_dev(z), _dev(t_len), _dev(u_len))
    if U % 4 == 0:
        assert product.fb_kernel_used() == want_kind
    _check_logits(got, z, want, t_len)
    # and the two-tensor call on the same log-sigmoids
    ref = product.forward_backward(_dev(le), _dev(ls), _dev(t_len), _dev(u_len))
    ge, gs = _np(ref[2]).astype(np.float64), _np(ref[3]).astype(np.float64)
    z64 = z.astype(np.float64)
    np.testing.assert_allclose(_np(got[2]), ge / (1.0 + np.exp(z64)) - gs / (1.0 + np.exp(-z64)), atol=2e-6, rtol=1e-4)


def test_logit_entry_host_pointers_and_extreme_logits(product, oracle_mod):
    z, le, ls, t_len, u_len = _logit_case(5, 200, 128, seed=77)
    z[0, :, :] *= 8.0          # saturated sigmoids: |z| up to ~35
    z[1, 10:20, 5:9] = -60.0   # emit practically forbidden in a patch
    z64 = z.astype(np.float64)
    le = (-np.logaddexp(0.0, -z64)).astype(np.float32)
    ls = (-np.logaddexp(0.0, z64)).astype(np.float32)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got = product.forward_backward_logits(z, t_len, u_len)     # host buffers
    _check_logits(got, z, want, t_len)
    got = product.forward_backward_logits(_dev(z), _dev(t_len), _dev(u_len))
    _check_logits(got, z, want, t_len)


# ---- 8-frame chunks of the time-parallel kernels (kind 10; the auto choice for wide lattices at medium batch sizes) ---
@pytest.mark.timeout(180)
@pytest.mark.parametrize("B,T,U", [(2, 700, 256), (1, 300, 256), (3, 2000, 256), (5, 333, 200), (2, 97, 132), (80, 120, 160)])
def test_short_chunk_time_parallel_kernels(product, oracle_mod, B, T, U):
    le, ls = make_inputs(B, T, U, seed=T * 3 + U + B)
    t_len, u_len = ragged_lengths(B, T, U, seed=T + B)
    t_len[0], u_len[0] = T, min(U, T)
    want = oracle_mod.forward_backward(le, ls, t_len, u_len)
    got, used = _run(product, le, ls, t_len, u_len, "device", 10)
    assert used == 10
    _check(got, want, t_len, u_len)
    if B == 80:   # more sweeps than SMs: the auto choice takes the short chunks too (kind 6 reported)
        fb0 = product.fb_fallback_count()
        got, used = _run(product, le, ls, t_len, u_len, "device")
        assert used == 6 and product.fb_fallback_count() == fb0
        _check(got, want, t_len, u_len)


@pytest.mark.parametrize("B,T,U", [(8, 800, 64), (8, 800, 96), (6, 600, 64), (4, 500, 48)])
def test_long_narrow_lattices_stay_on_the_time_parallel_kernels(product, oracle_mod, B, T, U):
    """T / U well beyond 6 with unbiased random rows (steep fronts): the boundary vectors take 16-token exponent groups
    there, and no utterance may need the log-domain re-run (with 32-token groups every one did at U=64 T=800)."""
    le, ls = make_inputs(B, T, U, seed=T + U)
    want = oracle_mod.forward_backward(le, ls)
    fb0 = product.fb_fallback_count()
    got, used = _run(product, le, ls, None, None, "device")
    assert used == 6 and product.fb_fallback_count() == fb0
    _check(got, want)

"""An independent pin of the authored lattice specification: tests/golden/lattice_autograd_golden.npz holds
log-likelihoods from a plain torch fp64 forward recursion and gradients from torch.autograd
(tests/golden/make_autograd_golden.py — no code shared with oracle/).  The oracle must reproduce it on the CPU; the
CUDA kernels, through the C-ABI, must reproduce it on the GPU, including the raw-logit entry point.

The reference has no forward-backward: parity of the lattice stays "unpinned by the reference";
this is the strongest pin available — two unrelated implementations of the written spec, and the kernels held to both."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "lattice_autograd_golden.npz"))


def _close(got, want, rtol, floor):
    """element-wise relative for entries >= floor, absolute rtol * floor below"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.all(np.abs(got - want) <= rtol * np.maximum(np.abs(want), floor))


def _chain(ge, gs, z):
    z = z.astype(np.float64)
    return ge / (1.0 + np.exp(z)) - gs / (1.0 + np.exp(-z))   # ge * sigmoid(-z) - gs * sigmoid(z)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_equals_torch_autograd(oracle_mod, tag):
    ll, loss, ge, gs = oracle_mod.forward_backward(G[tag + "_le"], G[tag + "_ls"], G[tag + "_t"], G[tag + "_u"])
    # (the oracle computes in fp64 and returns fp32 arrays: 2^-24 relative)
    assert _close(ll, G[tag + "_ll"], 2e-7, 1.0)
    assert _close(ge, G[tag + "_ge"], 2e-7, 1e-30) and _close(gs, G[tag + "_gs"], 2e-7, 1e-30)
    # the logit gradient the product chains: grad_emit * sigmoid(-z) - grad_shift * sigmoid(z).  (The fixture's le/ls are
    # the fp32 roundings of log sigmoid(+-z), its logit gradient is through exact fp64 log-sigmoids: 1e-6.)
    assert _close(_chain(ge, gs, G[tag + "_z"]), G[tag + "_gz"], 1e-5, 1e-6)
    assert _close(G[tag + "_llz"], G[tag + "_ll"], 1e-6, 1.0)


@pytest.mark.parametrize("tag", ["c", "d"])
def test_tone_oracle_equals_torch_autograd(oracle_mod, tag):
    r = oracle_mod.tone_latent_forward_backward(G[tag + "_le"], G[tag + "_ls"], G[tag + "_lt"], G[tag + "_t"], G[tag + "_u"])
    assert _close(r[0], G[tag + "_ll"], 2e-7, 1.0)
    for got, key in ((r[2], "_ge"), (r[3], "_gs"), (r[4], "_gt")):
        assert _close(got, G[tag + key], 2e-7, 1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_cuda_lattice_equals_torch_autograd(product, tag):
    import torch
    d = lambda k: torch.as_tensor(np.ascontiguousarray(G[k])).cuda()
    ll, loss, ge, gs = product.forward_backward(d(tag + "_le"), d(tag + "_ls"), d(tag + "_t"), d(tag + "_u"))
    assert _close(ll.cpu().numpy(), G[tag + "_ll"], 1e-5, 1e-3)
    assert _close(ge.cpu().numpy(), G[tag + "_ge"], 1e-4, 1e-6) and _close(gs.cpu().numpy(), G[tag + "_gs"], 1e-4, 1e-6)
    assert abs(float(loss.item()) + float(G[tag + "_ll"].sum())) <= 1e-5 * abs(float(G[tag + "_ll"].sum()))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("space", ["device", "host"])
def test_cuda_logit_entry_equals_torch_autograd(product, tag, space):
    import torch
    conv = (lambda k: torch.as_tensor(np.ascontiguousarray(G[k])).cuda()) if space == "device" else (lambda k: G[k])
    ll, loss, gz = product.forward_backward_logits(conv(tag + "_z"), conv(tag + "_t"), conv(tag + "_u"))
    np_ = lambda x: x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
    assert _close(np_(ll), G[tag + "_llz"], 1e-5, 1e-3)
    # gz is a difference of two occupancy terms: relative to the larger of |gz| and 1e-5
    assert _close(np_(gz), G[tag + "_gz"], 1e-4, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c", "d"])
def test_cuda_tone_lattice_equals_torch_autograd(product, tag):
    import torch
    d = lambda k: torch.as_tensor(np.ascontiguousarray(G[k])).cuda()
    r = product.tone_latent_forward_backward(d(tag + "_le"), d(tag + "_ls"), d(tag + "_lt"), d(tag + "_t"), d(tag + "_u"))
    assert _close(r[0].cpu().numpy(), G[tag + "_ll"], 1e-5, 1e-3)
    for got, key in ((r[2], "_ge"), (r[3], "_gs"), (r[4], "_gt")):
        assert _close(got.cpu().numpy(), G[tag + key], 1e-4, 1e-6)

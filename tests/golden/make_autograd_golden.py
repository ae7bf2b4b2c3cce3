"""Generates tests/golden/lattice_autograd_golden.npz: an INDEPENDENT pin of the authored lattice specification
(SURVEY.md §8 a-FB / a-TL).  Shares no code with oracle/: a plain torch fp64 log-space forward
recursion; the gradients come from torch.autograd, not from a backward recursion.  Run on CPU:

    python tests/golden/make_autograd_golden.py

The reference has no forward-backward, so this cannot make parity "pinned by the reference"; it
does show that two unrelated implementations of the written spec agree, and the CUDA kernels are held to both."""
import os

import numpy as np
import torch

NEG = -1.0e30  # finite stand-in for -inf: keeps autograd free of inf - inf


def lattice_ll(le, ls, T, U):
    """le, ls: [maxT, maxU] fp64.  alpha(0,0) = 0; a frame either emits (stay on u) or shifts (u -> u+1); no shift at
    the last token; the path ends with an emit at (T-1, U-1)."""
    if T <= 0 or U <= 0 or U > T:
        return None
    alpha = torch.full((U,), NEG, dtype=torch.float64)
    alpha = torch.cat([torch.zeros(1, dtype=torch.float64), alpha[1:]])
    for t in range(T - 1):
        stay = alpha + le[t, :U]
        move = torch.cat([torch.full((1,), NEG, dtype=torch.float64), (alpha + ls[t, :U])[:-1]])
        alpha = torch.logaddexp(stay, move)
    return alpha[U - 1] + le[T - 1, U - 1]


def tone_ll(le, ls, lt, T, U):
    """le, ls: [maxT, maxU, K]; lt: [maxU, K].  A stay keeps the token's tone, a shift draws the next token's tone."""
    K = lt.shape[-1]
    alpha = torch.cat([lt[0:1], torch.full((U - 1, K), NEG, dtype=torch.float64)])
    for t in range(T - 1):
        stay = alpha + le[t, :U]
        left = torch.logsumexp(alpha + ls[t, :U], dim=-1)          # [U]: mass leaving token u, any tone
        move = torch.cat([torch.full((1, K), NEG, dtype=torch.float64), lt[1:U] + left[:-1, None]])
        alpha = torch.logaddexp(stay, move)
    return torch.logsumexp(alpha[U - 1] + le[T - 1, U - 1], dim=-1)


def run_fb(le, ls, t_len, u_len):
    le = torch.tensor(le, dtype=torch.float64, requires_grad=True)
    ls = torch.tensor(ls, dtype=torch.float64, requires_grad=True)
    lls = []
    for b in range(le.shape[0]):
        ll = lattice_ll(le[b], ls[b], int(t_len[b]), int(u_len[b]))
        lls.append(ll)
    total = sum(x for x in lls if x is not None)
    total.backward()
    return (np.array([float(x) if x is not None else -np.inf for x in lls]), le.grad.numpy(), ls.grad.numpy())


def run_logits(z, t_len, u_len):
    z = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
    lls = [lattice_ll(le[b], ls[b], int(t_len[b]), int(u_len[b])) for b in range(z.shape[0])]
    sum(x for x in lls if x is not None).backward()
    return np.array([float(x) if x is not None else -np.inf for x in lls]), z.grad.numpy()


def run_tone(le, ls, lt, t_len, u_len):
    le = torch.tensor(le, dtype=torch.float64, requires_grad=True)
    ls = torch.tensor(ls, dtype=torch.float64, requires_grad=True)
    lt = torch.tensor(lt, dtype=torch.float64, requires_grad=True)
    lls = [tone_ll(le[b], ls[b], lt[b], int(t_len[b]), int(u_len[b])) for b in range(le.shape[0])]
    sum(lls).backward()
    return np.array([float(x) for x in lls]), le.grad.numpy(), ls.grad.numpy(), lt.grad.numpy()


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    # a: BASELINE configs[0] shape (B=1 U=32 T=120), full lengths; b: a ragged batch; both also through raw logits
    for tag, (B, T, U), ragged in (("a", (1, 120, 32), False), ("b", (4, 57, 20), True)):
        z = rng.standard_normal((B, T, U)).astype(np.float32)          # the fp32 values ARE the inputs
        t_len = np.full(B, T, np.int32)
        u_len = np.full(B, min(T, U), np.int32)
        if ragged:
            t_len = rng.integers(30, T + 1, B).astype(np.int32)
            u_len = np.array([rng.integers(8, min(U, t) + 1) for t in t_len], np.int32)
            t_len[0], u_len[0] = T, U
        le = torch.nn.functional.logsigmoid(torch.tensor(z, dtype=torch.float64)).numpy().astype(np.float32)
        ls = torch.nn.functional.logsigmoid(-torch.tensor(z, dtype=torch.float64)).numpy().astype(np.float32)
        ll, ge, gs = run_fb(le, ls, t_len, u_len)
        llz, gz = run_logits(z, t_len, u_len)
        out.update({f"{tag}_z": z, f"{tag}_le": le, f"{tag}_ls": ls, f"{tag}_t": t_len, f"{tag}_u": u_len,
                    f"{tag}_ll": ll, f"{tag}_ge": ge, f"{tag}_gs": gs, f"{tag}_llz": llz, f"{tag}_gz": gz})
    # c: tone-latent, K = 4 and d: K = 3 ragged
    for tag, (B, T, U, K), ragged in (("c", (2, 30, 12, 4), False), ("d", (3, 26, 10, 3), True)):
        z = rng.standard_normal((B, T, U, K))
        le = torch.nn.functional.logsigmoid(torch.tensor(z)).numpy().astype(np.float32)
        ls = torch.nn.functional.logsigmoid(-torch.tensor(z)).numpy().astype(np.float32)
        lt = torch.log_softmax(torch.tensor(rng.standard_normal((B, U, K))), dim=-1).numpy().astype(np.float32)
        t_len = np.full(B, T, np.int32)
        u_len = np.full(B, U, np.int32)
        if ragged:
            t_len = rng.integers(14, T + 1, B).astype(np.int32)
            u_len = np.array([rng.integers(4, min(U, t) + 1) for t in t_len], np.int32)
        ll, ge, gs, gt = run_tone(le, ls, lt, t_len, u_len)
        out.update({f"{tag}_le": le, f"{tag}_ls": ls, f"{tag}_lt": lt, f"{tag}_t": t_len, f"{tag}_u": u_len,
                    f"{tag}_ll": ll, f"{tag}_ge": ge, f"{tag}_gs": gs, f"{tag}_gt": gt})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lattice_autograd_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/lattice_golden.npz: small forward-backward and tone-latent cases with
the fp64 oracle's outputs.

The reference crate has no forward-backward, so these vectors are NOT reference
outputs; they freeze the authored specification (SURVEY.md §8 a-FB / a-TL) so that any later
change of the oracle or of the kernels is caught, and they carry the brute-force path-sum
log-likelihoods (an implementation that shares no code with either recursion).

    python tests/golden/make_lattice_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from lattice_util import brute_force_ll, brute_force_tone_ll, make_inputs  # noqa: E402


def main():
    oracle.build()
    out = {}
    # a-FB: config 1 of BASELINE.json (B=1 U=32 T=120) and a ragged small batch
    le, ls = make_inputs(1, 120, 32, seed=1234)
    ll, loss, ge, gs = oracle.forward_backward(le, ls, precision="f64")
    out.update(fb1_le=le, fb1_ls=ls, fb1_ll=ll, fb1_loss=np.float64(loss),
               fb1_ge=ge.astype(np.float32), fb1_gs=gs.astype(np.float32))
    le, ls = make_inputs(4, 9, 5, seed=7)
    t_len = np.array([9, 7, 5, 3], np.int32)
    u_len = np.array([5, 4, 5, 2], np.int32)
    ll, loss, ge, gs = oracle.forward_backward(le, ls, t_len, u_len, precision="f64")
    bf = np.array([brute_force_ll(le[b], ls[b], t_len[b], u_len[b]) for b in range(4)])
    out.update(fb2_le=le, fb2_ls=ls, fb2_t=t_len, fb2_u=u_len, fb2_ll=ll, fb2_bruteforce_ll=bf,
               fb2_ge=ge, fb2_gs=gs)
    # a-TL: K=2 small case with brute force, K=4 moderate case
    le, ls, lt = make_inputs(2, 6, 3, seed=11, K=2)
    r = oracle.tone_latent_forward_backward(le, ls, lt, precision="f64")
    bf = np.array([brute_force_tone_ll(le[b], ls[b], lt[b], 6, 3, 2) for b in range(2)])
    out.update(tl1_le=le, tl1_ls=ls, tl1_lt=lt, tl1_ll=r[0], tl1_bruteforce_ll=bf,
               tl1_ge=r[2], tl1_gs=r[3], tl1_gt=r[4])
    le, ls, lt = make_inputs(2, 20, 8, seed=12, K=4)
    r = oracle.tone_latent_forward_backward(le, ls, lt, precision="f64")
    out.update(tl2_le=le, tl2_ls=ls, tl2_lt=lt, tl2_ll=r[0], tl2_ge=r[2].astype(np.float32),
               tl2_gs=r[3].astype(np.float32), tl2_gt=r[4])
    np.savez_compressed(os.path.join(HERE, "lattice_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "lattice_golden.npz"), {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()

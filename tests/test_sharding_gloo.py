"""Multi-GPU host logic on CPU: two processes over ``gloo`` shard a batch of independent
utterances exactly as ``bench.py --gpus N`` does (contiguous batch shards, counter-based inputs
keyed by the GLOBAL utterance index, one all-reduce of the scalar loss and nothing else).  The
per-rank compute leg here is the CPU oracle — this test covers the partitioning and the
collective, not the kernels (those are the ``-m gpu`` tests)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, T, U, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_product
    from bench import synthetic_numpy
    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = load_product()
    lo, hi = P.shard_range(B, rank, world)
    le, ls = synthetic_numpy(lo, hi - lo, T, U)          # this rank's utterances only
    ll, loss, ge, gs = oracle.forward_backward(le, ls)
    t = torch.tensor([loss], dtype=torch.float64)
    P.all_reduce_loss(t)                                  # the path's only collective
    q.put((rank, lo, hi, float(t.item()), ll.tolist(), float(ge.sum() + gs.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_two_rank_batch_sharding_matches_single_rank(B):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from bench import synthetic_numpy
    import oracle
    oracle.build()

    T, U, world = 24, 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, T, U, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    le, ls = synthetic_numpy(0, B, T, U)
    ll, loss, ge, gs = oracle.forward_backward(le, ls)
    # shards tile the batch, in order, without overlap
    assert got[0][1] == 0 and got[-1][2] == B and got[0][2] == got[1][1]
    # every rank ends up with the same global loss = the single-process loss
    for r in got:
        assert abs(r[3] - loss) <= 1e-6 * abs(loss)   # the oracle returns the loss rounded to fp32
    # shard contents are independent of the rank count (inputs keyed by the global index)
    assert np.allclose(np.concatenate([r[4] for r in got]), ll, rtol=0, atol=0)
    assert abs(sum(r[5] for r in got) - float(ge.sum() + gs.sum())) < 1e-6


def test_shard_range_tiles_any_batch():
    sys.path.insert(0, ROOT)
    from conftest import load_product
    P = load_product()
    for B in (0, 1, 7, 32, 4096):
        for world in (1, 2, 3, 8):
            spans = [P.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1

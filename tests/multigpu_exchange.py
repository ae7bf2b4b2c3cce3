"""Multi-GPU check of the NVLink loss exchange (run under torchrun on a box with >= 2 GPUs; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_exchange.py
Every rank runs the lattice on its own batch shard; the exchanged, summed loss must equal the NCCL all-reduce of the
per-rank losses, eagerly and under CUDA-graph replay."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_product, synthetic_torch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P = load_product()
    P.lib()
    P.connect_loss_exchange()
    B, T, U = 6, 200, 128
    ok = True
    for it in range(3):
        le, ls = synthetic_torch((it * world + rank) * B, B, T, U, dev)
        ll, loss, ge, gs = P.forward_backward(le, ls)
        red = P.loss_allreduce()
        ref = loss.double().clone()
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        ok &= abs(float(red.item()) - float(ref.item())) <= 1e-5 * abs(float(ref.item()))
    # graph replay
    ws = torch.empty(P.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device=dev)
    out = (torch.empty(B, device=dev), torch.empty(1, device=dev), torch.empty(B, T, U, device=dev), torch.empty(B, T, U, device=dev))
    le, ls = synthetic_torch(rank * B, B, T, U, dev)
    P.forward_backward(le, ls, workspace=ws, out=out)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(4):
            P.forward_backward(le, ls, workspace=ws, out=out)
    for _ in range(5):
        g.replay()
    red = P.loss_allreduce()
    ref = out[1].double().clone()
    dist.all_reduce(ref)
    torch.cuda.synchronize()
    ok &= abs(float(red.item()) - float(ref.item())) <= 1e-5 * abs(float(ref.item()))
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multigpu_exchange:", "ok" if t.item() == 1.0 else "FAILED", "world", world, "loss", float(red.item()), flush=True)
    dist.barrier()
    P.disconnect_loss_exchange()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

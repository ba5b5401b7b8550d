"""Multi-GPU check of the data-parallel training path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py

Every rank trains `VISITS` label visits on its shard (rows keyed by GLOBAL row index); rank 0 also runs the same
visits on ONE GPU with the global batch and compares all parameters (the sums are re-associated across ranks, so
the multi-visit comparison is informational).  All ranks must hold bit-identical parameters afterwards, and ONE
optimiser step of each kind (no update) must reproduce the single-GPU gradients / statistics / losses (tools/dp_parity.py).  Also prints
the time per visit.  CVG_DISABLE_NVL=1 selects NCCL for the exchanges instead of the peer-memory all-reduce."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from cvae_gan_b200 import models
from cvae_gan_b200.engine import Engine

F_, K_, Z_ = 10, 5, 128
B_LOCAL = int(os.environ.get("DP_B", "1024"))
VISITS = int(os.environ.get("DP_VISITS", "4"))
# parameters with a mathematically zero gradient (pre-BatchNorm biases, one-hot label columns): Adam turns their round-off
# gradients into +-lr steps, so they are not comparable between two summation orders (tests/parity.py explains)
SKIP = {"encoder.0.bias", "encoder.3.bias", "encoder.6.bias", "main_model.0.bias", "main_model.3.bias", "main_model.6.bias",
        "encoder.0.weight", "main_model.0.weight"}


def init_engine(world, rank, dev, B):
    eng = Engine(F_, K_, Z_, max_batch=B, world_size=world, rank=rank)
    torch.manual_seed(0)
    mods = [models.CVAEGANEncoderModel(F_, K_, Z_), models.CVAEGANGeneratorModel(Z_, K_, F_),
            models.CVAEGANDiscriminatorModel(F_, K_), models.CVAEGANClassifierModel(F_, K_)]
    for net, m in enumerate(mods):
        eng.load_state(net, m.state_dict())
    return eng


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(3)
    tabs = [(torch.rand(F_, generator=g) + 0.08 * torch.randn(50000, F_, generator=g)).clamp(0, 1).to(dev) for _ in range(K_)]
    Bg = B_LOCAL * world
    eng = init_engine(world, rank, dev, B_LOCAL)
    eng.ctl_set(seed=77, counter=0, lambda_class=0.25)
    loss = torch.zeros(13, 4, device=dev)
    for v in range(VISITS):
        eng.visit(v % K_, Bg, class_rows=tabs[v % K_], loss_out=loss)
    torch.cuda.synchronize()
    # timing (eager launches; the bench uses CUDA graphs)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for v in range(10):
        eng.visit(v % K_, Bg, class_rows=tabs[v % K_], loss_out=loss)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    states = {net: eng.export_state(net) for net in range(4)}
    # bit-identical replicas
    ok_rep = True
    for net in range(4):
        for k, t in states[net].items():
            if not t.is_floating_point():
                continue
            ref = t.clone()
            dist.broadcast(ref, src=0)
            if not torch.equal(ref, t):
                ok_rep = False
    flag = torch.tensor([1 if ok_rep else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world} nvl={'off' if os.environ.get('CVG_DISABLE_NVL') == '1' else 'on'}: {dt * 1e3:.3f} ms per visit (eager), "
              f"replicas bit-identical: {bool(flag.item())}", flush=True)
        # single-GPU reference on the global batch
        one = init_engine(1, 0, dev, Bg)
        one.ctl_set(seed=77, counter=0, lambda_class=0.25)
        loss1 = torch.zeros(13, 4, device=dev)
        for v in list(range(VISITS)) + list(range(10)):
            one.visit(v % K_, Bg, class_rows=tabs[v % K_], loss_out=loss1)
        torch.cuda.synchronize()
        worst, worst_key = 0.0, ""
        for net in range(4):
            ref = one.export_state(net)
            for k, t in states[net].items():
                if not t.is_floating_point() or k in SKIP:
                    continue
                d = (t - ref[k]).abs().max().item() / (ref[k].abs().max().item() + 1e-12)
                if d > worst:
                    worst, worst_key = d, f"{net}/{k}"

        print(f"max relative deviation of any tensor vs the single-GPU run on the global batch: {worst:.3e} ({worst_key})", flush=True)
        print("losses dp :", [round(x, 5) for x in loss[-1].tolist()], flush=True)
        print("losses one:", [round(x, 5) for x in loss1[-1].tolist()], flush=True)
        print("(multi-visit trajectories differ by Adam-amplified round-off on both sides; the verdict is the one-step check below)", flush=True)
    # the verdict: one D / C / G step without update against a single-GPU run on the global batch, tensor by tensor
    from tools.dp_parity import run_dp_parity
    par = run_dp_parity(world, rank, dev, B_LOCAL)
    if rank == 0:
        print("dp_parity", par, flush=True)
        print("DP_CHECK", "OK" if (flag.item() == 1 and par["ok"]) else "FAIL", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

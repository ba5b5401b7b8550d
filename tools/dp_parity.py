"""One-step data-parallel parity on real GPUs (called by bench.py at world > 1 before timing, and by tools/dp_check.py).

Data: SURVEY 8(d) C3 shape - K = 4 classes with fractions [0.90, 0.06, 0.03, 0.01] (scaled down to 200 000 rows so the
check takes a second): label 0 has more rows than the global batch (cvae_gan.py:257-260, randperm branch), label 3 has
fewer (cvae_gan.py:250-253, randint-with-replacement branch).

For each of the two labels every rank runs ONE critic step, ONE classifier step and ONE encoder/generator step with
CVG_STEP_NO_UPDATE on its shard of the global batch (rows and noise keyed by GLOBAL row index); rank 0 runs the same three
steps on one GPU with the whole global batch.  Compared on rank 0, per tensor, relative to the tensor's largest entry:
the all-reduced gradients of every network, the BatchNorm running statistics and spectral-norm u / v the forward passes
mutated, and the losses.  Gradients that are mathematically zero (pre-BatchNorm biases, one-hot label columns of E / G) are
round-off on both sides and are skipped.
"""
from __future__ import annotations

import torch

F_, K_, Z_ = 10, 4, 128
FRACTIONS = (0.90, 0.06, 0.03, 0.01)
ROWS = 200_000
SKIP = ("encoder.0.bias", "encoder.3.bias", "encoder.6.bias", "main_model.0.bias", "main_model.3.bias", "main_model.6.bias")
ONE_HOT = {"encoder.0.weight": F_, "main_model.0.weight": Z_}      # columns from here on are the label columns


def _tables(dev):
    g = torch.Generator().manual_seed(11)
    tabs = []
    for k, fr in enumerate(FRACTIONS):
        n = int(round(ROWS * fr))
        c = torch.rand(F_, generator=g)
        tabs.append((c + 0.08 * torch.randn(n, F_, generator=g)).clamp(0, 1).to(dev))
    return tabs


def _engine(world, rank, B):
    from cvae_gan_b200 import models
    from cvae_gan_b200.engine import Engine
    eng = Engine(F_, K_, Z_, max_batch=B, world_size=world, rank=rank)
    torch.manual_seed(0)
    mods = [models.CVAEGANEncoderModel(F_, K_, Z_), models.CVAEGANGeneratorModel(Z_, K_, F_),
            models.CVAEGANDiscriminatorModel(F_, K_), models.CVAEGANClassifierModel(F_, K_)]
    for net, m in enumerate(mods):
        eng.load_state(net, m.state_dict())
    return eng


def _three_steps(eng, tabs, label, batch_global, counter0):
    """Returns {name: tensor} of everything the three steps produce."""
    from cvae_gan_b200._lib import STEP_NO_UPDATE
    out = {}
    seed = 4242
    for i, kind in enumerate("dcg"):
        c = counter0 + 2 * i
        x = eng.sample_rows(tabs[label], batch_global, seed=seed, counter=c)
        eng.zero_grads()
        loss = torch.zeros(4, device=eng.device)
        if kind == "d":
            eng.step_d(x, label, seed=seed, counter=c + 1, flags=STEP_NO_UPDATE, loss_out=loss)
            nets = (2,)
        elif kind == "c":
            eng.step_c(x, label, seed=seed, counter=c + 1, flags=STEP_NO_UPDATE, loss_out=loss)
            nets = (3,)
        else:
            eng.step_g(x, label, 0.25, seed=seed, counter=c + 1, flags=STEP_NO_UPDATE, loss_out=loss)
            nets = (0, 1)
        torch.cuda.synchronize()
        out[f"{kind}/loss"] = loss.clone()
        for net in nets:
            for key, (kd, shape, off) in eng.tables[net].items():
                if kd == 0:
                    out[f"{kind}/grad/{net}/{key}"] = eng.view(net, key, "grads").clone()
    for net in range(4):
        for key, (kd, shape, off) in eng.tables[net].items():
            if kd == 1:
                out[f"state/{net}/{key}"] = eng.view(net, key).clone()
    return out


def run_dp_parity(world: int, rank: int, dev, b_local: int, tol: float = 2e-5) -> dict:
    """Collective over all ranks.  Returns {"ok", "max_rel", "worst", ...} (identical on every rank)."""
    import torch.distributed as dist
    tabs = _tables(dev)
    bg = b_local * world
    dp = _engine(world, rank, b_local)
    one = _engine(1, 0, bg) if rank == 0 else None
    worst, worst_key, n_cmp = 0.0, "", 0
    for li, label in enumerate((0, 3)):
        got = _three_steps(dp, tabs, label, bg, counter0=100 * li)
        if rank == 0:
            ref = _three_steps(one, tabs, label, bg, counter0=100 * li)
            for k, r in ref.items():
                base = k.rsplit("/", 1)[-1]
                if "/grad/" in k and base in SKIP:
                    continue
                a, b = got[k].double(), r.double()
                if "/grad/" in k and base in ONE_HOT:
                    a, b = a[:, :ONE_HOT[base]], b[:, :ONE_HOT[base]]
                scale = float(b.abs().max())
                if scale == 0.0:
                    continue
                rel = float((a - b).abs().max()) / scale
                n_cmp += 1
                if rel > worst:
                    worst, worst_key = rel, f"label{label}/{k}"
    res = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.broadcast(res, src=0)
    worst = float(res.item())
    if rank == 0:
        one.close()
    dist.barrier()
    dp.close()
    return {"ok": bool(worst <= tol), "max_rel": worst, "worst": worst_key, "tol": tol, "tensors_compared": n_cmp,
            "world": world, "batch_global": bg, "labels": [0, 3],
            "what": "one D, C and E/G step (no update) per label on K=4 [0.90,0.06,0.03,0.01] data: all-reduced gradients, "
                    "BatchNorm running stats, spectral-norm u/v and losses vs a single-GPU run on the global batch",
            "executor": "program" if dp_mode() else "ffma"}


def dp_mode() -> int:
    import os
    return 1 if os.environ.get("CVG_TRAIN_MODE", "") == "mk" else 0

"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

The reference ships no golden vectors, so the pin for `oracle/cvae_gan_oracle.py`
is the reference's own outputs on seeded inputs:

  ref_fit_a.npz   CVAEGAN.fit for 2 epochs (e = 0,1  -> lambda_class = 0), F=10, K=5, B=64, class
                  sizes [200, 200, 64, 30, 200] (exercises the randperm / all-rows / randint
                  branches of _get_target_samples, cvae_gan.py:247-260), then generate_samples and
                  generate_qualified_samples at two thresholds.
  ref_fit_b.npz   same start state, 2 epochs with e = 350,351 (lambda_class = 0.25 ramp,
                  cvae_gan.py:198-204).  The reference hard-codes `range(epochs)`; we shadow the
                  name `range` in the reference module's *namespace* (source untouched) so the
                  loop variable starts at 350.
  ref_filter.npz  softmax/max/threshold decisions (cvae_gan.py:366-370) on seeded logits.

Stored: the start state (all four networks, reference state_dict keys), the data, the seeds,
and the reference's results (loss_history, final state, generated tensors).  The oracle replays
them with torch's CPU generator seeded identically (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import builtins
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.refload import load_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
F_, K, B = 10, 5, 64
CLASS_SIZES = [200, 200, 64, 30, 200]
FIT_SEED, GEN_SEED = 1234, 4321


def make_data(seed=0):
    from sklearn.datasets import make_blobs
    from sklearn.preprocessing import minmax_scale
    x, y = make_blobs(n_samples=CLASS_SIZES, n_features=F_, centers=None, random_state=seed)
    x = minmax_scale(x).astype(np.float32)
    perm = np.random.RandomState(seed).permutation(len(y))
    return x[perm], y[perm].astype(np.int64)


def flat_state(gan):
    out = {}
    for net in ("encoder", "generator", "discriminator", "classifier"):
        for k, v in getattr(gan, net).state_dict().items():
            out[f"{net}/{k}"] = v.detach().cpu().numpy().copy()
    return out


def run_fit(src, x, y, epoch_offset, init_state=None):
    src.datasets.feature_num, src.datasets.label_num = F_, K
    src.datasets.tr_samples, src.datasets.tr_labels = torch.from_numpy(x), torch.from_numpy(y)
    gc = src.config.gan_config
    gc.batch_size, gc.epochs = B, 2
    src.utils.set_random_state()  # seed 0, as tests/test_cvae_gan.py:18 does before construction
    gan = src.CVAEGAN()
    if init_state is not None:
        for net in ("encoder", "generator", "discriminator", "classifier"):
            sd = {k.split("/", 1)[1]: torch.from_numpy(v) for k, v in init_state.items()
                  if k.startswith(net + "/")}
            getattr(gan, net).load_state_dict(sd)
    init = flat_state(gan)
    mod = sys.modules["src.cvae_gan"]
    if epoch_offset:
        mod.range = lambda n: builtins.range(epoch_offset, epoch_offset + n)
    try:
        torch.manual_seed(FIT_SEED)
        gan.fit(src.datasets.TrDataset())
    finally:
        if "range" in vars(mod):
            del mod.range
    return gan, init


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    src = load_reference()
    torch.set_num_threads(1)
    x, y = make_data()

    # ---- A: e = 0,1 ---------------------------------------------------------------------------
    gan, init = run_fit(src, x, y, 0)
    out = {"x": x, "y": y, "meta": np.array([F_, K, B, FIT_SEED, GEN_SEED, 0], dtype=np.int64)}
    out.update({"init/" + k: v for k, v in init.items()})
    out.update({"final/" + k: v for k, v in flat_state(gan).items()})
    for k, v in gan.loss_history.items():
        out["loss/" + k] = np.array(v, dtype=np.float64)
    out["sample_keys"] = np.array(list(gan.samples.keys()), dtype=np.int64)
    # generation after fit (all nets in eval mode, cvae_gan.py:233-236)
    torch.manual_seed(GEN_SEED)
    out["gen/samples_l1_n37"] = gan.generate_samples(1, 37).numpy()
    for thr in (0.2, 0.5):
        for lab in (0, 3):
            q = gan.generate_qualified_samples(lab, 25, thr)
            out[f"gen/qualified_l{lab}_thr{thr}"] = q.numpy().reshape(-1, F_) if q.numel() else np.zeros((0, F_), np.float32)
    out["gen/classifier_training_after"] = np.array([int(gan.classifier.training)])
    np.savez_compressed(os.path.join(GOLDEN, "ref_fit_a.npz"), **out)
    print("A losses", {k: v for k, v in gan.loss_history.items()})

    # ---- B: e = 350,351 (lambda_class != 0), same start state -----------------------------------
    gan_b, init_b = run_fit(src, x, y, 350, init_state=init)
    for k in init:
        assert np.array_equal(init[k], init_b[k]), k
    out_b = {"meta": np.array([F_, K, B, FIT_SEED, GEN_SEED, 350], dtype=np.int64)}
    fin = flat_state(gan_b)
    # digests keep the fixture small: per tensor [sum, sum of squares, first 8, last 8]
    for k, v in fin.items():
        f = v.astype(np.float64).ravel()
        pad = np.zeros(16)
        pad[:min(8, f.size)] = f[:8]
        pad[8:8 + min(8, f.size)] = f[-8:]
        out_b["digest/" + k] = np.concatenate([[f.sum(), (f * f).sum()], pad])
    for k, v in gan_b.loss_history.items():
        out_b["loss/" + k] = np.array(v, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "ref_fit_b.npz"), **out_b)
    print("B losses", {k: v for k, v in gan_b.loss_history.items()})

    # ---- filter decisions -------------------------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(4096, K, generator=g) * 3.0
    logits[:64] = 0.0                                   # exact ties -> first index wins
    logits[64:128, 2] = logits[64:128, 0]               # pairwise ties
    fo = {"logits": logits.numpy()}
    for lab in range(K):
        for thr in (0.0, 0.2, 0.5, 0.9):
            probs = torch.softmax(logits, dim=1)
            mp, pr = torch.max(probs, dim=1)
            fo[f"keep_l{lab}_thr{thr}"] = np.packbits(((mp > thr) & (pr == lab)).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "ref_filter.npz"), **fo)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY - generates tests/golden/ref_cvae_*.npz by RUNNING THE UNMODIFIED REFERENCE's sibling trainer
`CVAE` (src/cvae.py), the pin for `OracleCVAEGAN.step_g_cvae / fit_cvae`.

    python oracle/make_golden_cvae.py          (build container only: needs /root/reference)

  ref_cvae_a.npz  `set_random_state(); CVAE()` starting state, CVAE.fit for 2 epochs (e = 0,1 -> lambda_class = 0), same data /
                  class sizes / batch as ref_fit_a.npz, then generate_samples, generate_qualified_samples and
                  reconstruct_samples (which, unlike the CVAE-GAN's, works: cvae.py:300-319).
  ref_cvae_b.npz  same start state, e = 350,351 (lambda_class ramp on the RECONSTRUCTION's class loss, cvae.py:145-151);
                  final state as per-tensor digests.
"""
from __future__ import annotations

import builtins
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.make_golden import B, F_, FIT_SEED, GEN_SEED, GOLDEN, K, make_data  # noqa: E402
from oracle.refload import load_reference  # noqa: E402

NETS3 = ("encoder", "generator", "classifier")


def flat_state(gan):
    return {f"{net}/{k}": v.detach().cpu().numpy().copy() for net in NETS3 for k, v in getattr(gan, net).state_dict().items()}


def run_fit(src, x, y, epoch_offset, init_state=None):
    src.datasets.feature_num, src.datasets.label_num = F_, K
    src.datasets.tr_samples, src.datasets.tr_labels = torch.from_numpy(x), torch.from_numpy(y)
    gc = src.config.gan_config
    gc.batch_size, gc.epochs = B, 2
    src.utils.set_random_state()
    mod = sys.modules["src.cvae"]
    gan = mod.CVAE()                      # src/__init__.py may not re-export the sibling class
    if init_state is not None:
        for net in NETS3:
            sd = {k.split("/", 1)[1]: torch.from_numpy(v) for k, v in init_state.items() if k.startswith(net + "/")}
            getattr(gan, net).load_state_dict(sd)
    init = flat_state(gan)
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
    src = load_reference()
    import importlib
    importlib.import_module("src.cvae")
    torch.set_num_threads(1)
    x, y = make_data()
    gan, init = run_fit(src, x, y, 0)
    out = {"x": x, "y": y, "meta": np.array([F_, K, B, FIT_SEED, GEN_SEED, 0], dtype=np.int64)}
    out.update({"init/" + k: v for k, v in init.items()})
    out.update({"final/" + k: v for k, v in flat_state(gan).items()})
    for k, v in gan.loss_history.items():
        out["loss/" + k] = np.array(v, dtype=np.float64)
    out["sample_keys"] = np.array(list(gan.samples.keys()), dtype=np.int64)
    torch.manual_seed(GEN_SEED)
    out["gen/samples_l1_n37"] = gan.generate_samples(1, 37).numpy()
    for thr in (0.2, 0.5):
        for lab in (0, 3):
            q = gan.generate_qualified_samples(lab, 25, thr)
            q = torch.stack(list(q)) if isinstance(q, (list, tuple)) and len(q) else q
            q = q if torch.is_tensor(q) else torch.zeros(0, F_)
            out[f"gen/qualified_l{lab}_thr{thr}"] = q.numpy().reshape(-1, F_) if q.numel() else np.zeros((0, F_), np.float32)
    # reconstruct_samples on mixed labels (eval-mode E and G; leaves both in train mode, cvae.py:315-316)
    xs, ys = torch.from_numpy(x[::37]).clone(), torch.from_numpy(y[::37]).clone()
    rec = gan.reconstruct_samples(xs, ys)
    out["rec/x"], out["rec/y"], out["rec/out"] = xs.numpy(), ys.numpy(), rec.numpy()
    out["rec/modes_after"] = np.array([gan.encoder.training, gan.generator.training, gan.classifier.training])
    np.savez_compressed(os.path.join(GOLDEN, "ref_cvae_a.npz"), **out)
    print("A losses", gan.loss_history)

    gan_b, init_b = run_fit(src, x, y, 350, init_state=init)
    for k in init:
        assert np.array_equal(init[k], init_b[k]), k
    out_b = {"meta": np.array([F_, K, B, FIT_SEED, GEN_SEED, 350], dtype=np.int64)}
    for k, v in flat_state(gan_b).items():
        f = v.astype(np.float64).ravel()
        pad = np.zeros(16)
        pad[:min(8, f.size)] = f[:8]
        pad[8:8 + min(8, f.size)] = f[-8:]
        out_b["digest/" + k] = np.concatenate([[f.sum(), (f * f).sum()], pad])
    for k, v in gan_b.loss_history.items():
        out_b["loss/" + k] = np.array(v, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "ref_cvae_b.npz"), **out_b)
    print("B losses", gan_b.loss_history)


if __name__ == "__main__":
    main()

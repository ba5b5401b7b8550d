"""TEST INFRASTRUCTURE ONLY - generates tests/golden/ref_vaegan.npz by RUNNING THE UNMODIFIED REFERENCE's sibling trainer
`VAEGAN` (src/vae_gan.py), the pin for `OracleCVAEGAN.step_g_vaegan / fit_vaegan` and for the unconditional forwards
(SURVEY.md 8 f4).

    python oracle/make_golden_vaegan.py          (build container only: needs /root/reference)

  ref_vaegan.npz  `set_random_state(); VAEGAN()` starting state, VAEGAN.fit for 4 epochs (each: 5 critic + 3 encoder/generator
                  steps on batches of 64 drawn from ALL rows), same data as ref_fit_a.npz (labels unused), then generate_samples
                  and reconstruct_samples.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.make_golden import B, F_, FIT_SEED, GEN_SEED, GOLDEN, K, make_data  # noqa: E402
from oracle.refload import load_reference  # noqa: E402

NETS3 = ("encoder", "generator", "discriminator")
EPOCHS = 4


def flat_state(gan):
    return {f"{net}/{k}": v.detach().cpu().numpy().copy() for net in NETS3 for k, v in getattr(gan, net).state_dict().items()}


def main():
    src = load_reference()
    mod = importlib.import_module("src.vae_gan")
    torch.set_num_threads(1)
    x, y = make_data()
    src.datasets.feature_num, src.datasets.label_num = F_, K
    src.datasets.tr_samples, src.datasets.tr_labels = torch.from_numpy(x), torch.from_numpy(y)
    gc = src.config.gan_config
    gc.batch_size, gc.epochs = B, EPOCHS
    src.utils.set_random_state()
    gan = mod.VAEGAN()
    out = {"x": x, "meta": np.array([F_, K, B, FIT_SEED, GEN_SEED, EPOCHS], dtype=np.int64)}
    out.update({"init/" + k: v for k, v in flat_state(gan).items()})
    torch.manual_seed(FIT_SEED)
    gan.fit(src.datasets.TrDataset())
    out.update({"final/" + k: v for k, v in flat_state(gan).items()})
    for k, v in gan.loss_history.items():
        out["loss/" + k] = np.array(v, dtype=np.float64)
    out["n_samples"] = np.array([len(gan.samples)], dtype=np.int64)
    torch.manual_seed(GEN_SEED)
    out["gen/samples_n41"] = gan.generate_samples(41).numpy()
    xs = torch.from_numpy(x[::29]).clone()
    out["rec/x"], out["rec/out"] = xs.numpy(), gan.reconstruct_samples(xs).numpy()
    out["rec/modes_after"] = np.array([gan.encoder.training, gan.generator.training, gan.discriminator.training])
    np.savez_compressed(os.path.join(GOLDEN, "ref_vaegan.npz"), **out)
    print("losses", gan.loss_history)


if __name__ == "__main__":
    main()

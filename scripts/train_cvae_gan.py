#!/usr/bin/env python
"""Counterpart of /root/reference/scripts/train_cvae_gan.py on the B200 engine.

    python scripts/train_cvae_gan.py --x-train x_train.csv --y-train y_train.csv --x-test x_test.csv --y-test y_test.csv
    python scripts/train_cvae_gan.py --synthetic            # OTIDS-shaped imbalanced blobs (no dataset needed)

CSV layout as produced by the reference's sample_can_hcrl_otids.py: feature rows, one-hot label rows, no header."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import cvae_gan_b200 as pkg
from cvae_gan_b200 import pipeline


def load_csv(path):
    import pandas as pd
    data = pd.read_csv(path, header=None, low_memory=False)
    for col in data.columns:
        data[col] = pd.to_numeric(data[col], errors='coerce')
    return torch.tensor(data.fillna(0).values, dtype=torch.float32)


TRAINERS = {"cvae_gan": "CVAEGAN", "cgan": "CGAN", "cvae": "CVAE"}


def main(default_trainer="cvae_gan"):
    ap = argparse.ArgumentParser()
    for a in ("--x-train", "--y-train", "--x-test", "--y-test"):
        ap.add_argument(a)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--clf-epochs", type=int, default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--trainer", default=default_trainer, choices=sorted(TRAINERS),
                    help="cvae_gan (scripts/train_cvae_gan.py), cgan (scripts/train_cgan.py) or cvae (scripts/train_cvae.py) of the reference")
    args = ap.parse_args()
    args.out = args.out or f"data_{args.trainer}.pkl"
    ds = pkg.datasets
    if args.synthetic:
        from sklearn.datasets import make_blobs
        x, y = make_blobs(n_samples=[18000, 1200, 600, 200], n_features=10, centers=None, cluster_std=1.5, random_state=0)
        perm = np.random.RandomState(0).permutation(len(y))
        x, y = torch.from_numpy(x[perm].astype(np.float32)), torch.from_numpy(y[perm].astype(np.int64))
        n_tr = int(0.8 * len(y))
        ds.tr_samples, ds.tr_labels, ds.te_samples, ds.te_labels = x[:n_tr], y[:n_tr], x[n_tr:], y[n_tr:]
    else:
        ds.tr_samples, ds.te_samples = load_csv(args.x_train), load_csv(args.x_test)
        ds.tr_labels = torch.argmax(load_csv(args.y_train), dim=1)
        ds.te_labels = torch.argmax(load_csv(args.y_test), dim=1)
    if args.epochs is not None:
        pkg.config.gan_config.epochs = args.epochs
    if args.clf_epochs is not None:
        pkg.config.classifier_config.epochs = args.clf_epochs
    gan, clf, rep = pipeline.run(ds, pkg.config, pickle_path=args.out, verbose=True, trainer=getattr(pkg, TRAINERS[args.trainer]))
    print("class counts before:", rep["class_counts_before"])
    print("generation:", rep["generation"])
    print("augmented train rows:", rep["train_rows"])
    print(rep["confusion_matrix"])
    print("multi-class:", {k: round(v, 4) for k, v in rep["metrics"].items()})
    print("binary     :", {k: round(v, 4) for k, v in rep["binary_metrics"].items()})


if __name__ == "__main__":
    main()

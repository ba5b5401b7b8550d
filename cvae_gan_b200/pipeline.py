"""The reference's driver around the hot path (/root/reference/scripts/train_cvae_gan.py; scripts/train_cgan.py and
scripts/train_cvae.py are the same driver around the sibling trainers - `run(trainer=CGAN | CVAE)`): min-max scaling of the
concatenated train+test features (:19-43), `CVAEGAN.fit` (:47-51), the class-balancing loop that tops every class up to
the majority count with `generate_qualified_samples` (:60-95), the pickle hand-off `(tr_x, tr_y, te_x, te_y)` as numpy
arrays (:131-140) and the downstream `Classifier` fine-tune + multi-class / binary test (:143-175).

Everything numerical runs on the CUDA engine; the scaling is two reductions and one elementwise pass on the device.
`run()` works on the package's dataset globals (cvae_gan_b200.datasets) exactly like the script works on `src.datasets`.
"""
from __future__ import annotations

import pickle
from typing import Optional

import torch

from . import config as _config
from . import datasets as _datasets
from .classifier import Classifier
from .cvae_gan import CVAEGAN


def set_random_state(config=_config):
    """utils.set_random_state (src/utils.py): seed python / numpy / torch with config.seed."""
    import random
    import numpy as np
    random.seed(config.seed)
    np.random.seed(config.seed)
    torch.manual_seed(config.seed)


def minmax_scale_(datasets=_datasets, device: Optional[str] = None):
    """train_cvae_gan.py:19-43: column-wise min-max over train+test together, then shift so the minimum is 0."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    n_tr = len(datasets.tr_samples)
    x = torch.cat([datasets.tr_samples, datasets.te_samples]).to(dev, torch.float32)
    lo, hi = x.min(dim=0).values, x.max(dim=0).values
    span = hi - lo
    span[span == 0] = 1.0                      # sklearn: constant columns scale by 1
    x = (x - lo) / span
    x = x - x.min()
    datasets.tr_samples, datasets.te_samples = x[:n_tr].cpu(), x[n_tr:].cpu()
    datasets.feature_num = int(x.shape[1])
    datasets.label_num = int(datasets.tr_labels.max()) + 1      # utils.set_dataset_values (utils.py:77-83): max label + 1


def balance(gan: CVAEGAN, datasets=_datasets, verbose: bool = False):
    """train_cvae_gan.py:60-95: every class is topped up to the largest class with qualified generated samples, which
    are appended to datasets.tr_samples / tr_labels.  Returns {label: {'target': n, 'actual': m}}."""
    max_cnt = max(len(gan.samples[i]) for i in gan.samples.keys())
    stats = {}
    for i in gan.samples.keys():
        need = max_cnt - len(gan.samples[i])
        stats[i] = {'target': need, 'actual': 0}
        if need <= 0:
            continue
        generated = gan.generate_qualified_samples(i, need)
        got = len(generated)
        stats[i]['actual'] = got
        if verbose:
            print(f"class {i}: target {need}, generated {got}")
        if got > 0:
            datasets.tr_samples = torch.cat([datasets.tr_samples, generated])
            datasets.tr_labels = torch.cat([datasets.tr_labels, torch.full([got], i, dtype=datasets.tr_labels.dtype)])
    assert len(datasets.tr_samples) == len(datasets.tr_labels)
    return stats


def dump_dataset(path: str, datasets=_datasets):
    """train_cvae_gan.py:131-140: the (tr_x, tr_y, te_x, te_y) numpy tuple other tools of the reference consume."""
    with open(path, 'wb') as f:
        pickle.dump((datasets.tr_samples.numpy(), datasets.tr_labels.numpy(), datasets.te_samples.numpy(),
                     datasets.te_labels.numpy()), f)


def run(datasets=_datasets, config=_config, pickle_path: Optional[str] = None, verbose: bool = False, trainer=None,
        name: Optional[str] = None):
    """The whole script; returns (gan, clf, report).  `trainer`: the host class to train - `CVAEGAN` (default), or a sibling
    with the same surface, `CGAN` / `CVAE`: the reference's scripts/train_cgan.py and scripts/train_cvae.py are this same driver
    around `src.CGAN()` / `src.CVAE()`; `name`: the `Classifier(name)` tag ('CVAE_GAN', 'CGAN', 'CVAE')."""
    trainer = trainer or CVAEGAN
    name = name or {"CVAEGAN": "CVAE_GAN"}.get(trainer.__name__, trainer.__name__)
    set_random_state(config)
    minmax_scale_(datasets)
    set_random_state(config)
    gan = trainer(config=config, datasets=datasets)
    gan.fit(datasets.TrDataset())
    before = {i: len(gan.samples[i]) for i in gan.samples.keys()}
    stats = balance(gan, datasets, verbose)
    if pickle_path:
        dump_dataset(pickle_path, datasets)
    set_random_state(config)
    clf = Classifier(name)
    clf.model = gan.classifier                      # train_cvae_gan.py:145
    clf.fit(datasets.TrDataset())
    clf.test(datasets.TeDataset())
    multi = dict(clf.metrics)
    cm = clf.confusion_matrix
    clf.binary_test(datasets.TeDataset())
    report = {"class_counts_before": before, "generation": stats, "metrics": multi, "confusion_matrix": cm,
              "binary_metrics": dict(clf.metrics), "train_rows": len(datasets.tr_samples)}
    return gan, clf, report

"""End-to-end driver (cvae_gan_b200/pipeline.py = /root/reference/scripts/train_cvae_gan.py) on small synthetic data:
scaling, CVAEGAN.fit, class balancing with qualified samples, the pickle hand-off format, Classifier fit / test."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("trainer", ["CVAEGAN", "CGAN", "CVAE"])
def test_pipeline_end_to_end(tmp_path, trainer):
    """`trainer`: scripts/train_cvae_gan.py, scripts/train_cgan.py and scripts/train_cvae.py of the reference are the same driver
    around the three trainer classes."""
    import cvae_gan_b200 as pkg
    from cvae_gan_b200 import pipeline
    from sklearn.datasets import make_blobs
    x, y = make_blobs(n_samples=[900, 300, 120], n_features=10, centers=None, cluster_std=1.0, random_state=3)
    perm = np.random.RandomState(3).permutation(len(y))
    x, y = torch.from_numpy((50 * x[perm]).astype(np.float32)), torch.from_numpy(y[perm].astype(np.int64))
    ds, cfg = pkg.datasets, pkg.config
    ds.tr_samples, ds.tr_labels, ds.te_samples, ds.te_labels = x[:1000], y[:1000], x[1000:], y[1000:]
    gc, cc = cfg.gan_config, cfg.classifier_config
    key = getattr(pkg, trainer)._CONFIG_KEY
    old = (gc.epochs, gc.batch_size, cc.epochs, dict(getattr(gc, key)))
    gc.epochs, gc.batch_size, cc.epochs = 6, 64, 8
    getattr(gc, key)['confidence_threshold'] = 0.0       # a 6-epoch classifier is not confident; keep argmax filtering
    out = str(tmp_path / "data.pkl")
    try:
        gan, clf, rep = pipeline.run(ds, cfg, pickle_path=out, trainer=getattr(pkg, trainer))
    finally:
        gc.epochs, gc.batch_size, cc.epochs = old[0], old[1], old[2]
        getattr(gc, key).update(old[3])
    assert type(gan).__name__ == trainer and clf.name == {"CVAEGAN": "CVAE_GAN"}.get(trainer, trainer) + "_classifier"
    # scaling: everything in [0, 1]
    assert float(ds.tr_samples.min()) >= 0.0 and float(ds.tr_samples.max()) <= 1.0 + 1e-6
    # balancing never exceeds the target and appends matching labels
    mx = max(rep["class_counts_before"].values())
    for lab, st in rep["generation"].items():
        assert 0 <= st["actual"] <= st["target"] == mx - rep["class_counts_before"][lab]
    assert rep["train_rows"] == 1000 + sum(st["actual"] for st in rep["generation"].values())
    assert len(ds.tr_samples) == len(ds.tr_labels) == rep["train_rows"]
    # pickle hand-off: (tr_x, tr_y, te_x, te_y) numpy arrays
    with open(out, "rb") as f:
        trx, try_, tex, tey = pickle.load(f)
    assert trx.shape == (rep["train_rows"], 10) and try_.shape == (rep["train_rows"],) and tex.shape[0] == tey.shape[0]
    assert trx.dtype == np.float32
    # the classifier learned the (well separated) classes
    assert rep["confusion_matrix"].sum() == len(ds.te_labels)
    assert rep["metrics"]["F1"] > 0.9, rep["metrics"]
    assert 0.0 <= rep["binary_metrics"]["F1"] <= 1.0
    gan.engine.close()

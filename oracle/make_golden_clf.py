"""TEST INFRASTRUCTURE ONLY - generates tests/golden/ref_clf.npz by RUNNING THE UNMODIFIED REFERENCE's downstream
classifier (/root/reference/src/classifier.py: Classifier.fit, .test) on seeded synthetic data.

    python oracle/make_golden_clf.py          (build container only: needs /root/reference)

Stored: data, the classifier's start state, the seed, the final state after `fit`, the test metrics and confusion
matrix.  tests/test_oracle_golden.py replays it with oracle/classifier_oracle.py (same seed -> same DataLoader
shuffles and dropout masks from torch's CPU generator)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.refload import load_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
F_, K = 10, 5
EPOCHS, LR, BS, SEED = 3, 1e-3, 64, 2468


def make_data(seed=5):
    from sklearn.datasets import make_blobs
    from sklearn.preprocessing import minmax_scale
    x, y = make_blobs(n_samples=[300, 250, 120, 60, 200], n_features=F_, centers=None, cluster_std=2.5, random_state=seed)
    x = minmax_scale(x).astype(np.float32)
    perm = np.random.RandomState(seed).permutation(len(y))
    x, y = x[perm], y[perm].astype(np.int64)
    n_tr = 630          # not a multiple of the batch size: the last batch is partial
    return x[:n_tr], y[:n_tr], x[n_tr:], y[n_tr:]


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    src = load_reference()
    torch.set_num_threads(1)
    xtr, ytr, xte, yte = make_data()
    src.datasets.feature_num, src.datasets.label_num = F_, K
    src.datasets.tr_samples, src.datasets.tr_labels = torch.from_numpy(xtr), torch.from_numpy(ytr)
    src.datasets.te_samples, src.datasets.te_labels = torch.from_numpy(xte), torch.from_numpy(yte)
    cc = src.config.classifier_config
    cc.epochs, cc.lr, cc.batch_size = EPOCHS, LR, BS
    src.utils.set_random_state()
    clf = src.Classifier("golden") if hasattr(src, "Classifier") else src.classifier.Classifier("golden")
    out = {"xtr": xtr, "ytr": ytr, "xte": xte, "yte": yte, "meta": np.array([F_, K, EPOCHS, BS, SEED], dtype=np.int64),
           "lr": np.array([LR])}
    for k, v in clf.model.state_dict().items():
        out["init/" + k] = v.detach().cpu().numpy().copy()
    torch.manual_seed(SEED)
    clf.fit(src.datasets.TrDataset())
    for k, v in clf.model.state_dict().items():
        out["final/" + k] = v.detach().cpu().numpy().copy()
    clf.test(src.datasets.TeDataset())
    out["metrics"] = np.array([clf.metrics["Precision"], clf.metrics["Recall"], clf.metrics["F1"]], dtype=np.float64)
    out["confusion"] = np.asarray(clf.confusion_matrix, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "ref_clf.npz"), **out)
    print("metrics", clf.metrics, "\n", clf.confusion_matrix)
    print("ref_clf.npz", os.path.getsize(os.path.join(GOLDEN, "ref_clf.npz")))


if __name__ == "__main__":
    main()

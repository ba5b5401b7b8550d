"""TEST INFRASTRUCTURE ONLY - CPU restatement of the downstream classifier fine-tuning of the reference
(/root/reference/src/classifier.py: Classifier.fit :24-45, predict :47-53, test :55-106) in functional torch-CPU code.

Pinned by oracle/make_golden.py (`ref_clf.npz`: the unmodified reference's Classifier.fit/test on seeded data) and
replayed by tests/test_oracle_golden.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product path (cvae_gan_b200/classifier.py) never does.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from oracle.cvae_gan_oracle import OracleAdam, TorchNoise, classifier_forward

KEYS = ["classifier_network.0.weight", "classifier_network.0.bias", "classifier_network.3.weight", "classifier_network.3.bias",
        "classifier_network.4.weight", "classifier_network.4.bias", "classifier_network.7.weight", "classifier_network.7.bias",
        "classifier_network.9.weight", "classifier_network.9.bias"]          # nn.Module.parameters() order


def classifier_step(sd: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor, noise, adam: OracleAdam = None):
    """classifier.py:36-44: zero_grad, prediction = model(samples), cross_entropy, backward, optimizer.step().
    Returns (loss, grads in KEYS order)."""
    params = [sd[k] for k in KEYS]
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    logits = classifier_forward(sd, x, True, noise)
    loss = F.cross_entropy(logits, y)
    grads = torch.autograd.grad(loss, params)
    for p in params:
        p.requires_grad_(False)
    if adam is not None:
        adam.step(list(grads))
    return float(loss.detach()), [g.detach() for g in grads]


def epoch_batches(n: int, batch_size: int) -> List[torch.Tensor]:
    """Index batches of one pass of DataLoader(dataset, batch_size, shuffle=True) drawing from torch's default CPU
    generator like torch.utils.data does: the iterator first draws its base seed, then RandomSampler seeds a private
    generator with one more int64 draw and takes torch.randperm(n, generator=g); the last batch may be partial."""
    torch.empty((), dtype=torch.int64).random_()                                  # _BaseDataLoaderIter base_seed
    g = torch.Generator()
    g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))        # RandomSampler.__iter__
    perm = torch.randperm(n, generator=g)
    return [perm[i:i + batch_size] for i in range(0, n, batch_size)]


def fit(sd: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor, epochs: int, lr: float, batch_size: int, noise=None):
    """Classifier.fit: a fresh Adam(lr, torch default betas), `epochs` shuffled passes; returns the last loss."""
    noise = noise or TorchNoise()
    adam = OracleAdam([sd[k] for k in KEYS], lr, betas=(0.9, 0.999), eps=1e-8)
    last = 0.0
    for _ in range(epochs):
        for idx in epoch_batches(x.size(0), batch_size):
            last, _ = classifier_step(sd, x[idx], y[idx], noise, adam)
    return last


@torch.no_grad()
def predict(sd, x):
    return torch.argmax(classifier_forward(sd, x, False), dim=1)


def macro_metrics(y_true: torch.Tensor, y_pred: torch.Tensor, K: int):
    """sklearn precision/recall/f1 with average='macro', zero_division=0, restated on a confusion matrix."""
    cm = torch.zeros(K, K, dtype=torch.float64)
    for t, p in zip(y_true.tolist(), y_pred.tolist()):
        cm[t, p] += 1
    present = sorted(set(y_true.tolist()) | set(y_pred.tolist()))     # sklearn averages over the labels that occur
    prec, rec, f1 = [], [], []
    for k in present:
        tp = cm[k, k]
        pp, ap = cm[:, k].sum(), cm[k, :].sum()
        p_ = float(tp / pp) if pp > 0 else 0.0
        r_ = float(tp / ap) if ap > 0 else 0.0
        prec.append(p_)
        rec.append(r_)
        f1.append(2 * p_ * r_ / (p_ + r_) if p_ + r_ > 0 else 0.0)
    n = max(len(present), 1)
    return {"Precision": sum(prec) / n, "Recall": sum(rec) / n, "F1": sum(f1) / n}, cm

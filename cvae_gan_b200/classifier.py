"""Drop-in for /root/reference/src/classifier.py (class Classifier :11-106): fine-tunes the CVAE-GAN's classifier
network on the (augmented) training set and reports macro precision / recall / F1 - the acceptance metric of
scripts/train_cvae_gan.py:143-165.

`fit` runs on the CUDA engine (cvg_step_classifier: train-mode forward with dropout, mean cross entropy with per-row
labels, backward, Adam(lr = classifier_config.lr, torch default betas) with a FRESH optimiser state, exactly what
`Adam(params=self.model.parameters(), lr=...)` gives the reference) whenever `self.model` is a classifier attached to
an engine - i.e. after `clf.model = gan.classifier` (train_cvae_gan.py:145).  Batches follow the reference's
`DataLoader(dataset, batch_size, shuffle=True)`: a new permutation per epoch from a generator seeded out of torch's
default CPU generator (the mechanism of `RandomSampler`), last partial batch included.  Dropout masks come from the
engine's Philox stream, whereas the reference's dropout draws interleave with the sampler's on the same generator -
so the trajectory matches the reference statistically, not bit for bit (tests compare F1; the bit-level replay of the
reference's own sequence lives in oracle/classifier_oracle.py).
"""
from __future__ import annotations

import numpy as np
import torch

from . import config, datasets, models


class Classifier:
    def __init__(self, name: str, engine=None):
        self.name = f'{name}_classifier'
        self.model = models.CVAEGANClassifierModel(datasets.feature_num, datasets.label_num)
        self._own_engine = None
        if engine is not None:
            self.model.attach(engine, 3)
        self.confusion_matrix: np.ndarray = None
        self.metrics = {'Precision': 0.0, 'Recall': 0.0, 'F1': 0.0}
        self.class_metrics = None
        self.loss_history = []

    # ---- engine plumbing ---------------------------------------------------------------------------------------
    def _engine(self):
        eng = getattr(self.model, "_engine", None)
        if eng is None:
            # a stand-alone classifier (reference: Classifier('x') without a GAN): give it its own engine
            from .engine import Engine
            eng = Engine(datasets.feature_num, datasets.label_num, config.gan_config.z_size,
                         max_batch=max(64, config.classifier_config.batch_size))
            self.model.attach(eng, 3)
            self._own_engine = eng
        return eng

    @staticmethod
    def _tensors(dataset):
        from .cvae_gan import dataset_tensors
        return dataset_tensors(dataset)

    # ---- classifier.py:24-45 -----------------------------------------------------------------------------------
    def fit(self, dataset):
        eng = self._engine()
        self.model.train()
        cc = config.classifier_config
        x_all, y_all = self._tensors(dataset)
        x_all = x_all.to(eng.device, torch.float32).contiguous()
        y_all = y_all.to(eng.device, torch.int64).contiguous()
        if y_all.numel() and (int(y_all.min()) < 0 or int(y_all.max()) >= eng.K):
            raise ValueError(f"labels must lie in [0, {eng.K}) - got [{int(y_all.min())}, {int(y_all.max())}]")
        n, bs = x_all.size(0), int(cc.batch_size)
        if bs > eng.max_batch:
            raise ValueError(f"classifier batch_size {bs} exceeds the engine's max_batch {eng.max_batch}")
        eng.reset_adam(3)                               # Adam(...) is constructed inside fit(): fresh state
        loss = torch.zeros(4, device=eng.device)
        step = 0
        seed = int(torch.empty((), dtype=torch.int64).random_().item())   # dropout stream of this fit
        for e in range(int(cc.epochs)):
            # RandomSampler.__iter__ (shuffle=True, no generator given): seed a fresh generator from the default RNG
            g = torch.Generator()
            g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            perm = torch.randperm(n, generator=g).to(eng.device)
            for i in range(0, n, bs):
                idx = perm[i:i + bs]
                eng.step_classifier(x_all[idx], y_all[idx], lr=cc.lr, seed=seed, counter=step, loss_out=loss)
                step += 1
            self.loss_history.append(float(loss[0].item()))
        self.model.eval()

    def predict(self, x: torch.Tensor, use_prob: bool = False) -> torch.Tensor:
        eng = self._engine()
        self.model.eval()
        prob = eng.classifier_forward(x.to(eng.device, torch.float32).contiguous())
        if use_prob:
            return prob.squeeze(dim=1).detach()
        return torch.argmax(prob, dim=1)

    # ---- classifier.py:57-106 ----------------------------------------------------------------------------------
    def test(self, dataset):
        from sklearn import metrics
        x_all, y_all = self._tensors(dataset)
        predicted = self.predict(x_all).cpu()
        real = y_all.cpu()
        labels = [i for i in range(datasets.label_num)]
        self.confusion_matrix = metrics.confusion_matrix(y_true=real, y_pred=predicted, labels=labels)
        self.metrics['Precision'] = metrics.precision_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.metrics['Recall'] = metrics.recall_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.metrics['F1'] = metrics.f1_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.class_metrics = metrics.classification_report(y_true=real, y_pred=predicted, labels=labels, output_dict=True,
                                                           zero_division=0)

    def binary_test(self, dataset):
        """classifier.py:108-150: every class > 0 collapses to 1, macro averages."""
        from sklearn import metrics
        x_all, y_all = self._tensors(dataset)
        predicted = (self.predict(x_all).cpu() > 0).long()
        real = (y_all.cpu() > 0).long()
        self.confusion_matrix = metrics.confusion_matrix(y_true=real, y_pred=predicted)
        self.metrics['Precision'] = metrics.precision_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.metrics['Recall'] = metrics.recall_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.metrics['F1'] = metrics.f1_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)

    def print_metrics(self, precision: int = 4):
        for k, v in self.metrics.items():
            print(f'{k}: {v:.{precision}f}')

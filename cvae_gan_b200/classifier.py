"""Drop-in for /root/reference/src/classifier.py (class Classifier :11-106): fine-tunes the CVAE-GAN's classifier
network on the (augmented) training set and reports macro precision / recall / F1 - the acceptance metric of
scripts/train_cvae_gan.py:143-165.

`fit` runs on the CUDA engine (cvg_step_classifier: train-mode forward with dropout, mean cross entropy with per-row
labels, backward, Adam(lr = classifier_config.lr, torch default betas) with a FRESH optimiser state, exactly what
`Adam(params=self.model.parameters(), lr=...)` gives the reference) whenever `self.model` is a classifier attached to
an engine - i.e. after `clf.model = gan.classifier` (train_cvae_gan.py:145).  Batches follow the reference's
`DataLoader(dataset, batch_size, shuffle=True)`: per epoch the iterator's base-seed draw, then a permutation from a generator
seeded out of torch's default CPU generator (the mechanism of `RandomSampler`), last partial batch included - given the same
generator state the batches ARE the loader's (tests/test_host_fit_logic.py).  Dropout masks come from the
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
            # iter(DataLoader(..., shuffle=True)) draws twice from the default CPU generator: the iterator's base seed
            # (_BaseDataLoaderIter), then RandomSampler.__iter__ seeds a private generator for the permutation
            torch.empty((), dtype=torch.int64).random_()
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
        """classifier.py:108-150: every class > 0 collapses to 1, macro averages, per-class report of the two classes."""
        from sklearn import metrics
        x_all, y_all = self._tensors(dataset)
        predicted = (self.predict(x_all).cpu() > 0).long()
        real = (y_all.cpu() > 0).long()
        self.confusion_matrix = metrics.confusion_matrix(y_true=real, y_pred=predicted)
        self.metrics['Precision'] = metrics.precision_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.metrics['Recall'] = metrics.recall_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.metrics['F1'] = metrics.f1_score(y_true=real, y_pred=predicted, average='macro', zero_division=0)
        self.class_metrics = metrics.classification_report(y_true=real, y_pred=predicted, output_dict=True, zero_division=0)

    def print_metrics(self, decimals: int = 4, print_class_metrics: bool = True):
        """classifier.py:152-200: the overall metrics as a rounded dict, then (optionally) the per-class rows of
        `class_metrics`, the macro / weighted averages and the accuracy."""
        print("整体评估指标:")
        print({k: round(v, decimals) for k, v in self.metrics.items()})
        if not print_class_metrics or self.class_metrics is None:
            return

        def rows(entry):
            for title, field in (("Precision", "precision"), ("Recall", "recall"), ("F1-Score", "f1-score")):
                print(f"  {title}: {round(entry[field], decimals)}")
            print(f"  Support: {entry['support']}")

        print("\n每个类别的评估指标:")
        for key, entry in self.class_metrics.items():
            if key in ('accuracy', 'macro avg', 'weighted avg') or not str(key).lstrip('-').isdigit():
                continue
            print(f"\n类别 {int(key)}:")
            rows(entry)
        for title, key in (("宏观平均", 'macro avg'), ("加权平均", 'weighted avg')):
            print(f"\n{title}:")
            if key in self.class_metrics:
                rows(self.class_metrics[key])
        if 'accuracy' in self.class_metrics:
            print(f"\nAccuracy: {round(self.class_metrics['accuracy'], decimals)}")

    # ---- classifier.py:202-303 ---------------------------------------------------------------------------------
    @staticmethod
    def roc_from_scores(scores: np.ndarray, labels: np.ndarray, label_num: int, is_binary: bool = False):
        """The numbers behind `plot_roc_curve`: the reference scores with the network's raw outputs (it calls them `prob`).
        Multi-class (more than 2 outputs and not `is_binary`): one-vs-rest, {class: (fpr, tpr, auc)}.  Otherwise the positive
        class is column 1 and every label > 0 counts as positive: {'binary': (fpr, tpr, auc)}."""
        from sklearn import metrics
        if not is_binary and scores.shape[1] > 2:
            from sklearn.preprocessing import label_binarize
            y_bin = label_binarize(labels, classes=list(range(label_num)))
            return {i: (*metrics.roc_curve(y_bin[:, i], scores[:, i])[:2], metrics.roc_auc_score(y_bin[:, i], scores[:, i]))
                    for i in range(y_bin.shape[1])}
        y_score = scores[:, 1] if scores.shape[1] > 1 else scores.reshape(-1)
        y_test = np.where(labels > 0, 1, 0)
        return {'binary': (*metrics.roc_curve(y_test, y_score)[:2], metrics.roc_auc_score(y_test, y_score))}

    def plot_roc_curve(self, dataset, is_binary: bool = False):
        """classifier.py:202-303 (needs matplotlib, which is not part of the hot path): scores from the CUDA classifier forward,
        curves from `roc_from_scores`, saved as `<name>_roc_curve_{binary|multiclass}.jpg` under `path_config.gan_outs`."""
        import matplotlib.pyplot as plt
        eng = self._engine()
        self.model.eval()
        x_all, y_all = self._tensors(dataset)
        scores = eng.classifier_forward(x_all.to(eng.device, torch.float32).contiguous()).cpu().numpy()
        curves = self.roc_from_scores(scores, y_all.cpu().numpy(), datasets.label_num, is_binary)
        plt.figure(figsize=(10, 8))
        colors = ['aqua', 'darkorange', 'cornflowerblue', 'green', 'red', 'purple']
        for n, (key, (fpr, tpr, auc)) in enumerate(curves.items()):
            if key == 'binary':
                plt.plot(fpr, tpr, color='darkorange', lw=2, label='ROC curve (area = %0.2f)' % auc)
            elif n < len(colors):
                plt.plot(fpr, tpr, color=colors[n], lw=2, label='ROC curve of class {0} (area = {1:0.2f})'.format(key, auc))
        plt.plot([0, 1], [0, 1], color='navy', lw=2, linestyle='--')
        plt.xlim([0.0, 1.0])
        plt.ylim([0.0, 1.05])
        plt.xlabel('False Positive Rate')
        plt.ylabel('True Positive Rate')
        plt.title(f'{self.name} Receiver Operating Characteristic (ROC) Curve')
        plt.legend(loc="lower right")
        plt.grid(True, alpha=0.3)
        out_dir = getattr(getattr(config, "path_config", None), "gan_outs", None)
        if out_dir is None:
            import pathlib
            out_dir = pathlib.Path(".")
        path = out_dir / f"{self.name.replace('_classifier', '')}_roc_curve_{'binary' if is_binary else 'multiclass'}.jpg"
        plt.savefig(path)
        plt.close()
        print(f"ROC曲线已保存至: {path}")
        return curves

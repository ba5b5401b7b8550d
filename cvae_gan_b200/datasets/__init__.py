"""Mirror of the module-level dataset globals of /root/reference/src/datasets/__init__.py:20-47 that
the trainer reads (feature_num, label_num, tr_samples, tr_labels) and of the Dataset protocol
(/root/reference/src/datasets/_dataset.py:4-22).  No CSV loading here: callers assign the tensors."""
import torch

tr_samples = torch.zeros(0, 0)
tr_labels = torch.zeros(0, dtype=torch.long)
te_samples = torch.zeros(0, 0)
te_labels = torch.zeros(0, dtype=torch.long)
feature_num = 0
label_num = 0


class Dataset:
    def __init__(self, training: bool = True):
        self.training = training

    def _xy(self):
        import sys
        m = sys.modules[__name__]
        return (m.tr_samples, m.tr_labels) if self.training else (m.te_samples, m.te_labels)

    def __len__(self):
        return len(self._xy()[1])

    def __getitem__(self, idx: int):
        x, y = self._xy()
        return x[idx], y[idx]

    def tensors(self):
        """(samples, labels) as whole tensors - lets `_divide_samples` partition without a Python loop."""
        return self._xy()


class TrDataset(Dataset):
    def __init__(self):
        super().__init__(training=True)


class TeDataset(Dataset):
    def __init__(self):
        super().__init__(training=False)

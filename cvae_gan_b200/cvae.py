"""`CVAE` - drop-in for the sibling trainer /root/reference/src/cvae.py:11-319: the CVAE-GAN without its critic.
Same encoder, generator and classifier as the CVAE-GAN (src/models/cvae_models.py defines them layer for layer like
cvae_gan_models.py, same state_dict keys), the same classifier step (cvae.py:89-115 is cvae_gan.py:131-157 statement for
statement), no critic step, and an encoder/generator step with ONE generator pass: x_recon = G(E(x)),
total = lambda_recon * MSE(x_recon, x) + lambda_kl * KL + lambda_class(e) * CE(C(x_recon), label) - the classification term is
taken on the RECONSTRUCTION, not on a prior sample -, Adam on encoder and generator (cvae.py:117-166; engine flag
CVG_STEP_CVAE, train.cu step_g_cvae).  A label visit is c_loop classifier steps + g_loop encoder/generator steps
(cvae.py:86-166).  Everything else - partition, device-side batch draws, CUDA-graph label visits, generation from the prior
and the confidence filter (cvae.py:263-298 is cvae_gan.py:339-378) - is the CVAEGAN host class's.

Surface kept from the reference class: `feature_num, label_num, encoder, generator, classifier, samples, lambda_recon,
lambda_kl, lambda_class, loss_history{recon_loss, kl_loss, class_loss}`; `fit, _divide_samples, _get_target_samples,
plot_loss_history, generate_samples, generate_qualified_samples, reconstruct_samples`.  Config:
`config.gan_config.cvae_config` read at call time.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import models
from ._lib import STEP_CVAE
from .cvae_gan import CVAEGAN

# src/models/cvae_models.py: same layers, sizes, checks and state_dict keys as the CVAE-GAN's three networks
CVAEEncoderModel = models.CVAEGANEncoderModel
CVAEGeneratorModel = models.CVAEGANGeneratorModel
CVAEClassifierModel = models.CVAEGANClassifierModel


class CVAE(CVAEGAN):
    _CONFIG_KEY = 'cvae_config'
    _HISTORY = (('recon_loss', 0), ('kl_loss', 1), ('class_loss', 3))
    _VISIT_FLAGS = STEP_CVAE
    _G_FORWARDS_PER_G_STEP = 1
    _USES_CRITIC = False
    _NAME = "CVAE"
    _BUILD_ORDER = ("encoder", "generator", "classifier")            # cvae.py:19-34

    def _loops(self, gc):
        """cvae.py:86-117: classifier steps, then encoder/generator steps; there is no critic."""
        return (0, int(gc.c_loop_num), int(gc.g_loop_num))

    def plot_loss_history(self):
        """cvae.py:203-261 (needs matplotlib, which is not part of the hot path)."""
        import matplotlib.pyplot as plt
        out_dir = getattr(getattr(self.config, "path_config", None), "gan_outs", None)
        if out_dir is None:
            import pathlib
            out_dir = pathlib.Path(".")
        titles = (('recon_loss', 'Reconstruction Loss', 'blue'), ('kl_loss', 'KL divergence loss', 'green'),
                  ('class_loss', 'Classification Loss', 'purple'))
        plt.figure(figsize=(12, 8))
        for i, (key, title, color) in enumerate(titles):
            plt.subplot(2, 2, i + 1)
            plt.plot(self.loss_history[key], color=color)
            plt.xlabel('Epoch')
            plt.ylabel('Loss')
            plt.title(title)
        plt.tight_layout()
        plt.savefig(out_dir / 'cvae_loss_history.jpg')
        plt.close()
        plt.figure(figsize=(12, 6))
        for key, label, color in (('recon_loss', '重构损失', 'blue'), ('kl_loss', 'KL散度', 'green'), ('class_loss', '分类损失', 'purple')):
            plt.plot(self.loss_history[key], label=label, color=color)
        plt.xlabel('Epoch')
        plt.ylabel('Loss')
        plt.title('CVAE损失曲线')
        plt.legend()
        plt.grid(True, alpha=0.3)
        plt.savefig(out_dir / 'cvae_combined_loss.jpg')
        plt.close()

    def reconstruct_samples(self, samples: torch.Tensor, labels: torch.Tensor):
        """cvae.py:300-319.  Unlike the CVAE-GAN's (which always raises), the reference's CVAE hands the generator a 2-D one-hot
        condition, so this one works: z_enc = E.encode(samples, labels) and x = G(z_enc, onehot(labels)) with both networks in
        eval mode (running statistics: rows are independent, so rows are grouped by label for the one-label-per-batch kernels),
        and - like the reference - encoder and generator are left in TRAIN mode afterwards."""
        eng = self.engine
        samples = samples.to(eng.device, torch.float32)
        labels = labels.to(eng.device).long()
        if labels.dim() == 2 and labels.size(1) == 1:
            labels = labels.squeeze(1)
        if labels.dim() != 1:
            raise ValueError(f"条件输入格式错误，期望1D或2D(单列)，实际: {labels.shape}")
        out = torch.empty(samples.size(0), self.feature_num, dtype=torch.float32, device=eng.device)
        for lab in torch.unique(labels).tolist():
            sel = (labels == lab).nonzero().flatten()
            mu, lv = eng.encoder_forward(samples[sel].contiguous(), int(lab))
            z = (mu + torch.randn_like(mu) * torch.exp(0.5 * lv)).contiguous()
            out[sel] = eng.generate(int(lab), z.size(0), z=z, train_mode=False)
        self.encoder.train()
        self.generator.train()
        return out.cpu()

    def reconstruct(self, samples: torch.Tensor, label: int):
        labels = torch.full((samples.size(0),), int(label), dtype=torch.long)
        return self.reconstruct_samples(samples, labels)

    def state_dict(self):
        self._sync_bn_counters()
        return OrderedDict((n, getattr(self, n).state_dict()) for n in ("encoder", "generator", "classifier"))

    def load_state_dict(self, sd):
        for net, n in ((0, "encoder"), (1, "generator"), (3, "classifier")):
            self.engine.load_state(net, sd[n])
            for k, v in sd[n].items():
                if k.endswith("num_batches_tracked"):
                    dict(getattr(self, n).named_buffers())[k].fill_(int(v))

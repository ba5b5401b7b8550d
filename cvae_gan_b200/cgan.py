"""`CGAN` - drop-in for the sibling trainer /root/reference/src/cgan.py:10-309: the CVAE-GAN without its VAE
branch.  Same three networks as the CVAE-GAN (src/models/cgan_models.py defines generator / critic / classifier layer for
layer like cvae_gan_models.py, same state_dict keys), the same critic and classifier steps (cgan.py:84-136 is
cvae_gan.py:104-157 statement for statement) and a generator step that only runs G(z_prior): total = lambda_adv * (-mean
D(x_fake)) + lambda_class(e) * CE(C(x_fake), label), Adam on the generator alone (cgan.py:138-178; engine flag
CVG_STEP_PRIOR_ONLY, train.cu step_g_prior).  Everything else - partition, device-side batch draws, CUDA-graph label
visits, generation and the confidence filter - is the CVAEGAN host class's.

Surface kept from the reference class: `feature_num, label_num, generator, discriminator, classifier, samples, lambda_adv,
lambda_class, loss_history{adv_loss, class_loss}`; `fit, _divide_samples, _get_target_samples, plot_loss_history,
generate_samples, generate_qualified_samples`.  Config: `config.gan_config.cgan_config` read at call time.
"""
from __future__ import annotations

from collections import OrderedDict

from . import models
from ._lib import STEP_PRIOR_ONLY
from .cvae_gan import CVAEGAN

# src/models/cgan_models.py: same layers, sizes, checks and state_dict keys as the CVAE-GAN's three networks
CGANGeneratorModel = models.CVAEGANGeneratorModel
CGANDiscriminatorModel = models.CVAEGANDiscriminatorModel
CGANClassifierModel = models.CVAEGANClassifierModel


class CGAN(CVAEGAN):
    _CONFIG_KEY = 'cgan_config'
    _HISTORY = (('adv_loss', 2), ('class_loss', 3))
    _VISIT_FLAGS = STEP_PRIOR_ONLY
    _G_FORWARDS_PER_G_STEP = 1
    _USES_ENCODER = False
    _NAME = "CGAN"
    _BUILD_ORDER = ("generator", "discriminator", "classifier")      # cgan.py:20-34

    def plot_loss_history(self):
        """cgan.py:216-262 (needs matplotlib, which is not part of the hot path)."""
        import matplotlib.pyplot as plt
        out_dir = getattr(getattr(self.config, "path_config", None), "gan_outs", None)
        if out_dir is None:
            import pathlib
            out_dir = pathlib.Path(".")
        plt.figure(figsize=(12, 6))
        for i, (key, title, color) in enumerate((('adv_loss', 'Adversarial Loss', 'red'), ('class_loss', 'Classification Loss', 'purple'))):
            plt.subplot(1, 2, i + 1)
            plt.plot(self.loss_history[key], color=color)
            plt.xlabel('Epoch')
            plt.ylabel('Loss')
            plt.title(title)
        plt.tight_layout()
        plt.savefig(out_dir / 'cgan_loss_history.jpg')
        plt.close()
        plt.figure(figsize=(12, 6))
        plt.plot([abs(v) for v in self.loss_history['adv_loss']], label='对抗损失(绝对值)', color='red')
        plt.plot(self.loss_history['class_loss'], label='分类损失', color='purple')
        plt.xlabel('Epoch')
        plt.ylabel('Loss')
        plt.title('CGAN损失曲线')
        plt.legend()
        plt.grid(True, alpha=0.3)
        plt.savefig(out_dir / 'cgan_combined_loss.jpg')
        plt.close()

    def reconstruct_samples(self, samples, labels):
        raise AttributeError("CGAN has no encoder (src/cgan.py defines no reconstruct_samples)")

    reconstruct = reconstruct_samples

    def state_dict(self):
        self._sync_bn_counters()
        return OrderedDict((n, getattr(self, n).state_dict()) for n in ("generator", "discriminator", "classifier"))

    def load_state_dict(self, sd):
        for net, n in ((1, "generator"), (2, "discriminator"), (3, "classifier")):
            self.engine.load_state(net, sd[n])
            for k, v in sd[n].items():
                if k.endswith("num_batches_tracked"):
                    dict(getattr(self, n).named_buffers())[k].fill_(int(v))

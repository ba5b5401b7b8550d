"""`VAEGAN` - drop-in for the sibling trainer /root/reference/src/vae_gan.py:10-261: the CVAE-GAN without labels.
Encoder, generator and critic are the CVAE-GAN's stacks WITHOUT the one-hot label columns (src/models/vae_gan_models.py;
engine: `CvgConfig.unconditional`), there is no classifier and no per-label partition: every epoch runs d_loop critic steps and
g_loop encoder/generator steps on batches drawn from ALL training rows (vae_gan.py:74-141), with
total = lambda_recon * MSE(G(z_enc), x) + lambda_kl * KL + lambda_adv * (-mean D(G(z_prior))).  On the engine that is one
`cvg_visit(label = 0, d_loop, c_loop = 0, g_loop, CVG_VISIT_LAMBDA_ZERO)` per epoch over the whole sample table - the same step
kernels as the CVAE-GAN with the label column switched off and the classification branch never run.

Surface kept from the reference class: `feature_num, encoder, generator, discriminator, samples, lambda_recon, lambda_kl,
lambda_adv, loss_history{recon_loss, kl_loss, adv_loss}`; `fit, _store_samples, _get_random_samples, plot_loss_history,
generate_samples(num), reconstruct_samples(samples)`.  Config: `config.gan_config.vae_gan_config`.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import config as _config
from . import datasets as _datasets
from . import models
from ._lib import NET_CLASSIFIER, NET_DISCRIMINATOR, NET_ENCODER, NET_GENERATOR, VISIT_LAMBDA_ZERO
from .cvae_gan import _dist_info, dataset_tensors
from .engine import Engine


class VAEGAN:
    def __init__(self, config=None, datasets=None, max_rows: int = None):
        self.config = config or _config
        self.datasets = datasets or _datasets
        gc = self.config.gan_config
        self.feature_num = self.datasets.feature_num
        self.rank, self.world_size = _dist_info()
        hidden = getattr(gc, "hidden", None)
        # same construction (and CPU-generator draw) order as vae_gan.py:14-26
        self.encoder = models.VAEGANEncoderModel(self.feature_num, gc.z_size, hidden=hidden)
        self.generator = models.VAEGANGeneratorModel(gc.z_size, self.feature_num, hidden=hidden)
        self.discriminator = models.VAEGANDiscriminatorModel(self.feature_num, hidden=hidden)
        label_num = max(1, int(getattr(self.datasets, "label_num", 1) or 1))
        with torch.random.fork_rng(devices=[]):      # the engine owns four networks; a VAE-GAN never runs this one
            self._classifier_module = models.CVAEGANClassifierModel(self.feature_num, label_num, hidden=hidden)
        self.samples = None
        cc = gc.vae_gan_config
        self.lambda_recon, self.lambda_kl, self.lambda_adv = cc['lambda_recon'], cc['lambda_kl'], cc['lambda_adv']
        self.loss_history = {'recon_loss': [], 'kl_loss': [], 'adv_loss': []}
        rows = max_rows or max(int(gc.batch_size) // self.world_size, 1 << 14)
        self.engine = Engine(self.feature_num, label_num, gc.z_size, rows, lambda_recon=self.lambda_recon,
                             lambda_kl=self.lambda_kl, lambda_adv=self.lambda_adv, g_lr=gc.g_lr, d_lr=gc.d_lr, c_lr=gc.c_lr,
                             world_size=self.world_size, rank=self.rank, hidden=hidden, unconditional=True)
        for net, mod in ((NET_ENCODER, self.encoder), (NET_GENERATOR, self.generator),
                         (NET_DISCRIMINATOR, self.discriminator), (NET_CLASSIFIER, self._classifier_module)):
            mod.attach(self.engine, net)
        self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._counter = 0
        self._gen_rows = 0
        self._bn_calls = {NET_ENCODER: 0, NET_GENERATOR: 0}
        self.use_cuda_graphs = True

    def _networks(self):
        return (self.encoder, self.generator, self.discriminator)

    def _sync_bn_counters(self):
        for net, mod in ((NET_ENCODER, self.encoder), (NET_GENERATOR, self.generator)):
            k = self._bn_calls[net]
            if k:
                for m in mod.modules():
                    if isinstance(m, torch.nn.BatchNorm1d):
                        m.num_batches_tracked += k
                self._bn_calls[net] = 0

    def fit(self, dataset):
        """vae_gan.py:42-157."""
        gc = self.config.gan_config
        eng = self.engine
        cc = gc.vae_gan_config       # the reference reads these at every step (vae_gan.py:131-135): refuse stale values
        want = (float(gc.g_lr), float(gc.d_lr), float(cc['lambda_recon']), float(cc['lambda_kl']), float(cc['lambda_adv']))
        have = (eng.cfg.g_lr, eng.cfg.d_lr, eng.cfg.lambda_recon, eng.cfg.lambda_kl, eng.cfg.lambda_adv)
        if any(abs(a - b) > 1e-12 + 1e-6 * abs(b) for a, b in zip(want, have)):
            raise ValueError(f"g_lr / d_lr / vae_gan_config changed after VAEGAN() was constructed: configured {want}, engine holds {have}")
        if self.world_size > 1:
            eng.verify_replicas()
        for m in self._networks():
            m.train()
        self._store_samples(dataset)
        for net in range(4):
            eng.adam_m[net].zero_()
            eng.adam_v[net].zero_()
            eng.grads[net].zero_()
            eng.set_adam_step(net, 0)
        loops = (int(gc.d_loop_num), 0, int(gc.g_loop_num))
        n_steps = sum(loops)
        batch = int(gc.batch_size)
        losses = torch.zeros(n_steps, 4, dtype=torch.float32, device=eng.device)
        eng.ctl_set(seed=self._seed, counter=self._counter, lambda_class=0.0)
        graph = None
        for e in range(gc.epochs):
            if self.use_cuda_graphs:
                if graph is None:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        eng.visit(0, batch, class_rows=self.samples, loops=loops, flags=VISIT_LAMBDA_ZERO, loss_out=losses)
                graph.replay()
            else:
                eng.visit(0, batch, class_rows=self.samples, loops=loops, flags=VISIT_LAMBDA_ZERO, loss_out=losses)
            self._counter += 2 * n_steps
            self._bn_calls[NET_GENERATOR] += loops[0] + 2 * loops[2]
            self._bn_calls[NET_ENCODER] += loops[2]
            row = losses[n_steps - 1].tolist()
            for key, col in (('recon_loss', 0), ('kl_loss', 1), ('adv_loss', 2)):
                self.loss_history[key].append(row[col])
            if e % 50 == 0:
                print(f"VAE-GAN训练轮次: {e}/{gc.epochs}, 重构损失: {row[0]:.4f}, KL损失: {row[1]:.4f}, 对抗损失: {row[2]:.4f}")
        torch.cuda.synchronize(eng.device)
        graph = None
        self._sync_bn_counters()
        for m in self._networks():
            m.eval()

    def _store_samples(self, dataset):
        """vae_gan.py:159-164: all training rows, labels ignored."""
        x, _ = dataset_tensors(dataset)
        self.samples = x.to(self.engine.device, torch.float32).contiguous()

    def _get_random_samples(self, num: int) -> torch.Tensor:
        """vae_gan.py:166-178 on the device (the three branches of `_get_target_samples`, over all rows)."""
        self._counter += 1
        return self.engine.sample_rows(self.samples, int(num), seed=self._seed, counter=self._counter)

    def plot_loss_history(self):
        """vae_gan.py:180-236 (needs matplotlib, which is not part of the hot path)."""
        import matplotlib.pyplot as plt
        out_dir = getattr(getattr(self.config, "path_config", None), "gan_outs", None)
        if out_dir is None:
            import pathlib
            out_dir = pathlib.Path(".")
        titles = (('recon_loss', 'Reconstruction Loss', 'blue'), ('kl_loss', 'KL divergence loss', 'green'),
                  ('adv_loss', 'Adversarial Loss', 'red'))
        plt.figure(figsize=(12, 8))
        for i, (key, title, color) in enumerate(titles):
            plt.subplot(2, 2, i + 1)
            plt.plot(self.loss_history[key], color=color)
            plt.xlabel('Epoch')
            plt.ylabel('Loss')
            plt.title(title)
        plt.tight_layout()
        plt.savefig(out_dir / 'vae_gan_loss_history.jpg')
        plt.close()
        plt.figure(figsize=(12, 6))
        for key, title, color in titles:
            vals = self.loss_history[key]
            plt.plot([abs(v) for v in vals] if key == 'adv_loss' else vals, label=title, color=color)
        plt.xlabel('Epoch')
        plt.ylabel('Loss')
        plt.legend()
        plt.grid(True, alpha=0.3)
        plt.savefig(out_dir / 'vae_gan_combined_loss.jpg')
        plt.close()

    def generate_samples(self, num: int):
        """vae_gan.py:238-241: G(randn[num, Z]) in whatever mode G is in, returned on the CPU."""
        out = self.engine.generate(0, int(num), seed=self._seed, row_offset=self._gen_rows, train_mode=self.generator.training)
        self._gen_rows += int(num)
        if self.generator.training:
            self._bn_calls[NET_GENERATOR] += 1
            self._sync_bn_counters()
        return out.cpu()

    def reconstruct_samples(self, samples: torch.Tensor):
        """vae_gan.py:243-261: eval-mode E and G, z_enc = mu + eps * std, both networks left in TRAIN mode afterwards."""
        eng = self.engine
        mu, lv = eng.encoder_forward(samples.to(eng.device, torch.float32).contiguous(), 0)
        z = (mu + torch.randn_like(mu) * torch.exp(0.5 * lv)).contiguous()
        out = eng.generate(0, z.size(0), z=z, train_mode=False)
        self.encoder.train()
        self.generator.train()
        return out.cpu()

    def state_dict(self):
        self._sync_bn_counters()
        return OrderedDict((n, getattr(self, n).state_dict()) for n in ("encoder", "generator", "discriminator"))

    def load_state_dict(self, sd):
        for net, n in ((0, "encoder"), (1, "generator"), (2, "discriminator")):
            self.engine.load_state(net, sd[n])
            for k, v in sd[n].items():
                if k.endswith("num_batches_tracked"):
                    dict(getattr(self, n).named_buffers())[k].fill_(int(v))

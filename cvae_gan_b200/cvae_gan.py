"""`CVAEGAN` - drop-in for /root/reference/src/cvae_gan.py:9-397 whose training and generation run
in hand-written sm_100a CUDA kernels behind the C ABI (include/cvaegan_b200.h).

Same surface as the reference class: construct with no arguments after the dataset/config globals
are set, `fit(dataset)`, `generate_samples`, `generate_qualified_samples`, `samples`, `loss_history`,
`encoder / generator / discriminator / classifier`, `lambda_*`, `plot_loss_history`,
`reconstruct_samples`.  Config names are read from `config.gan_config` at CALL time like the
reference (cvae_gan.py:100,104,108).

Randomness: the reference draws from torch's generators; this implementation draws everything inside
the kernels from Philox4x32-10 keyed by `torch.initial_seed()` (set by `set_random_state`) and a
step counter, with rows keyed by their GLOBAL index, so a run is reproducible and independent of the
number of GPUs.  Parity against the reference is checked with injected noise (tests/).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import config as _config
from . import datasets as _datasets
from . import models
from ._lib import NET_CLASSIFIER, NET_DISCRIMINATOR, NET_ENCODER, NET_GENERATOR, VISIT_LAMBDA_ZERO
from .engine import Engine, patience_scan


def lambda_class_at(e: int, lambda_class: float) -> float:
    """cvae_gan.py:198-204: 0 for e < 200, linear ramp over [200, 500), then the full weight."""
    if e < 200:
        return 0.0
    if e < 500:
        return lambda_class * ((e - 200) / 300)
    return lambda_class


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def dataset_tensors(dataset):
    """(samples [N, F], labels [N]) of the dataset that was PASSED IN (the reference iterates `for sample, label in
    dataset`, cvae_gan.py:239).  Fast paths: this package's own Dataset (`tensors()` method) and torch's TensorDataset
    (`tensors` tuple); anything else is iterated."""
    t = getattr(dataset, "tensors", None)
    if callable(t):
        return t()
    if isinstance(t, (tuple, list)) and len(t) == 2:
        return t[0], t[1]
    xs, ys = zip(*[dataset[i] for i in range(len(dataset))])
    return torch.stack([torch.as_tensor(v) for v in xs]), torch.stack([torch.as_tensor(v) for v in ys])


class CVAEGAN:
    # what the sibling trainers (cgan.py) change: the config dict, the loss_history rows taken from a step's loss_out
    # {recon, kl, adv, class}, the visit flags, and how many generator forwards a generator step runs
    _CONFIG_KEY = 'cvae_gan_config'
    _HISTORY = (('recon_loss', 0), ('kl_loss', 1), ('adv_loss', 2), ('class_loss', 3))
    _VISIT_FLAGS = 0
    _G_FORWARDS_PER_G_STEP = 2
    _USES_ENCODER = True
    _USES_CRITIC = True
    _NAME = "CVAE-GAN"
    # the networks the reference class constructs, in ITS order (= order of the CPU-generator draws of their initialisation:
    # cvae_gan.py:19-39); the engine always owns all four, the rest are built afterwards without touching the RNG stream
    _BUILD_ORDER = ("encoder", "generator", "discriminator", "classifier")

    @classmethod
    def _build_networks(cls, feature_num: int, label_num: int, z_size: int, hidden=None):
        """{name: module} for all four engine networks; same-seed construction gives the reference class's starting
        parameters for the networks it has (tests/test_abi_and_host.py).  `hidden`: see `gan_config.hidden`."""
        ctor = {"encoder": lambda: models.CVAEGANEncoderModel(feature_num, label_num, z_size, hidden=hidden),
                "generator": lambda: models.CVAEGANGeneratorModel(z_size, label_num, feature_num, hidden=hidden),
                "discriminator": lambda: models.CVAEGANDiscriminatorModel(feature_num, label_num, hidden=hidden),
                "classifier": lambda: models.CVAEGANClassifierModel(feature_num, label_num, hidden=hidden)}
        nets = {n: ctor[n]() for n in cls._BUILD_ORDER}
        rest = [n for n in ctor if n not in nets]
        if rest:
            with torch.random.fork_rng(devices=[]):
                for n in rest:
                    nets[n] = ctor[n]()
        return nets

    def _loops(self, gc):
        """(critic, classifier, encoder/generator) steps per label visit (cvae_gan.py:104,131,160)."""
        return (int(gc.d_loop_num), int(gc.c_loop_num), int(gc.g_loop_num))

    def __init__(self, config=None, datasets=None, max_rows: int = None):
        """`config` / `datasets`: modules with the reference's names (defaults: this package's mirrors;
        pass the reference's own `src.config`, `src.datasets` to run inside its scripts)."""
        self.config = config or _config
        self.datasets = datasets or _datasets
        gc = self.config.gan_config
        self.feature_num = self.datasets.feature_num
        self.label_num = self.datasets.label_num
        self.rank, self.world_size = _dist_info()

        # same construction (and CPU-generator draw) order as the reference class (cvae_gan.py:19-39)
        hidden = getattr(gc, "hidden", None)        # None: the reference's widths; (h1, h2, h3): the widened model (configs[4])
        nets = self._build_networks(self.feature_num, self.label_num, gc.z_size, hidden)
        encoder, critic = nets["encoder"], nets["discriminator"]
        if self._USES_ENCODER:
            self.encoder = encoder
        self._encoder_module = encoder      # the engine owns four networks; CGAN never runs (or exposes) this one,
        if self._USES_CRITIC:
            self.discriminator = critic
        self._critic_module = critic        # and CVAE never runs (or exposes) the critic
        self.generator = nets["generator"]
        self.classifier = nets["classifier"]

        self.samples = dict()
        cc = getattr(gc, self._CONFIG_KEY)
        if 'lambda_recon' in cc:
            self.lambda_recon = cc['lambda_recon']
            self.lambda_kl = cc['lambda_kl']
        if 'lambda_adv' in cc:
            self.lambda_adv = cc['lambda_adv']
        self.lambda_class = cc['lambda_class']
        self.loss_history = {k: [] for k, _ in self._HISTORY}

        rows = max_rows or max(int(gc.batch_size) // self.world_size, 1 << 14)
        self.engine = Engine(self.feature_num, self.label_num, gc.z_size, rows,
                             lambda_recon=cc.get('lambda_recon', 0.0), lambda_kl=cc.get('lambda_kl', 0.0),
                             lambda_adv=cc.get('lambda_adv', 0.0),
                             g_lr=gc.g_lr, d_lr=gc.d_lr, c_lr=gc.c_lr, world_size=self.world_size, rank=self.rank, hidden=hidden)
        for net, mod in ((NET_ENCODER, encoder), (NET_GENERATOR, self.generator),
                         (NET_DISCRIMINATOR, critic), (NET_CLASSIFIER, self.classifier)):
            mod.attach(self.engine, net)
        self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._counter = 0          # one Philox counter value per optimiser step / sampling call
        self._gen_rows = 0         # rows of the generation noise stream consumed so far
        self._bn_calls = {NET_ENCODER: 0, NET_GENERATOR: 0}
        self.use_cuda_graphs = True   # False: same kernels launched eagerly (debugging aid; identical numerics)

    # ------------------------------------------------------------------------------------------------
    def _networks(self):
        nets = (self.generator,) + ((self.discriminator,) if self._USES_CRITIC else ()) + (self.classifier,)
        return ((self.encoder,) + nets) if self._USES_ENCODER else nets

    def _next(self) -> int:
        self._counter += 1
        return self._counter

    def _sync_bn_counters(self):
        for net, mod in ((NET_ENCODER, self._encoder_module), (NET_GENERATOR, self.generator)):
            k = self._bn_calls[net]
            if k:
                for m in mod.modules():
                    if isinstance(m, torch.nn.BatchNorm1d):
                        m.num_batches_tracked += k
                self._bn_calls[net] = 0

    def fit(self, dataset):
        """cvae_gan.py:59-236."""
        gc = self.config.gan_config
        eng = self.engine
        # the learning rates and loss weights were handed to the engine at construction (CvgConfig); the reference re-reads them
        # here (cvae_gan.py:75-97, 207-212) - refuse to train silently with stale values
        want = (float(gc.g_lr), float(gc.d_lr), float(gc.c_lr), float(getattr(self, "lambda_recon", 0.0)),
                float(getattr(self, "lambda_kl", 0.0)), float(getattr(self, "lambda_adv", 0.0)))
        have = (eng.cfg.g_lr, eng.cfg.d_lr, eng.cfg.c_lr, eng.cfg.lambda_recon, eng.cfg.lambda_kl, eng.cfg.lambda_adv)
        if any(abs(a - b) > 1e-12 + 1e-6 * abs(b) for a, b in zip(want, have)):
            raise ValueError("g_lr / d_lr / c_lr / lambda_recon / lambda_kl / lambda_adv changed after CVAEGAN() was constructed: "
                             f"configured {want}, engine holds {have}; construct a new CVAEGAN after changing them")
        if self.world_size > 1:
            eng.verify_replicas()
        for m in self._networks():
            m.train()
        self._divide_samples(dataset)
        # fresh optimisers every fit(), like cvae_gan.py:75-97
        for net in range(4):
            eng.adam_m[net].zero_()
            eng.adam_v[net].zero_()
            eng.grads[net].zero_()
            eng.set_adam_step(net, 0)
        lam = getattr(gc, self._CONFIG_KEY)['lambda_class']
        loops = self._loops(gc)
        n_steps = sum(loops)
        batch = int(gc.batch_size)
        losses = torch.zeros(n_steps, 4, dtype=torch.float32, device=eng.device)
        # Philox key + starting counter go to the device control block; every step advances it there
        eng.ctl_set(seed=self._seed, counter=self._counter)
        graphs = {}
        for e in range(gc.epochs):
            lam_e = lambda_class_at(e, lam)
            eng.ctl_set(lambda_class=lam_e)
            flags = (VISIT_LAMBDA_ZERO if lam_e == 0.0 else 0) | self._VISIT_FLAGS
            for target_label in self.samples.keys():
                key = (target_label, flags, batch, loops)
                if self.use_cuda_graphs:
                    g = graphs.get(key)
                    if g is None:
                        # a label visit is ~400 kernel launches of a few microseconds each: captured once per
                        # (label, lambda_class != 0) and replayed, the host cost per visit is one graph launch
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            eng.visit(target_label, batch, class_rows=self.samples[target_label], loops=loops,
                                      flags=flags, loss_out=losses)
                        graphs[key] = g
                    g.replay()
                else:
                    eng.visit(target_label, batch, class_rows=self.samples[target_label], loops=loops, flags=flags,
                              loss_out=losses)
                self._counter += 2 * n_steps
                self._bn_calls[NET_GENERATOR] += loops[0] + loops[1] + self._G_FORWARDS_PER_G_STEP * loops[2]
                if self._USES_ENCODER:
                    self._bn_calls[NET_ENCODER] += loops[2]
            # one read-back per epoch: the last label's last generator step (cvae_gan.py:219-222)
            row = losses[n_steps - 1].tolist()
            for key, col in self._HISTORY:
                self.loss_history[key].append(row[col])
            if e % 50 == 0:
                names = {'recon_loss': '重构损失', 'kl_loss': 'KL损失', 'adv_loss': '对抗损失', 'class_loss': '分类损失'}
                print(f"{self._NAME}训练轮次: {e}/{gc.epochs}, " + ", ".join(f"{names[k]}: {row[c]:.4f}" for k, c in self._HISTORY))
        torch.cuda.synchronize(eng.device)
        graphs.clear()
        self._sync_bn_counters()
        for m in self._networks():
            m.eval()

    def _divide_samples(self, dataset) -> None:
        """cvae_gan.py:238-245, without the O(N^2) per-row torch.cat: one stable partition on the device.
        Key order = first occurrence in the data, rows keep their order, exactly like the reference."""
        x, y = dataset_tensors(dataset)
        x = x.to(self.engine.device, torch.float32)
        y = y.to(self.engine.device).long()
        labels, first = [], {}
        uniq = torch.unique(y)
        for lab in uniq.tolist():
            first[lab] = int((y == lab).nonzero()[0])
        labels = sorted(first, key=first.get)
        for lab in labels:
            rows = x[y == lab].contiguous()
            if lab not in self.samples:
                self.samples[lab] = rows
            else:
                self.samples[lab] = torch.cat([self.samples[lab], rows])

    def _get_target_samples(self, label: int, num: int) -> torch.Tensor:
        """cvae_gan.py:247-260 on the device (three branches by class size); returns this rank's shard."""
        return self.engine.sample_rows(self.samples[label], int(num), seed=self._seed, counter=self._next())

    # ------------------------------------------------------------------------------------------------
    def plot_loss_history(self):
        """cvae_gan.py:263-337 (needs matplotlib, which is not part of the hot path)."""
        import matplotlib.pyplot as plt
        out_dir = getattr(getattr(self.config, "path_config", None), "gan_outs", None)
        if out_dir is None:
            import pathlib
            out_dir = pathlib.Path(".")
        titles = (('recon_loss', 'Reconstruction Loss', 'blue'), ('kl_loss', 'KL divergence loss', 'green'),
                  ('adv_loss', 'Adversarial Loss', 'red'), ('class_loss', 'Classification Loss', 'purple'))
        plt.figure(figsize=(12, 8))
        for i, (key, title, color) in enumerate(titles):
            plt.subplot(2, 2, i + 1)
            plt.plot(self.loss_history[key], color=color)
            plt.xlabel('Epoch')
            plt.ylabel('Loss')
            plt.title(title)
        plt.tight_layout()
        plt.savefig(out_dir / 'cvae_gan_loss_history.jpg')
        plt.close()
        plt.figure(figsize=(12, 6))
        for key, title, color in titles:
            vals = self.loss_history[key]
            plt.plot([abs(v) for v in vals] if key == 'adv_loss' else vals, label=title, color=color)
        plt.xlabel('Epoch')
        plt.ylabel('Loss')
        plt.legend()
        plt.grid(True, alpha=0.3)
        plt.savefig(out_dir / 'cvae_gan_combined_loss.jpg')
        plt.close()

    # ------------------------------------------------------------------------------------------------
    def generate_samples(self, target_label: int, num: int):
        """cvae_gan.py:339-345: G(randn[num, Z], onehot) in whatever mode G is in, returned on the CPU."""
        out = self.engine.generate(int(target_label), int(num), seed=self._seed, row_offset=self._gen_rows,
                                   train_mode=self.generator.training)
        self._gen_rows += int(num)
        if self.generator.training:
            self._bn_calls[NET_GENERATOR] += 1
            self._sync_bn_counters()
        return out.cpu()

    def generate_qualified_samples(self, target_label: int, num: int, confidence_threshold: float = None):
        """cvae_gan.py:347-378.  The reference draws chunks of <= 10 rows and stops after 20 empty chunks;
        rows are independent in eval mode, so here the noise stream is generated, classified and filtered
        in large fused batches and the chunk/patience bookkeeping is replayed on the keep mask
        (cvg_patience_scan): the result is the same prefix of accepted rows of the stream."""
        if confidence_threshold is None:
            confidence_threshold = getattr(self.config.gan_config, self._CONFIG_KEY)['confidence_threshold']
        eng = self.engine
        num = int(num)
        if self.generator.training:
            return self._generate_qualified_train_mode(int(target_label), num, float(confidence_threshold))
        keeps, xs, idxs = [], [], []
        produced, accepted = 0, 0
        base = self._gen_rows
        consumed, got = 0, 0
        while True:
            rate = (accepted + 1) / (produced + 2)
            want = int(min(max(1024, 1.5 * (num - accepted) / rate + 256), 1 << 22))
            want = max(want, 200 + 10)   # patience 20 x chunk 10 of pure rejects must fit
            x, idx, cnt, _, keep = eng.generate_filter(int(target_label), want, float(confidence_threshold),
                                                       seed=self._seed, row_offset=base + produced, want_keep=True)
            c = int(cnt.item())
            keeps.append(keep.cpu())
            xs.append(x[:c])
            idxs.append(idx[:c])
            produced += want
            accepted += c
            try:
                consumed, got = patience_scan(torch.cat(keeps), num)
                break
            except Exception as ex:          # stream too short for the loop to terminate: extend it
                if "too short" not in str(ex):
                    raise
        self.classifier.train()              # the reference leaves C in train mode (cvae_gan.py:363)
        self._gen_rows = base + consumed
        if got == 0:
            return torch.tensor([])
        x = torch.cat(xs)
        idx = torch.cat(idxs)
        sel = idx < (base + consumed)
        x, idx = x[sel], idx[sel]
        order = torch.argsort(idx)
        return x[order][:got].cpu()

    def _generate_qualified_train_mode(self, target_label: int, num: int, confidence_threshold: float):
        """The same loop while the generator is still in TRAIN mode (called before `fit()`, or after `generator.train()`): the
        reference then normalises every chunk of <= 10 rows with that chunk's own batch statistics (and updates the running
        statistics per chunk), so rows are not independent and the loop is run literally, chunk by chunk, like
        cvae_gan.py:355-376: train-mode generator forward, eval-mode classifier, filter, one host read-back per chunk.
        A final chunk of one row fails like torch's BatchNorm does ("Expected more than 1 value per channel when training")."""
        eng = self.engine
        result, patience = [], 20
        while len(result) < num and patience > 0:
            n = min(10, num - len(result))
            if n < 2:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size torch.Size([{n}, "
                                 f"{self.generator.main_model[0].out_features}])")
            x = eng.generate(target_label, n, seed=self._seed, row_offset=self._gen_rows, train_mode=True)
            self._gen_rows += n
            self._bn_calls[NET_GENERATOR] += 1
            keep = eng.filter_logits(eng.classifier_forward(x), target_label, confidence_threshold)
            valid = x[keep].cpu()
            result.extend(valid)
            if len(valid) == 0:
                patience -= 1
        self._sync_bn_counters()
        self.classifier.train()              # the reference leaves C in train mode (cvae_gan.py:363)
        return torch.stack(result) if result else torch.tensor([])

    def reconstruct_samples(self, samples: torch.Tensor, labels: torch.Tensor):
        """cvae_gan.py:380-397.  The reference ALWAYS raises here: it hands 1-D labels to the generator,
        whose forward requires a 2-D one-hot condition (cvae_gan_models.py:142-143).  Kept as is; use
        `reconstruct()` for a working E -> G round trip."""
        raise ValueError(f"条件应为2D张量，实际: {labels.shape}")

    def reconstruct(self, samples: torch.Tensor, label: int):
        """Working variant: x -> E (eval) -> z = mu + eps*std -> G (eval), all rows share one label."""
        eng = self.engine
        mu, lv = eng.encoder_forward(samples, int(label))
        z = (mu + torch.randn_like(mu) * torch.exp(0.5 * lv)).contiguous()
        return eng.generate(int(label), z.size(0), z=z, train_mode=False).cpu()

    # ------------------------------------------------------------------------------------------------
    def state_dict(self):
        self._sync_bn_counters()
        return OrderedDict((n, getattr(self, n).state_dict()) for n in
                           ("encoder", "generator", "discriminator", "classifier"))

    def load_state_dict(self, sd):
        for net, n in enumerate(("encoder", "generator", "discriminator", "classifier")):
            self.engine.load_state(net, sd[n])
            for k, v in sd[n].items():
                if k.endswith("num_batches_tracked"):
                    dict(getattr(self, n).named_buffers())[k].fill_(int(v))

"""ORACLE - TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference CVAE-GAN hot path.

This file restates, in plain functional PyTorch on the CPU, the arithmetic of
    /root/reference/src/cvae_gan.py                 (trainer, losses, sampling, filter)
    /root/reference/src/models/cvae_gan_models.py   (encoder / generator / critic / classifier)
    /root/reference/src/utils.py:95-102             (init_weights)
so that the CUDA product can be checked against it on a box where the reference itself is
absent.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
(`cvae_gan_b200/`) never does, and fails loudly when its CUDA library is missing.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference (imported from a
scratch copy of /root/reference) and stores its inputs/outputs in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` replays them through this file (same torch CPU generator, same
draw order) and requires agreement to float32 round-off.  The reference
ships no golden vectors of its own, so running it is the only pin.
The sibling trainers restated here are pinned the same way, each by its own generating script:
    /root/reference/src/cgan.py     step_g_prior / fit_cgan          oracle/make_golden_cgan.py   -> ref_cgan_{a,b}.npz
    /root/reference/src/cvae.py     step_g_cvae / fit_cvae / reconstruct_samples_cvae
                                                                     oracle/make_golden_cvae.py   -> ref_cvae_{a,b}.npz
    /root/reference/src/vae_gan.py  step_g_vaegan / fit_vaegan / reconstruct_samples_vaegan (unconditional forwards)
                                                                     oracle/make_golden_vaegan.py -> ref_vaegan.npz
`OracleConfig.hidden` (layer widths the reference hard-codes; BASELINE.json configs[4]) is NOT expressible in the reference: it
is pinned only at the reference's own widths (the golden fit replayed with hidden = (256, 128, 64)); at other widths the
restatement is the same code path with different shapes and has no reference output to be compared with.

Third-party arithmetic: everything the reference computes is a `torch` op (un-pinned by the
reference; this image has torch 2.11.0+cu128).  Spectral-norm semantics follow the installed
`torch/nn/utils/parametrizations.py:_SpectralNorm` (1 power iteration per train-mode forward,
eps 1e-12, sigma = u.(W v) with u,v detached, W/sigma).

Two noise modes:
  * `TorchNoise`     draws from torch's global CPU generator with exactly the calls, shapes and
                     order the reference makes (so a seeded replay reproduces the reference);
  * `InjectedNoise`  hands out tensors the caller prepared (the same tensors are given to the
                     CUDA kernels through the C ABI), because torch's mt19937 stream cannot be
                     reproduced by an in-kernel Philox generator.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

NETS = ("encoder", "generator", "discriminator", "classifier")

# ---------------------------------------------------------------------------------------------
# configuration (values of /root/reference/src/config/gan_config.py:1-21)
# ---------------------------------------------------------------------------------------------


class OracleConfig:
    def __init__(self, **kw):
        self.epochs = 500
        self.batch_size = 128
        self.z_size = 128
        self.g_lr = 2e-4
        self.g_loop_num = 3
        self.d_lr = 2e-4
        self.d_loop_num = 5
        self.c_lr = 1e-4
        self.c_loop_num = 5
        self.lambda_recon = 1.0
        self.lambda_kl = 0.1
        self.lambda_adv = 1.0
        self.lambda_class = 0.5
        self.confidence_threshold = 0.5
        self.epoch_offset = 0  # first value of `e` (the reference always starts at 0)
        self.hidden = None     # (h1, h2, h3): widened restatement (BASELINE.json configs[4]); None = the reference's widths
        self.unconditional = False   # True: E / G / D without label columns (sibling trainer VAE-GAN, src/vae_gan.py)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def lambda_class_schedule(e: int, lambda_class: float) -> float:
    """cvae_gan.py:198-204."""
    if e < 200:
        return 0.0
    if e < 500:
        return lambda_class * ((e - 200) / 300)
    return lambda_class


# ---------------------------------------------------------------------------------------------
# layer shapes (cvae_gan_models.py:14-18, 83-87, 171-175, 257-259)
# ---------------------------------------------------------------------------------------------


def hidden_sizes(total_in: int, fixed3: bool, hidden=None):
    """`hidden` = None: the reference's formulas.  (h1, h2, h3): the widened restatement SURVEY 7.1 / 8(d) C5 asks for (the
    reference hard-codes its widths, so this is NOT expressible there): the same three widths in every network; with the
    reference's own values for a shape it reproduces the reference exactly (tests/test_oracle_golden.py)."""
    if hidden:
        return tuple(int(v) for v in hidden)
    h1 = max(256, total_in)
    h2 = max(128, total_in // 2)
    h3 = 64 if fixed3 else max(64, total_in // 4)
    return h1, h2, h3


def tensor_table(net: str, F_: int, K: int, Z: int, hidden=None, unconditional=False):
    """(key, shape, kind) in reference `state_dict()` order; kind in {param, buffer}."""
    t = []
    Kc = 0 if unconditional else K      # label columns of E / G / D (the VAE-GAN sibling's networks have none)
    if net == "encoder":
        tin = F_ + Kc
        h = hidden_sizes(tin, False, hidden)
        dims = [tin, *h]
        for i, li in enumerate((0, 3, 6)):
            t.append((f"encoder.{li}.weight", (dims[i + 1], dims[i]), "param"))
            t.append((f"encoder.{li}.bias", (dims[i + 1],), "param"))
            t.append((f"encoder.{li + 1}.weight", (dims[i + 1],), "param"))
            t.append((f"encoder.{li + 1}.bias", (dims[i + 1],), "param"))
            t.append((f"encoder.{li + 1}.running_mean", (dims[i + 1],), "buffer"))
            t.append((f"encoder.{li + 1}.running_var", (dims[i + 1],), "buffer"))
            t.append((f"encoder.{li + 1}.num_batches_tracked", (), "buffer_i64"))
        t.append(("fc_mu.weight", (Z, h[2]), "param"))
        t.append(("fc_mu.bias", (Z,), "param"))
        t.append(("fc_logvar.weight", (Z, h[2]), "param"))
        t.append(("fc_logvar.bias", (Z,), "param"))
    elif net == "generator":
        tin = Z + Kc
        h = hidden_sizes(tin, False, hidden)
        dims = [tin, *h]
        for i, li in enumerate((0, 3, 6)):
            t.append((f"main_model.{li}.weight", (dims[i + 1], dims[i]), "param"))
            t.append((f"main_model.{li}.bias", (dims[i + 1],), "param"))
            t.append((f"main_model.{li + 1}.weight", (dims[i + 1],), "param"))
            t.append((f"main_model.{li + 1}.bias", (dims[i + 1],), "param"))
            t.append((f"main_model.{li + 1}.running_mean", (dims[i + 1],), "buffer"))
            t.append((f"main_model.{li + 1}.running_var", (dims[i + 1],), "buffer"))
            t.append((f"main_model.{li + 1}.num_batches_tracked", (), "buffer_i64"))
        t.append(("last_layer.0.weight", (F_, h[2]), "param"))
        t.append(("last_layer.0.bias", (F_,), "param"))
    elif net == "discriminator":
        tin = F_ + Kc
        h = hidden_sizes(tin, True, hidden)
        dims = [tin, *h, 1]
        for i, li in enumerate((0, 3, 6, 8)):
            p = f"discriminator_network.{li}"
            t.append((f"{p}.bias", (dims[i + 1],), "param"))
            t.append((f"{p}.parametrizations.weight.original", (dims[i + 1], dims[i]), "param"))
            t.append((f"{p}.parametrizations.weight.0._u", (dims[i + 1],), "buffer"))
            t.append((f"{p}.parametrizations.weight.0._v", (dims[i],), "buffer"))
    elif net == "classifier":
        h = hidden_sizes(F_, True, hidden)
        dims = [F_, *h, K]
        for i, li in enumerate((0, 3, 7, 9)):
            t.append((f"classifier_network.{li}.weight", (dims[i + 1], dims[i]), "param"))
            t.append((f"classifier_network.{li}.bias", (dims[i + 1],), "param"))
            if li == 3:
                t.append(("classifier_network.4.weight", (dims[i + 1],), "param"))
                t.append(("classifier_network.4.bias", (dims[i + 1],), "param"))
    else:
        raise KeyError(net)
    return t


# ---------------------------------------------------------------------------------------------
# noise sources
# ---------------------------------------------------------------------------------------------


class TorchNoise:
    """Draws from torch's global CPU generator exactly like the reference does on device 'cpu'.

    randn            -> torch.randn(B, Z)                     (cvae_gan.py:114,140,173,344)
    randn_like       -> torch.randn_like(std)                 (cvae_gan_models.py:68)
    dropout_mask     -> empty_like(x).bernoulli_(1-p)         (what F.dropout does on CPU)
    randperm/randint -> torch.randperm(n) / torch.randint     (cvae_gan.py:252,259)
    """

    def randn(self, rows: int, cols: int, tag: str = "") -> torch.Tensor:
        return torch.randn(rows, cols)

    def randn_like(self, t: torch.Tensor, tag: str = "") -> torch.Tensor:
        return torch.randn_like(t)

    def dropout_mask(self, rows: int, cols: int, p: float, tag: str = "") -> torch.Tensor:
        return torch.empty(rows, cols).bernoulli_(1.0 - p)

    def randperm(self, n: int) -> torch.Tensor:
        return torch.randperm(n)

    def randint(self, n: int, size: int) -> torch.Tensor:
        return torch.randint(0, n, (size,))


class InjectedNoise:
    """Hands out caller-prepared tensors by tag, in FIFO order per tag.

    Tags used by the oracle:
      'z'        prior noise [B,Z]         'eps'  reparameterisation noise [B,Z]
      'd_mask1'  critic dropout keep-mask [B,256]   'd_mask2'  [B,128]
      'c_mask1'  classifier keep-mask [B,256]       'c_mask2'  [B,128]
      'idx'      row indices for _get_target_samples (any branch that draws)
    Keep-masks hold 0.0 / 1.0 (1 = kept); the oracle divides by (1-p) like F.dropout.
    """

    def __init__(self, dtype=torch.float32):
        self.q: Dict[str, List[torch.Tensor]] = {}
        self.dtype = dtype          # float64: the twin run that measures the reference's fp32 round-off (tests)

    def push(self, tag: str, t: torch.Tensor):
        self.q.setdefault(tag, []).append(t)
        return self

    def _pop(self, tag: str) -> torch.Tensor:
        if tag not in self.q or not self.q[tag]:
            raise RuntimeError(f"InjectedNoise: no tensor queued for tag '{tag}'")
        return self.q[tag].pop(0)

    def randn(self, rows, cols, tag="z"):
        t = self._pop(tag)
        assert tuple(t.shape) == (rows, cols), (tag, t.shape, rows, cols)
        return t.to(self.dtype)

    def randn_like(self, ref, tag="eps"):
        t = self._pop(tag)
        assert t.shape == ref.shape
        return t.to(self.dtype)

    def dropout_mask(self, rows, cols, p, tag=""):
        t = self._pop(tag)
        assert tuple(t.shape) == (rows, cols), (tag, t.shape, rows, cols)
        return t.to(self.dtype)

    def randperm(self, n):
        return self._pop("idx")

    def randint(self, n, size):
        return self._pop("idx")


# ---------------------------------------------------------------------------------------------
# networks (functional; parameters live in dicts keyed by the reference's state_dict names)
# ---------------------------------------------------------------------------------------------

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LN_EPS = 1e-5
SN_EPS = 1e-12
LRELU = 0.2
DROP_P = 0.3


def _one_hot(label: int, rows: int, K: int, dtype=torch.float32) -> torch.Tensor:
    return F.one_hot(torch.full([rows], int(label), dtype=torch.long), num_classes=K).to(dtype)


def _cat_label(x, label, K: int):
    """cat(x, onehot(label)) - or x itself for the unconditional networks of the VAE-GAN sibling (vae_gan_models.py:37,95,144),
    whose first Linear has no label columns (K == 0)."""
    if K == 0:
        return x
    return torch.cat([x, _one_hot(label, x.shape[0], K, x.dtype)], dim=1)


def _bn(x, sd, prefix, train: bool, dp=None):
    """BatchNorm1d, SURVEY appendix A.2.  In train mode updates running stats in `sd` (also under
    no_grad, cvae_gan.py:113-115).  `dp` optionally all-reduces the batch moments (data parallel)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if not train:
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b,
                            False, BN_MOMENTUM, BN_EPS)
    if dp is None:
        # let torch do it, exactly as nn.BatchNorm1d.forward does (updates running stats in place)
        sd[prefix + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b,
                            True, BN_MOMENTUM, BN_EPS)
    # data-parallel restatement: global moments from all-reduced [sum, sumsq, count]
    n_local = x.shape[0]
    s = torch.cat([x.sum(0), (x * x).sum(0), x.new_tensor([float(n_local)])]).double()
    s = dp.all_reduce_sum_autograd(s)
    n = s[-1]
    C = x.shape[1]
    mean = s[:C] / n
    var = (s[C:2 * C] / n - mean * mean).clamp_min(0.0)
    with torch.no_grad():
        sd[prefix + ".num_batches_tracked"] += 1
        sd[prefix + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.float())
        sd[prefix + ".running_var"].mul_(1 - BN_MOMENTUM).add_(
            BN_MOMENTUM * (var * n / (n - 1)).float())
    xh = (x - mean.float()) * torch.rsqrt(var.float() + BN_EPS)
    return xh * w + b


def encoder_forward(sd, x, label: int, train: bool, dp=None):
    """cvae_gan_models.py:49-64: cat(x, onehot) -> [Lin, BN, LReLU]x3 -> fc_mu, fc_logvar."""
    K = sd["encoder.0.weight"].shape[1] - x.shape[1]
    h = _cat_label(x, label, K)
    for li in (0, 3, 6):
        h = F.linear(h, sd[f"encoder.{li}.weight"], sd[f"encoder.{li}.bias"])
        h = _bn(h, sd, f"encoder.{li + 1}", train, dp)
        h = F.leaky_relu(h, LRELU)
    mu = F.linear(h, sd["fc_mu.weight"], sd["fc_mu.bias"])
    log_var = F.linear(h, sd["fc_logvar.weight"], sd["fc_logvar.bias"])
    return mu, log_var


def generator_forward(sd, z, label: int, train: bool, dp=None):
    """cvae_gan_models.py:136-156: cat(z, onehot) -> [Lin, BN, LReLU]x3 -> Lin -> Sigmoid."""
    K = sd["main_model.0.weight"].shape[1] - z.shape[1]
    h = _cat_label(z, label, K)
    for li in (0, 3, 6):
        h = F.linear(h, sd[f"main_model.{li}.weight"], sd[f"main_model.{li}.bias"])
        h = _bn(h, sd, f"main_model.{li + 1}", train, dp)
        h = F.leaky_relu(h, LRELU)
    return torch.sigmoid(F.linear(h, sd["last_layer.0.weight"], sd["last_layer.0.bias"]))


def _sn_weight(sd, prefix: str, train: bool):
    """torch _SpectralNorm.forward: one in-place power iteration in train mode, then W/sigma."""
    W = sd[prefix + ".parametrizations.weight.original"]
    u = sd[prefix + ".parametrizations.weight.0._u"]
    v = sd[prefix + ".parametrizations.weight.0._v"]
    if train:
        with torch.no_grad():
            Wm = W.detach()
            u.copy_(F.normalize(torch.mv(Wm, v), dim=0, eps=SN_EPS))
            v.copy_(F.normalize(torch.mv(Wm.t(), u), dim=0, eps=SN_EPS))
    uu, vv = u.clone(), v.clone()
    sigma = torch.vdot(uu, torch.mv(W, vv))
    return W / sigma


def discriminator_forward(sd, x, label: int, train: bool, noise=None, mask_tags=("d_mask1", "d_mask2")):
    """cvae_gan_models.py:215-230: cat(x, onehot) -> SN-Lin, LReLU, Drop -> SN-Lin, LReLU, Drop
    -> SN-Lin, LReLU -> SN-Lin.  Raw critic score [B,1]."""
    K = sd["discriminator_network.0.parametrizations.weight.original"].shape[1] - x.shape[1]
    h = _cat_label(x, label, K)
    for i, li in enumerate((0, 3, 6)):
        p = f"discriminator_network.{li}"
        h = F.linear(h, _sn_weight(sd, p, train), sd[p + ".bias"])
        h = F.leaky_relu(h, LRELU)
        if i < 2 and train:
            m = noise.dropout_mask(h.shape[0], h.shape[1], DROP_P, tag=mask_tags[i])
            h = h * (m / (1.0 - DROP_P))
    p = "discriminator_network.8"
    return F.linear(h, _sn_weight(sd, p, train), sd[p + ".bias"])


def classifier_forward(sd, x, train: bool, noise=None, mask_tags=("c_mask1", "c_mask2")):
    """cvae_gan_models.py:280-283: Lin, ReLU, Drop -> Lin, LayerNorm, ReLU, Drop -> Lin, ReLU -> Lin."""
    h = F.linear(x, sd["classifier_network.0.weight"], sd["classifier_network.0.bias"])
    h = F.relu(h)
    if train:
        m = noise.dropout_mask(h.shape[0], h.shape[1], DROP_P, tag=mask_tags[0])
        h = h * (m / (1.0 - DROP_P))
    h = F.linear(h, sd["classifier_network.3.weight"], sd["classifier_network.3.bias"])
    h = F.layer_norm(h, (h.shape[1],), sd["classifier_network.4.weight"], sd["classifier_network.4.bias"], LN_EPS)
    h = F.relu(h)
    if train:
        m = noise.dropout_mask(h.shape[0], h.shape[1], DROP_P, tag=mask_tags[1])
        h = h * (m / (1.0 - DROP_P))
    h = F.relu(F.linear(h, sd["classifier_network.7.weight"], sd["classifier_network.7.bias"]))
    return F.linear(h, sd["classifier_network.9.weight"], sd["classifier_network.9.bias"])


# ---------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam single-tensor path; betas (0.5, 0.999), eps 1e-8, no decay/amsgrad)
# ---------------------------------------------------------------------------------------------


class OracleAdam:
    def __init__(self, params: List[torch.Tensor], lr: float, betas=(0.5, 0.999), eps=1e-8):
        self.params, self.lr, self.b1, self.b2, self.eps = params, lr, betas[0], betas[1], eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    @torch.no_grad()
    def step(self, grads: List[Optional[torch.Tensor]]):
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if g is None:
                continue
            m.lerp_(g, 1.0 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


# ---------------------------------------------------------------------------------------------
# filter (cvae_gan.py:366-371)
# ---------------------------------------------------------------------------------------------


def filter_logits(logits: torch.Tensor, label: int, thr: float) -> torch.Tensor:
    probs = torch.softmax(logits, dim=1)
    max_probs, preds = torch.max(probs, dim=1)
    return (max_probs > thr) & (preds == label)


# ---------------------------------------------------------------------------------------------
# data-parallel helper used by the gloo world_size-2 CPU tests
# ---------------------------------------------------------------------------------------------


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        import torch.distributed as dist
        out = t.clone()
        dist.all_reduce(out)
        return out

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        g = g.clone()
        dist.all_reduce(g)
        return g


class DataParallelCtx:
    """Batch rows are sharded over ranks; BN moments and gradients are summed over ranks."""

    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world

    def all_reduce_sum_autograd(self, t):
        return _AllReduceSum.apply(t)

    def all_reduce_sum_(self, t):
        import torch.distributed as dist
        dist.all_reduce(t)
        return t

    def shard(self, t: torch.Tensor) -> torch.Tensor:
        n = t.shape[0]
        per = n // self.world
        assert per * self.world == n
        return t[self.rank * per:(self.rank + 1) * per]


# ---------------------------------------------------------------------------------------------
# trainer
# ---------------------------------------------------------------------------------------------


class OracleCVAEGAN:
    """Restates `CVAEGAN` (cvae_gan.py:9-378).  State is four dicts keyed like the reference's
    `state_dict()`s; it must be initialised from a reference-format state (`load_state`) or by
    `init_like_reference()` (same distributions as the reference, not the same draws)."""

    def __init__(self, feature_num: int, label_num: int, cfg: Optional[OracleConfig] = None, dp=None):
        self.feature_num, self.label_num = feature_num, label_num
        self.cfg = cfg or OracleConfig()
        self.dp = dp
        self.sd: Dict[str, "OrderedDict[str, torch.Tensor]"] = {n: OrderedDict() for n in NETS}
        self.samples: "OrderedDict[int, torch.Tensor]" = OrderedDict()
        self.loss_history = {"recon_loss": [], "kl_loss": [], "adv_loss": [], "class_loss": []}
        self.training = {n: True for n in NETS}
        self.opt: Dict[str, OracleAdam] = {}
        self.last_losses = {}

    # ---- state ------------------------------------------------------------------------------
    def load_state(self, states: Dict[str, Dict[str, torch.Tensor]]):
        for net in NETS:
            tab = tensor_table(net, self.feature_num, self.label_num, self.cfg.z_size, self.cfg.hidden, self.cfg.unconditional)
            sd = OrderedDict()
            for key, shape, kind in tab:
                t = torch.as_tensor(states[net][key]).clone()
                assert tuple(t.shape) == tuple(shape), (net, key, t.shape, shape)
                if kind == "param":
                    t = t.float().requires_grad_(True)
                elif kind == "buffer":
                    t = t.float()
                else:
                    t = t.long()
                sd[key] = t
            self.sd[net] = sd
        return self

    def twin64(self):
        """The same model in float64 (parameters, buffers, Adam moments): running identical steps on it measures how far
        the reference's OWN float32 arithmetic is from the exact result - the yardstick for tolerances (tests/parity.py)."""
        import copy
        t = copy.copy(self)
        t.sd = {n: OrderedDict() for n in NETS}
        for n in NETS:
            for k, v in self.sd[n].items():
                if v.is_floating_point():
                    w = v.detach().double().clone()
                    t.sd[n][k] = w.requires_grad_(True) if v.requires_grad else w
                else:
                    t.sd[n][k] = v.clone()
        t.samples = OrderedDict((k, v.double()) for k, v in self.samples.items())
        t.loss_history = {k: [] for k in self.loss_history}
        t.training = dict(self.training)
        t.last_losses = {}
        t.opt = {}
        if self.opt:
            t.make_optimizers()
            for n in NETS:
                t.opt[n].t = self.opt[n].t
                t.opt[n].m = [m.double().clone() for m in self.opt[n].m]
                t.opt[n].v = [v.double().clone() for v in self.opt[n].v]
        return t

    def init_like_reference(self, generator: Optional[torch.Generator] = None):
        """Same *distributions* as the reference constructor (utils.py:95-102 + torch defaults for
        the parametrised critic layers + SN u,v init with 15 power iterations)."""
        g = generator
        states = {}
        for net in NETS:
            sd = {}
            for key, shape, kind in tensor_table(net, self.feature_num, self.label_num, self.cfg.z_size, self.cfg.hidden, self.cfg.unconditional):
                if kind == "buffer_i64":
                    sd[key] = torch.zeros((), dtype=torch.long)
                elif key.endswith("running_mean"):
                    sd[key] = torch.zeros(shape)
                elif key.endswith("running_var"):
                    sd[key] = torch.ones(shape)
                elif net == "discriminator":
                    if key.endswith("original"):
                        bound = 1.0 / math.sqrt(shape[1])
                        sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
                    elif key.endswith("bias"):
                        sd[key] = None  # filled below (needs fan_in)
                    else:
                        sd[key] = F.normalize(torch.randn(shape, generator=g), dim=0, eps=SN_EPS)
                elif key == "classifier_network.4.weight":
                    sd[key] = torch.ones(shape)
                elif len(shape) == 2:
                    sd[key] = torch.randn(shape, generator=g) * 0.02
                elif key.endswith("weight"):  # BN gamma
                    sd[key] = 1.0 + torch.randn(shape, generator=g) * 0.02
                else:
                    sd[key] = torch.zeros(shape)
            if net == "discriminator":
                for li in (0, 3, 6, 8):
                    p = f"discriminator_network.{li}"
                    W = sd[p + ".parametrizations.weight.original"]
                    bound = 1.0 / math.sqrt(W.shape[1])
                    sd[p + ".bias"] = (torch.rand(W.shape[0], generator=g) * 2 - 1) * bound
                    u, v = sd[p + ".parametrizations.weight.0._u"], sd[p + ".parametrizations.weight.0._v"]
                    for _ in range(15):
                        u = F.normalize(torch.mv(W, v), dim=0, eps=SN_EPS)
                        v = F.normalize(torch.mv(W.t(), u), dim=0, eps=SN_EPS)
                    sd[p + ".parametrizations.weight.0._u"], sd[p + ".parametrizations.weight.0._v"] = u, v
            states[net] = sd
        return self.load_state(states)

    def state(self) -> Dict[str, Dict[str, torch.Tensor]]:
        return {n: OrderedDict((k, v.detach().clone()) for k, v in self.sd[n].items()) for n in NETS}

    def params(self, net: str) -> List[torch.Tensor]:
        """In `module.parameters()` order (== tensor_table order restricted to params)."""
        return [t for t in self.sd[net].values() if t.requires_grad]

    def param_keys(self, net: str) -> List[str]:
        return [k for k, t in self.sd[net].items() if t.requires_grad]

    # ---- data -------------------------------------------------------------------------------
    def divide_samples(self, x: torch.Tensor, y: torch.Tensor):
        """cvae_gan.py:238-245 (result only; key order = first occurrence)."""
        self.samples = OrderedDict()
        for lab in y.tolist():
            if lab not in self.samples:
                self.samples[lab] = x[y == lab]
        return self.samples

    def get_target_samples(self, label: int, num: int, noise) -> torch.Tensor:
        """cvae_gan.py:247-260."""
        avail = self.samples[label]
        if len(avail) < num:
            return avail[noise.randint(len(avail), num)]
        if len(avail) == num:
            return avail
        return avail[noise.randperm(len(avail))[:num]]

    # ---- optimisers (cvae_gan.py:75-97) -------------------------------------------------------
    def make_optimizers(self):
        c = self.cfg
        self.opt = {
            "encoder": OracleAdam(self.params("encoder"), c.g_lr),
            "generator": OracleAdam(self.params("generator"), c.g_lr),
            "discriminator": OracleAdam(self.params("discriminator"), c.d_lr),
            "classifier": OracleAdam(self.params("classifier"), c.c_lr),
        }

    def _backward(self, loss, nets, apply_update=True):
        ps = [p for n in nets for p in self.params(n)]
        grads = torch.autograd.grad(loss, ps, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, ps)]
        if self.dp is not None:
            for g in grads:
                self.dp.all_reduce_sum_(g)
        out, i = {}, 0
        for n in nets:
            k = len(self.params(n))
            out[n] = list(grads[i:i + k])
            i += k
            if apply_update:
                self.opt[n].step(out[n])
        return out

    def _mean(self, t):
        """Mean over the *global* batch: local sum / global rows (autograd-safe)."""
        if self.dp is None:
            return t.mean()
        return t.sum() / (t.shape[0] * self.dp.world * (t.numel() // t.shape[0]))

    # ---- the three step types (cvae_gan.py:104-216) -------------------------------
    def step_d(self, x_real, label: int, noise, apply_update=True):
        B = x_real.shape[0]
        with torch.no_grad():
            z = noise.randn(B, self.cfg.z_size, tag="z")
            x_gen = generator_forward(self.sd["generator"], z, label, self.training["generator"], self.dp)
        D = self.sd["discriminator"]
        tr = self.training["discriminator"]
        d_real = discriminator_forward(D, x_real, label, tr, noise)
        d_fake = discriminator_forward(D, x_gen.detach(), label, tr, noise)
        d_loss = -self._mean(d_real) + self._mean(d_fake)
        grads = self._backward(d_loss, ["discriminator"], apply_update)
        self.last_losses["d_loss"] = self._global_scalar(d_loss)
        return d_loss.detach(), grads

    def step_c(self, x_real, label: int, noise, apply_update=True):
        B = x_real.shape[0]
        tgt = torch.full([B], int(label), dtype=torch.long)
        with torch.no_grad():
            z = noise.randn(B, self.cfg.z_size, tag="z")
            x_gen = generator_forward(self.sd["generator"], z, label, self.training["generator"], self.dp)
        C = self.sd["classifier"]
        tr = self.training["classifier"]
        ce_real = self._ce(classifier_forward(C, x_real, tr, noise), tgt)
        ce_fake = self._ce(classifier_forward(C, x_gen, tr, noise), tgt)
        c_loss = ce_real + ce_fake
        grads = self._backward(c_loss, ["classifier"], apply_update)
        self.last_losses["c_loss"] = self._global_scalar(c_loss)
        return c_loss.detach(), grads

    def _ce(self, logits, tgt):
        if self.dp is None:
            return F.cross_entropy(logits, tgt)
        return F.cross_entropy(logits, tgt, reduction="sum") / (logits.shape[0] * self.dp.world)

    def _global_scalar(self, t):
        t = t.detach().clone()
        if self.dp is not None:
            self.dp.all_reduce_sum_(t)
        return float(t)

    def step_g(self, x_real, label: int, noise, lambda_class_now: float, apply_update=True):
        c = self.cfg
        B = x_real.shape[0]
        tgt = torch.full([B], int(label), dtype=torch.long)
        E, G = self.sd["encoder"], self.sd["generator"]
        mu, log_var = encoder_forward(E, x_real, label, self.training["encoder"], self.dp)
        std = torch.exp(0.5 * log_var)
        eps = noise.randn_like(std, tag="eps")
        z_enc = mu + eps * std
        z_prior = noise.randn(B, c.z_size, tag="z")
        x_recon = generator_forward(G, z_enc, label, self.training["generator"], self.dp)
        x_fake = generator_forward(G, z_prior, label, self.training["generator"], self.dp)
        if self.dp is None:
            recon = F.mse_loss(x_recon, x_real)
            kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / mu.size(0)
        else:
            Bg = B * self.dp.world
            recon = ((x_recon - x_real) ** 2).sum() / (Bg * x_real.shape[1])
            kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / Bg
        d_fake = discriminator_forward(self.sd["discriminator"], x_fake, label,
                                       self.training["discriminator"], noise)
        adv = -self._mean(d_fake)
        cls = self._ce(classifier_forward(self.sd["classifier"], x_fake,
                                          self.training["classifier"], noise), tgt)
        total = c.lambda_recon * recon + c.lambda_kl * kl + c.lambda_adv * adv + lambda_class_now * cls
        grads = self._backward(total, ["encoder", "generator"], apply_update)
        losses = {k: self._global_scalar(v) for k, v in
                  (("recon_loss", recon), ("kl_loss", kl), ("adv_loss", adv), ("class_loss", cls))}
        self.last_losses.update(losses)
        return losses, grads

    # ---- sibling trainer CGAN: cgan.py:138-178, the generator step without the VAE branch ------------
    def step_g_prior(self, label: int, B: int, noise, lambda_class_now: float, apply_update=True):
        """CGAN generator step: x_fake = G(z_prior, onehot) only; total = lambda_adv * (-mean D(x_fake)) + lambda_now * CE;
        Adam on the generator alone (cgan.py:159-178).  No real batch is drawn in this step."""
        c = self.cfg
        tgt = torch.full([B], int(label), dtype=torch.long)
        z_prior = noise.randn(B, c.z_size, tag="z")
        x_fake = generator_forward(self.sd["generator"], z_prior, label, self.training["generator"], self.dp)
        d_fake = discriminator_forward(self.sd["discriminator"], x_fake, label, self.training["discriminator"], noise)
        adv = -self._mean(d_fake)
        cls = self._ce(classifier_forward(self.sd["classifier"], x_fake, self.training["classifier"], noise), tgt)
        total = c.lambda_adv * adv + lambda_class_now * cls
        grads = self._backward(total, ["generator"], apply_update)
        losses = {"adv_loss": self._global_scalar(adv), "class_loss": self._global_scalar(cls)}
        self.last_losses.update(losses)
        return losses, grads

    def fit_cgan(self, x: torch.Tensor, y: torch.Tensor, noise=None, step_hook=None):
        """CGAN.fit (cgan.py:51-196): the critic and classifier steps are the CVAE-GAN's (cgan.py:84-136 is
        cvae_gan.py:104-157 statement for statement); the generator step is `step_g_prior`; loss_history keeps adv / class."""
        noise = noise or TorchNoise()
        c = self.cfg
        self.loss_history = {"adv_loss": [], "class_loss": []}
        for n in ("generator", "discriminator", "classifier"):
            self.training[n] = True
        self.divide_samples(x, y)
        self.make_optimizers()
        for e in range(c.epoch_offset, c.epoch_offset + c.epochs):
            losses = None
            for label in self.samples.keys():
                for _ in range(c.d_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    out = self.step_d(xr, label, noise)
                    if step_hook:
                        step_hook("d", e, label, out)
                for _ in range(c.c_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    out = self.step_c(xr, label, noise)
                    if step_hook:
                        step_hook("c", e, label, out)
                for _ in range(c.g_loop_num):
                    losses, g = self.step_g_prior(label, c.batch_size, noise, lambda_class_schedule(e, c.lambda_class))
                    if step_hook:
                        step_hook("g", e, label, (losses, g))
            for k in self.loss_history:
                self.loss_history[k].append(losses[k])
        for n in ("generator", "discriminator", "classifier"):
            self.training[n] = False
        return self

    # ---- sibling trainer CVAE: cvae.py:117-166, the encoder/generator step without critic and prior pass -----
    def step_g_cvae(self, x_real, label: int, noise, lambda_class_now: float, apply_update=True):
        """CVAE encoder/generator step: mu, logvar = E(x); z_enc = mu + eps * std; x_recon = G(z_enc, onehot);
        total = lambda_recon * MSE + lambda_kl * KL + lambda_now * CE(C(x_recon), label) - the classifier reads the
        RECONSTRUCTION (cvae.py:141-142) -; Adam on encoder and generator (cvae.py:153-166)."""
        c = self.cfg
        B = x_real.shape[0]
        tgt = torch.full([B], int(label), dtype=torch.long)
        mu, log_var = encoder_forward(self.sd["encoder"], x_real, label, self.training["encoder"], self.dp)
        std = torch.exp(0.5 * log_var)
        eps = noise.randn_like(std, tag="eps")
        z_enc = mu + eps * std
        x_recon = generator_forward(self.sd["generator"], z_enc, label, self.training["generator"], self.dp)
        if self.dp is None:
            recon = F.mse_loss(x_recon, x_real)
            kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / mu.size(0)
        else:
            Bg = B * self.dp.world
            recon = ((x_recon - x_real) ** 2).sum() / (Bg * x_real.shape[1])
            kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / Bg
        cls = self._ce(classifier_forward(self.sd["classifier"], x_recon, self.training["classifier"], noise), tgt)
        total = c.lambda_recon * recon + c.lambda_kl * kl + lambda_class_now * cls
        grads = self._backward(total, ["encoder", "generator"], apply_update)
        losses = {k: self._global_scalar(v) for k, v in (("recon_loss", recon), ("kl_loss", kl), ("class_loss", cls))}
        self.last_losses.update(losses)
        return losses, grads

    def fit_cvae(self, x: torch.Tensor, y: torch.Tensor, noise=None, step_hook=None):
        """CVAE.fit (cvae.py:51-179): per label c_loop classifier steps (cvae.py:89-115 is cvae_gan.py:131-157 statement for
        statement) then g_loop `step_g_cvae` steps; no critic; loss_history keeps recon / kl / class."""
        noise = noise or TorchNoise()
        c = self.cfg
        self.loss_history = {"recon_loss": [], "kl_loss": [], "class_loss": []}
        for n in ("encoder", "generator", "classifier"):
            self.training[n] = True
        self.divide_samples(x, y)
        self.make_optimizers()
        for e in range(c.epoch_offset, c.epoch_offset + c.epochs):
            losses = None
            for label in self.samples.keys():
                for _ in range(c.c_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    out = self.step_c(xr, label, noise)
                    if step_hook:
                        step_hook("c", e, label, out)
                for _ in range(c.g_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    losses, g = self.step_g_cvae(xr, label, noise, lambda_class_schedule(e, c.lambda_class))
                    if step_hook:
                        step_hook("g", e, label, (losses, g))
            for k in self.loss_history:
                self.loss_history[k].append(losses[k])
        for n in ("encoder", "generator", "classifier"):
            self.training[n] = False
        return self

    # ---- sibling trainer VAE-GAN: vae_gan.py:76-141 - unconditional networks, no classifier, no label visits ----
    def step_g_vaegan(self, x_real, noise, apply_update=True):
        """VAE-GAN encoder/generator step (vae_gan.py:103-141): mu, logvar = E(x); z_enc (randn_like) ; z_prior (randn);
        x_recon = G(z_enc); x_fake = G(z_prior); total = lambda_recon * MSE + lambda_kl * KL + lambda_adv * (-mean D(x_fake));
        Adam on encoder and generator.  The critic step is `step_d(x, None, ...)` (vae_gan.py:77-101 is cvae_gan.py:104-128
        without the label arguments)."""
        c = self.cfg
        B = x_real.shape[0]
        mu, log_var = encoder_forward(self.sd["encoder"], x_real, None, self.training["encoder"], self.dp)
        std = torch.exp(0.5 * log_var)
        eps = noise.randn_like(std, tag="eps")
        z_enc = mu + eps * std
        z_prior = noise.randn(B, c.z_size, tag="z")
        x_recon = generator_forward(self.sd["generator"], z_enc, None, self.training["generator"], self.dp)
        x_fake = generator_forward(self.sd["generator"], z_prior, None, self.training["generator"], self.dp)
        if self.dp is None:
            recon = F.mse_loss(x_recon, x_real)
            kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / mu.size(0)
        else:
            Bg = B * self.dp.world
            recon = ((x_recon - x_real) ** 2).sum() / (Bg * x_real.shape[1])
            kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / Bg
        d_fake = discriminator_forward(self.sd["discriminator"], x_fake, None, self.training["discriminator"], noise)
        adv = -self._mean(d_fake)
        total = c.lambda_recon * recon + c.lambda_kl * kl + c.lambda_adv * adv
        grads = self._backward(total, ["encoder", "generator"], apply_update)
        losses = {k: self._global_scalar(v) for k, v in (("recon_loss", recon), ("kl_loss", kl), ("adv_loss", adv))}
        self.last_losses.update(losses)
        return losses, grads

    def get_random_samples(self, num: int, noise) -> torch.Tensor:
        """vae_gan.py:166-178: `_get_target_samples` over ALL rows (no labels)."""
        avail = self.samples
        if len(avail) < num:
            return avail[noise.randint(len(avail), num)]
        if len(avail) == num:
            return avail
        return avail[noise.randperm(len(avail))[:num]]

    def fit_vaegan(self, x: torch.Tensor, noise=None, step_hook=None):
        """VAEGAN.fit (vae_gan.py:42-157): per EPOCH d_loop critic steps and g_loop encoder/generator steps on batches drawn
        from all rows; loss_history keeps recon / kl / adv of the epoch's last step."""
        assert self.cfg.unconditional
        noise = noise or TorchNoise()
        c = self.cfg
        self.loss_history = {"recon_loss": [], "kl_loss": [], "adv_loss": []}
        for n in ("encoder", "generator", "discriminator"):
            self.training[n] = True
        self.samples = x
        self.make_optimizers()
        for e in range(c.epoch_offset, c.epoch_offset + c.epochs):
            losses = None
            for _ in range(c.d_loop_num):
                out = self.step_d(self.get_random_samples(c.batch_size, noise), None, noise)
                if step_hook:
                    step_hook("d", e, None, out)
            for _ in range(c.g_loop_num):
                losses, g = self.step_g_vaegan(self.get_random_samples(c.batch_size, noise), noise)
                if step_hook:
                    step_hook("g", e, None, (losses, g))
            for k in self.loss_history:
                self.loss_history[k].append(losses[k])
        for n in ("encoder", "generator", "discriminator"):
            self.training[n] = False
        return self

    @torch.no_grad()
    def reconstruct_samples_vaegan(self, x: torch.Tensor, noise=None) -> torch.Tensor:
        """VAEGAN.reconstruct_samples (vae_gan.py:244-261): eval-mode E and G, one randn_like, both left in TRAIN mode."""
        noise = noise or TorchNoise()
        mu, lv = encoder_forward(self.sd["encoder"], x, None, False, None)
        std = torch.exp(0.5 * lv)
        out = generator_forward(self.sd["generator"], mu + noise.randn_like(std, tag="eps") * std, None, False, None)
        self.training["encoder"] = self.training["generator"] = True
        return out

    @torch.no_grad()
    def reconstruct_samples_cvae(self, x: torch.Tensor, labels: torch.Tensor, noise=None) -> torch.Tensor:
        """CVAE.reconstruct_samples (cvae.py:300-319): E and G in eval mode (running statistics, so rows are independent and
        may be grouped by label), ONE randn_like draw for the whole batch, and both networks left in TRAIN mode afterwards."""
        noise = noise or TorchNoise()
        labels = labels.long()
        mu = torch.empty(x.shape[0], self.cfg.z_size, dtype=x.dtype)
        lv = torch.empty_like(mu)
        groups = [(int(lab), (labels == lab).nonzero().flatten()) for lab in torch.unique(labels)]
        for lab, sel in groups:
            mu[sel], lv[sel] = encoder_forward(self.sd["encoder"], x[sel], lab, False, None)
        std = torch.exp(0.5 * lv)
        z = mu + noise.randn_like(std, tag="eps") * std
        out = torch.empty(x.shape[0], self.feature_num, dtype=x.dtype)
        for lab, sel in groups:
            out[sel] = generator_forward(self.sd["generator"], z[sel], lab, False, None)
        self.training["encoder"] = self.training["generator"] = True
        return out

    # ---- fit (cvae_gan.py:59-236) --------------------------------------------------------------
    def fit(self, x: torch.Tensor, y: torch.Tensor, noise=None, step_hook=None):
        noise = noise or TorchNoise()
        c = self.cfg
        for n in NETS:
            self.training[n] = True
        self.divide_samples(x, y)
        self.make_optimizers()
        for e in range(c.epoch_offset, c.epoch_offset + c.epochs):
            losses = None
            for label in self.samples.keys():
                for _ in range(c.d_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    out = self.step_d(xr, label, noise)
                    if step_hook:
                        step_hook("d", e, label, out)
                for _ in range(c.c_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    out = self.step_c(xr, label, noise)
                    if step_hook:
                        step_hook("c", e, label, out)
                for _ in range(c.g_loop_num):
                    xr = self.get_target_samples(label, c.batch_size, noise)
                    losses, g = self.step_g(xr, label, noise, lambda_class_schedule(e, c.lambda_class))
                    if step_hook:
                        step_hook("g", e, label, (losses, g))
            for k in self.loss_history:
                self.loss_history[k].append(losses[k])
        for n in NETS:
            self.training[n] = False
        return self

    # ---- generation (cvae_gan.py:339-378) -------------------------------------------------------
    @torch.no_grad()
    def generate_samples(self, label: int, num: int, noise=None) -> torch.Tensor:
        noise = noise or TorchNoise()
        z = noise.randn(num, self.cfg.z_size, tag="z")
        return generator_forward(self.sd["generator"], z, label, self.training["generator"], None).detach()

    @torch.no_grad()
    def classify_eval(self, x: torch.Tensor) -> torch.Tensor:
        return classifier_forward(self.sd["classifier"], x, False)

    @torch.no_grad()
    def generate_qualified_samples(self, label: int, num: int, thr: Optional[float] = None,
                                   noise=None, chunk: int = 10, patience: int = 20) -> torch.Tensor:
        """Literal restatement of the chunk-of-10 / patience-20 loop.  NB: the
        reference leaves the classifier in TRAIN mode afterwards; mirrored in `self.training`."""
        noise = noise or TorchNoise()
        thr = self.cfg.confidence_threshold if thr is None else thr
        result: List[torch.Tensor] = []
        while len(result) < num and patience > 0:
            s = self.generate_samples(label, min(chunk, num - len(result)), noise)
            logits = self.classify_eval(s)
            self.training["classifier"] = True
            keep = filter_logits(logits, label, thr)
            valid = s[keep]
            result.extend(valid)
            if len(valid) == 0:
                patience -= 1
        return torch.stack(result) if result else torch.tensor([])

    @torch.no_grad()
    def generate_filter_stream(self, label: int, z: torch.Tensor, thr: float):
        """Vectorised form used to check the fused kernel: rows of `z` are the stream of prior
        draws in order; returns (x, logits, keep_mask).  Row-independent because G and C run in
        eval mode (BN running stats, no dropout, no power iteration)."""
        x = generator_forward(self.sd["generator"], z, label, False, None)
        logits = classifier_forward(self.sd["classifier"], x, False)
        return x, logits, filter_logits(logits, label, thr)


def patience_scan(keep: torch.Tensor, num: int, chunk: int = 10, patience: int = 20):
    """Given the accept mask of an (unbounded) row stream, return how many rows the reference loop
    (cvae_gan.py:355-376) would consume and how many it would accept.  Pure integer logic."""
    keep = keep.to(torch.bool).tolist()
    pos, got = 0, 0
    while got < num and patience > 0:
        n = min(chunk, num - got)
        if pos + n > len(keep):
            raise ValueError("stream too short")
        k = sum(keep[pos:pos + n])
        pos += n
        got += k
        if k == 0:
            patience -= 1
    return pos, got

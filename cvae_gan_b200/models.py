"""Host-side mirrors of the four reference networks (/root/reference/src/models/cvae_gan_models.py).

These classes keep the reference's constructor signatures, attribute / state_dict key names, forward
signatures and ValueError checks so that code written against the reference (`gan.classifier` handed
to `Classifier`, `state_dict()` round trips, `.train()/.eval()`) keeps working.  They are STATE
CONTAINERS: after `attach()` every parameter and float buffer is a view into the flat device
buffers that the CUDA engine trains in place.  The trainer (`CVAEGAN.fit`, `generate_*`) never calls
their `forward`; the arithmetic of the hot path lives in libcvaegan_b200.so.  `forward` exists for
callers outside the hot path (the reference's own `Classifier.fit/test`, which needs autograd):
the classifier routes no-grad eval calls to the CUDA kernel, everything else is boundary plumbing.

Construction draws from torch's CPU generator in the same order as the reference constructor
(default Linear init, then `init_weights`, spectral-norm u/v init with 15 power iterations), so the
same seed gives the same starting parameters as `src.CVAEGAN()`.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.utils.parametrizations import spectral_norm

NEG_SLOPE = 0.2
DROP_P = 0.3


def init_weights(layer: nn.Module):
    """/root/reference/src/utils.py:95-102 - exact-type match: N(0, 0.02) Linear weights, zero bias,
    BatchNorm gamma ~ N(1, 0.02).  Parametrised (spectral-norm) Linears are a different type -> skipped."""
    if type(layer) is nn.Linear:
        nn.init.normal_(layer.weight, 0.0, 0.02)
        if layer.bias is not None:
            nn.init.constant_(layer.bias, 0)
    elif type(layer) is nn.BatchNorm1d:
        nn.init.normal_(layer.weight, 1.0, 0.02)
        nn.init.constant_(layer.bias, 0)


def _widths(total_in: int, fixed_last: bool, hidden=None):
    """cvae_gan_models.py:16-18,85-87,173-175,257-259.  `hidden` (not expressible in the reference, whose widths are hard-coded):
    the widened model of BASELINE.json configs[4], the same three widths for every network."""
    if hidden:
        return tuple(int(v) for v in hidden)
    return max(256, total_in), max(128, total_in // 2), (64 if fixed_last else max(64, total_in // 4))


def _bn_stack(dims):
    mods = []
    for i in range(3):
        mods += [nn.Linear(dims[i], dims[i + 1]), nn.BatchNorm1d(dims[i + 1]), nn.LeakyReLU(NEG_SLOPE)]
    return nn.Sequential(*mods)


class _Attached(nn.Module):
    """Mixin: alias parameters / float buffers to the engine's flat buffers."""

    _engine = None
    _net = -1

    def attach(self, engine, net: int, copy_in: bool = True):
        sd = self.state_dict()
        if copy_in:
            engine.load_state(net, sd)
        named = dict(self.named_parameters())
        named.update({k: v for k, v in self.named_buffers()})
        for key in engine.tables[net]:
            t = named[key]
            t.data = engine.view(net, key)
        for name, b in self.named_buffers():   # num_batches_tracked: not in the engine, keep on device
            if b.dtype == torch.int64:
                b.data = b.data.to(engine.device)
        object.__setattr__(self, "_engine", engine)
        self._net = net
        return self

    def _one_hot(self, cond: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.one_hot(cond, num_classes=self.num_classes).float()


class CVAEGANEncoderModel(_Attached):
    def __init__(self, input_dim: int, num_classes: int, latent_dim: int = 128, hidden=None):
        super().__init__()
        self.input_dim, self.num_classes, self.latent_dim = input_dim, num_classes, latent_dim
        tin = input_dim + num_classes
        h = _widths(tin, False, hidden)
        self.encoder = _bn_stack([tin, *h])
        self.fc_mu = nn.Linear(h[2], latent_dim)
        self.fc_logvar = nn.Linear(h[2], latent_dim)
        self.apply(init_weights)

    def _process_condition(self, condition: torch.Tensor) -> torch.Tensor:
        if condition.dim() == 1:
            condition = condition.long()
        elif condition.dim() == 2 and condition.size(1) == 1:
            condition = condition.squeeze(1).long()
        else:
            raise ValueError(f"条件输入格式错误，期望1D或2D(单列)，实际: {condition.shape}")
        return self._one_hot(condition)

    def forward(self, x: torch.Tensor, condition: torch.Tensor) -> tuple:
        if x.dim() != 2:
            raise ValueError(f"输入数据应为2D张量，实际: {x.shape}")
        if x.size(1) != self.input_dim:
            raise ValueError(f"输入特征维度不匹配，期望: {self.input_dim}，实际: {x.size(1)}")
        hid = self.encoder(torch.cat([x, self._process_condition(condition)], dim=1))
        return self.fc_mu(hid), self.fc_logvar(hid)

    def reparameterize(self, mu: torch.Tensor, log_var: torch.Tensor) -> torch.Tensor:
        std = torch.exp(0.5 * log_var)
        return mu + torch.randn_like(std) * std

    def encode(self, x: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        return self.reparameterize(*self.forward(x, condition))


class CVAEGANGeneratorModel(_Attached):
    def __init__(self, latent_dim: int, num_classes: int, output_dim: int, hidden=None):
        super().__init__()
        self.latent_dim, self.num_classes, self.output_dim = latent_dim, num_classes, output_dim
        tin = latent_dim + num_classes
        h = _widths(tin, False, hidden)
        self.main_model = _bn_stack([tin, *h])
        self.hidden_status: torch.Tensor = None
        self.last_layer = nn.Sequential(nn.Linear(h[2], output_dim), nn.Sigmoid())
        self.apply(init_weights)

    def _process_condition(self, condition: torch.Tensor, target_batch_size: int = None) -> torch.Tensor:
        if condition.dim() == 0:
            condition = condition.unsqueeze(0)
        if condition.dim() == 1:
            condition = condition.long()
            if target_batch_size and condition.size(0) == 1:
                condition = condition.repeat(target_batch_size)
        elif condition.dim() == 2 and condition.size(1) == 1:
            condition = condition.squeeze(1).long()
        else:
            raise ValueError(f"条件输入格式错误: {condition.shape}")
        return self._one_hot(condition)

    def generate_conditional_samples(self, num: int, condition: torch.Tensor) -> torch.Tensor:
        onehot = self._process_condition(condition, target_batch_size=num)
        if onehot.size(0) != num:
            raise ValueError(f"条件数量不匹配，期望: {num}，实际: {onehot.size(0)}")
        z = torch.randn(num, self.latent_dim, device=onehot.device)
        return self.forward(z, onehot)

    def forward(self, z: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        if z.dim() != 2:
            raise ValueError(f"潜在向量应为2D张量，实际: {z.shape}")
        if z.size(1) != self.latent_dim:
            raise ValueError(f"潜在维度不匹配，期望: {self.latent_dim}，实际: {z.size(1)}")
        if condition.dim() != 2:
            raise ValueError(f"条件应为2D张量，实际: {condition.shape}")
        if condition.size(1) != self.num_classes:
            raise ValueError(f"条件维度不匹配，期望: {self.num_classes}，实际: {condition.size(1)}")
        if z.size(0) != condition.size(0):
            raise ValueError(f"batch大小不匹配，潜在向量: {z.size(0)}，条件: {condition.size(0)}")
        hid = self.main_model(torch.cat([z, condition], dim=1))
        self.hidden_status = hid
        return self.last_layer(hid).view(-1, self.output_dim)

    def reconstruct(self, x: torch.Tensor, condition: torch.Tensor, encoder: nn.Module) -> torch.Tensor:
        with torch.no_grad():
            z = encoder.encode(x, condition)
            return self.forward(z, self._process_condition(condition, target_batch_size=x.size(0)))


class CVAEGANDiscriminatorModel(_Attached):
    def __init__(self, in_features: int, num_classes: int, hidden=None):
        super().__init__()
        self.in_features, self.num_classes = in_features, num_classes
        tin = in_features + num_classes
        h = _widths(tin, True, hidden)
        sn = lambda i, o: spectral_norm(nn.Linear(i, o))  # noqa: E731
        self.discriminator_network = nn.Sequential(
            sn(tin, h[0]), nn.LeakyReLU(NEG_SLOPE), nn.Dropout(DROP_P),
            sn(h[0], h[1]), nn.LeakyReLU(NEG_SLOPE), nn.Dropout(DROP_P),
            sn(h[1], h[2]), nn.LeakyReLU(NEG_SLOPE),
            sn(h[2], 1),
        )
        self.hidden_status: torch.Tensor = None
        self.apply(init_weights)

    def _process_condition(self, condition: torch.Tensor, target_batch_size: int = None) -> torch.Tensor:
        if condition.dim() == 0:
            condition = condition.unsqueeze(0)
        if condition.dim() == 1:
            condition = condition.long()
            if target_batch_size:
                if condition.size(0) == 1:
                    condition = condition.repeat(target_batch_size)
                elif condition.size(0) != target_batch_size:
                    raise ValueError(f"条件batch大小不匹配，期望: {target_batch_size}，实际: {condition.size(0)}")
        elif condition.dim() == 2 and condition.size(1) == 1:
            condition = condition.squeeze(1).long()
            if target_batch_size and condition.size(0) != target_batch_size:
                raise ValueError(f"条件batch大小不匹配，期望: {target_batch_size}，实际: {condition.size(0)}")
        else:
            raise ValueError(f"条件输入格式错误: {condition.shape}")
        return self._one_hot(condition)

    def _with_condition(self, x, condition):
        if condition is not None:
            onehot = self._process_condition(condition, target_batch_size=x.size(0))
        else:
            onehot = torch.zeros(x.size(0), self.num_classes, device=x.device)
        return torch.cat([x, onehot], dim=1)

    def forward(self, x: torch.Tensor, condition: torch.Tensor = None) -> torch.Tensor:
        if x.dim() > 2:
            x = x.view(x.size(0), -1)
        feats = self.discriminator_network[:-1](self._with_condition(x, condition))
        self.hidden_status = feats
        return self.discriminator_network[-1](feats)

    def get_feature_importance(self, x: torch.Tensor, condition: torch.Tensor = None):
        with torch.no_grad():
            first = self.discriminator_network[0]
            if hasattr(first, 'weight'):
                imp = torch.mean(torch.abs(first.weight.data), dim=0)
                return imp[:self.in_features], imp[self.in_features:]
        return None, None


class CVAEGANClassifierModel(_Attached):
    def __init__(self, in_features: int, num_classes: int, hidden=None):
        super().__init__()
        self.in_features, self.num_classes = in_features, num_classes
        h = _widths(in_features, True, hidden)
        self.classifier_network = nn.Sequential(
            nn.Linear(in_features, h[0]), nn.ReLU(), nn.Dropout(DROP_P),
            nn.Linear(h[0], h[1]), nn.LayerNorm(h[1]), nn.ReLU(), nn.Dropout(DROP_P),
            nn.Linear(h[1], h[2]), nn.ReLU(),
            nn.Linear(h[2], num_classes),
        )
        self.apply(init_weights)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() > 2:
            x = x.view(x.size(0), -1)
        eng = self._engine
        if (eng is not None and not self.training and not torch.is_grad_enabled() and x.is_cuda
                and self.classifier_network[0].weight.data_ptr() == eng.view(self._net, "classifier_network.0.weight").data_ptr()):
            return eng.classifier_forward(x)          # CUDA kernel path (generate_qualified_samples, predict)
        return self.classifier_network(x)             # autograd plumbing for out-of-scope callers

    def get_feature_importance(self, x: torch.Tensor):
        with torch.no_grad():
            first = self.classifier_network[0]
            if hasattr(first, 'weight'):
                return torch.mean(torch.abs(first.weight.data), dim=0)
        return None


# ---------------------------------------------------------------------------------------------------------------------
# Sibling trainer VAE-GAN (/root/reference/src/models/vae_gan_models.py:8-154): the same three stacks WITHOUT the label
# columns - first Linear over input_dim (E, D) / latent_dim (G) - and forwards that take no condition.  Same attribute and
# state_dict key names, same construction order of the layers (so the same seed gives the reference's starting parameters).
# ---------------------------------------------------------------------------------------------------------------------
class VAEGANEncoderModel(CVAEGANEncoderModel):
    def __init__(self, input_dim: int, latent_dim: int = 128, hidden=None):
        super().__init__(input_dim, 0, latent_dim, hidden=hidden)

    def forward(self, x: torch.Tensor) -> tuple:
        if x.dim() != 2:
            raise ValueError(f"输入数据应为2D张量，实际: {x.shape}")
        if x.size(1) != self.input_dim:
            raise ValueError(f"输入特征维度不匹配，期望: {self.input_dim}，实际: {x.size(1)}")
        hid = self.encoder(x)
        return self.fc_mu(hid), self.fc_logvar(hid)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.reparameterize(*self.forward(x))


class VAEGANGeneratorModel(CVAEGANGeneratorModel):
    def __init__(self, latent_dim: int, output_dim: int, hidden=None):
        super().__init__(latent_dim, 0, output_dim, hidden=hidden)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        if z.dim() != 2:
            raise ValueError(f"潜在向量应为2D张量，实际: {z.shape}")
        if z.size(1) != self.latent_dim:
            raise ValueError(f"潜在维度不匹配，期望: {self.latent_dim}，实际: {z.size(1)}")
        hid = self.main_model(z)
        self.hidden_status = hid
        return self.last_layer(hid).view(-1, self.output_dim)

    def reconstruct(self, x: torch.Tensor, encoder: nn.Module) -> torch.Tensor:
        with torch.no_grad():
            return self.forward(encoder.encode(x))


class VAEGANDiscriminatorModel(CVAEGANDiscriminatorModel):
    def __init__(self, in_features: int, hidden=None):
        super().__init__(in_features, 0, hidden=hidden)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() > 2:
            x = x.view(x.size(0), -1)
        feats = self.discriminator_network[:-1](x)
        self.hidden_status = feats
        return self.discriminator_network[-1](feats)

    def get_feature_importance(self, x: torch.Tensor):
        with torch.no_grad():
            first = self.discriminator_network[0]
            if hasattr(first, 'weight'):
                return torch.mean(torch.abs(first.weight.data), dim=0)
        return None

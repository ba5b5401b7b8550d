"""Mirror of /root/reference/src/config/__init__.py:14-23 (seed, device) for the hot path.
Unlike the reference, importing this package has no filesystem side effects."""
import torch

from . import classifier_config, gan_config  # noqa: F401

seed = 0

# the reference resolves 'auto' to cuda-if-available (config/__init__.py:17-23); this build is CUDA only
device: str = 'cuda' if torch.cuda.is_available() else 'cpu'

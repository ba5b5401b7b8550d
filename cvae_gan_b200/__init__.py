"""cvae_gan_b200 - B200-native (sm_100a) CVAE-GAN training + generation/filter hot path.

Drop-in for the reference's `src.CVAEGAN` (src/cvae_gan.py) and the four model classes
(src/models/cvae_gan_models.py), plus the sibling trainers `src.CGAN` (src/cgan.py), `src.CVAE`
(src/cvae.py) and `src.VAEGAN` (src/vae_gan.py).  Everything numerical runs in libcvaegan_b200.so
(hand-written CUDA behind the C ABI in include/cvaegan_b200.h); importing this package never touches a GPU, but
constructing `CVAEGAN` / `Engine` without the built library or without an sm_100 device raises.
"""
from . import config, datasets, models  # noqa: F401
from ._lib import CvgError  # noqa: F401
from .classifier import Classifier  # noqa: F401
from .cvae_gan import CVAEGAN, lambda_class_at  # noqa: F401
from .cgan import CGAN  # noqa: F401
from .cvae import CVAE  # noqa: F401
from .vae_gan import VAEGAN  # noqa: F401
from .engine import Engine, patience_scan  # noqa: F401

__all__ = ["CVAEGAN", "CGAN", "CVAE", "VAEGAN", "Classifier", "Engine", "CvgError", "config", "datasets", "models", "patience_scan", "lambda_class_at"]

#!/usr/bin/env python
"""Counterpart of /root/reference/scripts/train_cgan.py on the B200 engine: the driver of scripts/train_cvae_gan.py (scaling, fit,
class balancing with qualified samples, pickle hand-off, downstream classifier) around the sibling trainer
`cvae_gan_b200.CGAN`.  Same arguments as scripts/train_cvae_gan.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from train_cvae_gan import main  # noqa: E402

if __name__ == "__main__":
    main("cgan")

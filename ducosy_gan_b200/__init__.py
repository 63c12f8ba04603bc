"""ducosy_gan_b200 -- B200-native (sm_100a) hot path of DuCoSy-GAN behind the reference's modules/model.py API.

    from ducosy_gan_b200.modules.model import Generator, Discriminator, weights_init_normal   # drop-in
    from ducosy_gan_b200.synthesis import DualHUSynthesizer                                   # volume API

All arithmetic runs in libducosy_sm100.so (hand-written CUDA, see csrc/); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"

"""mpgan_b200: B200-native (sm_100a) implementation of the MPGAN / GAPT message-passing hot path.

Drop-in modules (same signatures and state_dict layout as rkansal47/MPGAN):
``MPGenerator``, ``MPDiscriminator``, ``MPNet``, ``MPLayer``, ``LinearNet``, ``SpectralNorm``,
``GAPT_G``, ``GAPT_D``, ``MAB``, ``SAB``, ``ISAB``, ``PMA``.
"""
from . import ops  # noqa: F401
from .gapt import GAPT_D, GAPT_G, ISAB, MAB, PMA, SAB  # noqa: F401
from .model import LinearNet, MPDiscriminator, MPGenerator, MPLayer, MPNet  # noqa: F401
from .spectral_normalization import SpectralNorm  # noqa: F401

__version__ = "0.1.0"

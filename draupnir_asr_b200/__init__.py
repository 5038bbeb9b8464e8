"""draupnir_asr_b200 — B200-native (sm_100a) drop-in for DRAUPNIR's tree-OU prior + ELBO likelihood path.

Importing the package loads `libdraupnir_b200.so`; if it is missing the import fails (no CPU / PyTorch fallback).
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)
from . import ops  # noqa: F401
from .backend import pyro, dist, USING_REAL_PYRO  # noqa: F401
from .models import DRAUPNIRModelClass, DRAUPNIRModel_classic, SamplingOutput  # noqa: F401
from .models_utils import (EmbedComplex, EmbedComplexEncoder, GPKernel, OUKernel_Fast, RNNDecoder_Tiling,  # noqa: F401
                           RNNEncoder)
from .guides import DRAUPNIRGUIDES  # noqa: F401
from .train import train  # noqa: F401

__version__ = "0.1.0"

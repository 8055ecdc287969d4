"""vector_quantization_b200 — B200-native (sm_100a) codebook-quantization hot path behind the
registry names / config keys / forward contract of magic-research/vector_quantization
(`vq/algorithms`).  See DESIGN.md and INTEGRATION.md."""
from . import _lib, ops  # noqa: F401

__version__ = '0.1.0'

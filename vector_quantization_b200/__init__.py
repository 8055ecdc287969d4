"""vector_quantization_b200 — B200-native (sm_100a) codebook-quantization hot path behind the
registry names / config keys / forward contract of magic-research/vector_quantization
(`vq/algorithms`).  See DESIGN.md and INTEGRATION.md."""
from . import _lib, ops  # noqa: F401
from . import functional, parallel, registry, tokenizer  # noqa: F401
from .anchors import *  # noqa: F401,F403
from .callbacks import *  # noqa: F401,F403
from .distances import *  # noqa: F401,F403
from .losses import *  # noqa: F401,F403
from .metrics import *  # noqa: F401,F403
from .quantizers import *  # noqa: F401,F403
from .registry import (AnchorRegistry, Config, VQITQuantizerCallbackRegistry,  # noqa: F401
                       VQITQuantizerDistanceRegistry, VQITQuantizerLossRegistry, VQITQuantizerRegistry)

__version__ = '0.1.0'


def build_quantizer(config, training: bool = True):
    """Build a quantizer from a reference-style `quantizer=dict(...)` config node
    (e.g. configs/vqgan/model.py:19-23 + configs/vq/*.py) and run its `init_weights`."""
    config = Config(config)
    init_weights = config.pop('init_weights', None) or Config()
    config['init_weights'] = None          # deferred: the mode must be set before the callbacks' before_init_weights
    q = VQITQuantizerRegistry.build(config)
    q.train(training)
    q.init_weights(Config(init_weights))
    return q

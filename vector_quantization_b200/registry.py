"""Registry / config plumbing of the drop-in layer.

The reference looks quantizer components up BY CLASS NAME in `todd` registry trees and builds them from
PyConfig dicts:
    VQITQuantizerRegistry            vq/tasks/image_tokenization/models/registries.py:20-21
    VQITQuantizerCallbackRegistry    vq/tasks/image_tokenization/models/quantizers/registries.py:9-10
    VQITQuantizerLossRegistry        vq/tasks/image_tokenization/models/quantizers/registries.py:13-14
    VQITQuantizerDistanceRegistry    vq/algorithms/vq/distances.py:17-18
    AnchorRegistry                   vq/algorithms/cvqvae/registries.py:8
`todd` is a third-party package that may be absent, so this module provides standalone registries with
the same names and the same build protocol (`type` key, `build_pre_hook(config, registry, item)`
classmethod, `build_or_return`, `register_(force=)`).  `vector_quantization_b200.plugin` additionally
force-registers the same classes into the reference's own registries when they are importable, which is
what a `custom_imports` entry triggers.
"""
from __future__ import annotations

import functools
import importlib
from typing import Any, Mapping

from torch import nn

__all__ = [
    'Config', 'Registry', 'VQITQuantizerRegistry', 'VQITQuantizerCallbackRegistry', 'VQITQuantizerLossRegistry',
    'VQITQuantizerDistanceRegistry', 'AnchorRegistry', 'ModelRegistry', 'InitRegistry', 'build_module_dict',
]


class Config(dict):
    """Attribute-style dict used for component configs (API subset of todd.Config)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.update(*args, **kwargs)

    @staticmethod
    def _wrap(v):
        if isinstance(v, Mapping) and not isinstance(v, Config):
            return Config(v)
        if isinstance(v, (list, tuple)):
            return type(v)(Config._wrap(i) for i in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v

    def get_config(self, k) -> 'Config':
        return self[k] if k in self and self[k] is not None else Config()


def get_config(config: Mapping, key: str) -> Mapping:
    """`config.get_config(key)` for both our Config and todd's."""
    v = config.get(key)
    return v if v is not None else Config()


class Registry:
    """A named table of classes/functions with todd's build protocol."""

    def __init__(self, name: str, parent: 'Registry | None' = None):
        self.name = name
        self._records: dict[str, Any] = {}
        self._fallbacks: list[Any] = []       # reference registries consulted after ours (plugin mode)
        self.parent = parent

    def register_(self, *names: str, force: bool = False):
        def deco(obj):
            for k in names or (obj.__name__,):
                if k in self._records and not force:
                    raise KeyError(f'{k} is already registered in {self.name}')
                self._records[k] = obj
            return obj

        return deco

    def add_fallback(self, registry: Any) -> None:
        if registry not in self._fallbacks:
            self._fallbacks.append(registry)

    def __contains__(self, key: str) -> bool:
        return key in self._records

    def lookup(self, key: str):
        leaf = key.rsplit('.', 1)[-1]  # dotted todd paths ('VQModelRegistry.X') address the same leaf name
        if leaf in self._records:
            return self._records[leaf]
        for fb in self._fallbacks:
            for attr in ('lookup', '_lookup', 'get'):
                fn = getattr(fb, attr, None)
                if fn is None:
                    continue
                try:
                    item = fn(key)
                except Exception:  # noqa: BLE001 - foreign registry API
                    item = None
                if item is not None:
                    return item
        if leaf.startswith('torch_'):  # todd registers torch modules as e.g. torch_nn_modules_sparse_Embedding
            parts = leaf.split('_')
            for i in range(len(parts) - 1, 0, -1):
                try:
                    return getattr(importlib.import_module('.'.join(parts[:i])), '_'.join(parts[i:]))
                except (ImportError, AttributeError):
                    continue
        raise KeyError(f'{key!r} is not registered in {self.name}')

    def build(self, config: Mapping, **defaults):
        config = Config(config)
        for k, v in defaults.items():
            config.setdefault(k, v)
        type_ = config.pop('type')
        item = self.lookup(type_) if isinstance(type_, str) else type_
        # nn.Module items with an `init_weights` method get it called with the config's `init_weights` node, an
        # EMPTY Config when the node is absent (as todd does: the VQ-KD config has none, configs/vqkd/model.py:20-26,
        # yet its callbacks create the k-means lazy init / `_probability` buffer in before_init_weights);
        # `init_weights=None` defers the call to the caller (vqb.build_quantizer sets the train/eval mode first).
        init_weights = None
        if isinstance(item, type) and issubclass(item, nn.Module) and hasattr(item, 'init_weights'):
            init_weights = config.pop('init_weights', Config())
        hook = getattr(item, 'build_pre_hook', None)
        if hook is not None:
            config = hook(config, self, item)
        obj = item(**config)
        if init_weights is not None:
            obj.init_weights(Config(init_weights))
        return obj

    def build_or_return(self, x, **defaults):
        return self.build(x, **defaults) if isinstance(x, Mapping) else x


class _InitRegistry(Registry):
    """`InitRegistry.build(Config(type='uniform_', a=.., b=..))` -> partial(nn.init.uniform_, a=.., b=..)."""

    def build(self, config: Mapping, **defaults):
        config = Config(config)
        type_ = config.pop('type')
        if type_ in self._records:
            return functools.partial(self._records[type_], **config)
        return functools.partial(getattr(nn.init, type_), **config)


ModelRegistry = Registry('ModelRegistry')
VQITQuantizerRegistry = Registry('VQITQuantizerRegistry', ModelRegistry)
VQITQuantizerCallbackRegistry = Registry('VQITQuantizerCallbackRegistry', VQITQuantizerRegistry)
VQITQuantizerLossRegistry = Registry('VQITQuantizerLossRegistry', VQITQuantizerRegistry)
VQITQuantizerDistanceRegistry = Registry('VQITQuantizerDistanceRegistry', VQITQuantizerRegistry)
AnchorRegistry = Registry('AnchorRegistry')
InitRegistry = _InitRegistry('InitRegistry')


class ModuleDict(nn.ModuleDict):
    """nn.ModuleDict whose call returns {name: module(*args)} (todd.patches.torch.ModuleDict)."""

    def forward(self, *args, **kwargs) -> dict:
        return {k: m(*args, **kwargs) for k, m in self.items()}


def build_module_dict(registry: Registry, config: Mapping, **kwargs) -> ModuleDict:
    """vq/utils/builders.py:24-34 equivalent."""
    return ModuleDict({k: registry.build_or_return(v, **kwargs) for k, v in config.items() if v is not None})

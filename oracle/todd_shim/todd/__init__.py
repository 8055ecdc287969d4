"""Minimal stand-in for the third-party `todd` (todd_ai) package.  TEST INFRASTRUCTURE ONLY.

The reference pins `todd_ai @ git+https://github.com/LutingWang/todd.git@ed2a3ae75a66…`
(/root/reference `.todd_version:1`, `setup.py:6-14`); it is neither vendored in /root/reference nor
installable offline.  This shim provides just enough of its *plumbing* surface (Config, Registry,
BuildPreHookMixin, ModuleDict, HolderMixin, PriorityQueue, Store, dist helpers) for the reference's
own hot-path source files to be imported UNMODIFIED from /root/reference by `oracle/make_golden.py`.

Arithmetic that lives in todd — `todd.utils.EMA/ema` and `todd.models.losses.MSELoss` — is restated
from its published semantics and is UNVERIFIED against the real package: parity is
"unpinned" at exactly this boundary.  Nothing under `vector_quantization_b200/` imports this package;
if the real `todd` is installed, it takes precedence and this shim is not put on sys.path.
"""
from __future__ import annotations

import functools
import importlib
import logging
import sys
import types
from typing import Any, Generic, Iterable, Mapping, TypeVar

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

__shim__ = True
logger = logging.getLogger('todd_shim')


# ---------------------------------------------------------------------------------------------
# Config
# ---------------------------------------------------------------------------------------------
class Config(dict):
    """Attribute-access dict; nested mappings become Configs."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, Config):
            return v
        if isinstance(v, Mapping):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(i) for i in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def get_config(self, k):
        return self.get(k) if k in self else Config()

    def copy(self):
        return Config(self)


# ---------------------------------------------------------------------------------------------
# Registry
# ---------------------------------------------------------------------------------------------
class Item:  # todd.bases.registries.Item is only used as a type annotation by the reference
    pass


class RegistryMeta(type):
    """Class-style registries: subclassing a registry creates a child; lookup walks down."""

    def __init__(cls, name, bases, ns):
        super().__init__(name, bases, ns)
        cls._records = {}
        cls._children = {}
        for b in bases:
            if isinstance(b, RegistryMeta):
                b._children[name] = cls

    def register_(cls, *names, force: bool = False):
        def deco(obj):
            keys = names or (obj.__name__,)
            for k in keys:
                if k in cls._records and not force:
                    raise KeyError(f'{k} already registered in {cls.__name__}')
                cls._records[k] = obj
            return obj

        return deco

    def _lookup(cls, key: str):
        if '.' in key:
            head, rest = key.split('.', 1)
            if head == cls.__name__:
                return cls._lookup(rest)
            if head in cls._children:
                return cls._children[head]._lookup(rest)
        if key in cls._records:
            return cls._records[key]
        for child in cls._children.values():
            try:
                return child._lookup(key)
            except KeyError:
                pass
        raise KeyError(key)

    def lookup(cls, key: str):
        try:
            return cls._lookup(key)
        except KeyError:
            if key.startswith('torch_'):  # todd registers torch modules as e.g. torch_nn_modules_sparse_Embedding
                parts = key.split('_')
                for i in range(len(parts) - 1, 0, -1):
                    try:
                        mod = importlib.import_module('.'.join(parts[:i]))
                        return getattr(mod, '_'.join(parts[i:]))
                    except (ImportError, AttributeError):
                        continue
            raise KeyError(f'{key} not found in {cls.__name__}') from None

    def _build(cls, item, config: Config):
        return item(**config)

    def build(cls, config, **defaults):
        config = Config(config)
        for k, v in defaults.items():
            config.setdefault(k, v)
        type_ = config.pop('type')
        item = cls.lookup(type_) if isinstance(type_, str) else type_
        init_weights = config.pop('init_weights', None) if cls._pops_init_weights(item) else None
        hook = getattr(item, 'build_pre_hook', None)
        if hook is not None:
            config = hook(config, cls, item)
        obj = cls._build(item, config)
        if init_weights is not None and hasattr(obj, 'init_weights'):
            obj.init_weights(Config(init_weights))
        return obj

    @staticmethod
    def _pops_init_weights(item) -> bool:
        return isinstance(item, type) and issubclass(item, nn.Module)

    def build_or_return(cls, x, **defaults):
        if isinstance(x, Mapping):
            return cls.build(x, **defaults)
        return x


class Registry(metaclass=RegistryMeta):
    pass


class BuildPreHookMixin:

    @classmethod
    def build_pre_hook(cls, config: Config, registry: RegistryMeta, item) -> Config:
        return config


# ---------------------------------------------------------------------------------------------
# stores / misc
# ---------------------------------------------------------------------------------------------
class _Store:
    DRY_RUN = False
    cuda = False
    PRETRAINED = ''
    DEBUG = False


Store = _Store()

T = TypeVar('T')


class HolderMixin(Generic[T]):

    def __init__(self, *args, instance=None, **kwargs):
        super().__init__(*args, **kwargs)
        if instance is not None:
            self._instance = instance

    def bind(self, instance) -> None:
        self._instance = instance


def ema(x, y, decay):
    """UNVERIFIED restatement of todd.utils.ema: x * decay + y * (1 - decay)."""
    return x * decay + y * (1 - decay)


class EMA:
    """UNVERIFIED restatement of todd.utils.EMA (default decay assumed 0.99)."""

    def __init__(self, decay: float = 0.99):
        self._decay = decay

    @property
    def decay(self):
        return self._decay

    def __call__(self, x, y):
        return y if x is None else ema(x, y, self._decay)


def is_sync(x: torch.Tensor) -> bool:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
        return True
    xs = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(xs, x)
    return all(torch.equal(xs[0], t) for t in xs)


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def all_gather(x: torch.Tensor):
    if get_world_size() <= 1:
        return [x]
    xs = [torch.empty_like(x) for _ in range(get_world_size())]
    dist.all_gather(xs, x.contiguous())
    return xs


class ModuleDict(nn.ModuleDict):

    def forward(self, *args, **kwargs) -> dict:
        return {k: m(*args, **kwargs) for k, m in self.items()}


class ModuleList(nn.ModuleList):

    def forward(self, *args, **kwargs) -> list:
        return [m(*args, **kwargs) for m in self]


class Sequential(nn.Sequential):
    pass


class PriorityQueue:
    """UNVERIFIED: `queue(key)` yields items ordered by their priority for `key` (default 0), stable."""

    def __init__(self, priorities: Iterable[Mapping[str, int]], items: Iterable[Any]):
        self._priorities = list(priorities)
        self._items = list(items)

    def __call__(self, key: str):
        order = sorted(range(len(self._items)), key=lambda i: -self._priorities[i].get(key, 0))
        return [self._items[i] for i in order]


class BaseLoss(BuildPreHookMixin, nn.Module):
    """UNVERIFIED restatement of todd.models.losses.BaseLoss (constant weight, mean reduction)."""

    def __init__(self, *args, reduction: str = 'mean', weight: float = 1.0, **kwargs):
        super().__init__(*args, **kwargs)
        self._reduction = reduction
        self._loss_weight = weight

    def _reduce(self, loss: torch.Tensor) -> torch.Tensor:
        if self._reduction == 'mean':
            loss = loss.mean()
        elif self._reduction == 'sum':
            loss = loss.sum()
        return loss * self._loss_weight if self._loss_weight != 1.0 else loss


class MSELoss(BaseLoss):
    """UNVERIFIED restatement of todd.models.losses.MSELoss(norm=False)."""

    def __init__(self, *args, norm: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self._norm = norm

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self._norm:
            pred = F.normalize(pred)
            target = F.normalize(target)
        return self._reduce(F.mse_loss(pred, target, reduction='none'))


class BaseMetric(Generic[T]):

    def __init__(self, *args, **kwargs):
        pass


# ---------------------------------------------------------------------------------------------
# registries that exist in todd
# ---------------------------------------------------------------------------------------------
class ModelRegistry(Registry):
    pass


class DatasetRegistry(Registry):
    pass


class RunnerRegistry(Registry):
    pass


class TaskRegistry(Registry):
    pass


class LossRegistry(ModelRegistry):
    pass


class EnvRegistry(Registry):
    pass


class InitRegistry(Registry):

    @classmethod
    def build(cls, config, **defaults):  # e.g. Config(type='uniform_', a=.., b=..) -> partial(nn.init.uniform_, ..)
        config = Config(config)
        type_ = config.pop('type')
        return functools.partial(getattr(nn.init, type_), **config)


class PyConfig(Config):

    @classmethod
    def load(cls, *args, **kwargs):
        raise NotImplementedError('PyConfig.load is outside the oracle shim')


# ---------------------------------------------------------------------------------------------
# expose the sub-module layout the reference imports from
# ---------------------------------------------------------------------------------------------
def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition('.')
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


_module('todd.bases')
_module('todd.bases.registries', BuildPreHookMixin=BuildPreHookMixin, Item=Item)
_module('todd.bases.registries.base', BuildPreHookMixin=BuildPreHookMixin, Item=Item)
_module('todd.patches')
_module('todd.patches.torch', ModuleDict=ModuleDict, ModuleList=ModuleList, Sequential=Sequential,
        get_rank=get_rank, get_world_size=get_world_size, all_gather=all_gather,
        load_state_dict=lambda m, sd, *a, **k: m.load_state_dict(sd, *a, **k),
        load_state_dict_=lambda fs: {k: v for f in fs for k, v in torch.load(f, 'cpu').items()})
_module('todd.patches.py_', get_=lambda obj, path: eval('obj' + path, {'obj': obj}))  # noqa: S307
_module('todd.runners', Memo=dict)
_module('todd.runners.utils', PriorityQueue=PriorityQueue)
_module('todd.runners.metrics', BaseMetric=BaseMetric)
_module('todd.models')
_module('todd.models.losses', BaseLoss=BaseLoss, MSELoss=MSELoss)
sys.modules['todd.models'].losses = sys.modules['todd.models.losses']
_module('todd.utils', EMA=EMA, ema=ema, is_sync=is_sync, HolderMixin=HolderMixin, EnvRegistry=EnvRegistry,
        todd_version=lambda: 'shim')
_module('todd.registries', ModelRegistry=ModelRegistry, DatasetRegistry=DatasetRegistry,
        RunnerRegistry=RunnerRegistry, TaskRegistry=TaskRegistry, InitRegistry=InitRegistry,
        LossRegistry=LossRegistry)
_module('todd.configs', PyConfig=PyConfig)

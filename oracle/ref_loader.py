"""Import the reference's OWN hot-path source files, unmodified, from /root/reference.
TEST INFRASTRUCTURE ONLY (used by `oracle/make_golden.py` and `tests/test_oracle_vs_reference.py`
in the dev container; /root/reference does not exist on the GPU box).

`import vq` cannot run as-is: `vq/__init__.py:3` eagerly imports runners/datasets/algorithms that need
most of the third-party `todd` package (absent).  We therefore
  1. put the minimal `todd` shim (`oracle/todd_shim`) on sys.path unless a real `todd` is importable;
  2. register *bare* packages (`__path__` only, `__init__` NOT executed) for `vq`, `vq.utils`,
     `vq.models`, `vq.tasks`, `vq.tasks.image_tokenization`, `...image_tokenization.models`,
     `vq.algorithms`, `vq.algorithms.vqkd`, `vq.algorithms.vqgan`;
  3. execute the reference's registry files and `vq/utils/{builders,misc}.py` into those packages;
  4. import the real hot-path packages: `...models.quantizers`, `vq.algorithms.{vq,sq,fsq,cvqvae}`,
     `vq.algorithms.vqkd.quantizers`, `vq.algorithms.vqgan.quantizer`.
No reference source is copied into this repository.
"""
from __future__ import annotations

import importlib
import importlib.util
import pathlib
import sys
import types

REF = pathlib.Path('/root/reference')
_loaded = None


def available() -> bool:
    return (REF / 'vq' / 'algorithms' / 'vq' / 'quantizers.py').exists()


def _bare(name: str, path: pathlib.Path) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [str(path)]
    m.__package__ = name
    sys.modules[name] = m
    parent, _, leaf = name.rpartition('.')
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


def _exec_into(pkg: types.ModuleType, file: pathlib.Path, modname: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(modname, file)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    for k in getattr(mod, '__all__', []):
        setattr(pkg, k, getattr(mod, k))
    setattr(pkg, modname.rpartition('.')[2], mod)
    return mod


def load():
    """Returns a namespace with the reference's registries and classes."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('/root/reference is not present (dev container only)')
    try:
        import todd  # noqa: F401
    except ImportError:
        sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent / 'todd_shim'))
        import todd  # noqa: F401
    import torch
    if not hasattr(torch.serialization, 'FILE_LIKE'):  # annotation-only name removed in recent torch (misc.py:67)
        torch.serialization.FILE_LIKE = object
    vq = _bare('vq', REF / 'vq')
    _exec_into(vq, REF / 'vq' / 'registries.py', 'vq.registries')
    utils = _bare('vq.utils', REF / 'vq' / 'utils')
    _exec_into(utils, REF / 'vq' / 'utils' / 'builders.py', 'vq.utils.builders')
    _exec_into(utils, REF / 'vq' / 'utils' / 'misc.py', 'vq.utils.misc')
    models = _bare('vq.models', REF / 'vq' / 'models')
    import todd.models as tm
    if not hasattr(tm, 'LossRegistry'):
        tm.LossRegistry = sys.modules['todd.registries'].LossRegistry
    _exec_into(models, REF / 'vq' / 'models' / 'registries.py', 'vq.models.registries')
    tasks = _bare('vq.tasks', REF / 'vq' / 'tasks')
    _exec_into(tasks, REF / 'vq' / 'tasks' / 'registries.py', 'vq.tasks.registries')
    it = _bare('vq.tasks.image_tokenization', REF / 'vq' / 'tasks' / 'image_tokenization')
    _exec_into(it, REF / 'vq' / 'tasks' / 'image_tokenization' / 'registries.py',
               'vq.tasks.image_tokenization.registries')
    itm = _bare('vq.tasks.image_tokenization.models', REF / 'vq' / 'tasks' / 'image_tokenization' / 'models')
    _exec_into(itm, REF / 'vq' / 'tasks' / 'image_tokenization' / 'models' / 'registries.py',
               'vq.tasks.image_tokenization.models.registries')
    quantizers = importlib.import_module('vq.tasks.image_tokenization.models.quantizers')
    _bare('vq.algorithms', REF / 'vq' / 'algorithms')
    a_vq = importlib.import_module('vq.algorithms.vq')
    a_sq = importlib.import_module('vq.algorithms.sq')
    a_fsq = importlib.import_module('vq.algorithms.fsq')
    a_cvq = importlib.import_module('vq.algorithms.cvqvae')
    _bare('vq.algorithms.vqkd', REF / 'vq' / 'algorithms' / 'vqkd')
    a_vqkd = importlib.import_module('vq.algorithms.vqkd.quantizers')
    _bare('vq.algorithms.vqgan', REF / 'vq' / 'algorithms' / 'vqgan')
    a_vqgan = importlib.import_module('vq.algorithms.vqgan.quantizer')
    import todd as todd_mod
    _loaded = types.SimpleNamespace(
        todd=todd_mod, todd_is_shim=bool(getattr(todd_mod, '__shim__', False)), quantizers=quantizers,
        vq=a_vq, sq=a_sq, fsq=a_fsq, cvqvae=a_cvq, vqkd=a_vqkd, vqgan=a_vqgan,
        VQITQuantizerRegistry=itm.VQITQuantizerRegistry)
    return _loaded


def build_quantizer(config: dict, training: bool = True):
    """Build a reference quantizer from a reference-style config dict (the `quantizer=dict(...)` node of
    configs/vqgan/model.py:19-23 etc.) exactly as todd's registry would."""
    ref = load()
    cfg = ref.todd.Config(config)
    init_weights = cfg.pop('init_weights', None)
    q = ref.VQITQuantizerRegistry.build(cfg)
    q.train(training)
    if init_weights is not None:
        q.init_weights(ref.todd.Config(init_weights))
    else:
        q.init_weights(ref.todd.Config())
    return q

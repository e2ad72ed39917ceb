"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference from /root/reference.

Works only in the build container (the GPU box has no /root/reference); used by
oracle/make_golden.py to generate tests/golden/* and by tests that are skipped
when the reference tree is absent. Nothing in the product path, `-m gpu` tests,
smoke() or bench.py imports this module.

The reference needs two import roots: `/root/reference`
(for `from python.commons import *`, TrainingNeural.py:25) and
`/root/reference/python` (notebook spelling). We import under isolated names so
the repo's own mirror modules (`Training.TrainingNeural`, ...) are never shadowed.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GMC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "python", "Training", "TrainingNeural.py"))


class _Ref(types.SimpleNamespace):
    pass


_cached = None


def load() -> _Ref:
    """Import commons / GraphCreator / graphExtender / TrainingNeural /
    TestingNeuralNetwork / RandomizedMaxCut of the reference, unmodified."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    from . import dgl_shim
    dgl_shim.install()

    # The reference does `from python.commons import *`; give it a private
    # `python` package that points at the reference tree and is removed again
    # afterwards so it cannot leak into the product's import namespace.
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "python" or k.startswith("python.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        commons = importlib.import_module("python.commons")
        training = importlib.import_module("python.Training.TrainingNeural")
        extender = importlib.import_module("python.DataGenerator.graphExtender")
        creator = importlib.import_module("python.DataGenerator.GraphCreator")
        testing = importlib.import_module("python.Testing.TestingNeuralNetwork")
        randomized = importlib.import_module("python.RandomAlgorithm.RandomizedMaxCut")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "python" or k.startswith("python.")]:
            mod = sys.modules.pop(k)
            sys.modules["_gmc_reference." + k] = mod
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _cached = _Ref(commons=commons, training=training, extender=extender, creator=creator,
                   testing=testing, randomized=randomized)
    return _cached


@contextlib.contextmanager
def active():
    """Temporarily expose the reference under its real module names
    (`python.Training.TrainingNeural`, ...) -- needed while it pickles its own
    `TrainingConfig` (torch.save in TrainingNeural.py:447-482)."""
    load()
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "python" or k.startswith("python.")}
    mine = {k[len("_gmc_reference."):]: v for k, v in sys.modules.items()
            if k.startswith("_gmc_reference.")}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(mine)
    try:
        yield _cached
    finally:
        for k in mine:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v

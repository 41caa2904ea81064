"""Import the UNMODIFIED reference (TuringQ/deepquantum) from /root/reference/src.

TEST INFRASTRUCTURE ONLY.  This module exists so that golden fixtures can be generated in the
build container, where /root/reference is mounted read-only.  Nothing in the product package
(`deepquantum_b200/`), in the `-m gpu` tests, in `smoke()` or in `bench.py` imports it: the
reference tree does not exist on the GPU box.

Four cosmetic third-party modules that the reference imports at module scope (only for drawing /
Bayesian optimisation, never on the statevector path) are not installed in this image; they are
replaced by empty stub modules before the import.
"""
import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get('B200Q_REFERENCE_SRC', '/root/reference/src')


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'deepquantum'))


def load_reference():
    """Return the reference `deepquantum` module (raises if the tree is not mounted)."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_SRC}')
    if 'deepquantum' in sys.modules:
        return sys.modules['deepquantum']

    class _Dummy:  # placeholder class for `from x import Y` statements
        def __init__(self, *a, **k):
            raise RuntimeError('stubbed cosmetic dependency')

    for name in ('qiskit', 'svgwrite', 'bayes_opt'):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name, QuantumCircuit=_Dummy, BayesianOptimization=_Dummy, UtilityFunction=_Dummy)
    try:
        importlib.import_module('matplotlib.pyplot')
    except Exception:
        cm = _stub('matplotlib.cm')
        patches = _stub('matplotlib.patches')
        pyplot = _stub('matplotlib.pyplot')
        _stub('matplotlib', cm=cm, patches=patches, pyplot=pyplot)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    return importlib.import_module('deepquantum')

"""TEST INFRASTRUCTURE ONLY — imports the real reference (mit-han-lab/torchquantum) from
/root/reference so that golden vectors can be generated and the oracle port can be pinned.

The reference imports qiskit/torchpack/matplotlib/... at package-import time but never touches
them on the statevector path; none of them is installed here, so this module fabricates those
top-level packages (a meta-path finder that hands out MagicMocks) and then repairs the
undefined-global ``mat_dict`` bug that makes nine functional wrappers raise NameError at this
commit (functional/hadamard.py:163, paulix.py:511,555,599, single_excitation.py:88,
u1.py:113,160, u2.py:112,159).

Nothing on the product path, in bench.py or in the ``-m gpu`` tests may import this: the
reference tree does not exist on the GPU box.  ``load_reference()`` returns ``None`` when
/root/reference is absent.
"""
import importlib
import importlib.abc
import importlib.machinery
import logging
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("TQ_REFERENCE_ROOT", "/root/reference")

_FAKE_ROOTS = {
    "qiskit", "qiskit_aer", "qiskit_ibm_runtime", "qiskit_ibm_provider", "torchpack",
    "opt_einsum", "pathos", "pylatexenc", "matplotlib", "torchdiffeq", "pyscf",
    "tensorflow", "qutip",
}


class _FakeModule(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith("__") and key.endswith("__"):
            raise AttributeError(key)
        value = MagicMock(name=f"{self.__name__}.{key}")
        setattr(self, key, value)
        return value


class _FakeFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _FAKE_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _FakeModule(spec.name)
        mod.__path__ = []
        mod.__file__ = "/nonexistent/" + spec.name.replace(".", "/") + "/__init__.py"
        return mod

    def exec_module(self, mod):
        if mod.__name__ == "torchpack.utils.logging":
            mod.logger = logging.getLogger("tq-reference")
        if mod.__name__ == "torchpack.utils.config":
            mod.Config = type("Config", (dict,), {})


_loaded = None


def load_reference():
    """Return (tq, tqf) of the real reference, or None if the tree is not mounted."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "torchquantum")):
        return None
    import warnings

    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if not any(isinstance(f, _FakeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _FakeFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torchquantum as tq
    import torchquantum.functional as tqf

    for name in ("hadamard", "paulix", "single_excitation", "u1", "u2"):
        importlib.import_module("torchquantum.functional." + name).mat_dict = tqf.mat_dict
    _loaded = (tq, tqf)
    return _loaded

"""The reference's OWN qiskit-free tests, executed against this package (VERDICT r1 "missing" #5).

Each listed file of /root/reference/test is loaded from where it lies (nothing is copied into the repo) and run with
`torchquantum` resolving to `torchquantum_b200`, so the literal state vectors / matrices / groupings those tests hold
(test/functional/test_controlled_unitary.py:30-764, test/operator/test_ControlledU.py, test/functional/
test_func_mat_exp.py:30-95, test/algorithm/test_hamiltonian.py, test/device/test_init.py:32-43, test/operator/
test_op_hamil_exp.py:32-108, test/measurement/test_find_observable_groups.py, test/measurement/
test_expval_joint_sampling_grouping.py) are checked against our implementation as they stand.  The reference tests use
CPU tensors: they reach the kernels through the host-emulation build of the kernel source (tests/emu).  Skipped where
the reference tree is not mounted (the GPU box); the same behaviours are pinned there by tests/golden/reftests.npz
(tests/cases_api.py::check_reference_test_vectors)."""
import importlib.util
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from tests.test_cpu_emulated import emulated  # noqa: F401  (autouse fixture: host emulation build)

REF_TESTS = "/root/reference/test"
FILES = [
    "device/test_init.py",
    "functional/test_controlled_unitary.py",
    "functional/test_func_mat_exp.py",
    "operator/test_ControlledU.py",
    "operator/test_op_hamil_exp.py",
    "algorithm/test_hamiltonian.py",
    "measurement/test_find_observable_groups.py",
    "measurement/test_expval_joint_sampling_grouping.py",
]


def _alias_package():
    """`import torchquantum...` inside the reference tests resolves to this package"""
    import torchquantum_b200 as tq
    import torchquantum_b200.algorithm
    import torchquantum_b200.functional
    import torchquantum_b200.macro
    import torchquantum_b200.measurement
    import torchquantum_b200.operator

    saved = {k: sys.modules[k] for k in list(sys.modules) if k == "torchquantum" or k.startswith("torchquantum.")
             or k == "test" or k.startswith("test.")}
    for k in saved:   # (the live-reference tests of test_oracle_golden.py may have imported the real package)
        del sys.modules[k]
    sys.modules["torchquantum"] = tq
    for name, mod in list(sys.modules.items()):
        if name.startswith("torchquantum_b200."):
            sys.modules["torchquantum." + name[len("torchquantum_b200."):]] = mod
    pkg = types.ModuleType("test")
    pkg.__path__ = [REF_TESTS]
    sys.modules["test"] = pkg
    spec = importlib.util.spec_from_file_location("test.utils", os.path.join(REF_TESTS, "utils.py"))
    utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(utils)
    sys.modules["test.utils"] = utils
    return saved


def _restore(saved):
    for k in [k for k in sys.modules if k == "torchquantum" or k.startswith("torchquantum.") or k == "test" or k.startswith("test.")]:
        del sys.modules[k]
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v


@pytest.mark.parametrize("rel", FILES)
def test_reference_test_file_passes_against_this_package(rel):
    path = os.path.join(REF_TESTS, rel)
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    saved = _alias_package()
    cwd = os.getcwd()
    try:
        os.chdir(os.path.dirname(REF_TESTS))   # the reference tests open their data files relative to the repo root
        random.seed(0)
        np.random.seed(0)
        torch.manual_seed(0)
        spec = importlib.util.spec_from_file_location("reftest_" + rel.replace("/", "_")[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ran = 0
        for name in dir(mod):
            fn = getattr(mod, name)
            if name.startswith("test") and callable(fn) and getattr(fn, "__module__", None) == mod.__name__:
                fn()
                ran += 1
        assert ran >= 1, f"no test function found in {rel}"
    finally:
        os.chdir(cwd)
        _restore(saved)

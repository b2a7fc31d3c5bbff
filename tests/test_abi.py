"""The C-ABI library: exports every symbol include/tqb200.h declares; the Python constants mirror
the wire-format header; the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tqb200.h")).read()
    return sorted(set(re.findall(r"\b(tqb_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from torchquantum_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(_native.LIB_PATH)  # loads without a GPU (no compute calls here)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tqb200.h but not exported"
    # every entry point is a torch op, except the two introspection calls and the workspace-size query of the GEMM
    assert set(_native.OP_NAMES) | {"tqb_abi_version", "tqb_last_error", "tqb_gemm_workspace_floats"} == set(syms)
    lib.tqb_abi_version.restype = ctypes.c_int
    assert lib.tqb_abi_version() == _native.ABI_VERSION


def test_every_abi_entry_is_a_registered_torch_op():
    """north_star / SURVEY §8b: the front end reaches the kernels through torch.library custom ops registered by the
    thin C++ extension (csrc/tqb_torch_ops.cpp); schema check only, no compute."""
    from torchquantum_b200 import _native

    ops = _native._load_ops()
    assert set(ops) == set(_native.OP_NAMES)
    for name in _native.OP_NAMES:
        schema = getattr(torch.ops.tqb200, name).default._schema
        assert schema.name == f"tqb200::{name}"
        assert any(a.alias_info is not None and a.alias_info.is_write for a in schema.arguments), f"{name}: no output argument"
        keys = torch._C._dispatch_dump(f"tqb200::{name}")
        if name != "tqb_host_uniforms":   # the one HOST helper of the ABI (Python's Mersenne Twister stream)
            assert "CUDA" in keys, f"{name}: no CUDA kernel registered"


def test_python_constants_mirror_header():
    from torchquantum_b200 import _codes

    hdr = open(os.path.join(ROOT, "torchquantum_b200", "csrc", "tqb_program.h")).read()
    for name, val in re.findall(r"#define TQB_(\w+) (0x[0-9a-fA-F]+|\d+)\b", hdr):
        assert getattr(_codes, name) == int(val, 0), name
    enum = re.sub(r"//.*", "", re.search(r"enum \{(.*?)\};", hdr, re.S).group(1))
    for name, val in re.findall(r"TQB_(K_\w+) = (\d+)", enum):
        assert getattr(_codes, name) == int(val), name


def test_no_cpu_fallback():
    """A CPU tensor must be refused loudly by the product path."""
    import torchquantum_b200 as tq
    from torchquantum_b200 import _native

    _native.use_library_for_testing(None)
    q = tq.QuantumDevice(n_wires=2, bsz=1, device="cpu")
    q.h(wires=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q.get_states_1d()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "torchquantum_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"

"""Binding of libtqb200.so (the C ABI in include/tqb200.h) as torch.library custom ops.

`lib/libtqb200_torch.so` (csrc/tqb_torch_ops.cpp, a thin C++ extension) registers one op per ABI entry point in
the `tqb200` namespace — `torch.ops.tqb200.tqb_pass_forward(src, dst, n, bsz, prog, desc, msh, mps, stride, ext_hi)` —
and forwards to the C ABI with raw pointers and the current CUDA stream.  `call(name, *args)` below is how the
package reaches them: the argument order of include/tqb200.h, Tensors for device pointers, int lists / CPU tensors
for host arrays, no stream argument.

There is no CPU fallback: if either library is missing or does not load, every entry point raises, and a CPU
tensor raises inside the op.  (The CPU test-suite injects a host *emulation* build of the same kernel source
through ``use_library_for_testing``; the package itself never looks for it.)
"""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TQB200_LIB") or os.path.join(_HERE, "lib", "libtqb200.so")  # env: kernel A/B experiments
OPS_LIB_PATH = os.path.join(_HERE, "lib", "libtqb200_torch.so")
ABI_VERSION = 7

# every ABI entry point registered as torch.ops.tqb200.<name> (mirrors include/tqb200.h)
OP_NAMES = (
    "tqb_build_mats", "tqb_build_mats_bwd", "tqb_pass_forward", "tqb_pass_backward", "tqb_apply_dense",
    "tqb_apply_dense_tc", "tqb_block_build", "tqb_apply_blocks", "tqb_quadphase", "tqb_dense_grad", "tqb_expval_pauli",
    "tqb_expval_pauli_bwd", "tqb_expval_z", "tqb_expval_z_bwd", "tqb_probs", "tqb_sample_counts", "tqb_host_uniforms",
    "tqb_exchange_p2p", "tqb_split_tf32", "tqb_gemm_prep", "tqb_gemm_tf32x3",
)

_ops = None          # name -> OpOverload
_cuda_bound = False
_emulated = False


def _load_ops():
    global _ops
    if _ops is None:
        if not os.path.exists(OPS_LIB_PATH):
            raise RuntimeError(
                f"torchquantum_b200: op library {OPS_LIB_PATH} not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        torch.ops.load_library(OPS_LIB_PATH)
        _ops = {name: getattr(torch.ops.tqb200, name).default for name in OP_NAMES}
    return _ops


def lib():
    """Loads both libraries (once) and binds the CUDA one; -> the op table."""
    global _cuda_bound
    ops = _load_ops()
    if not _cuda_bound and not _emulated:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"torchquantum_b200: CUDA library {LIB_PATH} not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        version = torch.ops.tqb200._bind(LIB_PATH, False)
        if version != ABI_VERSION:
            raise RuntimeError("libtqb200: ABI version mismatch")
        _cuda_bound = True
    return ops


def abi_version():
    lib()
    return ABI_VERSION


def use_library_for_testing(path):
    """Tests only: route CPU tensors to another build of the ABI (the host emulation of the kernel source)."""
    global _emulated
    _load_ops()
    if path is None:
        torch.ops.tqb200._bind("", True)
        _emulated = False
        return
    torch.ops.tqb200._bind(path, True)
    _emulated = True


def is_emulated():
    return _emulated


def call(name, *args):
    ops = _ops if (_cuda_bound or _emulated) else lib()
    ops[name](*args)


def check_device(t):
    if t.is_cuda:
        if _emulated:
            raise RuntimeError("emulation library active but tensor is on CUDA")
        return
    if not _emulated:
        raise RuntimeError(
            "torchquantum_b200 runs on CUDA (sm_100a) only: got a CPU tensor and there is no CPU fallback")

"""ctypes binding of libtqb200.so (the C ABI in include/tqb200.h).

There is no CPU fallback: if the CUDA library is missing or does not load, every entry point
raises.  (The CPU test-suite injects a host *emulation* build of the same kernel source through
``use_library_for_testing``; the package itself never looks for it.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TQB200_LIB") or os.path.join(_HERE, "lib", "libtqb200.so")  # env: kernel A/B experiments
ABI_VERSION = 4

_lib = None
_emulated = False

_vp, _i, _ll, _u = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_uint

# name -> argtypes (all return int); mirrors include/tqb200.h
SIGNATURES = {
    "tqb_build_mats": [_vp, _i, _i, _ll, _vp, _vp, _i, _vp, _vp, _vp, _ll, _vp],
    "tqb_build_mats_bwd": [_vp, _i, _ll, _vp, _vp, _i, _vp, _ll, _vp, _i, _vp, _vp],
    "tqb_pass_forward": [_vp, _vp, _i, _ll, _vp, _vp, _vp, _vp, _ll, _u, _vp],
    "tqb_pass_backward": [_vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp, _vp, _ll, _vp, _ll, _u, _vp],
    "tqb_apply_dense": [_vp, _vp, _i, _ll, _vp, _ll, _i, _vp, _i, _vp],
    "tqb_apply_dense_tc": [_vp, _vp, _i, _ll, _vp, _i, _vp, _i, _vp],
    "tqb_block_build": [_vp, _i, _vp, _vp, _vp, _u, _vp, _vp],
    "tqb_apply_blocks": [_vp, _vp, _i, _ll, _vp, _vp, _i, _i, _vp],
    "tqb_dense_grad": [_vp, _vp, _i, _ll, _i, _vp, _vp, _vp],
    "tqb_expval_pauli": [_vp, _i, _ll, _vp, _vp, _vp, _i, _vp, _vp],
    "tqb_expval_pauli_bwd": [_vp, _i, _ll, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "tqb_expval_z": [_vp, _i, _ll, _vp, _vp],
    "tqb_expval_z_bwd": [_vp, _i, _ll, _vp, _vp, _vp],
    "tqb_probs": [_vp, _i, _ll, _vp, _vp],
    "tqb_sample_counts": [_vp, _i, _ll, _vp, _i, _vp, _vp, _vp],
}


def _bind(handle):
    handle.tqb_abi_version.restype = ctypes.c_int
    handle.tqb_last_error.restype = ctypes.c_char_p
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    if handle.tqb_abi_version() != ABI_VERSION:
        raise RuntimeError("libtqb200: ABI version mismatch")
    return handle


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"torchquantum_b200: CUDA library {LIB_PATH} not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        _lib = _bind(ctypes.CDLL(LIB_PATH))
    return _lib


def use_library_for_testing(path):
    """Tests only: route the ABI to another build (the host emulation of the kernel source)."""
    global _lib, _emulated
    if path is None:
        _lib, _emulated = None, False
        return
    _lib = _bind(ctypes.CDLL(path))
    _emulated = True


def is_emulated():
    return _emulated


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib().tqb_last_error().decode()}")


def check_device(t):
    if t.is_cuda:
        if _emulated:
            raise RuntimeError("emulation library active but tensor is on CUDA")
        return
    if not _emulated:
        raise RuntimeError(
            "torchquantum_b200 runs on CUDA (sm_100a) only: got a CPU tensor and there is no CPU fallback")


def stream_ptr(t):
    if t.is_cuda:
        import torch

        return torch.cuda.current_stream(t.device).cuda_stream
    return 0

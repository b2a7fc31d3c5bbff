"""CPU/GPU ORACLE IN DOUBLE PRECISION — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The fp64 twin of ``oracle/tq_oracle.py``: the SAME source text (the line-by-line restatement of the reference's
gate-application path, pinned to the reference's golden vectors by ``tests/test_oracle_golden.py``) executed with
``C64 = complex128`` / ``F32 = float64``.  It exists to judge *accumulated* rounding at depth: a depth-100 circuit run in
complex64 (by the reference, by the oracle or by the CUDA kernels) carries ~2e-6 of its own rounding noise, so
two complex64 results cannot be compared to 1e-5 against each other with any margin; both are compared with this twin
instead.  ``tests/test_oracle_golden.py::test_fp64_twin_agrees_with_the_pinned_oracle`` ties it back to the pinned oracle.

``run_ops64(..., device="cuda")`` lets the twin use torch's CUDA ops for the 2^20-2^28-amplitude cases (a CPU run of
3 550 gates on 2^24 complex128 amplitudes takes a quarter of an hour); that is torch's own permute/bmm, none of the
product's kernels, and the only callers are tests and tools.

Only ``tests/``, ``tools/`` checkers and ``__graft_entry__.smoke()`` may import this module.
"""
import os
import types

import torch

_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tq_oracle.py")


def _load():
    src = open(_SRC).read()
    assert "C64 = torch.complex64" in src and "F32 = torch.float32" in src
    src = src.replace("C64 = torch.complex64", "C64 = torch.complex128").replace("F32 = torch.float32", "F32 = torch.float64")
    mod = types.ModuleType("oracle._tq_oracle_fp64")
    mod.__file__ = _SRC
    exec(compile(src, _SRC, "exec"), mod.__dict__)
    return mod


O64 = _load()


def run_ops64(n_wires, bsz, ops, device="cpu", init_states=None, keep_graph=False):
    """ops as for tq_oracle.run_ops (parameters are read as the float32 values the engine sees, then promoted).
    Returns [bsz, 2**n] complex128 on `device`.  keep_graph: parameters stay attached to autograd (the tiny gate
    matrices are built on the CPU and moved, differentiably, to `device`), so gradients come out in float64 too."""
    dev = torch.device(device)
    if init_states is None:
        s = torch.zeros(bsz, 2 ** n_wires, dtype=torch.complex128, device=dev)
        s[:, 0] = 1
    else:
        s = init_states.to(device=dev, dtype=torch.complex128)
    s = s.reshape([s.shape[0]] + [2] * n_wires)
    for op in ops:
        wires = [op["wires"]] if isinstance(op["wires"], int) else [int(w) for w in op["wires"]]
        p = op.get("params")
        if p is not None:
            p = torch.as_tensor(p)
            if not keep_graph:
                p = p.detach()
            p = p.to("cpu")
            p = p.to(torch.complex128 if p.is_complex() else torch.float64)
        m = O64.gate_matrix(op["name"], p, n_wires=op.get("n_wires") if op.get("n_wires") is not None else len(wires),
                            inverse=op.get("inverse", False))
        s = O64.apply_unitary_bmm(s, m.to(dev), wires).contiguous()
    return s.reshape(s.shape[0], -1)

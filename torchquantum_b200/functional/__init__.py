"""`tqf` — drop-in for torchquantum.functional (functional/__init__.py:25-218).

Every name of the reference's `func_name_dict` exists with the same signature
`(q_device, wires, params=None, n_wires=None, static=False, parent_graph=None, inverse=False,
comp_method="bmm")`.  Instead of building a torch matrix and calling bmm per gate
(functional/gate_wrapper.py:324-446) a call validates and *queues* the gate on the device; the
fused sm_100a passes run when the state is next read.  `static` / `parent_graph` /
`comp_method` are accepted and ignored: the engine always fuses.
"""
import numpy as np
import torch

from .. import engine
from .. import gates as G
from ..macro import C_DTYPE, F_DTYPE
from .matrices import (mat_dict, multirz_eigvals, multirz_matrix, multicnot_matrix, multixcnot_matrix, qft_matrix,
                       qubitunitary_matrix, qubitunitaryfast_matrix, qubitunitarystrict_matrix, globalphase_matrix)

__all__ = ["func_name_dict", "mat_dict", "gate_wrapper", "apply_unitary_bmm", "apply_unitary_einsum"]


_BORROW = [0]


class borrowed_params:
    """Inside this context gate calls keep the caller's parameter tensors by reference instead of snapshotting them
    (encoders slicing one input batch into many gates: the batch outlives the flush; the version guard of the queue
    still catches an in-place edit)."""

    def __enter__(self):
        _BORROW[0] += 1

    def __exit__(self, *exc):
        _BORROW[0] -= 1


def _snapshot(params):
    """The reference consumes parameter VALUES when the gate is called (functional/gate_wrapper.py:361-377 builds the
    matrix immediately); a lazily executed gate must not see later in-place edits of the caller's tensor.  Tensors
    that take part in autograd are kept by reference (their in-place edits are caught by the queue's version guard);
    plain value tensors are tiny and copied."""
    if params.requires_grad or _BORROW[0]:
        return params
    return params.clone()


def _normalize_params(name, params):
    """Shape rules of the reference (functional/gate_wrapper.py:361-377)."""
    if params is None:
        return None
    unitary_like = name in G.UNITARY_LIKE
    if not isinstance(params, torch.Tensor):
        params = torch.tensor(params, dtype=C_DTYPE if unitary_like else F_DTYPE)
    if unitary_like:
        return params.unsqueeze(0) if params.dim() == 2 else params
    if params.dim() == 1:
        params = params.unsqueeze(-1)
    elif params.dim() == 0:
        params = params.unsqueeze(-1).unsqueeze(-1)
    return params


def check_wires(name, wires_t, n_wires):
    """Wire indices are validated at the gate call (the reference fails there too: `permute` raises on an index out of
    range or a repeated dim, gate_wrapper.py:96-140), not when the lazy queue runs."""
    for w in wires_t:
        if not 0 <= w < n_wires:
            raise IndexError(f"{name}: wire {w} is out of range for a device with {n_wires} wires")
    if len(wires_t) > 1 and len(set(wires_t)) != len(wires_t):
        raise RuntimeError(f"{name}: repeated wire in {list(wires_t)}")


def gate_wrapper(name, mat, method, q_device, wires, params=None, n_wires=None, static=False, parent_graph=None,
                 inverse=False):
    """Queue one gate on `q_device` (reference: functional/gate_wrapper.py:324-446).

    `mat` may be a constant tensor / callable from `mat_dict` (ignored for catalogue gates: the
    device builds its own matrix from `params`) or, for names outside the catalogue, an explicit
    matrix that is applied as `qubitunitaryfast` would."""
    cname = G.canonical(name)
    params = _normalize_params(cname, params)
    wires = [wires] if isinstance(wires, (int, np.integer)) else list(wires)
    wires_t = tuple(int(w) for w in wires)
    check_wires(name, wires_t, q_device.n_wires)

    if q_device.record_op:
        q_device._record({
            "name": name,
            "wires": np.array(wires_t).squeeze().tolist(),
            "params": params,  # converted to a python list lazily (QuantumDevice.op_history)
            "inverse": inverse,
            "trainable": params.requires_grad if params is not None else False,
        })

    if cname not in G.SPECS:
        # explicit matrix (controlled_unitary, user gates): behave like qubitunitaryfast
        assert isinstance(mat, torch.Tensor), f"unknown gate {name}"
        cname, params = "qubitunitaryfast", (mat.unsqueeze(0) if mat.dim() == 2 else mat)
    spec = G.SPECS[cname]
    if cname in G.UNITARY_LIKE:
        if cname == "qubitunitary":
            params = qubitunitary_matrix(params)
        elif cname == "qubitunitarystrict":
            params = qubitunitarystrict_matrix(params)
        else:
            params = params.squeeze(0)
        assert np.log2(params.shape[-1]) == len(wires_t)  # functional/gate_wrapper.py:433
        if params.device != q_device.compute_device:
            params = params.to(q_device.compute_device)
        q_device.enqueue(cname, wires_t, _snapshot(params.to(C_DTYPE)), bool(inverse))
        return
    expected = spec.n_wires if spec.n_wires is not None else len(wires_t)
    assert expected == len(wires_t), f"{name}: expected {expected} wires, got {len(wires_t)}"
    if spec.n_params > 0 and params is None:
        raise TypeError(f"{name}: missing params")
    if spec.n_params == 0:
        params = None
    elif params.device != q_device.compute_device:
        params = params.to(q_device.compute_device)
    else:
        params = _snapshot(params)
    q_device.enqueue(cname, wires_t, params, bool(inverse))


def _make(name):
    cname = G.canonical(name)
    spec = G.SPECS.get(cname)
    simple = spec is not None and cname not in G.UNITARY_LIKE and spec.kind not in (G.RESET,)
    n_wires_fixed = spec.n_wires if spec is not None else None
    n_params = spec.n_params if spec is not None else 0

    def gate(q_device, wires, params=None, n_wires=None, static=False, parent_graph=None, inverse=False,
             comp_method="bmm"):
        if simple and not q_device.record_op:
            # hot path of gate_wrapper for catalogue gates (same checks, no per-call dictionary lookups)
            if isinstance(wires, (int, np.integer)):
                wires_t = (int(wires),)
            else:
                wires_t = tuple([int(w) for w in wires])
            if n_wires_fixed is not None and n_wires_fixed != len(wires_t):
                raise AssertionError(f"{name}: expected {n_wires_fixed} wires, got {len(wires_t)}")
            check_wires(name, wires_t, q_device.n_wires)
            if n_params == 0:
                params = None
            else:
                if params is None:
                    raise TypeError(f"{name}: missing params")
                if not isinstance(params, torch.Tensor) or params.dim() != 2:
                    params = _normalize_params(cname, params)
                dev = q_device._flat.device
                if params.device != dev:
                    params = params.to(dev)
                elif not params.requires_grad and not _BORROW[0]:
                    params = params.clone()  # _snapshot
            q_device._queue.append((cname, wires_t, params, bool(inverse)))
            return
        gate_wrapper(name=cname, mat=mat_dict.get(cname), method=comp_method, q_device=q_device, wires=wires,
                     params=params, n_wires=n_wires, static=static, parent_graph=parent_graph, inverse=inverse)

    gate.__name__ = gate.__qualname__ = name
    gate.__doc__ = f"Apply `{name}` (canonical `{cname}`) to `q_device`; see functional.gate_wrapper."
    return gate


_NAMES = [
    "hadamard", "h", "sh", "shadamard", "paulix", "pauliy", "pauliz", "i", "s", "t", "sx", "cnot", "cz", "cy", "rx",
    "ry", "rz", "rxx", "xx", "ryy", "yy", "rzz", "zz", "rzx", "zx", "swap", "sswap", "cswap", "toffoli", "phaseshift",
    "p", "cp", "rot", "multirz", "crx", "cry", "crz", "crot", "u1", "u2", "u3", "u", "cu1", "cphase", "cr", "cu2",
    "cu3", "cu", "qubitunitary", "qubitunitaryfast", "qubitunitarystrict", "multicnot", "multixcnot", "x", "y", "z",
    "cx", "ccnot", "ccx", "singleexcitation", "ecr", "echoedcrossresonance", "qft", "sdg", "tdg", "sxdg", "ch",
    "ccz", "iswap", "cs", "csdg", "csx", "chadamard", "dcx", "xxminyy", "xxplusyy", "c3x", "r", "globalphase",
    "c3sx", "rccx", "rc3x", "c4x",
]
func_name_dict = {}
for _n in _NAMES:
    func_name_dict[_n] = _make(_n)
    globals()[_n] = func_name_dict[_n]


def reset(q_device, wires, inverse=False):
    """functional/reset.py:19-42: project the wires on |0> and renormalise (non-unitary)."""
    wires = [wires] if isinstance(wires, (int, np.integer)) else list(wires)
    check_wires("reset", tuple(int(w) for w in wires), q_device.n_wires)
    q_device.enqueue("reset", tuple(int(w) for w in wires), None, False)


func_name_dict["reset"] = reset


def matrix_exp(qdev, wires, params, n_wires=None, inverse=False, comp_method="bmm"):
    """functional/func_mat_exp.py:34-72: apply exp(params) as an explicit matrix."""
    if isinstance(params, np.ndarray):
        params = torch.from_numpy(params)
    qubitunitaryfast(qdev, wires=wires, params=torch.matrix_exp(params))  # noqa: F821


def controlled_unitary(qdev, c_wires, t_wires, params):
    """functional/func_controlled_unitary.py:31-133: identity on every control pattern except
    all-ones, where kron copies of `params` act on the target sets."""
    if isinstance(params, np.ndarray):
        params = torch.from_numpy(params)
    c_wires, t_wires = np.array(c_wires), np.array(t_wires)
    u = params.to(C_DTYPE)
    u_wires = int(np.log2(u.shape[-1]))
    wires = [c_wires.item()] if c_wires.ndim == 0 else [int(w) for w in c_wires]
    n_c = len(wires)
    t_flat = [t_wires.item()] if t_wires.ndim == 0 else [int(w) for w in t_wires.flatten()]
    wires += t_flat
    n_sets = max(len(t_flat) // u_wires, 1)
    block = u
    for _ in range(n_sets - 1):
        block = torch.kron(block, u)
    dim = 2 ** (n_c + len(t_flat))
    unitary = torch.eye(dim, dtype=C_DTYPE, device=u.device)
    unitary[dim - block.shape[-1]:, dim - block.shape[-1]:] = block
    gate_wrapper(name="qubitunitaryfast", mat=unitary, method="bmm", q_device=qdev, wires=wires, params=unitary,
                 n_wires=len(wires))


func_name_dict_collect = {"matrix_exp": matrix_exp, "controlled_unitary": controlled_unitary}


def _apply_explicit(state, mat, wires):
    n = state.dim() - 1
    bsz = state.shape[0]
    flat = state.reshape(bsz, -1)
    mat = mat.to(C_DTYPE)
    if mat.dim() == 3 and mat.shape[0] != bsz:
        if bsz == 1:  # the einsum path of the reference broadcasts a single state over a batched matrix
            flat = flat.expand(mat.shape[0], -1)
            bsz = mat.shape[0]
        else:
            raise RuntimeError("batch of the matrix does not match the batch of the state")
    out = engine.run_circuit(flat, n, [("qubitunitaryfast", tuple(int(w) for w in wires), mat, False)])
    return out.reshape([bsz] + [2] * n)


def apply_unitary_bmm(state, mat, wires):
    """Pure function `state, mat, wires -> new state` (functional/gate_wrapper.py:96-140)."""
    if mat.dim() == 3 and mat.shape[0] != state.shape[0]:
        raise RuntimeError("bmm: batch of the matrix does not match the batch of the state")
    return _apply_explicit(state, mat, wires)


def apply_unitary_einsum(state, mat, wires):
    """functional/gate_wrapper.py:18-93 (same contraction; broadcasts a bsz=1 state)."""
    return _apply_explicit(state, mat, wires)

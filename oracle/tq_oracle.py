"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A torch-CPU restatement of the gate-application hot path of mit-han-lab/torchquantum
(reference @ 00039cc).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product package
(``torchquantum_b200``) never does and has no CPU fallback.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the *real* reference (imported from
/root/reference through ``oracle/refshim.py``) and stores its inputs/outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against those
fixtures, and (when /root/reference is mounted) directly against the live reference.

The arithmetic follows the reference step by step (same dtype casts, same torch primitives:
permute → reshape → bmm, kron → mm, float32 |.|**2 marginals) so that results agree to float32
rounding, and it is differentiable with plain autograd exactly like the reference
(reference backward == autograd through gate_wrapper).

All ``file:line`` citations are relative to /root/reference/torchquantum/.
"""
import bisect
import itertools
import math
import random
from collections import OrderedDict

import numpy as np
import torch

C64 = torch.complex64  # macro.py:41
F32 = torch.float32  # macro.py:42
INV_SQRT2 = 1 / math.sqrt(2)  # macro.py:47


# --------------------------------------------------------------------------------------
# gate matrices (functional/*.py builders and constant tables)
# --------------------------------------------------------------------------------------
def _col(params, k):
    """k-th parameter column as complex64 [b] (the reference casts θ to complex first,
    e.g. functional/rx.py:29, so trig runs in complex64 == float32 precision)."""
    return params[:, k].type(C64)


def _assemble(entries, b, dim):
    """entries: {(row, col): tensor[b] | python scalar} → [b, dim, dim] complex64."""
    rows = []
    for r in range(dim):
        cols = []
        for c in range(dim):
            v = entries.get((r, c), 0.0)
            if not isinstance(v, torch.Tensor):
                v = torch.full((b,), complex(v), dtype=C64)
            cols.append(v.type(C64))
        rows.append(torch.stack(cols, dim=-1))
    return torch.stack(rows, dim=-2)


def _controlled(block, b):
    """2x2 block {(r,c): v} placed on rows/cols 2..3 of a 4x4 identity-topped matrix
    (functional/rx.py:97-115, ry.py:29-47, rz.py:132-147, rot.py:67-89, u3.py:90-112)."""
    entries = {(0, 0): 1.0, (1, 1): 1.0}
    for (r, c), v in block.items():
        entries[(r + 2, c + 2)] = v
    return _assemble(entries, b, 4)


def _rx_block(p):  # functional/rx.py:29-46
    t = _col(p, 0)
    co, jsi = torch.cos(t / 2), 1j * torch.sin(-t / 2)
    return {(0, 0): co, (0, 1): jsi, (1, 0): jsi, (1, 1): co}


def _ry_block(p):  # functional/ry.py:62-67
    t = _col(p, 0)
    co, si = torch.cos(t / 2), torch.sin(t / 2)
    return {(0, 0): co, (0, 1): -si, (1, 0): si, (1, 1): co}


def _rz_block(p):  # functional/rz.py:160-171
    e = torch.exp(-0.5j * _col(p, 0))
    return {(0, 0): e, (1, 1): torch.conj(e)}


def _phase_block(p):  # functional/phase_shift.py:29-44, u1.py:29-44
    return {(0, 0): 1.0, (1, 1): torch.exp(1j * _col(p, 0))}


def _rot_block(p):  # functional/rot.py:29-54
    phi, theta, omega = _col(p, 0), _col(p, 1), _col(p, 2)
    co, si = torch.cos(theta / 2), torch.sin(theta / 2)
    return {
        (0, 0): torch.exp(-0.5j * (phi + omega)) * co,
        (0, 1): -torch.exp(0.5j * (phi - omega)) * si,
        (1, 0): torch.exp(-0.5j * (phi - omega)) * si,
        (1, 1): torch.exp(0.5j * (phi + omega)) * co,
    }


def _u2_block(p, scale):  # functional/u2.py:29-41 (scale 1/sqrt2) and :54-71 (cu2: no scale!)
    phi, lam = _col(p, 0), _col(p, 1)
    return {
        (0, 0): scale * torch.ones_like(phi),
        (0, 1): -scale * torch.exp(1j * lam),
        (1, 0): scale * torch.exp(1j * phi),
        (1, 1): scale * torch.exp(1j * (phi + lam)),
    }


def _u3_block(p, gamma=None):  # functional/u3.py:52-77 ; cu (gamma) :26-49
    theta, phi, lam = _col(p, 0), _col(p, 1), _col(p, 2)
    co, si = torch.cos(theta / 2), torch.sin(theta / 2)
    g = 0 if gamma is None else gamma
    blk = {
        (0, 0): co * torch.exp(1j * g) if gamma is not None else co,
        (0, 1): -si * torch.exp(1j * (lam + g)),
        (1, 0): si * torch.exp(1j * (phi + g)),
        (1, 1): co * torch.exp(1j * (phi + lam + g)),
    }
    return blk


def _r_block(p):  # functional/r.py:30-53
    theta, phi = _col(p, 0), _col(p, 1)
    e = torch.exp(-1j * phi)
    co, jsi = torch.cos(theta / 2), 1j * torch.sin(-theta / 2)
    return {(0, 0): co, (0, 1): e * jsi, (1, 0): torch.conj(e) * jsi, (1, 1): co}


def _one(block_fn):
    return lambda p: _assemble(block_fn(p), p.shape[0], 2)


def _ctl(block_fn):
    return lambda p: _controlled(block_fn(p), p.shape[0])


def _cu2(p):
    # functional/u2.py:54-71 — the reference cu2 block is [[1, -e^{iλ}], [e^{iφ}, e^{i(φ+λ)}]]
    # with NO 1/sqrt2 and entry (2,2) left at 1 (not unitary); restated as is.
    return _controlled(_u2_block(p, 1.0), p.shape[0])


def _rxx(p):  # functional/rx.py:60-84
    t = _col(p, 0)
    co, mjs = torch.cos(t / 2), -1j * torch.sin(t / 2)
    return _assemble({(0, 0): co, (1, 1): co, (2, 2): co, (3, 3): co,
                      (0, 3): mjs, (1, 2): mjs, (2, 1): mjs, (3, 0): mjs}, p.shape[0], 4)


def _ryy(p):  # functional/ry.py:80-104
    t = _col(p, 0)
    co, js = torch.cos(t / 2), 1j * torch.sin(t / 2)
    return _assemble({(0, 0): co, (1, 1): co, (2, 2): co, (3, 3): co,
                      (0, 3): js, (1, 2): -js, (2, 1): -js, (3, 0): js}, p.shape[0], 4)


def _rzz(p):  # functional/rz.py:61-80
    e = torch.exp(-0.5j * _col(p, 0))
    ce = torch.conj(e)
    return _assemble({(0, 0): e, (1, 1): ce, (2, 2): ce, (3, 3): e}, p.shape[0], 4)


def _rzx(p):  # functional/rz.py:93-119
    t = _col(p, 0)
    co, js = torch.cos(t / 2), 1j * torch.sin(t / 2)
    return _assemble({(0, 0): co, (0, 1): -js, (1, 0): -js, (1, 1): co,
                      (2, 2): co, (2, 3): js, (3, 2): js, (3, 3): co}, p.shape[0], 4)


def _cu1(p):  # functional/u1.py:57-72
    return _assemble({(0, 0): 1.0, (1, 1): 1.0, (2, 2): 1.0,
                      (3, 3): torch.exp(1j * _col(p, 0))}, p.shape[0], 4)


def _single_excitation(p):  # functional/single_excitation.py:29-48
    t = _col(p, 0)
    co, si = torch.cos(t / 2), torch.sin(t / 2)
    return _assemble({(0, 0): 1.0, (3, 3): 1.0, (1, 1): co, (1, 2): -si, (2, 1): si, (2, 2): co},
                     p.shape[0], 4)


def _xxminyy(p):  # functional/xx_min_yy.py:29-75
    theta, beta = _col(p, 0), _col(p, 1)
    co, si = torch.cos(theta / 2), torch.sin(theta / 2)
    return _assemble({(0, 0): co, (0, 3): -1j * si * torch.exp(-1j * beta), (1, 1): 1.0,
                      (2, 2): 1.0, (3, 0): -1j * si * torch.exp(1j * beta), (3, 3): co},
                     p.shape[0], 4)


def _xxplusyy(p):  # functional/xx_plus_yy.py:29-75
    theta, beta = _col(p, 0), _col(p, 1)
    co, si = torch.cos(theta / 2), torch.sin(theta / 2)
    return _assemble({(0, 0): 1.0, (1, 1): co, (1, 2): -1j * si * torch.exp(1j * beta),
                      (2, 1): -1j * si * torch.exp(-1j * beta), (2, 2): co, (3, 3): 1.0},
                     p.shape[0], 4)


def _globalphase(p):  # functional/global_phase.py:26-30
    return torch.exp(1j * _col(p, 0)).reshape(-1, 1, 1)


def pauli_signs(n):
    """(+1,-1) eigenvalue table of Z^{⊗n}: (-1)^popcount(k) (util/utils.py:82-96)."""
    k = np.arange(2 ** n)
    pop = np.zeros_like(k)
    for s in range(n):
        pop += (k >> s) & 1
    return 1 - 2 * (pop & 1)


def _multirz(p, n_wires):  # functional/rz.py:19-48
    eig = torch.exp(-1j * _col(p, 0).unsqueeze(-1) / 2 * torch.tensor(pauli_signs(n_wires)))
    return torch.diag_embed(eig).type(C64)


PARAM_BUILDERS = {
    # name: (num_wires, num_params, builder)
    "rx": (1, 1, _one(_rx_block)), "ry": (1, 1, _one(_ry_block)), "rz": (1, 1, _one(_rz_block)),
    "phaseshift": (1, 1, _one(_phase_block)), "u1": (1, 1, _one(_phase_block)),
    "rot": (1, 3, _one(_rot_block)), "u2": (1, 2, _one(lambda p: _u2_block(p, INV_SQRT2))),
    "u3": (1, 3, _one(_u3_block)), "r": (1, 2, _one(_r_block)),
    "crx": (2, 1, _ctl(_rx_block)), "cry": (2, 1, _ctl(_ry_block)), "crz": (2, 1, _ctl(_rz_block)),
    "crot": (2, 3, _ctl(_rot_block)), "cu3": (2, 3, _ctl(_u3_block)),
    "cu": (2, 4, _ctl(lambda p: _u3_block(p, gamma=_col(p, 3)))),
    "cu1": (2, 1, _cu1), "cu2": (2, 2, _cu2),
    "rxx": (2, 1, _rxx), "ryy": (2, 1, _ryy), "rzz": (2, 1, _rzz), "rzx": (2, 1, _rzx),
    "singleexcitation": (2, 1, _single_excitation),
    "xxminyy": (2, 2, _xxminyy), "xxplusyy": (2, 2, _xxplusyy),
    "globalphase": (0, 1, _globalphase),
}


def _c(rows):
    return torch.tensor(rows, dtype=C64)


def _ctrl_const(base, n_ctrl):
    """identity with `base` (2x2) in the bottom-right corner."""
    d = 2 ** (n_ctrl + 1)
    m = torch.eye(d, dtype=C64)
    m[d - 2:, d - 2:] = base
    return m


_X = _c([[0, 1], [1, 0]])
_Y = _c([[0, -1j], [1j, 0]])
_Z = _c([[1, 0], [0, -1]])
_H = _c([[INV_SQRT2, INV_SQRT2], [INV_SQRT2, -INV_SQRT2]])
_S = _c([[1, 0], [0, 1j]])
_SX = 0.5 * _c([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])


def _rccx():  # functional/paulix.py:116-127
    m = torch.eye(8, dtype=C64)
    m[5, 5] = -1
    m[6, 6] = 0
    m[7, 7] = 0
    m[6, 7] = -1j
    m[7, 6] = 1j
    return m


def _rc3x():  # functional/paulix.py:129-149
    m = torch.eye(16, dtype=C64)
    m[12, 12] = 1j
    m[13, 13] = -1j
    m[14, 14] = 0
    m[15, 15] = 0
    m[14, 15] = 1
    m[15, 14] = -1
    return m


FIXED_MATRICES = {
    # functional/hadamard.py:12-32, paulix.py:91-150, pauliy.py:18-23, pauliz.py:18-36, i.py,
    # s.py:19-28, t.py:19-22, sx.py:19-50, swap.py:19-48, ecr.py:18-23
    "hadamard": _H,
    "shadamard": _c([[math.cos(math.pi / 8), -math.sin(math.pi / 8)],
                     [math.sin(math.pi / 8), math.cos(math.pi / 8)]]),
    "chadamard": _ctrl_const(_H, 1),
    "paulix": _X, "pauliy": _Y, "pauliz": _Z, "i": torch.eye(2, dtype=C64),
    "s": _S, "sdg": _S.conj(),
    "t": _c([[1, 0], [0, np.exp(1j * np.pi / 4)]]), "tdg": _c([[1, 0], [0, np.exp(-1j * np.pi / 4)]]),
    "sx": _SX, "sxdg": _SX.conj(),
    "cnot": _ctrl_const(_X, 1), "cy": _ctrl_const(_Y, 1), "cz": _ctrl_const(_Z, 1),
    "cs": _ctrl_const(_S, 1), "csdg": _ctrl_const(_S.conj(), 1), "csx": _ctrl_const(_SX, 1),
    "swap": _c([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]),
    "sswap": _c([[1, 0, 0, 0], [0, (1 + 1j) / 2, (1 - 1j) / 2, 0],
                 [0, (1 - 1j) / 2, (1 + 1j) / 2, 0], [0, 0, 0, 1]]),
    "iswap": _c([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]),
    "dcx": _c([[1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0, 1, 0, 0]]),
    "ecr": INV_SQRT2 * _c([[0, 0, 1, 1j], [0, 0, 1j, 1], [1, -1j, 0, 0], [-1j, 1, 0, 0]]),
    "toffoli": _ctrl_const(_X, 2), "ccz": _ctrl_const(_Z, 2),
    "cswap": torch.eye(8, dtype=C64)[[0, 1, 2, 3, 4, 6, 5, 7]],
    "rccx": _rccx(), "rc3x": _rc3x(),
    "c3x": _ctrl_const(_X, 3), "c4x": _ctrl_const(_X, 4), "c3sx": _ctrl_const(_SX, 3),
}

# functional-name aliases (functional/__init__.py:125-210 and the `x = paulix` style bindings)
ALIASES = {
    "h": "hadamard", "sh": "shadamard", "ch": "chadamard", "x": "paulix", "y": "pauliy",
    "z": "pauliz", "cx": "cnot", "ccnot": "toffoli", "ccx": "toffoli", "xx": "rxx", "yy": "ryy",
    "zz": "rzz", "zx": "rzx", "p": "phaseshift", "cp": "cu1", "cr": "cu1", "cphase": "cu1",
    "u": "u3", "echoedcrossresonance": "ecr",
}


def multicnot_matrix(n_wires):  # functional/paulix.py:19-35
    return _ctrl_const(_X, n_wires - 1)


def multixcnot_matrix(n_wires):  # functional/paulix.py:38-55 (flip last wire iff all others 0)
    m = torch.eye(2 ** n_wires, dtype=C64)
    m[:2, :2] = _X
    return m


def qft_matrix(n_wires):  # functional/qft.py:19-33
    d = 2 ** n_wires
    idx = np.arange(d)
    omega = np.exp(2 * np.pi * 1j / d)
    return torch.tensor(omega ** np.outer(idx, idx) / np.sqrt(d), dtype=C64)


def normalize_params(name, params):
    """Shape rules of gate_wrapper (functional/gate_wrapper.py:361-377)."""
    name = ALIASES.get(name, name)
    unitary_like = name in ("qubitunitary", "qubitunitaryfast", "qubitunitarystrict")
    if params is None:
        return None
    if not isinstance(params, torch.Tensor):
        params = torch.tensor(params, dtype=C64 if unitary_like else F32)
    if unitary_like:
        return params.unsqueeze(0) if params.dim() == 2 else params
    if params.dim() == 1:
        params = params.unsqueeze(-1)
    elif params.dim() == 0:
        params = params.unsqueeze(-1).unsqueeze(-1)
    return params


def gate_matrix(name, params=None, n_wires=None, inverse=False):
    """Matrix exactly as gate_wrapper builds it (functional/gate_wrapper.py:406-432):
    returns [d,d] for shared parameters (leading batch of 1 squeezed) or [b,d,d]."""
    name = ALIASES.get(name, name)
    params = normalize_params(name, params)
    if name in FIXED_MATRICES:
        m = FIXED_MATRICES[name]
    elif name == "multicnot":
        m = multicnot_matrix(n_wires)
    elif name == "multixcnot":
        m = multixcnot_matrix(n_wires)
    elif name == "qft":
        m = qft_matrix(n_wires)
    elif name == "multirz":
        m = _multirz(params, n_wires).squeeze(0)
    elif name in ("qubitunitary", "qubitunitaryfast"):  # functional/qubit_unitary.py:19-62
        m = params.squeeze(0)
    elif name == "qubitunitarystrict":  # functional/qubit_unitary.py:65-86 (U @ V, no conj)
        u, _, v = torch.svd(params)
        m = u.matmul(v)
    else:
        m = PARAM_BUILDERS[name][2](params).squeeze(0)
    if inverse:
        m = m.conj().transpose(-1, -2)
    return m


def gate_arity(name, n_wires=None):
    name = ALIASES.get(name, name)
    if name in FIXED_MATRICES:
        return int(round(math.log2(FIXED_MATRICES[name].shape[-1])))
    if name in PARAM_BUILDERS:
        return PARAM_BUILDERS[name][0]
    return n_wires


# --------------------------------------------------------------------------------------
# gate application (functional/gate_wrapper.py:96-140) and the state container
# --------------------------------------------------------------------------------------
def apply_unitary_bmm(state, mat, wires):
    """state [bsz,2,...,2]; mat [d,d] or [bsz,d,d]; matrix index = wires in the given order,
    first wire most significant."""
    mat = mat.type(C64)
    dims = [w + 1 for w in wires]
    rest = [d for d in range(1, state.dim()) if d not in dims]
    order = [0] + dims + rest
    back = list(np.argsort(order))
    shape = state.shape
    flat = state.permute(order).reshape(shape[0], mat.shape[-1], -1)
    if mat.dim() == 2:
        mat = mat.expand(shape[0], *mat.shape)
    out = mat.bmm(flat)
    permuted_shape = [shape[i] for i in order]
    return out.view(permuted_shape).permute(back)


def zero_states(n_wires, bsz):  # device/devices.py:64-71
    s = torch.zeros(bsz, 2 ** n_wires, dtype=C64)
    s[:, 0] = 1
    return s.reshape([bsz] + [2] * n_wires)


def normalize_statevector(states):  # util/utils.py:789-814
    flat = states.reshape(states.shape[0], -1).clone()
    energy = (flat.abs() ** 2).sum(-1)
    flat[energy == 0, 0] = 1
    factors = torch.sqrt(1 / (flat.abs() ** 2).sum(-1)).unsqueeze(-1)
    return (flat * factors).reshape(states.shape)


def reset_wire(states, wires):  # functional/reset.py:19-42 (single-wire semantics)
    states = states.clone()
    for w in wires:
        idx = [slice(None)] * states.dim()
        idx[w + 1] = 1
        states[tuple(idx)] = 0
    return normalize_statevector(states)


def apply_op(states, name, wires, params=None, n_wires=None, inverse=False):
    """One gate_wrapper call in dynamic mode."""
    wires = [wires] if isinstance(wires, int) else [int(w) for w in wires]
    if name == "reset":
        return reset_wire(states, wires)
    m = gate_matrix(name, params, n_wires=n_wires if n_wires is not None else len(wires),
                    inverse=inverse)
    assert 2 ** len(wires) == m.shape[-1]  # functional/gate_wrapper.py:433
    return apply_unitary_bmm(states, m, wires)


def run_ops(n_wires, bsz, ops, init_states=None):
    """ops: iterable of dicts {name, wires, params?, n_wires?, inverse?}. Returns
    [bsz, 2**n] complex64."""
    s = zero_states(n_wires, bsz) if init_states is None else init_states.reshape(
        [init_states.shape[0]] + [2] * n_wires)
    for op in ops:
        s = apply_op(s, op["name"], op["wires"], op.get("params"), op.get("n_wires"),
                     op.get("inverse", False))
    return s.reshape(s.shape[0], -1)


# --------------------------------------------------------------------------------------
# measurement (measurement/measurements.py)
# --------------------------------------------------------------------------------------
_PAULI = {"X": _X, "Y": _Y, "Z": _Z, "I": torch.eye(2, dtype=C64)}


def pauli_string_matrix(observable):
    """Kronecker product of the string's Pauli matrices (util/utils.py:990-1009, algorithm/hamiltonian.py:139-145)."""
    observable = observable.upper()
    ham = _PAULI[observable[0]]
    for ch in observable[1:]:
        ham = torch.kron(ham, _PAULI[ch])
    return ham


def expval_joint_analytical(states_1d, observable):
    """Dense-kron form, exactly measurements.py:232-279. O(4^n): small n only."""
    observable = observable.upper()
    n = int(round(math.log2(states_1d.shape[-1])))
    assert len(observable) == n
    ham = _PAULI[observable[0]]
    for ch in observable[1:]:
        ham = torch.kron(ham, _PAULI[ch])
    return (states_1d.conj() * torch.mm(ham, states_1d.transpose(0, 1)).transpose(0, 1)).sum(-1).real


# diagonalizing gates (operator/standard_gates/paulix.py:28-29, pauliy.py:28-29, hadamard.py:30-31)
DIAGONALIZING = {
    "X": [("hadamard", None)],
    "Y": [("pauliz", None), ("s", None), ("hadamard", None)],
    "Z": [],
    "I": [],
    "H": [("ry", -math.pi / 4)],
}


def expval_joint_rotated(states_1d, observable):
    """Large-n restatement: rotate every non-I wire to the Z basis (as
    measurements.py:203-209 does for the sampling estimator) and take
    sum_k |psi_k|^2 (-1)^popcount(k & zmask).  O(2^n).  Cross-checked against the dense form
    for n <= 10 in tests/test_oracle_golden.py."""
    observable = observable.upper()
    n = len(observable)
    s = states_1d.reshape([states_1d.shape[0]] + [2] * n)
    zmask = 0
    for w, ch in enumerate(observable):
        for g, p in DIAGONALIZING[ch]:
            s = apply_op(s, g, [w], p)
        if ch != "I":
            zmask |= 1 << (n - 1 - w)
    probs = s.reshape(s.shape[0], -1).abs() ** 2
    k = np.arange(2 ** n)
    par = np.zeros_like(k)
    m = k & zmask
    for sft in range(n):
        par ^= (m >> sft) & 1
    sign = torch.tensor(1.0 - 2.0 * par, dtype=F32)
    return probs.type(F32) @ sign


_EIGVALS = {"X": [1.0, -1.0], "Y": [1.0, -1.0], "Z": [1.0, -1.0], "I": [1.0, 1.0], "H": [1.0, -1.0]}


def expval(states, wires, observables):
    """measurements.py:282-312.  `observables` are single-letter names; returns
    (expectations [bsz, len(wires)], rotated_states) — the reference mutates the device."""
    n = states.dim() - 1
    for w, ob in zip(wires, observables):
        for g, p in DIAGONALIZING[ob.upper()]:
            states = apply_op(states, g, [w], p)
    mag = torch.abs(states) ** 2
    out = []
    for w, ob in zip(wires, observables):
        dims = [d for d in range(1, n + 1) if d != w + 1]
        probs = mag.sum(dims) if dims else mag
        out.append(probs.mv(torch.tensor(_EIGVALS[ob.upper()], dtype=F32)))
    return torch.stack(out, dim=-1), states


def measure(states_1d, n_shots=1024, rng=random):
    """measurements.py:37-74 restated without the 2^n·n_shots string traffic: per batch element
    the float32 weights |psi|^2, then CPython's random.choices arithmetic (Lib/random.py:
    cum_weights = accumulate(weights); total = cum[-1] + 0.0; bisect(cum, random()*total, 0, n-1))
    under NumPy>=2 scalar rules, i.e. float32 running sums and float32(u)*total products
    (SURVEY Appendix C).  Returns list[OrderedDict[bitstring -> count]] with all 2^n keys."""
    n = int(round(math.log2(states_1d.shape[-1])))
    keys = ["{:0{}b}".format(k, n) for k in range(2 ** n)]
    mags = states_1d.abs().detach().cpu().numpy()
    res = []
    for mag in mags:
        probs = np.abs(mag) ** 2
        cum = list(itertools.accumulate(probs))
        total = cum[-1] + 0.0
        hi = len(cum) - 1
        counts = np.zeros(2 ** n, dtype=np.int64)
        for _ in range(n_shots):
            counts[bisect.bisect(cum, rng.random() * total, 0, hi)] += 1
        res.append(OrderedDict((k, int(c)) for k, c in zip(keys, counts)))
    return res

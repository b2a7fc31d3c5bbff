"""Gate matrices as torch tensors — the `mat_dict` of the reference (functional/__init__.py:97-123).

These are NOT on the hot path (the kernels build their matrices on the device, see
csrc/tqb_core.cuh `build_gate`); they back `Operator.matrix`, `mat_dict[...]` lookups and the
explicit-matrix entry points.  Conventions: SURVEY Appendix A.  Input params are `[b, p]` float
tensors; output is `[b, d, d]` complex64 with a leading batch of 1 squeezed, like the reference.
"""
import math

import numpy as np
import torch

from .. import gates as G
from ..macro import C_DTYPE


def _cx(re, im=None):
    return torch.complex(re, torch.zeros_like(re) if im is None else im)


def _expi(a):
    return torch.complex(torch.cos(a), torch.sin(a))


def _stack(entries, dim):
    """entries: dict (r, c) -> complex tensor [b]; missing -> 0."""
    some = next(iter(entries.values()))
    zero = torch.zeros_like(some)
    rows = [torch.stack([entries.get((r, c), zero) for c in range(dim)], dim=-1) for r in range(dim)]
    return torch.stack(rows, dim=-2)


def _one_qubit_block(kind, p):
    half = 0.5
    if kind == "rx":
        c, s = torch.cos(p[:, 0] * half), torch.sin(p[:, 0] * half)
        z = torch.zeros_like(c)
        return {(0, 0): _cx(c), (0, 1): _cx(z, -s), (1, 0): _cx(z, -s), (1, 1): _cx(c)}
    if kind == "ry":
        c, s = torch.cos(p[:, 0] * half), torch.sin(p[:, 0] * half)
        return {(0, 0): _cx(c), (0, 1): _cx(-s), (1, 0): _cx(s), (1, 1): _cx(c)}
    if kind == "rz":
        return {(0, 0): _expi(-p[:, 0] * half), (1, 1): _expi(p[:, 0] * half)}
    if kind == "phase":
        return {(0, 0): _cx(torch.ones_like(p[:, 0])), (1, 1): _expi(p[:, 0])}
    if kind == "rot":
        phi, theta, omega = p[:, 0], p[:, 1], p[:, 2]
        c, s = _cx(torch.cos(theta * half)), _cx(torch.sin(theta * half))
        return {(0, 0): _expi(-(phi + omega) * half) * c, (0, 1): -_expi((phi - omega) * half) * s,
                (1, 0): _expi(-(phi - omega) * half) * s, (1, 1): _expi((phi + omega) * half) * c}
    if kind in ("u2", "u2raw"):
        k = G.INV_SQRT2 if kind == "u2" else 1.0
        phi, lam = p[:, 0], p[:, 1]
        return {(0, 0): _cx(torch.full_like(phi, k)), (0, 1): -k * _expi(lam), (1, 0): k * _expi(phi),
                (1, 1): k * _expi(phi + lam)}
    if kind in ("u3", "cu"):
        theta, phi, lam = p[:, 0], p[:, 1], p[:, 2]
        gam = p[:, 3] if kind == "cu" else torch.zeros_like(theta)
        c, s = _cx(torch.cos(theta * half)), _cx(torch.sin(theta * half))
        return {(0, 0): c * _expi(gam), (0, 1): -s * _expi(lam + gam), (1, 0): s * _expi(phi + gam),
                (1, 1): c * _expi(phi + lam + gam)}
    if kind == "r":
        theta, phi = p[:, 0], p[:, 1]
        c, s = torch.cos(theta * half), torch.sin(theta * half)
        mis = _cx(torch.zeros_like(s), -s)
        return {(0, 0): _cx(c), (0, 1): mis * _expi(-phi), (1, 0): mis * _expi(phi), (1, 1): _cx(c)}
    raise KeyError(kind)


def _controlled(block):
    some = next(iter(block.values()))
    one = torch.ones_like(some)
    ent = {(0, 0): one, (1, 1): one}
    for (r, c), v in block.items():
        ent[(r + 2, c + 2)] = v
    return _stack(ent, 4)


def _two_qubit(kind, p):
    half = 0.5
    c, s = torch.cos(p[:, 0] * half), torch.sin(p[:, 0] * half)
    z = torch.zeros_like(c)
    one = _cx(torch.ones_like(c))
    mis, pis, cc = _cx(z, -s), _cx(z, s), _cx(c)
    if kind == "rxx":
        return _stack({(0, 0): cc, (1, 1): cc, (2, 2): cc, (3, 3): cc, (0, 3): mis, (1, 2): mis, (2, 1): mis,
                       (3, 0): mis}, 4)
    if kind == "ryy":
        return _stack({(0, 0): cc, (1, 1): cc, (2, 2): cc, (3, 3): cc, (0, 3): pis, (1, 2): mis, (2, 1): mis,
                       (3, 0): pis}, 4)
    if kind == "rzz":
        e, ce = _expi(-p[:, 0] * half), _expi(p[:, 0] * half)
        return _stack({(0, 0): e, (1, 1): ce, (2, 2): ce, (3, 3): e}, 4)
    if kind == "rzx":
        return _stack({(0, 0): cc, (0, 1): mis, (1, 0): mis, (1, 1): cc, (2, 2): cc, (2, 3): pis, (3, 2): pis,
                       (3, 3): cc}, 4)
    if kind == "singleexcitation":
        return _stack({(0, 0): one, (3, 3): one, (1, 1): cc, (1, 2): _cx(-s), (2, 1): _cx(s), (2, 2): cc}, 4)
    if kind == "xxminyy":
        beta = p[:, 1]
        return _stack({(0, 0): cc, (3, 3): cc, (1, 1): one, (2, 2): one, (0, 3): mis * _expi(-beta),
                       (3, 0): mis * _expi(beta)}, 4)
    if kind == "xxplusyy":
        beta = p[:, 1]
        return _stack({(0, 0): one, (3, 3): one, (1, 1): cc, (2, 2): cc, (1, 2): mis * _expi(beta),
                       (2, 1): mis * _expi(-beta)}, 4)
    raise KeyError(kind)


def _mk1(kind):
    return lambda params: _stack(_one_qubit_block(kind, params), 2).squeeze(0)


def _mkc(kind):
    return lambda params: _controlled(_one_qubit_block(kind, params)).squeeze(0)


def _mk2(kind):
    return lambda params: _two_qubit(kind, params).squeeze(0)


def pauli_eigs(n):
    """(-1)^popcount table, util/utils.py:82-96."""
    if n == 1:
        return np.array([1, -1])
    return np.concatenate([pauli_eigs(n - 1), -pauli_eigs(n - 1)])


def multirz_eigvals(params, n_wires):
    signs = torch.tensor(pauli_eigs(n_wires), dtype=torch.float32, device=params.device)
    return _expi(-params[:, :1] * 0.5 * signs)


def multirz_matrix(params, n_wires):
    return torch.diag_embed(multirz_eigvals(params, n_wires)).squeeze(0)


def globalphase_matrix(params):
    return _expi(params[:, 0]).reshape(-1, 1, 1).squeeze(0) if params.shape[0] > 1 else _expi(params[:, 0]).reshape(1, 1)


def _ctrl_np(base, n_ctrl):
    d = 2 ** (n_ctrl + int(round(math.log2(base.shape[0]))))
    m = np.eye(d, dtype=np.complex64)
    m[d - base.shape[0]:, d - base.shape[0]:] = base
    return m


def multicnot_matrix(n_wires):
    return torch.from_numpy(_ctrl_np(G.SPECS["paulix"].const, n_wires - 1))


def multixcnot_matrix(n_wires):
    m = np.eye(2 ** n_wires, dtype=np.complex64)
    m[:2, :2] = G.SPECS["paulix"].const
    return torch.from_numpy(m)


def qft_matrix(n_wires):
    return torch.from_numpy(G.qft_matrix_np(n_wires))


def qubitunitary_matrix(params):
    """functional/qubit_unitary.py:19-54: square + unitary check (host sync, as the reference)."""
    matrix = params.squeeze(0)
    assert matrix.shape[-1] == matrix.shape[-2], "Operator must be a square matrix."
    u = matrix.detach().cpu().numpy()
    eye = np.identity(u.shape[-1])
    prod = np.matmul(u, np.swapaxes(u.conj(), -1, -2))
    assert np.allclose(prod, np.broadcast_to(eye, prod.shape), atol=1e-5), "Operator must be unitary."
    return matrix


def qubitunitaryfast_matrix(params):
    return params.squeeze(0)


def qubitunitarystrict_matrix(params):
    u, _, v = torch.svd(params)  # functional/qubit_unitary.py:83-86 (U @ V, not U @ V^H)
    return u.matmul(v)


def _const_tensor(name):
    spec = G.SPECS[name]
    c = spec.const
    if spec.kind == G.DIAG:
        c = np.diag(c)
    n_ctrl = spec.n_ctrl
    return torch.from_numpy(np.ascontiguousarray(_ctrl_np(c, n_ctrl) if n_ctrl > 0 else c).astype(np.complex64))


mat_dict = {
    "rx": _mk1("rx"), "ry": _mk1("ry"), "rz": _mk1("rz"), "phaseshift": _mk1("phase"), "u1": _mk1("phase"),
    "rot": _mk1("rot"), "u2": _mk1("u2"), "u3": _mk1("u3"), "r": _mk1("r"),
    "crx": _mkc("rx"), "cry": _mkc("ry"), "crz": _mkc("rz"), "crot": _mkc("rot"), "cu1": _mkc("phase"),
    "cu2": _mkc("u2raw"), "cu3": _mkc("u3"), "cu": _mkc("cu"),
    "rxx": _mk2("rxx"), "ryy": _mk2("ryy"), "rzz": _mk2("rzz"), "rzx": _mk2("rzx"),
    "singleexcitation": _mk2("singleexcitation"), "xxminyy": _mk2("xxminyy"), "xxplusyy": _mk2("xxplusyy"),
    "multirz": multirz_matrix, "globalphase": globalphase_matrix,
    "multicnot": multicnot_matrix, "multixcnot": multixcnot_matrix, "qft": qft_matrix,
    "qubitunitary": qubitunitary_matrix, "qubitunitaryfast": qubitunitaryfast_matrix,
    "qubitunitarystrict": qubitunitarystrict_matrix,
    "i": torch.eye(2, dtype=C_DTYPE),
}
for _name, _spec in G.SPECS.items():
    if _spec.const is not None and _spec.n_wires is not None and _name not in mat_dict:
        mat_dict[_name] = _const_tensor(_name)

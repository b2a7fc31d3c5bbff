"""Gate catalogue: how every functional name of the reference (functional/__init__.py:125-210)
lowers to the three in-tile micro-ops of the engine.

  U1Q  — (multi-)controlled 2x2 on one target           (rx, h, cnot, crx, toffoli, multicnot ...)
  DIAG — (multi-)controlled phase pair selected by the parity of a bit mask
                                                          (rz, cz, cu1, rzz, multirz, s, t ...)
  U2Q  — (multi-)controlled 4x4 on two targets           (rxx, swap, cswap, ecr, rccx ...)
  DENSE — anything else (qft, qubitunitary on >= 3 wires): out-of-place generic kernel

Matrix conventions are the reference's: the matrix index is built from the
wires in the given order, first wire most significant; controlled gates list controls first.
"""
import math

import numpy as np

from . import _codes as K

INV_SQRT2 = 1 / math.sqrt(2)

U1Q, DIAG, U2Q, DENSE, RESET, NOP = "u1q", "diag", "u2q", "dense", "reset", "nop"


def _m(rows):
    return np.array(rows, dtype=np.complex64)


_X = _m([[0, 1], [1, 0]])
_Y = _m([[0, -1j], [1j, 0]])
_H = _m([[INV_SQRT2, INV_SQRT2], [INV_SQRT2, -INV_SQRT2]])
_SH = _m([[math.cos(math.pi / 8), -math.sin(math.pi / 8)], [math.sin(math.pi / 8), math.cos(math.pi / 8)]])
_SX = (0.5 * _m([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])).astype(np.complex64)
_SXDG = _m([[0.5 - 0.5j, 0.5 + 0.5j], [0.5 + 0.5j, 0.5 - 0.5j]])
_SWAP = _m([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
_SSWAP = _m([[1, 0, 0, 0], [0, (1 + 1j) / 2, (1 - 1j) / 2, 0], [0, (1 - 1j) / 2, (1 + 1j) / 2, 0], [0, 0, 0, 1]])
_ISWAP = _m([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])
_DCX = _m([[1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0, 1, 0, 0]])
_ECR = (INV_SQRT2 * _m([[0, 0, 1, 1j], [0, 0, 1j, 1], [1, -1j, 0, 0], [-1j, 1, 0, 0]])).astype(np.complex64)
_RCCX_BLOCK = _m([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]])
_RC3X_BLOCK = _m([[1j, 0, 0, 0], [0, -1j, 0, 0], [0, 0, 0, 1], [0, 0, -1, 0]])


class Spec:
    """kind: micro-op; n_wires (None = any); n_params; n_ctrl leading control wires (-1: all but
    the targets); const: constant matrix (2x2, 4x4 or diagonal pair) or None; builder: device
    builder kind for parameterised gates; unitary: safe for the adjoint un-apply."""

    def __init__(self, kind, n_wires, n_params=0, n_ctrl=0, const=None, builder=None, unitary=True,
                 neg_ctrl=False, parity_all=False):
        self.kind, self.n_wires, self.n_params, self.n_ctrl = kind, n_wires, n_params, n_ctrl
        self.const, self.builder, self.unitary = const, builder, unitary
        self.neg_ctrl, self.parity_all = neg_ctrl, parity_all


def _d(d0, d1):
    return np.array([d0, d1], dtype=np.complex64)


SPECS = {
    # fixed one-qubit
    "hadamard": Spec(U1Q, 1, const=_H), "shadamard": Spec(U1Q, 1, const=_SH),
    "paulix": Spec(U1Q, 1, const=_X), "pauliy": Spec(U1Q, 1, const=_Y),
    "pauliz": Spec(DIAG, 1, const=_d(1, -1)), "i": Spec(NOP, 1),
    "s": Spec(DIAG, 1, const=_d(1, 1j)), "sdg": Spec(DIAG, 1, const=_d(1, -1j)),
    "t": Spec(DIAG, 1, const=_d(1, np.exp(1j * np.pi / 4))),
    "tdg": Spec(DIAG, 1, const=_d(1, np.exp(-1j * np.pi / 4))),
    "sx": Spec(U1Q, 1, const=_SX), "sxdg": Spec(U1Q, 1, const=_SXDG),
    # fixed controlled
    "cnot": Spec(U1Q, 2, n_ctrl=1, const=_X), "cy": Spec(U1Q, 2, n_ctrl=1, const=_Y),
    "cz": Spec(DIAG, 2, n_ctrl=1, const=_d(1, -1)), "chadamard": Spec(U1Q, 2, n_ctrl=1, const=_H),
    "cs": Spec(DIAG, 2, n_ctrl=1, const=_d(1, 1j)), "csdg": Spec(DIAG, 2, n_ctrl=1, const=_d(1, -1j)),
    "csx": Spec(U1Q, 2, n_ctrl=1, const=_SX),
    "toffoli": Spec(U1Q, 3, n_ctrl=2, const=_X), "ccz": Spec(DIAG, 3, n_ctrl=2, const=_d(1, -1)),
    "c3x": Spec(U1Q, 4, n_ctrl=3, const=_X), "c4x": Spec(U1Q, 5, n_ctrl=4, const=_X),
    "c3sx": Spec(U1Q, 4, n_ctrl=3, const=_SX),
    "multicnot": Spec(U1Q, None, n_ctrl=-1, const=_X),
    "multixcnot": Spec(U1Q, None, n_ctrl=-1, const=_X, neg_ctrl=True),
    # fixed two-qubit
    "swap": Spec(U2Q, 2, const=_SWAP), "sswap": Spec(U2Q, 2, const=_SSWAP), "iswap": Spec(U2Q, 2, const=_ISWAP),
    "dcx": Spec(U2Q, 2, const=_DCX), "ecr": Spec(U2Q, 2, const=_ECR),
    "cswap": Spec(U2Q, 3, n_ctrl=1, const=_SWAP), "rccx": Spec(U2Q, 3, n_ctrl=1, const=_RCCX_BLOCK),
    "rc3x": Spec(U2Q, 4, n_ctrl=2, const=_RC3X_BLOCK),
    # parameterised one-qubit
    "rx": Spec(U1Q, 1, 1, builder=K.K_RX), "ry": Spec(U1Q, 1, 1, builder=K.K_RY),
    "rz": Spec(DIAG, 1, 1, builder=K.K_RZ), "phaseshift": Spec(DIAG, 1, 1, builder=K.K_PHASE),
    "u1": Spec(DIAG, 1, 1, builder=K.K_PHASE), "rot": Spec(U1Q, 1, 3, builder=K.K_ROT),
    "u2": Spec(U1Q, 1, 2, builder=K.K_U2), "u3": Spec(U1Q, 1, 3, builder=K.K_U3),
    "r": Spec(U1Q, 1, 2, builder=K.K_R),
    # parameterised controlled
    "crx": Spec(U1Q, 2, 1, n_ctrl=1, builder=K.K_RX), "cry": Spec(U1Q, 2, 1, n_ctrl=1, builder=K.K_RY),
    "crz": Spec(DIAG, 2, 1, n_ctrl=1, builder=K.K_RZ), "crot": Spec(U1Q, 2, 3, n_ctrl=1, builder=K.K_ROT),
    "cu1": Spec(DIAG, 2, 1, n_ctrl=1, builder=K.K_PHASE),
    # reference cu2 block is [[1,-e^{il}],[e^{ip},e^{i(p+l)}]] without 1/sqrt2 (functional/u2.py:54-71): not unitary
    "cu2": Spec(U1Q, 2, 2, n_ctrl=1, builder=K.K_U2RAW, unitary=False),
    "cu3": Spec(U1Q, 2, 3, n_ctrl=1, builder=K.K_U3), "cu": Spec(U1Q, 2, 4, n_ctrl=1, builder=K.K_CU),
    # parameterised two-qubit
    "rxx": Spec(U2Q, 2, 1, builder=K.K_RXX), "ryy": Spec(U2Q, 2, 1, builder=K.K_RYY),
    "rzx": Spec(U2Q, 2, 1, builder=K.K_RZX), "rzz": Spec(DIAG, 2, 1, builder=K.K_RZ, parity_all=True),
    "singleexcitation": Spec(U2Q, 2, 1, builder=K.K_SEXC),
    "xxminyy": Spec(U2Q, 2, 2, builder=K.K_XXMYY), "xxplusyy": Spec(U2Q, 2, 2, builder=K.K_XXPYY),
    # any number of wires
    "multirz": Spec(DIAG, None, 1, builder=K.K_RZ, parity_all=True),
    "globalphase": Spec(DIAG, 0, 1, builder=K.K_GPHASE),
    "qft": Spec(DENSE, None),
    "qubitunitary": Spec(DENSE, None), "qubitunitaryfast": Spec(DENSE, None, unitary=False),
    "qubitunitarystrict": Spec(DENSE, None),
    "reset": Spec(RESET, None),
}

# functional/__init__.py:125-210 aliases
ALIASES = {
    "h": "hadamard", "sh": "shadamard", "ch": "chadamard", "x": "paulix", "y": "pauliy", "z": "pauliz",
    "cx": "cnot", "ccnot": "toffoli", "ccx": "toffoli", "xx": "rxx", "yy": "ryy", "zz": "rzz",
    "zx": "rzx", "p": "phaseshift", "cp": "cu1", "cr": "cu1", "cphase": "cu1", "u": "u3",
    "echoedcrossresonance": "ecr",
}

UNITARY_LIKE = ("qubitunitary", "qubitunitaryfast", "qubitunitarystrict")


def canonical(name):
    return ALIASES.get(name, name)


def qft_matrix_np(k):
    d = 2 ** k
    idx = np.arange(d)
    return (np.exp(2j * np.pi / d) ** np.outer(idx, idx) / np.sqrt(d)).astype(np.complex64)

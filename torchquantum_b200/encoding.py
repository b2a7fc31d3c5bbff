"""Encoders — drop-in for encoding/encodings.py (GeneralEncoder + the named tables, PhaseEncoder,
MultiPhaseEncoder, StateEncoder).  Encoders only build gate lists with per-sample parameters
`x[:, idx]`; the arithmetic happens in the fused kernels when the device state is read."""
import torch

from . import functional as tqf
from .functional import borrowed_params, func_name_dict
from .module import QuantumModule, static_support
from .operator import op_name_dict

__all__ = ["Encoder", "GeneralEncoder", "PhaseEncoder", "MultiPhaseEncoder", "StateEncoder",
           "encoder_op_list_name_dict"]


class Encoder(QuantumModule):
    def forward(self, qdev, x):
        raise NotImplementedError


class GeneralEncoder(Encoder):
    """func_list: [{'input_idx': [0], 'func': 'rx', 'wires': [0]}, ...] (encoding/encodings.py:50-102)."""

    def __init__(self, func_list):
        super().__init__()
        self.func_list = func_list

    def _columns(self, x, idx):
        """x[:, idx] without a host->device index upload per call: a consecutive run of columns is a
        view; any other list goes through a device index tensor cached per device (also keeps the
        step capturable in a CUDA graph)."""
        if isinstance(idx, (list, tuple)) and len(idx) > 0 and all(isinstance(i, int) for i in idx):
            if all(idx[k + 1] == idx[k] + 1 for k in range(len(idx) - 1)) and idx[0] >= 0:
                return x[:, idx[0]:idx[-1] + 1]
            cache = self.__dict__.setdefault("_idx_cache", {})
            key = (tuple(idx), x.device)
            t = cache.get(key)
            if t is None:
                t = cache[key] = torch.tensor(idx, dtype=torch.long, device=x.device)
            return x.index_select(1, t)
        return x[:, idx]

    @static_support
    def forward(self, qdev, x):
        with borrowed_params():
            for info in self.func_list:
                if op_name_dict[info["func"]].num_params > 0:
                    params = self._columns(x, info["input_idx"])
                else:
                    params = None
                func_name_dict[info["func"]](qdev, wires=info["wires"], params=params, static=self.static_mode,
                                             parent_graph=self.graph)


class PhaseEncoder(Encoder):
    """One rotation `func` per wire, x[:, k] on wire k (encoding/encodings.py:133-161)."""

    def __init__(self, func):
        super().__init__()
        self.func = func

    @static_support
    def forward(self, qdev, x):
        for k in range(x.shape[1]):  # x[:, [k]] is a fresh tensor already
            self.func(qdev, wires=k, params=x[:, [k]], static=self.static_mode, parent_graph=self.graph)


class MultiPhaseEncoder(Encoder):
    """funcs[k] consumes the next num_params columns of x (encoding/encodings.py:163-206)."""

    def __init__(self, funcs, wires=None):
        super().__init__()
        self.funcs = funcs if isinstance(funcs, (list, tuple)) else [funcs]
        self.wires = wires

    @static_support
    def forward(self, qdev, x):
        wires = self.wires if self.wires is not None else list(range(qdev.n_wires)) * (len(self.funcs) // qdev.n_wires)
        x_id = 0
        with borrowed_params():
            for k, func in enumerate(self.funcs):
                stride = 3 if func in ("u3", "cu3") else (2 if func in ("u2",) else 1)
                func_name_dict[func](qdev, wires=wires[k], params=x[:, x_id:x_id + stride], static=self.static_mode,
                                     parent_graph=self.graph)
                x_id += stride


class StateEncoder(Encoder):
    """Amplitude encoding: normalised x padded with zeros becomes the state (encodings.py:208-241)."""

    def forward(self, qdev, x):
        state = torch.cat((x, torch.zeros(x.shape[0], 2 ** qdev.n_wires - x.shape[1], device=x.device)), dim=-1)
        state = state / (torch.sqrt((state.abs() ** 2).sum(dim=-1))).unsqueeze(-1)
        qdev.states = state.view([x.shape[0]] + [2] * qdev.n_wires).type(torch.complex64)


def _rounds(seq):
    """[(func, count), ...] -> consecutive inputs, `count` gates of `func` on wires 0..count-1."""
    out, idx = [], 0
    for func, count in seq:
        for w in range(count):
            out.append({"input_idx": [idx], "func": func, "wires": [w]})
            idx += 1
    return out


def _u3rx(n, with_h):
    out = []
    for w in range(n):
        out.append({"input_idx": [4 * w, 4 * w + 1, 4 * w + 2], "func": "u3", "wires": [w]})
        out.append({"input_idx": [4 * w + 3], "func": "rx", "wires": [w]})
        if with_h and w == 0:
            out += [{"func": "h", "wires": [k]} for k in range(n)]
    return out


def _rzsx(n, rounds):
    out = []
    for r in range(rounds):
        for w in range(n):
            out.append({"input_idx": [r * n + w], "func": "rz", "wires": [w]})
            out.append({"func": "sx", "wires": [w]})
    return out


encoder_op_list_name_dict = {
    "1x1_ry": _rounds([("ry", 1)]),
    "2x1_ryry": _rounds([("ry", 2)]),
    "2x8_rxryrzrxryrzrxry": _rounds([(f, 2) for f in ("rx", "ry", "rz", "rx", "ry", "rz", "rx", "ry")]),
    "3x1_ryryry": _rounds([("ry", 3)]),
    "3x1_rxrxrx": _rounds([("rx", 3)]),
    "4_ry": _rounds([("ry", 4)]),
    "4x4_u3rx": _u3rx(4, False),
    "4x4_u3_h_rx": _u3rx(4, True),
    "4x4_ryzxy": _rounds([("ry", 4), ("rz", 4), ("rx", 4), ("ry", 4)]),
    "8x2_ry": _rounds([("ry", 8), ("ry", 8)]),
    "16_ry": _rounds([("ry", 16)]),
    "4x4_rzsx": _rzsx(4, 4),
    "15_ryrz": _rounds([("ry", 15), ("rz", 1)]),
    "2x8_ryzxyzxyz": _rounds([(f, 2) for f in ("ry", "rz", "rx", "ry", "rz", "rx", "ry", "rz")]),
    "10_ryzx": _rounds([("ry", 4), ("rz", 4), ("rx", 2)]),
    "10_ry": _rounds([("ry", 10)]),
    "25_ry": _rounds([("ry", 25)]),
    "25_ryrz": _rounds([("ry", 21), ("rz", 4)]),
    "6x6_ryzxy": _rounds([("ry", 10), ("rz", 10), ("rx", 10), ("ry", 6)]),
    "6x6_ryrz": _rounds([("ry", 21), ("rz", 15)]),
    "6x6_ryrzrx": _rounds([("ry", 16), ("rz", 16), ("rx", 4)]),
    "10x10_ryzxyzxyzxy": _rounds([(f, 10) for f in ("ry", "rz", "rx", "ry", "rz", "rx", "ry", "rz", "rx", "ry")]),
    "8x8_ryzxyzxy": _rounds([("ry", 10), ("rz", 10), ("rx", 10), ("ry", 10), ("rz", 10), ("rx", 10), ("ry", 4)]),
    "8x2_ryz": _rounds([("ry", 8), ("rz", 8)]),
    "16x1_ry": _rounds([("ry", 16)]),
}

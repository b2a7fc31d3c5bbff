"""`tq.Operator` hierarchy and the gate classes — drop-in for operator/op_types.py:53-443 and
operator/standard_gates/*.py (66 classes, `op_name_dict`, `fixed_ops`, `parameterized_ops`).

An operator is an nn.Module shell: it owns `params` (nn.Parameter `[1, num_params]`, U(-pi, pi)
or `init_params`), `wires`, `inverse`, and its `forward` calls the functional of the same name,
which queues the gate on the device.  No arithmetic happens here.
"""
from abc import ABCMeta
from collections.abc import Iterable
from enum import IntEnum

import numpy as np
import torch
import torch.nn as nn

from . import functional as tqf
from . import gates as G
from .functional import mat_dict
from .functional import matrices as M
from .macro import C_DTYPE, F_DTYPE
from .module import MUTATION_EPOCH, QuantumModule

__all__ = ["op_name_dict", "Operator", "Operation", "DiagonalOperation", "Observable", "WiresEnum", "NParamsEnum",
           "AnyNParams", "AllWires", "AnyWires", "fixed_ops", "parameterized_ops"]


class WiresEnum(IntEnum):
    AnyWires = -1
    AllWires = 0


class NParamsEnum(IntEnum):
    AnyNParams = -1


AnyNParams = NParamsEnum.AnyNParams
AllWires = WiresEnum.AllWires
AnyWires = WiresEnum.AnyWires


class Operator(QuantumModule):
    fixed_ops = []       # filled below
    parameterized_ops = []

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        self._name = value

    def __init__(self, has_params: bool = False, trainable: bool = False, init_params=None, n_wires=None,
                 wires=None, inverse=False):
        super().__init__()
        self.params = None
        self.n_wires = n_wires
        self.wires = wires
        self._name = self.__class__.__name__
        self.static_matrix = None
        self.inverse = inverse
        self.clifford_quantization = False
        if trainable and not has_params:
            has_params = True  # the reference logs a warning and switches has_params on
        self.has_params = has_params
        self.trainable = trainable
        if self.has_params:
            self.params = self.build_params(trainable=self.trainable)
            self.reset_params(init_params)

    @classmethod
    def _matrix(cls, params):
        raise NotImplementedError

    @property
    def matrix(self):
        return self._matrix(self.params)

    @classmethod
    def _eigvals(cls, params):
        return None

    @property
    def eigvals(self):
        return self._eigvals(self.params)

    def _get_unitary_matrix(self):
        return self.matrix

    def set_wires(self, wires):
        self.wires = [wires] if isinstance(wires, int) else wires

    def forward(self, q_device, wires=None, params=None, inverse=None):
        """operator/op_types.py:182-266: external `params` overwrite self.params; wires update;
        then the functional queues the gate.  (Attributes are only re-assigned when they change:
        nn.Module.__setattr__ costs more than queueing the gate.)"""
        if inverse is not None and inverse != self.inverse:
            self.inverse = inverse
        if params is not None:
            self.params = params
        p = self.params
        if p is not None and p.dim() == 1:
            p = p.unsqueeze(-1)
            self.params = p
        if wires is not None:
            wires = [wires] if isinstance(wires, (int, np.integer)) else wires
            if wires is not self.wires:
                self.wires = wires
        noise = self.noise_model_tq
        if p is None:
            if self.n_wires is None:
                self.func(q_device, self.wires, inverse=self.inverse)
            else:
                self.func(q_device, self.wires, n_wires=self.n_wires, inverse=self.inverse)
        else:
            if noise is not None and hasattr(noise, "add_noise"):
                p = noise.add_noise(p)
            if self.n_wires is None:
                self.func(q_device, self.wires, params=p, inverse=self.inverse)
            else:
                self.func(q_device, self.wires, params=p, n_wires=self.n_wires, inverse=self.inverse)
        if noise is not None and getattr(noise, "is_add_noise", False):
            for noise_op in noise.sample_noise_op(self):
                noise_op(q_device)

    def fast_record(self, q_device, wires):
        """The queue record `(name, wires, params, inverse)` of this operator on `wires`, or None when the
        full `forward` has to run (noise model, op recording, unusual parameters, non-catalogue gate)."""
        p = self.params
        cname = self._fast_cname
        if (cname is None or self.noise_model_tq is not None or q_device.record_op
                or (p is not None and (p.dim() != 2 or p.device != q_device._flat.device))):
            return None
        wires_t = (int(wires),) if isinstance(wires, (int, np.integer)) else tuple([int(w) for w in wires])
        tqf.check_wires(cname, wires_t, q_device.n_wires)
        return (cname, wires_t, p, bool(self.inverse))

    def queue_on(self, q_device, wires):
        """Hot path for this package's own layer templates: same effect as `self(q_device, wires=wires)`
        for a catalogue gate with 2-D (or no) parameters when no noise model, static graph or op
        recording is involved; anything else goes through the full `forward`."""
        rec = self.fast_record(q_device, wires)
        if rec is None:
            self(q_device, wires=wires)
            return
        if wires is not self.wires:
            self.__dict__["wires"] = wires
        q_device._queue.append(rec)

    @property
    def _fast_cname(self):
        cls = type(self)
        c = cls.__dict__.get("_fast_cname_cache", 0)
        if c == 0:
            name = G.canonical(getattr(self.func, "__name__", ""))
            spec = G.SPECS.get(name)
            ok = (spec is not None and name not in G.UNITARY_LIKE and spec.kind != G.RESET
                  and spec.n_wires is not None and type(cls.forward) is type(Operator.forward)
                  and cls.forward is Operator.forward)
            c = name if ok else None
            cls._fast_cname_cache = c
        return c

    def __repr__(self):
        return f" class: {self.name} \n parameters: {self.params} \n wires: {self.wires} \n inverse: {self.inverse}"

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() may swap Parameter objects without going through __setattr__
        self.__dict__["_ver"] = self.__dict__.get("_ver", 0) + 1
        MUTATION_EPOCH[0] += 1
        return super()._apply(fn, *args, **kwargs)

    _PLAIN_ATTRS = frozenset(("wires", "inverse", "n_wires", "static_matrix", "clifford_quantization"))

    def __setattr__(self, name, value):
        # every attribute write bumps the operator's version: layer templates cache the records they queue
        # and re-validate them by version (layer._RecordCache)
        d = self.__dict__
        d["_ver"] = d.get("_ver", 0) + 1
        MUTATION_EPOCH[0] += 1
        if name in Operator._PLAIN_ATTRS:  # plain python state: skip nn.Module's type dispatch
            d[name] = value
            return
        # nn.Module would refuse to overwrite a registered Parameter with a plain tensor
        if name == "params" and "_parameters" in self.__dict__ and "params" in self._parameters \
                and not isinstance(value, nn.Parameter):
            del self._parameters["params"]
        super().__setattr__(name, value)


class Observable(Operator, metaclass=ABCMeta):
    def __init__(self, has_params: bool = False, trainable: bool = False, init_params=None, n_wires=None,
                 wires=None, inverse=False):
        super().__init__(has_params=has_params, trainable=trainable, init_params=init_params, n_wires=n_wires,
                         wires=wires, inverse=inverse)
        self.return_type = None

    def diagonalizing_gates(self):
        raise NotImplementedError


class Operation(Operator, metaclass=ABCMeta):
    def __init__(self, has_params: bool = False, trainable: bool = False, init_params=None, n_wires=None,
                 wires=None, inverse=False):
        super().__init__(has_params=has_params, trainable=trainable, init_params=init_params, n_wires=n_wires,
                         wires=wires, inverse=inverse)
        if type(self.num_wires) == int:
            self.n_wires = self.num_wires

    @property
    def matrix(self):
        return self._matrix(self.params)

    @property
    def eigvals(self):
        return self._eigvals(self.params)

    def init_params(self):
        raise NotImplementedError

    def build_params(self, trainable):
        parameters = nn.Parameter(torch.empty([1, self.num_params], dtype=F_DTYPE))
        parameters.requires_grad = True if trainable else False
        return parameters

    def reset_params(self, init_params=None):
        if init_params is not None:
            if isinstance(init_params, Iterable):
                for k, init_param in enumerate(init_params):
                    torch.nn.init.constant_(self.params[:, k], init_param)
            else:
                torch.nn.init.constant_(self.params, init_params)
        else:
            torch.nn.init.uniform_(self.params, -np.pi, np.pi)


# Observables keep the reference's behaviour of building params through Operation's helpers
Observable.build_params = Operation.build_params
Observable.reset_params = Operation.reset_params


class DiagonalOperation(Operation, metaclass=ABCMeta):
    @classmethod
    def _eigvals(cls, params):
        raise NotImplementedError

    @classmethod
    def _matrix(cls, params):
        return torch.diag(cls._eigvals(params))


# ---------------------------------------------------------------------------------------
# gate classes, generated from the catalogue (torchquantum_b200/gates.py)
# ---------------------------------------------------------------------------------------
def _fixed(name, op_name, base=Operation, eigvals=None, extra=None):
    mat = mat_dict[op_name]
    ns = {"num_params": 0, "num_wires": G.SPECS[op_name].n_wires, "op_name": op_name, "matrix": mat,
          "func": staticmethod(getattr(tqf, op_name)), "_matrix": classmethod(lambda cls, params: cls.matrix),
          "__doc__": f"Class for the {name} gate."}
    if eigvals is not None:
        ns["eigvals"] = torch.tensor(eigvals, dtype=C_DTYPE)
        ns["_eigvals"] = classmethod(lambda cls, params: cls.eigvals)
    if extra:
        ns.update(extra)
    return ABCMeta(name, (base,), ns)


def _param(name, op_name, base=Operation, extra=None):
    spec = G.SPECS[op_name]
    builder = mat_dict[op_name]
    ns = {"num_params": spec.n_params, "num_wires": spec.n_wires, "op_name": op_name,
          "func": staticmethod(getattr(tqf, op_name)),
          "_matrix": classmethod(lambda cls, params: builder(params)),
          "__doc__": f"Class for the {name} gate."}
    if extra:
        ns.update(extra)
    return ABCMeta(name, (base,), ns)


I = _fixed("I", "i", Observable, [1, 1], {"diagonalizing_gates": lambda self: []})
Hadamard = _fixed("Hadamard", "hadamard", Observable, [1, -1], {
    "diagonalizing_gates": lambda self: [RY(has_params=True, trainable=False, init_params=-np.pi / 4)]})
PauliX = _fixed("PauliX", "paulix", Observable, [1, -1], {"diagonalizing_gates": lambda self: [Hadamard()]})
PauliY = _fixed("PauliY", "pauliy", Observable, [1, -1],
                {"diagonalizing_gates": lambda self: [PauliZ(), S(), Hadamard()]})
PauliZ = _fixed("PauliZ", "pauliz", Observable, [1, -1], {"diagonalizing_gates": lambda self: []})
SHadamard = _fixed("SHadamard", "shadamard")
CHadamard = _fixed("CHadamard", "chadamard")
S = _fixed("S", "s", DiagonalOperation, [1, 1j])
SDG = _fixed("SDG", "sdg", DiagonalOperation, [1, -1j])
T = _fixed("T", "t", DiagonalOperation, [1, np.exp(1j * np.pi / 4)])
TDG = _fixed("TDG", "tdg", DiagonalOperation, [1, np.exp(-1j * np.pi / 4)])
SX = _fixed("SX", "sx", Operation, [1, 1j])
SXDG = _fixed("SXDG", "sxdg", Operation, [1, 1j])
CNOT = _fixed("CNOT", "cnot")
CY = _fixed("CY", "cy")
CZ = _fixed("CZ", "cz", DiagonalOperation, [1, 1, 1, -1])
CCZ = _fixed("CCZ", "ccz", DiagonalOperation, [1, 1, 1, 1, 1, 1, 1, -1])
CS = _fixed("CS", "cs", DiagonalOperation, [1, 1, 1, 1j])
CSDG = _fixed("CSDG", "csdg", DiagonalOperation, [1, 1, 1, -1j])
CSX = _fixed("CSX", "csx")
C3SX = _fixed("C3SX", "c3sx")
C3X = _fixed("C3X", "c3x")
C4X = _fixed("C4X", "c4x")
DCX = _fixed("DCX", "dcx")
SWAP = _fixed("SWAP", "swap")
SSWAP = _fixed("SSWAP", "sswap")
ISWAP = _fixed("ISWAP", "iswap")
CSWAP = _fixed("CSWAP", "cswap")
Toffoli = _fixed("Toffoli", "toffoli")
RCCX = _fixed("RCCX", "rccx")
RC3X = _fixed("RC3X", "rc3x")
ECR = _fixed("ECR", "ecr")

RX = _param("RX", "rx")
RY = _param("RY", "ry")
RZ = _param("RZ", "rz", DiagonalOperation, {
    "_eigvals": classmethod(lambda cls, params: torch.diagonal(mat_dict["rz"](params), dim1=-2, dim2=-1))})
RXX = _param("RXX", "rxx")
RYY = _param("RYY", "ryy")
RZZ = _param("RZZ", "rzz", DiagonalOperation, {
    "_eigvals": classmethod(lambda cls, params: torch.diagonal(mat_dict["rzz"](params), dim1=-2, dim2=-1))})
RZX = _param("RZX", "rzx")
CRX = _param("CRX", "crx")
CRY = _param("CRY", "cry")
CRZ = _param("CRZ", "crz")
PhaseShift = _param("PhaseShift", "phaseshift", DiagonalOperation, {
    "_eigvals": classmethod(lambda cls, params: torch.diagonal(mat_dict["phaseshift"](params), dim1=-2, dim2=-1))})
U1 = _param("U1", "u1", DiagonalOperation, {
    "_eigvals": classmethod(lambda cls, params: torch.diagonal(mat_dict["u1"](params), dim1=-2, dim2=-1))})
CU1 = _param("CU1", "cu1", DiagonalOperation, {
    "_eigvals": classmethod(lambda cls, params: torch.diagonal(mat_dict["cu1"](params), dim1=-2, dim2=-1))})
Rot = _param("Rot", "rot")
CRot = _param("CRot", "crot")
U2 = _param("U2", "u2")
CU2 = _param("CU2", "cu2")
U3 = _param("U3", "u3")
CU3 = _param("CU3", "cu3")
CU = _param("CU", "cu")
# `R` subclasses DiagonalOperation in the reference but overrides _matrix
R = _param("R", "r", DiagonalOperation)
SingleExcitation = _param("SingleExcitation", "singleexcitation")
# the reference points XXMINYY.func / XXPLUSYY.func at the *_matrix builders, which makes the class
# forward unusable there; here the class forward works (intended semantics)
XXMINYY = _param("XXMINYY", "xxminyy")
XXPLUSYY = _param("XXPLUSYY", "xxplusyy")
GlobalPhase = _param("GlobalPhase", "globalphase")


class MultiRZ(DiagonalOperation, metaclass=ABCMeta):
    """Class for the multi-qubit RZ gate (operator/standard_gates/rz.py)."""
    num_params = 1
    num_wires = AnyWires
    op_name = "multirz"
    func = staticmethod(tqf.multirz)

    @classmethod
    def _eigvals(cls, params, n_wires):
        return M.multirz_eigvals(params, n_wires)

    @property
    def eigvals(self):
        return self._eigvals(self.params, self.n_wires)

    @classmethod
    def _matrix(cls, params, n_wires):
        return M.multirz_matrix(params, n_wires)

    @property
    def matrix(self):
        return self._matrix(self.params, self.n_wires)


def _anywire_fixed(name, op_name, builder):
    return ABCMeta(name, (Operation,), {
        "num_params": 0, "num_wires": AnyWires, "op_name": op_name, "func": staticmethod(getattr(tqf, op_name)),
        "_matrix": classmethod(lambda cls, params, n_wires: builder(n_wires)),
        "matrix": property(lambda self: self._matrix(self.params, self.n_wires)),
        "__doc__": f"Class for the {name} gate."})


MultiCNOT = _anywire_fixed("MultiCNOT", "multicnot", M.multicnot_matrix)
MultiXCNOT = _anywire_fixed("MultiXCNOT", "multixcnot", M.multixcnot_matrix)
QFT = _anywire_fixed("QFT", "qft", M.qft_matrix)


class Reset(Operator, metaclass=ABCMeta):
    num_params = 0
    num_wires = AnyWires
    op_name = "reset"
    func = staticmethod(tqf.reset)

    @classmethod
    def _matrix(cls, params):
        return None


class QubitUnitary(Operation, metaclass=ABCMeta):
    num_params = AnyNParams
    num_wires = AnyWires
    op_name = "qubitunitary"
    func = staticmethod(tqf.qubitunitary)

    @classmethod
    def _matrix(cls, params):
        return M.qubitunitary_matrix(params)

    def build_params(self, trainable):
        return None

    def reset_params(self, init_params=None):
        self.params = torch.tensor(init_params, dtype=C_DTYPE)
        self.register_buffer(f"{self.name}_unitary", self.params)


class QubitUnitaryFast(Operation, metaclass=ABCMeta):
    num_params = AnyNParams
    num_wires = AnyWires
    op_name = "qubitunitaryfast"
    func = staticmethod(tqf.qubitunitaryfast)

    def __init__(self, has_params: bool = False, trainable: bool = False, init_params=None, n_wires=None,
                 wires=None):
        super().__init__(has_params=True, trainable=trainable, init_params=init_params, n_wires=n_wires,
                         wires=wires)

    @classmethod
    def from_controlled_operation(cls, op, c_wires, t_wires, trainable):
        """Controlled version of `op` (operator/standard_gates/qubit_unitary.py:57-143): identity unless
        all control wires are 1; then kron copies of op.matrix act on the target sets."""
        c_wires, t_wires = np.array(c_wires), np.array(t_wires)
        orig_u = op.matrix
        wires = [c_wires.item()] if c_wires.ndim == 0 else [int(w) for w in c_wires]
        n_c = len(wires)
        t_flat = [t_wires.item()] if t_wires.ndim == 0 else [int(w) for w in t_wires.flatten()]
        wires += t_flat
        # (the operator's own n_wires may be None for fixed gates such as tq.PauliX(): its matrix tells the arity)
        n_sets = max(len(t_flat) // int(np.log2(orig_u.shape[-1])), 1)
        block = orig_u
        for _ in range(n_sets - 1):
            block = torch.kron(block, orig_u)
        n_wires = n_c + len(t_flat)
        unitary = torch.eye(2 ** n_wires, dtype=C_DTYPE)
        unitary[2 ** n_wires - block.shape[-1]:, 2 ** n_wires - block.shape[-1]:] = block.detach()
        return cls(has_params=True, trainable=trainable, init_params=unitary, n_wires=n_wires, wires=wires)

    @classmethod
    def _matrix(cls, params):
        return M.qubitunitaryfast_matrix(params)

    def build_params(self, trainable):
        return None

    def reset_params(self, init_params=None):
        self.params = torch.as_tensor(init_params).to(C_DTYPE).clone()
        self.register_buffer(f"{self.name}_unitary", self.params)


class TrainableUnitary(Operation, metaclass=ABCMeta):
    num_params = AnyNParams
    num_wires = AnyWires
    op_name = "trainableunitary"
    func = staticmethod(tqf.qubitunitaryfast)

    def build_params(self, trainable):
        parameters = nn.Parameter(torch.empty(1, 2 ** self.n_wires, 2 ** self.n_wires, dtype=C_DTYPE))
        parameters.requires_grad = True if trainable else False
        return parameters

    def reset_params(self, init_params=None):
        mat = torch.randn((1, 2 ** self.n_wires, 2 ** self.n_wires), dtype=C_DTYPE)
        u, _, v = torch.svd(mat)
        self.params.data.copy_(u.matmul(v.permute(0, 2, 1)))

    @classmethod
    def _matrix(cls, params):
        return M.qubitunitaryfast_matrix(params)


class TrainableUnitaryStrict(TrainableUnitary, metaclass=ABCMeta):
    op_name = "trainableunitarystrict"
    func = staticmethod(tqf.qubitunitarystrict)


# aliases of the reference (classes with the same behaviour under another name)
H, SH, CH = Hadamard, SHadamard, CHadamard
EchoedCrossResonance = ECR
CCX = Toffoli
U = U3

all_variables = [
    EchoedCrossResonance, ECR, GlobalPhase, I, ISWAP, PauliX, CNOT, C4X, C3X, DCX, MultiCNOT, MultiXCNOT, PauliY, CY,
    PauliZ, CZ, CCZ, Hadamard, SHadamard, CHadamard, H, SH, CH, PhaseShift, QFT, R, Reset, Rot, CRot, RX, RXX, CRX,
    RY, RYY, CRY, RZ, MultiRZ, RZZ, RZX, CRZ, Toffoli, CCX, RC3X, RCCX, S, SDG, CS, CSDG, SingleExcitation, SWAP,
    SSWAP, CSWAP, SX, CSX, C3SX, SXDG, T, TDG, TrainableUnitary, TrainableUnitaryStrict, U1, CU1, U2, CU2, U3, CU3,
    CU, U, XXMINYY, XXPLUSYY,
]

op_name_dict = {x.op_name: x for x in all_variables}
op_name_dict.update({
    "h": H, "sh": SH, "u": U, "qubitunitary": QubitUnitary, "qubitunitaryfast": QubitUnitaryFast, "x": PauliX,
    "y": PauliY, "z": PauliZ, "cx": CNOT, "xx": RXX, "yy": RYY, "zz": RZZ, "zx": RZX, "ccx": Toffoli, "p": U1,
    "cp": CU1, "cr": CU1,
})

fixed_ops = sorted({a.__name__ for a in all_variables if a.num_params == 0})
parameterized_ops = sorted({a.__name__ for a in all_variables if a.num_params != 0})
Operator.fixed_ops = fixed_ops
Operator.parameterized_ops = parameterized_ops

__all__ += sorted({a.__name__ for a in all_variables}) + ["H", "SH", "CH", "CCX", "U", "EchoedCrossResonance",
                                                          "QubitUnitary", "QubitUnitaryFast"]
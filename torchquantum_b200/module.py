"""QuantumModule / QuantumModuleList / QuantumModuleDict — drop-in for module/modules.py:36-350.

Static mode: the reference records ops into a QuantumGraph and multiplies them into
`wires_per_block`-qubit unitaries (graph/graphs.py).  Here every gate is already queued on the
device and fused by the engine's planner, so `static_on(wires_per_block=...)` only sets the flags
the reference's user code reads; results are identical in both modes by construction.
"""
from abc import ABCMeta
from collections.abc import Iterable

import torch.nn as nn

__all__ = ["QuantumModule", "QuantumModuleList", "QuantumModuleDict", "static_support", "QuantumGraph"]


class QuantumGraph:
    """Placeholder with the attributes user code touches (`q_device`, `device`, `is_list_finish`);
    fusion happens in the device's gate queue, not here."""

    def __init__(self):
        self.is_list_finish = False
        self.module_list = []
        self.func_list = []
        self.q_device = None
        self.device = None

    def add_op(self, op):  # kept for API compatibility; never needed
        self.module_list.append(op)

    def add_func(self, **kwargs):
        self.func_list.append(kwargs)


def static_support(f):
    """Decorator of the reference (graph/graphs.py:60-82).  It stores the device on the module (user
    code reads `self.q_device`) and runs the wrapped forward: gates queue on the device and are
    fused when the state is read, in static and dynamic mode alike."""
    import functools

    @functools.wraps(f)
    def forward_register_graph(*args, **kwargs):
        module, q_device = args[0], (args[1] if len(args) > 1 else kwargs.get("q_device", kwargs.get("qdev")))
        if getattr(module, "static_mode", False) and getattr(module, "is_graph_top", False):
            module.q_device = q_device
        return f(*args, **kwargs)

    return forward_register_graph


class QuantumModule(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.static_mode = False
        self.graph = None
        self.parent_graph = None
        self.is_graph_top = False
        self.unitary = None
        self.wires = None
        self.n_wires = None
        self.q_device = None
        self.device = None
        self.wires_per_block = None
        self.qiskit_processor = None
        self.noise_model_tq = None
        self.Operator_list = None

    def load_op_history(self, op_history):
        """Rebuild operators from the op-list wire format (module/modules.py:92-134)."""
        from . import operator as op_mod

        ops = []
        for op in op_history:
            cls = op_mod.op_name_dict[op["name"]]
            trainable = op.get("trainable", False)
            has_params = (op.get("params", None) is not None) or trainable
            n_wires = len(op["wires"]) if isinstance(op["wires"], Iterable) else 1
            ops.append(cls(has_params=has_params, trainable=trainable, init_params=op.get("params", None),
                           n_wires=n_wires, wires=op["wires"], inverse=op.get("inverse", False)))
        self.Operator_list = QuantumModuleList(ops)

    @classmethod
    def from_op_history(cls, op_history):
        qmodule = cls()
        qmodule.load_op_history(op_history)
        qmodule.forward = qmodule.forward_Operators_list
        return qmodule

    def forward_Operators_list(self, qdev):
        assert self.Operator_list is not None, "Operator_list should not contain nothing"
        for oper in self.Operator_list:
            oper(qdev)

    def set_noise_model_tq(self, noise_model_tq):
        for module in self.modules():
            module.noise_model_tq = noise_model_tq

    def set_qiskit_processor(self, processor):
        for module in self.modules():
            module.qiskit_processor = processor

    def set_wires_per_block(self, wires_per_block):
        self.wires_per_block = wires_per_block

    def static_on(self, is_graph_top=True, wires_per_block=3):
        from .device import QuantumDevice

        self.wires_per_block = wires_per_block
        self.static_mode = True
        self.is_graph_top = is_graph_top
        if self.graph is None:
            self.graph = QuantumGraph()
        for module in self.children():
            if isinstance(module, (nn.ModuleList, nn.ModuleDict)):
                module.graph = self.graph
            module.parent_graph = self.graph
            if not isinstance(module, QuantumDevice) and hasattr(module, "static_on"):
                module.static_on(is_graph_top=False)

    def static_off(self):
        from .device import QuantumDevice

        self.static_mode = False
        self.graph = None
        for module in self.children():
            if not isinstance(module, QuantumDevice) and hasattr(module, "static_off"):
                module.static_off()

    def set_graph_build_finish(self):
        if self.graph is not None:
            self.graph.is_list_finish = True

    def static_forward(self, q_device):
        self.q_device = q_device
        self.device = q_device.state.device
        q_device.flush()

    def get_unitary(self, x=None, device=None):
        """Unitary of the module: run it on the identity batch (module/modules.py:326-350)."""
        from .device import QuantumDevice

        assert self.n_wires is not None, "n_wires should not be None, specify it in the Quantum Module " \
                                         "before calling get_unitary()"
        if device is None:
            p = next(self.parameters(), None)
            device = p.device if p is not None else (x.device if x is not None else "cuda")
        qdev = QuantumDevice(n_wires=self.n_wires, device=device)
        qdev.reset_identity_states()
        if x is None:
            self.forward(qdev)
        else:
            self.forward(qdev, x)
        return qdev.get_states_1d().T


# Global mutation epoch: bumped by every structural change the layer templates' record caches validate against (an
# attribute write on an operator or a template layer, a mutation of a QuantumModuleList, `.to()` on an operator).  While
# it stands still, a whole template (LayerTemplate0._chunk) can reuse its cached records without re-walking its layers.
MUTATION_EPOCH = [0]


class QuantumModuleList(nn.ModuleList, QuantumModule, metaclass=ABCMeta):
    """nn.ModuleList of quantum modules.  Every mutation bumps `_list_ver`, which the layer templates' queue-record
    caches (layer._RecordCache) compare instead of re-walking the list every step."""

    def __init__(self, *args, **kwargs):
        self.__dict__["_list_ver"] = 0
        super().__init__(*args, **kwargs)

    def _touch(self):
        self.__dict__["_list_ver"] = self.__dict__.get("_list_ver", 0) + 1
        MUTATION_EPOCH[0] += 1

    def __setitem__(self, idx, module):
        self._touch()
        return super().__setitem__(idx, module)

    def __delitem__(self, idx):
        self._touch()
        return super().__delitem__(idx)

    def __iadd__(self, modules):
        self._touch()
        return super().__iadd__(modules)

    def append(self, module):
        self._touch()
        return super().append(module)

    def extend(self, modules):
        self._touch()
        return super().extend(modules)

    def insert(self, index, module):
        self._touch()
        return super().insert(index, module)

    def pop(self, key):
        self._touch()
        return super().pop(key)


class QuantumModuleDict(nn.ModuleDict, QuantumModule, metaclass=ABCMeta):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

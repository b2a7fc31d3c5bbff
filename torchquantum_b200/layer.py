"""Circuit-template layers used by the BASELINE configs — drop-in for the corresponding classes of
torchquantum/layer (RandomLayer, Op1QAllLayer, Op2QAllLayer, Op2QDenseLayer, LayerTemplate0,
NLocal, TwoLocal, EfficientSU2, RealAmplitudes, QFTLayer).  Pure gate-list builders: they call
operators, which queue gates on the device."""
from collections.abc import Iterable

import numpy as np
import torch

from . import operator as op
from .module import MUTATION_EPOCH, QuantumModule, QuantumModuleList, static_support
from .operator import op_name_dict

__all__ = ["RandomLayer", "RandomLayerAllTypes", "Op1QAllLayer", "Op2QAllLayer", "Op2QDenseLayer",
           "LayerTemplate0", "NLocal", "TwoLocal", "EfficientSU2", "RealAmplitudes", "QFTLayer"]


class _RecordCache:
    """Template layers queue the same records (same Parameter objects, wires, inverse flags) every step.
    `records(q_device)` returns them from a cache that stays valid while (a) the layer's operator list is the same object
    with the same mutation counter (QuantumModuleList._list_ver), (b) no attribute of the layer itself was assigned
    (`__setattr__` below drops the cache: a new op list, another jump, ...), (c) the operators' version counters did not
    move (any attribute write on an operator bumps its version) and (d) the parameters still live on the device; it
    returns None whenever some operator needs its full `forward` (noise model, op recording, ...)."""
    _CACHE_KEEPS = frozenset(("q_device", "device", "static_mode", "graph", "parent_graph", "is_graph_top", "training"))

    def _layer_ops_wires(self):
        raise NotImplementedError

    def _op_container(self):
        """the QuantumModuleList this layer's operators live in"""
        raise NotImplementedError

    def __setattr__(self, name, value):
        if name not in self._CACHE_KEEPS:
            self.__dict__["_rec_cache"] = None
            MUTATION_EPOCH[0] += 1
        super().__setattr__(name, value)

    def records(self, q_device):
        if q_device.record_op or not getattr(q_device, "accepts_record_chunks", False):
            return None
        dev = q_device._flat.device
        c = self.__dict__.get("_rec_cache")
        if c is not None and c[0] == dev and c[7] == q_device.n_wires:   # (wire indices were validated for that width)
            cont = self._op_container()
            if cont is c[5] and cont.__dict__["_list_ver"] == c[6]:
                ver = 0
                for o in c[2]:
                    ver += o.__dict__["_ver"]
                if ver == c[1] and all(p.device == dev for p in c[4]):
                    return c[3]
        ops, recs, ver = [], [], 0
        cont = self._op_container()
        for o, wires in self._layer_ops_wires():
            r = o.fast_record(q_device, wires)
            if r is None:
                self.__dict__["_rec_cache"] = None
                return None
            if wires is not o.wires:
                o.__dict__["wires"] = wires
            ops.append(o)
            recs.append(r)
            ver += o.__dict__["_ver"]
        self.__dict__["_rec_cache"] = (dev, ver, ops, recs, [r[2] for r in recs if r[2] is not None], cont,
                                       cont.__dict__["_list_ver"], q_device.n_wires)
        return recs


class Op1QAllLayer(QuantumModule, _RecordCache):
    """The same one-qubit operation on every wire (layer/layers/layers.py:71-97)."""

    def __init__(self, op, n_wires: int, has_params=False, trainable=False):
        super().__init__()
        self.n_wires = n_wires
        self.op = op
        self.ops_all = QuantumModuleList()
        for _ in range(n_wires):
            self.ops_all.append(op(has_params=has_params, trainable=trainable))

    def _layer_ops_wires(self):
        return [(op_k, k) for k, op_k in enumerate(self.ops_all)]

    def _op_container(self):
        return self.ops_all

    @static_support
    def forward(self, q_device):
        recs = self.records(q_device)
        if recs is not None:
            q_device._queue.extend(recs)
            return
        for k, op_k in enumerate(self.ops_all):
            op_k.queue_on(q_device, k)


class Op2QAllLayer(QuantumModule, _RecordCache):
    """The same two-qubit operation on wire pairs [k, k+jump] (layer/entanglement/op2_layer.py:43-110)."""

    def __init__(self, op, n_wires: int, has_params=False, trainable=False, wire_reverse=False, jump=1,
                 circular=False):
        super().__init__()
        self.n_wires = n_wires
        self.jump = jump
        self.circular = circular
        self.op = op
        self.ops_all = QuantumModuleList()
        self.wire_reverse = wire_reverse
        n_ops = n_wires if circular else n_wires - jump
        for _ in range(n_ops):
            self.ops_all.append(op(has_params=has_params, trainable=trainable))

    def _layer_ops_wires(self):
        out = []
        for k, op_k in enumerate(self.ops_all):
            wires = [k, (k + self.jump) % self.n_wires]
            if self.wire_reverse:
                wires.reverse()
            out.append((op_k, wires))
        return out

    def _op_container(self):
        return self.ops_all

    @static_support
    def forward(self, q_device):
        recs = self.records(q_device)
        if recs is not None:
            q_device._queue.extend(recs)
            return
        for op_k, wires in self._layer_ops_wires():
            op_k.queue_on(q_device, wires)


class Op2QDenseLayer(QuantumModule):
    """The same two-qubit operation on every pair of wires (op2_layer.py `Op2QDenseLayer`)."""

    def __init__(self, op, n_wires: int, has_params=False, trainable=False, wire_reverse=False):
        super().__init__()
        self.n_wires = n_wires
        self.op = op
        self.ops_all = QuantumModuleList()
        self.wire_reverse = wire_reverse
        for _ in range(self.n_wires * (self.n_wires - 1) // 2):
            self.ops_all.append(op(has_params=has_params, trainable=trainable))

    @static_support
    def forward(self, q_device):
        k = 0
        for i in range(self.n_wires - 1):
            for j in range(i + 1, self.n_wires):
                wires = [i, j]
                if self.wire_reverse:
                    wires.reverse()
                self.ops_all[k](q_device, wires=wires)
                k += 1


class RandomLayer(QuantumModule, _RecordCache):
    """Random op soup (layer/layers/random_layers.py:80-229).  Determinism is the NumPy global RNG
    call order of the reference: per op one `np.random.choice(op_types, p=ratios)` then one
    `np.random.choice(wires, size=k, replace=False)`."""

    def __init__(self, wires, n_ops=None, n_params=None, op_ratios=None,
                 op_types=(op.RX, op.RY, op.RZ, op.CNOT), seed=None, qiskit_compatible=False):
        super().__init__()
        self.n_ops = n_ops
        self.n_params = n_params
        assert n_params is not None or n_ops is not None
        self.wires = wires if isinstance(wires, Iterable) else [wires]
        self.n_wires = len(wires)
        op_types = op_types if isinstance(op_types, Iterable) else [op_types]
        if op_ratios is None:
            op_ratios = [1] * len(op_types)
        else:
            op_ratios = op_ratios if isinstance(op_ratios, Iterable) else [op_ratios]
        self.op_types = list(op_types)
        self.op_ratios = np.array(op_ratios) / sum(op_ratios)
        self.seed = seed
        self.op_list = QuantumModuleList()
        if seed is not None:
            np.random.seed(seed)
        self.build_random_layer()

    def rebuild_random_layer_from_op_list(self, n_ops_in, wires_in, op_list_in):
        self.n_ops = n_ops_in
        self.wires = wires_in
        self.op_list = QuantumModuleList()
        for op_in in op_list_in:
            self.op_list.append(op_name_dict[op_in.name.lower()](
                has_params=op_in.has_params, trainable=op_in.trainable, wires=op_in.wires, n_wires=op_in.n_wires))

    def build_random_layer(self):
        op_cnt = param_cnt = 0
        while True:
            cls = np.random.choice(self.op_types, p=self.op_ratios)
            n_op_wires = cls.num_wires
            if n_op_wires > self.n_wires:
                continue
            any_wire = n_op_wires == -1
            if any_wire:
                n_op_wires = self.n_wires
            op_wires = list(np.random.choice(self.wires, size=n_op_wires, replace=False))
            if any_wire:
                if cls.__name__ == "MultiRZ":
                    operation = cls(has_params=True, trainable=True, n_wires=n_op_wires, wires=op_wires)
                else:
                    operation = cls(n_wires=n_op_wires, wires=op_wires)
            elif cls.__name__ in op.parameterized_ops:
                operation = cls(has_params=True, trainable=True, wires=op_wires)
            else:
                operation = cls(wires=op_wires)
            self.op_list.append(operation)
            op_cnt += 1
            param_cnt += cls.num_params
            if self.n_ops is not None and op_cnt == self.n_ops:
                break
            elif self.n_ops is None and self.n_params is not None:
                if param_cnt == self.n_params:
                    break
                elif param_cnt > self.n_params:
                    op_cnt -= 1
                    param_cnt -= cls.num_params
                    del self.op_list[-1]

    def _layer_ops_wires(self):
        return [(operation, operation.wires) for operation in self.op_list]

    def _op_container(self):
        return self.op_list

    @static_support
    def forward(self, q_device):
        recs = self.records(q_device)
        if recs is not None:
            q_device._queue.extend(recs)
            return
        for operation in self.op_list:
            operation(q_device)


class RandomLayerAllTypes(RandomLayer):
    """RandomLayer over (almost) every gate type (random_layers.py `RandomLayerAllTypes`)."""

    def __init__(self, wires, n_ops=None, n_params=None, op_ratios=None,
                 op_types=(op.Hadamard, op.SHadamard, op.PauliX, op.PauliY, op.PauliZ, op.S, op.T, op.SX, op.CNOT,
                           op.CZ, op.CY, op.RX, op.RY, op.RZ, op.RZZ, op.SWAP, op.CSWAP, op.Toffoli, op.PhaseShift,
                           op.Rot, op.MultiRZ, op.CRX, op.CRY, op.CRZ, op.CRot, op.U1, op.U2, op.U3, op.MultiCNOT,
                           op.MultiXCNOT),
                 seed=None, qiskit_compatible=False):
        super().__init__(wires=wires, n_ops=n_ops, n_params=n_params, op_ratios=op_ratios, op_types=op_types,
                         seed=seed, qiskit_compatible=qiskit_compatible)


class LayerTemplate0(QuantumModule):
    name = None

    def __init__(self, arch: dict = None):
        super().__init__()
        self.n_wires = arch["n_wires"]
        self.arch = arch
        self.n_blocks = arch.get("n_blocks", None)
        self.n_layers_per_block = arch.get("n_layers_per_block", None)
        self.layers_all = self.build_layers()

    def build_layers(self):
        raise NotImplementedError

    def __setattr__(self, name, value):
        if name not in _RecordCache._CACHE_KEEPS:    # (forward() assigns `q_device` every step)
            MUTATION_EPOCH[0] += 1
        super().__setattr__(name, value)

    def _chunk(self, q_device):
        """All layers' records as one engine.RecordChunk when every layer is a cacheable template."""
        c = self.__dict__.get("_chunk_cache")
        if c is not None and c[2] == MUTATION_EPOCH[0]:
            # nothing any of the caches below validates against has changed since this chunk was last validated (no
            # attribute write on an operator or layer, no list mutation, no `.to()`): skip the walk over the layers.
            # What has no epoch is checked here: hooks registered on a layer, the device the parameters live on
            if (not q_device.record_op and getattr(q_device, "accepts_record_chunks", False) and not any(c[3])
                    and c[4] == q_device.n_wires):
                dev = q_device._flat.device
                if all(p.device == dev for p in c[1].params):
                    return c[1]
        epoch = MUTATION_EPOCH[0]
        parts = []
        for layer in self.layers_all:
            if not isinstance(layer, _RecordCache) or layer._forward_hooks or layer._forward_pre_hooks:
                return None
            recs = layer.records(q_device)
            if recs is None:
                return None
            parts.append(recs)
        hooks = [h for layer in self.layers_all for h in (layer._forward_hooks, layer._forward_pre_hooks)]
        if c is not None and len(c[0]) == len(parts) and all(a is b for a, b in zip(c[0], parts)):
            self.__dict__["_chunk_cache"] = (c[0], c[1], epoch, hooks, q_device.n_wires)
            return c[1]
        from .engine import RecordChunk

        chunk = RecordChunk([r for recs in parts for r in recs])
        self.__dict__["_chunk_cache"] = (parts, chunk, epoch, hooks, q_device.n_wires)
        return chunk

    @static_support
    def forward(self, q_device):
        self.q_device = q_device
        chunk = self._chunk(q_device)
        if chunk is not None:
            q_device._queue.append(chunk)
            return
        for k in range(len(self.layers_all)):
            self.layers_all[k](q_device)


class NLocal(LayerTemplate0):
    """Alternating rotation / entanglement blocks (layer/nlocal/nlocal.py:37-134)."""

    def __init__(self, rotation_ops=None, entanglement_ops=None, arch=None, rotation_layer=Op1QAllLayer,
                 entanglement_layer=Op2QAllLayer, reps=1, rotation_layer_params=None,
                 entanglement_layer_params=None, initial_circuit=None, skip_final_rotation_layer=False):
        self.rotation_ops = rotation_ops
        self.rotation_layer = rotation_layer
        self.rotation_layer_params = rotation_layer_params or {}
        self.entanglement_ops = entanglement_ops
        self.entanglement_layer = entanglement_layer
        self.entanglement_layer_params = entanglement_layer_params or {}
        self.initial_circuit = initial_circuit
        self.skip_final_rotation_layer = skip_final_rotation_layer
        self.reps = reps
        super().__init__(arch)

    def build_rotation_block(self):
        return [self.rotation_layer(op=rot, n_wires=self.n_wires, **self.rotation_layer_params)
                for rot in self.rotation_ops]

    def build_entanglement_block(self):
        return [self.entanglement_layer(op=ent, n_wires=self.n_wires, **self.entanglement_layer_params)
                for ent in self.entanglement_ops]

    def build_layers(self):
        layers_all = QuantumModuleList()
        if self.initial_circuit is not None:
            layers_all.append(self.initial_circuit)
        for _ in range(self.reps):
            layers_all.extend(self.build_rotation_block())
            layers_all.extend(self.build_entanglement_block())
        if not self.skip_final_rotation_layer:
            layers_all.extend(self.build_rotation_block())
        return layers_all


def _as_op_list(ops):
    if isinstance(ops, str):
        return [op_name_dict[ops]]
    if isinstance(ops, Iterable):
        return [op_name_dict[o] if isinstance(o, str) else o for o in ops]
    return [ops]


class TwoLocal(NLocal):
    """layer/nlocal/two_local.py:42-111."""

    def __init__(self, n_wires=1, rotation_ops=None, entanglement_ops=None, rotation_layer=Op1QAllLayer,
                 entanglement_layer="linear", reps=1, entanglement_layer_params=None, initial_circuit=None,
                 skip_final_rotation_layer=False):
        entanglement_layer_params = entanglement_layer_params or {}
        if entanglement_layer == "linear":
            entanglement_layer = Op2QAllLayer
        elif entanglement_layer == "reverse_linear":
            entanglement_layer, entanglement_layer_params = Op2QAllLayer, {"wire_reverse": True}
        elif entanglement_layer == "circular":
            entanglement_layer, entanglement_layer_params = Op2QAllLayer, {"circular": True}
        elif entanglement_layer == "full":
            entanglement_layer = Op2QDenseLayer
        super().__init__(arch={"n_wires": n_wires}, rotation_ops=_as_op_list(rotation_ops),
                         rotation_layer=rotation_layer,
                         rotation_layer_params={"has_params": True, "trainable": True},
                         entanglement_ops=_as_op_list(entanglement_ops), entanglement_layer=entanglement_layer,
                         entanglement_layer_params=entanglement_layer_params, initial_circuit=initial_circuit,
                         reps=reps, skip_final_rotation_layer=skip_final_rotation_layer)


class EfficientSU2(TwoLocal):
    """RY + RZ rotation layers with CNOT entanglement (layer/nlocal/efficient_su2.py:33-58)."""

    def __init__(self, n_wires=None, entanglement_layer="reverse_linear", reps=3, skip_final_rotation_layer=False):
        super().__init__(n_wires=n_wires, rotation_ops=[op.RY, op.RZ], entanglement_ops=[op.CNOT],
                         entanglement_layer=entanglement_layer, reps=reps,
                         skip_final_rotation_layer=skip_final_rotation_layer)


class RealAmplitudes(TwoLocal):
    """RY rotation layers with CNOT entanglement (layer/nlocal/real_amplitudes.py)."""

    def __init__(self, n_wires=None, entanglement_layer="reverse_linear", reps=3, skip_final_rotation_layer=False):
        super().__init__(n_wires=n_wires, rotation_ops=[op.RY], entanglement_ops=[op.CNOT],
                         entanglement_layer=entanglement_layer, reps=reps,
                         skip_final_rotation_layer=skip_final_rotation_layer)


class QFTLayer(QuantumModule):
    """H + CU1(pi/2^d) ladder (+ swaps) (layer/layers/qft_layer.py:11-115)."""

    def __init__(self, n_wires=None, wires=None, do_swaps=True, inverse=False):
        super().__init__()
        assert n_wires is not None or wires is not None
        if n_wires is None:
            n_wires = len(wires)
        if wires is None:
            wires = range(n_wires)
        self.n_wires = n_wires
        self.wires = wires
        self.do_swaps = do_swaps
        self.gates_all = self.build_inverse_circuit() if inverse else self.build_circuit()

    def _swaps(self):
        return [{"name": "swap", "wires": [self.wires[w], self.wires[self.n_wires - w - 1]]}
                for w in range(self.n_wires // 2)] if self.do_swaps else []

    def build_circuit(self):
        ops = []
        for top in range(self.n_wires):
            ops.append({"name": "hadamard", "wires": self.wires[top]})
            for w in range(top + 1, self.n_wires):
                ops.append({"name": "cu1", "params": torch.pi / (2 ** (w - top)),
                            "wires": [self.wires[w], self.wires[top]]})
        return QuantumModule.from_op_history(ops + self._swaps())

    def build_inverse_circuit(self):
        ops = self._swaps()
        for top in range(self.n_wires)[::-1]:
            for w in range(top + 1, self.n_wires)[::-1]:
                ops.append({"name": "cu1", "params": -torch.pi / (2 ** (w - top)),
                            "wires": [self.wires[w], self.wires[top]]})
            ops.append({"name": "hadamard", "wires": self.wires[top]})
        return QuantumModule.from_op_history(ops)

    @static_support
    def forward(self, q_device):
        self.gates_all(q_device)

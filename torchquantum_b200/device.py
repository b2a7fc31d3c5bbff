"""QuantumDevice — drop-in for torchquantum/device/devices.py:36-136.

Same constructor, buffers (`state`, `states`), reset/clone/set helpers and gate methods, but gates
are *queued* and run as fused passes when the state is read (`.states`, `get_states_1d()`,
measurement): the lazy queue is what lets many gates share one trip through HBM.
The exposed `states` tensor is always contiguous `[bsz] + [2] * n_wires` complex64.
"""
import weakref
from typing import Union

import numpy as np
import torch
import torch.nn as nn

from . import engine
from .macro import C_DTYPE

__all__ = ["QuantumDevice"]


class GateQueue(list):
    """The lazy gate queue of a device.  Two pieces of bookkeeping keep lazy execution equal to the reference's
    eager one (functional/gate_wrapper.py:324-446 consumes parameter values and the autograd mode at call time):

    * `grad`: the autograd mode the queued gates were *issued* under.  A gate issued under another mode first
      runs what is queued (under the mode it was issued with), so `with torch.no_grad(): qdev.get_states_1d()` in
      the middle of a circuit neither drops the graph of the gates before it nor tracks gates issued inside it;
    * `guard`: (tensor, version) of every parameter tensor queued through the public gate calls; `check()` raises
      when one was modified in place between the gate call and the flush (leaf tensors without grad are snapshotted
      at the call instead, see functional._snapshot)."""
    __slots__ = ("owner", "grad", "guard")

    def __init__(self, owner):
        super().__init__()
        # weak: a device <-> queue cycle would keep a dropped device (its state, and the autograd graph behind it)
        # alive until the cycle collector runs
        self.owner, self.grad, self.guard = weakref.ref(owner), torch.is_grad_enabled(), []

    def _mode(self):
        g = torch.is_grad_enabled()
        if g != self.grad:
            if len(self):
                owner = self.owner()
                owner.flush()
                owner._queue._mode()
                return owner._queue
            self.grad = g
        return self

    def append(self, rec):
        q = self._mode()
        list.append(q, rec)
        if type(rec) is tuple:
            if rec[2] is not None:
                q.guard.append((rec[2], rec[2]._version))
        else:   # engine.RecordChunk of a layer template: its Parameters are queued by reference
            q.guard.append((rec.params, [p._version for p in rec.params]))

    def extend(self, recs):
        q = self._mode()
        list.extend(q, recs)
        ps = [r[2] for r in recs if r[2] is not None]
        if ps:
            q.guard.append((ps, [p._version for p in ps]))

    def check(self):
        for t, v in self.guard:
            if (t._version != v) if type(v) is int else ([p._version for p in t] != v):
                raise RuntimeError(
                    "a gate parameter tensor was modified in place after the gate was queued and before the state "
                    "was read (gates run lazily): read the state (`qdev.states`) before modifying it, or pass a clone")


class QuantumDevice(nn.Module):
    accepts_record_chunks = True  # the queue may hold engine.RecordChunk items (layer templates)

    def __init__(self, n_wires: int, device_name: str = "default", bsz: int = 1,
                 device: Union[torch.device, str] = "cpu", record_op: bool = False):
        super().__init__()
        # plain python state goes straight into __dict__ (what nn.Module.__setattr__ does after its type dispatch): user code
        # builds a device per forward pass, as the reference's examples do
        self.__dict__.update(n_wires=n_wires, device_name=device_name, bsz=bsz, device=device, record_op=record_op,
                             _op_history=[], _history_pending=[], _flat=None)
        _state = torch.zeros(2 ** n_wires, dtype=C_DTYPE, device=device)  # built on the device: no H2D copy
        _state[:1].fill_(1)
        _state = torch.reshape(_state, [2] * n_wires)
        self.register_buffer("state", _state)
        self.__dict__["_queue"] = GateQueue(self)
        self._buffers["states"] = None
        self.reset_states(bsz)

    # --- state container --------------------------------------------------------------
    def _set_flat(self, flat):
        d = self.__dict__
        d["_queue"] = GateQueue(self)
        d["_flat"] = flat
        self._buffers["states"] = None  # materialised lazily as a view

    @property
    def states(self):
        self.flush()
        v = self._flat.reshape([self._flat.shape[0]] + [2] * self.n_wires)
        self._buffers["states"] = v
        return v

    @states.setter
    def states(self, value):
        bsz = value.shape[0]
        self._set_flat(value.reshape(bsz, 2 ** self.n_wires))

    def __setattr__(self, name, value):
        # `states` is a registered buffer name (checkpoint compatibility), so nn.Module would
        # store the tensor itself; route it through the property instead
        if name == "states":
            type(self).states.fset(self, value)
        else:
            super().__setattr__(name, value)

    def _apply(self, fn, *args, **kwargs):
        self.flush()
        self._buffers["states"] = None
        super()._apply(fn, *args, **kwargs)
        self._flat = fn(self._flat)
        self.device = self.state.device
        return self

    @property
    def compute_device(self):
        return self._flat.device

    def flush(self):
        """Run every queued gate (fused) and materialise the state."""
        if self._queue:
            q, self._queue = self._queue, GateQueue(self)
            q.check()
            if q.grad == torch.is_grad_enabled():
                self._flat = engine.run_circuit(self._flat, self.n_wires, q)
            else:
                with torch.set_grad_enabled(q.grad):
                    self._flat = engine.run_circuit(self._flat, self.n_wires, q)

    # --- checkpoints: `states` is always a key of the state dict, as in the reference (device/devices.py:57-71) ----
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        self.states  # flush + materialise the [bsz, 2, ..., 2] view as the `states` buffer
        super()._save_to_state_dict(destination, prefix, keep_vars)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        key = prefix + "states"
        if key in state_dict:
            v = state_dict[key]
            if v.numel() % (2 ** self.n_wires) != 0:
                error_msgs.append(f"size mismatch for {key}: {tuple(v.shape)} is not [bsz] + [2] * {self.n_wires}")
            else:
                self.states = v.detach().to(device=self._flat.device, dtype=C_DTYPE).reshape(-1, 2 ** self.n_wires).clone()
                self.bsz = self._flat.shape[0]
            state_dict = {k: t for k, t in state_dict.items() if k != key}
        elif strict:
            missing_keys.append(key)
        self.flush()
        self._buffers["states"] = None
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def enqueue(self, name, wires, params, inverse):
        self._queue.append((name, wires, params, inverse))

    def reset_op_history(self):
        self._op_history = []
        self._history_pending = []

    @property
    def op_history(self):
        # the reference converts params to python lists per gate (a D2H sync each,
        # functional/gate_wrapper.py:380-391); here the conversion is deferred to the first read
        for entry in self._history_pending:
            p = entry["params"]
            if p is not None:
                entry["params"] = p.squeeze().detach().cpu().numpy().tolist()
        self._history_pending = []
        return self._op_history

    @op_history.setter
    def op_history(self, value):
        self._op_history = value
        self._history_pending = []

    def _record(self, entry):
        self._op_history.append(entry)
        self._history_pending.append(entry)

    def clone_states(self, existing_states: torch.Tensor):
        self.states = existing_states.clone()

    def set_states(self, states: torch.Tensor):
        self.states = states

    def reset_states(self, bsz: int):
        flat = torch.zeros(bsz, 2 ** self.n_wires, dtype=C_DTYPE, device=self.state.device)
        flat[:, 0].fill_(1)  # (a fill kernel; `flat[:, 0] = 1` uploads a host scalar, which a CUDA-graph capture rejects)
        self._set_flat(flat)

    def reset_identity_states(self):
        self._set_flat(torch.eye(2 ** self.n_wires, device=self.state.device, dtype=C_DTYPE))

    def reset_all_eq_states(self, bsz: int):
        energy = np.sqrt(1 / (2 ** self.n_wires) / 2)
        flat = torch.ones(bsz, 2 ** self.n_wires, dtype=C_DTYPE, device=self.state.device) * (energy + energy * 1j)
        self._set_flat(flat)

    def get_states_1d(self):
        self.flush()
        return self._flat

    def get_state_1d(self):
        return torch.reshape(self.state, [2 ** self.n_wires])

    @property
    def name(self):
        return self.__class__.__name__

    def __repr__(self):
        return (f" class: {self.name} \n device name: {self.device_name} \n number of qubits: {self.n_wires} \n"
                f" batch size: {self.bsz} \n current computing device: {self.state.device} \n"
                f" recording op history: {self.record_op} \n"
                f" current states: {repr(self.get_states_1d().cpu().detach().numpy())}")


def _bind_functionals():
    from .functional import func_name_dict, func_name_dict_collect

    for name, func in func_name_dict.items():
        setattr(QuantumDevice, name, func)
    for name, func in func_name_dict_collect.items():
        setattr(QuantumDevice, name, func)


_bind_functionals()

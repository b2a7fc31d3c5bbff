"""Gradient estimators that do not use backpropagation.

* `parameter_shift_gradient` — drop-in for torchquantum/util/utils.py:1017-1075: for every trainable tensor of the model,
  0.5 * (<O>(theta + s) - <O>(theta - s)) with both expectation values ESTIMATED FROM SHOTS
  (`expval_joint_sampling`), in the reference's order, so that under a shared `random` stream the result is the
  reference's bit for bit.  The reference deep-copies the model twice through `type(model)()` + `load_state_dict` and
  re-loads a state dict per shift; here the live parameters are set to the shifted values and restored afterwards
  (same circuits at the same angles, same order of measurements).
* `hadamard_grad` — the Hadamard-test gradient of examples/hadamard_grad/hadamard_grad.py:5-83 for circuits recorded
  in `op_history`: d<O>/dtheta of a trainable `OpPauliExp` = - sum_k c_k <O (x) Y_ancilla> of the circuit with a
  controlled-P_k inserted behind the gate.  The reference builds and runs one (n+1)-qubit circuit per (gate, Pauli term);
  here ALL of them run as ONE batched circuit: sample b of the batch is the circuit of term b, the gates of the
  original circuit are shared by the whole batch and the controlled Pauli of term b is a per-sample explicit matrix
  (identity for every other sample), so the whole gradient is one trip through the fused gate passes.
"""
import numpy as np
import torch

from . import functional as tqf
from .algorithm import pauli_string_to_matrix
from .device import QuantumDevice
from .measurement import expval_joint_analytical, expval_joint_sampling
from .module import QuantumModule
from .op_hamil_exp import OpPauliExp

__all__ = ["parameter_shift_gradient", "hadamard_grad"]


def parameter_shift_gradient(model, input_data, expectation_operator, shift_rate=np.pi * 0.5, shots=1024):
    """model: a QuantumModule whose `forward(input_data)` leaves its result in `model.q_device`; returns one gradient
    per entry of the model's state dict after the device's two buffers (`q_device.state`, `q_device.states`), i.e. per
    parameter tensor, as the reference does."""
    entries = [t for _, t in list(model.state_dict(keep_vars=True).items())[2:]]
    grads = torch.zeros(len(entries))
    with torch.no_grad():
        original = [t.detach().clone() for t in entries]
        # the reference keeps one model per sign and shifts each entry back with `-=` / `+=`, so an entry that has had its
        # turn carries the rounding of (theta + s) - s in the "+" model and (theta - s) + s in the "-" model from then on;
        # the two value sets are tracked here so that every circuit is evaluated at the reference's exact angles
        values = {+1.0: [t.clone() for t in original], -1.0: [t.clone() for t in original]}
        try:
            for k in range(len(entries)):
                estimates = {}
                for sign in (+1.0, -1.0):
                    values[sign][k] += sign * shift_rate
                    for t, v in zip(entries, values[sign]):
                        t.copy_(v)
                    model.forward(input_data)
                    estimates[sign] = expval_joint_sampling(model.q_device, observable=expectation_operator, n_shots=shots)
                    values[sign][k] -= sign * shift_rate
                grads[k] = (estimates[+1.0] - estimates[-1.0]) * 0.5
        finally:
            for t, v in zip(entries, original):
                t.copy_(v)
    return grads


def _controlled_pauli(pauli, n_targets):
    """|0><0| (x) 1 + |1><1| (x) P on wires [ancilla] + targets"""
    dim = 2 ** n_targets
    m = torch.eye(2 * dim, dtype=torch.complex64)
    m[dim:, dim:] = pauli_string_to_matrix(pauli)
    return m


def hadamard_grad(op_history, n_wires, observable, device=None):
    """-> list ordered as `op_history`: None for non-trainable entries, the gradient (0-dim tensor) of
    <observable> with respect to the gate's theta for trainable `OpPauliExp` entries.  Parameters are assumed unique
    and ordered, as in the reference (test/hadamard_grad/test_hadamard_grad.py)."""
    if device is None:
        device = "cuda"
    terms = []   # (position in op_history, coefficient, pauli string)
    for k, op in enumerate(op_history):
        if op["trainable"]:
            if "coeffs" not in op:
                raise KeyError(f"hadamard_grad: trainable op {op['name']!r} carries no Pauli decomposition "
                               "(only OpPauliExp entries are differentiated, as in the reference)")
            terms += [(k, float(c), p) for c, p in zip(op["coeffs"], op["paulis"])]
    grads = [None] * len(op_history)
    if not terms:
        return grads
    ancilla = n_wires
    dev = QuantumDevice(n_wires=n_wires + 1, bsz=len(terms), device=device)
    with torch.no_grad():
        dev.h(wires=ancilla)
        for k, op in enumerate(op_history):
            if "coeffs" in op:
                OpPauliExp(coeffs=op["coeffs"], paulis=op["paulis"], theta=op["params"], trainable=False)(dev, wires=op["wires"])
            else:
                QuantumModule.from_op_history([op]).to(device)(dev)
            mine = [b for b, t in enumerate(terms) if t[0] == k]
            if mine:
                t_wires = op["wires"] if isinstance(op["wires"], (list, tuple)) else [op["wires"]]
                mats = torch.eye(2 ** (1 + len(t_wires)), dtype=torch.complex64).repeat(len(terms), 1, 1)
                for b in mine:
                    mats[b] = _controlled_pauli(terms[b][2], len(t_wires))
                tqf.qubitunitaryfast(dev, wires=[ancilla] + list(t_wires), params=mats.to(device))
        # the closing Hadamard of the ancilla commutes with every gate on the system wires: once, at the end, for all terms
        dev.h(wires=ancilla)
        expval = expval_joint_analytical(dev, observable + "Y")
    for b, (k, coeff, _) in enumerate(terms):
        g = -coeff * expval[b]
        grads[k] = g if grads[k] is None else grads[k] + g
    return grads

"""`OpHamilExp` / `OpPauliExp` — exp(-i theta H / 2) as an operation (drop-in for operator/op_hamil_exp.py:40-150).
The matrix exponential is tiny host-side torch code (2^k x 2^k); the gate application goes through
`tqf.qubitunitaryfast`, i.e. the fused / tensor-core dense path of this package."""
from typing import List

import numpy as np
import torch

from . import functional as tqf
from .algorithm import Hamiltonian
from .module import QuantumModule

__all__ = ["OpHamilExp", "OpPauliExp"]


class OpHamilExp(QuantumModule):
    """Matrix exponential operation exp(-i * theta * H / 2); the default theta is 0.0."""

    def __init__(self, hamil: Hamiltonian, trainable: bool = True, theta: float = 0.0):
        super().__init__()
        if trainable:
            self.theta = torch.nn.parameter.Parameter(torch.tensor(theta))
        else:
            self.theta = torch.tensor(theta)
        self.hamil = hamil

    def get_exponent_matrix(self):
        return self.hamil.matrix.to(self.theta.device) * -1j * self.theta / 2

    @property
    def exponent_matrix(self):
        return self.get_exponent_matrix()

    def get_matrix(self):
        return torch.matrix_exp(self.exponent_matrix)

    @property
    def matrix(self):
        return self.get_matrix()

    def forward(self, qdev, wires):
        matrix = self.matrix.to(qdev.compute_device)
        tqf.qubitunitaryfast(q_device=qdev, wires=wires, params=matrix)


class OpPauliExp(OpHamilExp):
    """OpHamilExp built from coefficients and Pauli strings; records itself in `op_history` (op_hamil_exp.py:96-150)."""

    def __init__(self, coeffs: List[float], paulis: List[str], endianness: str = "big", trainable: bool = True,
                 theta: float = 0.0):
        super().__init__(hamil=Hamiltonian(coeffs, paulis, endianness), trainable=trainable, theta=theta)
        self.coeffs, self.paulis, self.trainable = coeffs, paulis, trainable

    def forward(self, qdev, wires):
        if qdev.record_op:
            entry = dict(name=type(self).__name__, wires=np.array(wires).squeeze().tolist(), coeffs=self.coeffs,
                         paulis=self.paulis, inverse=False, trainable=self.trainable, params=self.theta.item())
            qdev.op_history.append(entry)
        # ONE history entry per gate: the reference also records the qubitunitaryfast it queues (op_hamil_exp.py:132-150),
        # which replays the gate twice through from_op_history and is why its hadamard_grad test is skipped
        recording, qdev.record_op = qdev.record_op, False
        try:
            super().forward(qdev, wires)
        finally:
            qdev.record_op = recording

"""`tq.Hamiltonian` and `tq.VQE` — the immediate caller of `expval_joint_analytical` (drop-in for
torchquantum/algorithm/hamiltonian.py:93-198 and algorithm/vqe.py:35-113).

The reference evaluates the energy with one `expval_joint_analytical` call per Pauli term (each a dense 2^n x 2^n
Kronecker product and a matmul); here the whole Pauli sum is ONE launch of the multi-term read-out kernel (the terms
share every load of the state) followed by a dot product with the coefficients, and its backward is one launch too.
"""
from typing import Any, List

import torch

from .device import QuantumDevice
from .measurement import expval_joint_analytical_multi

__all__ = ["Hamiltonian", "VQE", "parse_hamiltonian_file"]

_PAULI = {
    "X": torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64),
    "Y": torch.tensor([[0, -1j], [1j, 0]], dtype=torch.complex64),
    "Z": torch.tensor([[1, 0], [0, -1]], dtype=torch.complex64),
    "I": torch.tensor([[1, 0], [0, 1]], dtype=torch.complex64),
}


def pauli_string_to_matrix(pauli: str, device=torch.device("cpu")) -> torch.Tensor:
    """Kronecker product of the string's Pauli matrices, first character = most significant (util/utils.py:990-1009)."""
    pauli = pauli.upper()
    m = _PAULI[pauli[0]].to(device)
    for ch in pauli[1:]:
        m = torch.kron(m, _PAULI[ch].to(device))
    return m


def parse_hamiltonian_file(filename: str) -> dict:
    """File format of examples/vqe/h2.txt (algorithm/hamiltonian.py:37-91): a header `name method n_wires`, then one
    line per term `coeff P0 P1 ...` with P = X|Y|Z|I followed by the wire index; wires not named are identity."""
    info = {"paulis": [], "coeffs": []}
    with open(filename, "r") as f:
        lines = [ln.strip() for ln in f.read().split("\n")]
    first = True
    for line in lines:
        if not line:
            continue
        if first:
            name, method, n_wires = line.split()
            info["name"], info["method"], info["n_wires"] = name, method, int(n_wires)
            first = False
            continue
        fields = line.split()
        info["coeffs"].append(float(fields[0]))
        chars = ["I"] * info["n_wires"]
        for tok in fields[1:]:
            assert tok[0].upper() in "XYZI", tok
            chars[int(tok[1:])] = tok[0].upper()
        info["paulis"].append("".join(chars))
    return info


class Hamiltonian(object):
    """Weighted sum of Pauli strings (algorithm/hamiltonian.py:93-198)."""

    def __init__(self, coeffs: List[float], paulis: List[str], endianness: str = "big") -> None:
        # same checks and messages as the reference constructor (algorithm/hamiltonian.py:116-137)
        self._validate(coeffs, paulis, endianness)
        self.endianness = endianness
        self.coeffs = coeffs
        self.paulis = list(paulis) if endianness == "big" else [p[::-1] for p in paulis]
        self.n_wires = len(paulis[0])
        self._coeff_cache = {}

    @staticmethod
    def _validate(coeffs, paulis, endianness):
        if endianness not in ("big", "little"):
            raise ValueError("Endianness must be either big or little.")
        if len(coeffs) != len(paulis):
            raise ValueError("The number of coefficients and operators must be the same.")
        width = len(paulis[0])
        if any(len(p) != width for p in paulis):
            raise ValueError("The length of each operator must be the same.")
        if any(ch not in "XYZI" for p in paulis for ch in p):
            raise ValueError("The operator must be a string of X, Y, Z, and I.")

    @property
    def matrix(self) -> torch.Tensor:
        return self.get_matrix()

    def get_matrix(self) -> torch.Tensor:
        """Dense 2^n x 2^n matrix sum_t c_t P_t (small n only; the engine never builds it)."""
        terms = [c * pauli_string_to_matrix(p) for c, p in zip(self.coeffs, self.paulis)]
        return torch.stack(terms).sum(0)

    @property
    def hamil_info(self) -> dict:
        """The dictionary form the reference's VQE reads (algorithm/vqe.py:88-95)."""
        return {"n_wires": self.n_wires,
                "hamil_list": [{"coeff": c, "pauli_string": p} for c, p in zip(self.coeffs, self.paulis)]}

    def coeff_tensor(self, device) -> torch.Tensor:
        key = (device, tuple(self.coeffs))   # `coeffs` is a public, mutable attribute in the reference: follow edits
        t = self._coeff_cache.get(key)
        if t is None:
            self._coeff_cache.clear()
            t = self._coeff_cache[key] = torch.tensor(self.coeffs, dtype=torch.float32, device=device)
        return t

    def expval(self, qdev) -> torch.Tensor:
        """<psi_b| H |psi_b> -> [bsz] float32: all terms in one kernel launch."""
        ev = expval_joint_analytical_multi(qdev, self.paulis)
        return ev @ self.coeff_tensor(ev.device)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.coeffs}, {self.paulis}, {self.endianness})"

    def __len__(self) -> int:
        return len(self.coeffs)

    @classmethod
    def from_file(cls, file_path: str) -> Any:
        info = parse_hamiltonian_file(file_path)
        return cls(info["coeffs"], info["paulis"])


class VQE(object):
    """Variational quantum eigensolver (algorithm/vqe.py:35-113): same constructor, attributes and methods."""

    def __init__(self, hamil, ansatz, train_configs) -> None:
        self.hamil = hamil
        self.ansatz = ansatz
        self.train_configs = train_configs
        self.n_wires = hamil.n_wires
        self.n_epochs = self.train_configs.get("n_epochs", 100)
        self.n_steps = self.train_configs.get("n_steps", 10)
        self.optimizer_name = self.train_configs.get("optimizer", "Adam")
        self.scheduler_name = self.train_configs.get("scheduler", "CosineAnnealingLR")
        self.lr = self.train_configs.get("lr", 0.1)
        self.device = self.train_configs.get("device", "cuda")
        self.ansatz = self.ansatz.to(self.device)

    def get_expval(self, qdev):
        return self.hamil.expval(qdev)

    def get_loss(self):
        qdev = QuantumDevice(n_wires=self.n_wires, bsz=1, device=self.device)
        self.ansatz(qdev)
        return self.get_expval(qdev)

    def train(self):
        """Adam / cosine schedule by default (train_configs), `n_epochs` x `n_steps` steps; returns the last energy."""
        opt = getattr(torch.optim, self.optimizer_name)(self.ansatz.parameters(), lr=self.lr)
        sched = getattr(torch.optim.lr_scheduler, self.scheduler_name)(opt, T_max=self.n_epochs)
        energy = None
        for epoch in range(self.n_epochs):
            for step in range(self.n_steps):
                opt.zero_grad()
                energy = self.get_loss()
                energy.backward()
                opt.step()
                print(f"Epoch: {epoch}, Step: {step}, Loss: {energy}")
            sched.step()
        return energy.detach().cpu().item()

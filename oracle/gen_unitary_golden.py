"""TEST INFRASTRUCTURE — golden vectors for `qubitunitary` and `qubitunitarystrict` (functional/qubit_unitary.py:19-86;
round-1 review: the 229 gate cases of tests/golden/gates.npz cover `qubitunitaryfast` only), produced by running the REAL
reference (/root/reference through oracle/refshim.py): explicit matrices on 1-3 wires of a 4-wire device (random initial
state, bsz 2), wires in any order, plain and inverse, shared and per-sample matrices, and for the strict variant
NON-unitary inputs (the reference projects them with U @ V of the SVD, not U @ V^H).

    python oracle/gen_unitary_golden.py        # writes tests/golden/unitary_gates.npz

Checked on the CPU emulation and on the GPU by tests/cases.py::check_unitary_gates."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refshim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "unitary_gates.npz")
N, BSZ = 4, 2
WIRES = [[0], [3], [1, 2], [3, 0], [2, 0, 3], [0, 1, 2], [3, 2, 1, 0]]


def random_unitary(rng, dim, batch=None):
    shape = (dim, dim) if batch is None else (batch, dim, dim)
    a = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    q, r = np.linalg.qr(a)
    d = np.diagonal(r, axis1=-2, axis2=-1)
    return (q * (d / np.abs(d))[..., None, :]).astype(np.complex64)


def main():
    tq, tqf = load_reference()
    rng = np.random.default_rng(2024)
    init = rng.normal(size=(BSZ, 2 ** N)) + 1j * rng.normal(size=(BSZ, 2 ** N))
    init = (init / np.linalg.norm(init, axis=1, keepdims=True)).astype(np.complex64)
    arrays, cases = {"init": init}, []
    for name in ("qubitunitary", "qubitunitarystrict"):
        for wires in WIRES:
            dim = 2 ** len(wires)
            for inverse in (False, True):
                for batched in (False, True):
                    bsz = BSZ
                    if name == "qubitunitarystrict":
                        # a generic (non-unitary) complex matrix: the projection is part of the gate.  The reference
                        # drops its own `params.squeeze(0)` (qubit_unitary.py:83), so a shared matrix only works there on
                        # a bsz-1 device; per-sample matrices work at any batch size
                        shape = (BSZ, dim, dim) if batched else (dim, dim)
                        mat = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
                        bsz = BSZ if batched else 1
                    else:
                        mat = random_unitary(rng, dim, BSZ if batched else None)
                    q = tq.QuantumDevice(n_wires=N, bsz=bsz)
                    q.set_states(torch.tensor(init[:bsz]).reshape([bsz] + [2] * N))
                    getattr(tqf, name)(q, wires=wires, params=torch.tensor(mat), inverse=inverse)
                    k = len(cases)
                    arrays[f"mat{k}"] = mat
                    arrays[f"out{k}"] = q.get_states_1d().detach().numpy().astype(np.complex64)
                    cases.append({"name": name, "wires": wires, "inverse": inverse, "batched": batched, "bsz": bsz})
    # the unitarity check of qubitunitary (AssertionError) is behaviour too
    arrays["manifest"] = np.array(json.dumps({"n": N, "bsz": BSZ, "cases": cases}))
    np.savez_compressed(OUT, **arrays)
    print(f"{len(cases)} cases -> {OUT}")


if __name__ == "__main__":
    main()

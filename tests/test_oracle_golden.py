"""Pins the CPU oracle (oracle/tq_oracle.py) against fixtures produced by the real reference
(oracle/gen_golden.py) and, when /root/reference is mounted, against the live reference."""
import random

import numpy as np
import pytest
import torch

from oracle import tq_oracle as O
from tests.helpers import GRAD_RTOL, assert_close, load, op_params


def test_every_gate_matches_reference():
    z, man = load("gates.npz")
    n, init = man["n"], torch.tensor(z["init"])
    for case in man["cases"]:
        s = init.reshape([init.shape[0]] + [2] * n)
        if case["name"] == "qubitunitaryfast":
            out = O.apply_unitary_bmm(s, torch.tensor(z[case["mat"]]), case["wires"])
        else:
            out = O.apply_op(s, case["name"], case["wires"], op_params(case), case.get("n_wires"),
                             case["inverse"])
        assert_close(out.reshape(init.shape[0], -1), z[case["out"]], what=str(case["name"]))


def test_explicit_unitaries_match_reference():
    """qubitunitary / qubitunitarystrict (functional/qubit_unitary.py:19-86): the oracle against the reference's outputs"""
    z, man = load("unitary_gates.npz")
    n = man["n"]
    for k, case in enumerate(man["cases"]):
        bsz = case["bsz"]
        s = torch.tensor(z["init"][:bsz]).reshape([bsz] + [2] * n)
        out = O.apply_op(s, case["name"], case["wires"], torch.tensor(z[f"mat{k}"]), None, case["inverse"])
        assert_close(out.reshape(bsz, -1), z[f"out{k}"], what=f"{case['name']} {case['wires']}")


def _run_circuit(z, c, requires_grad):
    pre = c["prefix"]
    init = torch.tensor(z[pre + "init"]).requires_grad_(requires_grad)
    ops, leaves = [], []
    for op in c["ops"]:
        op = dict(op)
        if op.get("params") is not None:
            op["params"] = op_params(op, requires_grad)
            leaves.append(op["params"])
        ops.append(op)
    states = O.run_ops(c["n"], c["bsz"], ops, init)
    return init, leaves, states


def test_random_circuits_states_expvals_grads():
    z, man = load("circuits.npz")
    for c in man:
        pre = c["prefix"]
        init, leaves, states = _run_circuit(z, c, True)
        assert_close(states, z[pre + "states"], what=pre + "states")
        ev = torch.stack([O.expval_joint_analytical(states, o) for o in c["obs"]], dim=-1)
        assert_close(ev, z[pre + "ev"], what=pre + "ev")
        ev2 = torch.stack([O.expval_joint_rotated(states, o) for o in c["obs"]], dim=-1)
        assert_close(ev2, z[pre + "ev"], what=pre + "ev(rotated form)")
        loss = (ev * torch.tensor(z[pre + "w"])).sum() + (states * torch.tensor(z[pre + "proj"]).conj()).real.sum()
        assert_close(loss, z[pre + "loss"], what=pre + "loss")
        grads = torch.autograd.grad(loss, leaves + [init])
        for gi, g in enumerate(grads[:-1]):
            assert_close(g, z[pre + f"g{gi}"], rtol=GRAD_RTOL, what=pre + f"g{gi}")
        assert_close(grads[-1], z[pre + "ginit"], rtol=GRAD_RTOL, what=pre + "ginit")


def test_expval_per_wire_and_state_mutation():
    z, man = load("circuits.npz")
    for c in man:
        pre = c["prefix"]
        _, _, states = _run_circuit(z, c, False)
        s = states.reshape([c["bsz"]] + [2] * c["n"])
        ev, after = O.expval(s, list(range(c["n"])), c["obs_wires"])
        assert_close(ev, z[pre + "ev_wires"], what=pre + "ev_wires")
        assert_close(after.reshape(c["bsz"], -1), z[pre + "states_after_expval"], what=pre + "after")


def test_measure_histograms_bit_exact_and_rng_stream():
    z, man = load("measure.npz")
    for ci, c in enumerate(man):
        random.seed(c["seed"])
        res = O.measure(torch.tensor(z[f"m{ci}_states"]), n_shots=c["shots"])
        counts = np.array([list(d.values()) for d in res])
        assert np.array_equal(counts, z[f"m{ci}_counts"])
        assert list(res[0].keys())[1] == "{:0{}b}".format(1, c["n"])
        assert random.random() == c["next_random"]  # shared MT19937 stream advanced identically


def test_mnist_cfg1_forward_backward():
    import torch.nn.functional as F

    z, man = load("mnist_cfg1.npz")
    n, bsz = man["n"], man["bsz"]
    x = torch.tensor(z["x"]).requires_grad_(True)
    ops, leaves = [], []
    for op in man["ops"]:
        o = {"name": op["name"], "wires": op["wires"], "inverse": op["inverse"]}
        if "input_idx" in op:
            o["params"] = x[:, op["input_idx"]]
        elif op["params"] is not None:
            t = torch.tensor(op["params"], dtype=torch.float32).reshape(1, -1)
            if op["trainable"]:
                t.requires_grad_(True)
                leaves.append(t)
            o["params"] = t
        ops.append(o)
    states = O.run_ops(n, bsz, ops)
    assert_close(states, z["states"], what="cfg1 states")
    out, _ = O.expval(states.reshape([bsz] + [2] * n), list(range(n)), ["Z"] * n)
    assert_close(out, z["out"], what="cfg1 MeasureAll")
    loss = F.nll_loss(F.log_softmax(out.reshape(bsz, 2, 2).sum(-1), dim=1), torch.tensor(z["target"]))
    assert_close(loss, z["loss"], what="cfg1 loss")
    grads = torch.autograd.grad(loss, [x] + leaves)
    assert_close(grads[0], z["grad_x"], rtol=GRAD_RTOL, what="cfg1 grad_x")
    want = [g["grad"] for g in man["param_grads"]]
    assert len(want) == len(leaves) == 45
    for g, w in zip(grads[1:], want):
        assert_close(g.reshape(-1), torch.tensor(w), rtol=GRAD_RTOL, what="cfg1 param grad")


def test_known_answers_from_reference_tests():
    """Literal vectors held by the reference's own tests: device init (test/device/test_init.py),
    Bell/GHZ-style docstring examples of expval_joint_analytical (measurements.py:243-255)."""
    s = O.zero_states(3, 2).reshape(2, -1)
    assert s[0, 0] == 1 and s.abs().sum() == 2
    s = O.run_ops(2, 1, [{"name": "hadamard", "wires": [0]}, {"name": "x", "wires": [1]},
                         {"name": "cnot", "wires": [0, 1]}])
    assert_close(O.expval_joint_analytical(s, "II"), torch.tensor([1.0]))
    assert_close(O.expval_joint_analytical(s, "XX"), torch.tensor([1.0]))
    assert_close(O.expval_joint_analytical(s, "ZZ"), torch.tensor([-1.0]))


def test_live_reference_if_mounted():
    from oracle.refshim import load_reference

    ref = load_reference()
    if ref is None:
        pytest.skip("reference tree not mounted")
    tq, tqf = ref
    torch.manual_seed(5)
    n, bsz = 5, 3
    qdev = tq.QuantumDevice(n_wires=n, bsz=bsz)
    ops = []
    for name, wires, p in [("h", [0], None), ("u3", [1], torch.rand(bsz, 3)), ("crx", [4, 2], torch.rand(1, 1)),
                           ("rzz", [3, 0], torch.rand(bsz, 1)), ("toffoli", [2, 4, 1], None),
                           ("qft", [3, 1, 0], None)]:
        kw = {"params": p} if p is not None else {}
        if name == "qft":
            kw["n_wires"] = 3
        getattr(tqf, name)(qdev, wires=wires, **kw)
        ops.append({"name": name, "wires": wires, "params": p, "n_wires": kw.get("n_wires")})
    assert_close(O.run_ops(n, bsz, ops), qdev.get_states_1d(), what="live reference")

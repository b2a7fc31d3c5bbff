"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REAL reference
(/root/reference, imported through oracle/refshim.py) on seeded inputs.

Run here (the container that mounts /root/reference):  python oracle/gen_golden.py
The fixtures are committed; nothing reads /root/reference at test time on the GPU box.

Every fixture stores a JSON manifest (op lists: name / wires / params / inverse) plus the
reference's outputs: final states, expectation values, gradients, histograms.
"""
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refshim import load_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# reference functional names with (arity, n_params); arity None = any
PARAM_GATES = {
    "rx": (1, 1), "ry": (1, 1), "rz": (1, 1), "phaseshift": (1, 1), "u1": (1, 1), "rot": (1, 3),
    "u2": (1, 2), "u3": (1, 3), "r": (1, 2), "crx": (2, 1), "cry": (2, 1), "crz": (2, 1),
    "crot": (2, 3), "cu1": (2, 1), "cu2": (2, 2), "cu3": (2, 3), "cu": (2, 4), "rxx": (2, 1),
    "ryy": (2, 1), "rzz": (2, 1), "rzx": (2, 1), "singleexcitation": (2, 1), "xxminyy": (2, 2),
    "xxplusyy": (2, 2), "p": (1, 1), "u": (1, 3), "cp": (2, 1), "cr": (2, 1), "cphase": (2, 1),
    "xx": (2, 1), "yy": (2, 1), "zz": (2, 1), "zx": (2, 1),
}
# reference crashes with per-sample params for these (probed): broadcasting bugs in the builders
NO_PER_SAMPLE = {"cu1", "cu2", "cu3", "cu", "singleexcitation", "xxminyy", "xxplusyy", "cp", "cr",
                 "cphase"}
FIXED_GATES = {
    "hadamard": 1, "h": 1, "sh": 1, "paulix": 1, "pauliy": 1, "pauliz": 1, "i": 1, "s": 1, "t": 1,
    "sx": 1, "x": 1, "y": 1, "z": 1, "sdg": 1, "tdg": 1, "sxdg": 1, "cnot": 2, "cx": 2, "cz": 2,
    "cy": 2, "swap": 2, "sswap": 2, "iswap": 2, "ch": 2, "chadamard": 2, "cs": 2, "csdg": 2,
    "csx": 2, "dcx": 2, "ecr": 2, "echoedcrossresonance": 2, "cswap": 3, "toffoli": 3, "ccnot": 3,
    "ccx": 3, "ccz": 3, "rccx": 3, "c3x": 4, "c3sx": 4, "rc3x": 4, "c4x": 5,
}


def rand_state(bsz, n, gen):
    re = torch.randn(bsz, 2 ** n, generator=gen)
    im = torch.randn(bsz, 2 ** n, generator=gen)
    s = torch.complex(re, im)
    return (s / s.abs().pow(2).sum(-1, keepdim=True).sqrt()).type(torch.complex64)


def run_ref(tq, tqf, n, states, ops):
    qdev = tq.QuantumDevice(n_wires=n, bsz=states.shape[0])
    qdev.set_states(states.clone())
    for op in ops:
        kw = {}
        if op.get("params") is not None:
            kw["params"] = op["params"]
        if op.get("n_wires") is not None:
            kw["n_wires"] = op["n_wires"]
        getattr(tqf, op["name"])(qdev, wires=op["wires"], inverse=op.get("inverse", False), **kw)
    return qdev


def jsonable_ops(ops):
    out = []
    for op in ops:
        o = dict(op)
        if o.get("params") is not None:
            o["params"] = torch.as_tensor(o["params"]).detach().tolist()
        out.append(o)
    return out


def gen_gates(tq, tqf):
    """Every functional gate × {shared, per-sample, inverse} on random wires of a 5/6-wire
    random state (bsz 3)."""
    gen = torch.Generator().manual_seed(1234)
    rng = np.random.default_rng(1234)
    cases, arrays = [], {}
    n, bsz = 6, 3
    init = rand_state(bsz, n, gen)
    arrays["init"] = init.numpy()

    def add(name, wires, params, inverse, n_wires=None):
        op = {"name": name, "wires": wires, "inverse": inverse}
        if params is not None:
            op["params"] = params
        if n_wires is not None:
            op["n_wires"] = n_wires
        out = run_ref(tq, tqf, n, init, [op]).get_states_1d()
        key = f"out{len(cases)}"
        arrays[key] = out.numpy()
        case = jsonable_ops([op])[0]
        case["out"] = key
        cases.append(case)

    for name, (k, p) in PARAM_GATES.items():
        for inverse in (False, True):
            wires = rng.choice(n, size=k, replace=False).tolist()
            add(name, wires, torch.rand(1, p, generator=gen) * 6 - 3, inverse)
            if name not in NO_PER_SAMPLE:
                wires = rng.choice(n, size=k, replace=False).tolist()
                add(name, wires, torch.rand(bsz, p, generator=gen) * 6 - 3, inverse)
    for name, k in FIXED_GATES.items():
        for inverse in (False, True):
            add(name, rng.choice(n, size=k, replace=False).tolist(), None, inverse)
    for k in (2, 3, 4):
        for inverse in (False, True):
            wires = rng.choice(n, size=k, replace=False).tolist()
            add("multicnot", wires, None, inverse, n_wires=k)
            add("multixcnot", wires, None, inverse, n_wires=k)
            add("qft", wires, None, inverse, n_wires=k)
            add("multirz", wires, torch.rand(1, 1, generator=gen) * 6 - 3, inverse, n_wires=k)
            add("multirz", wires, torch.rand(bsz, 1, generator=gen) * 6 - 3, inverse, n_wires=k)
    # globalphase on zero wires (shared only; per-sample crashes in the reference)
    add("globalphase", [], torch.rand(1, 1, generator=gen), False)
    # arbitrary (unitary and non-unitary) matrices through qubitunitaryfast
    for k in (1, 2, 3):
        wires = rng.choice(n, size=k, replace=False).tolist()
        re = torch.randn(2 ** k, 2 ** k, generator=gen)
        im = torch.randn(2 ** k, 2 ** k, generator=gen)
        q, _ = torch.linalg.qr(torch.complex(re, im))
        op = {"name": "qubitunitaryfast", "wires": wires, "inverse": False}
        qdev = tq.QuantumDevice(n_wires=n, bsz=bsz)
        qdev.set_states(init.clone())
        tqf.qubitunitaryfast(qdev, wires=wires, params=q.type(torch.complex64))
        key = f"out{len(cases)}"
        arrays[key] = qdev.get_states_1d().numpy()
        arrays[key + "_mat"] = q.type(torch.complex64).numpy()
        op["out"] = key
        op["mat"] = key + "_mat"
        cases.append(op)
    # reset (non-unitary)
    qdev = tq.QuantumDevice(n_wires=n, bsz=bsz)
    qdev.set_states(init.clone())
    tqf.reset(qdev, wires=[2])
    arrays["out_reset"] = qdev.get_states_1d().numpy()
    cases.append({"name": "reset", "wires": [2], "inverse": False, "out": "out_reset"})
    arrays["manifest"] = np.array(json.dumps({"n": n, "bsz": bsz, "cases": cases}))
    np.savez_compressed(os.path.join(GOLDEN, "gates.npz"), **arrays)
    print("gates:", len(cases), "cases")


CIRCUIT_POOL_1P = ["rx", "ry", "rz", "phaseshift", "u1", "crx", "cry", "crz", "rxx", "ryy", "rzz",
                   "rzx", "cu1"]
CIRCUIT_POOL_NP = ["rot", "u2", "u3", "r", "crot", "cu3", "cu", "xxminyy", "xxplusyy",
                   "singleexcitation"]
CIRCUIT_POOL_FIXED = ["h", "sh", "x", "y", "z", "s", "t", "sx", "sdg", "tdg", "sxdg", "cnot", "cz",
                      "cy", "swap", "sswap", "iswap", "ch", "cs", "csx", "dcx", "ecr", "cswap",
                      "toffoli", "ccz", "rccx", "c3x"]


def random_circuit(n, bsz, n_ops, rng, gen):
    """op soup in the spirit of test/static_mode_test.py: shared trainable params, per-sample
    inputs, random inverses."""
    ops, leaves = [], []
    for _ in range(n_ops):
        kind = rng.choice(["1p", "np", "fixed"], p=[0.5, 0.2, 0.3])
        pool = {"1p": CIRCUIT_POOL_1P, "np": CIRCUIT_POOL_NP, "fixed": CIRCUIT_POOL_FIXED}[kind]
        name = str(rng.choice(pool))
        k = PARAM_GATES[name][0] if name in PARAM_GATES else FIXED_GATES[name]
        if k > n:
            continue
        wires = rng.choice(n, size=k, replace=False).tolist()
        op = {"name": name, "wires": wires, "inverse": bool(rng.random() < 0.3)}
        if name in PARAM_GATES:
            p = PARAM_GATES[name][1]
            per_sample = name not in NO_PER_SAMPLE and rng.random() < 0.4
            t = (torch.rand(bsz if per_sample else 1, p, generator=gen) * 6 - 3).requires_grad_(True)
            op["params"] = t
            leaves.append(t)
        ops.append(op)
    return ops, leaves


def gen_circuits(tq, tqf):
    from torchquantum.measurement import expval_joint_analytical, expval

    gen = torch.Generator().manual_seed(99)
    rng = np.random.default_rng(99)
    arrays, circuits = {}, []
    for ci, (n, bsz, n_ops) in enumerate([(2, 3, 30), (3, 4, 60), (4, 5, 80), (5, 3, 120),
                                          (7, 2, 200), (10, 2, 150)]):
        ops, leaves = random_circuit(n, bsz, n_ops, rng, gen)
        init = rand_state(bsz, n, gen).requires_grad_(True)
        qdev = run_ref(tq, tqf, n, init, ops)
        states = qdev.get_states_1d()
        obs = ["".join(rng.choice(list("XYZI"), size=n)) for _ in range(4)]
        ev = torch.stack([expval_joint_analytical(qdev, o) for o in obs], dim=-1)  # [bsz, 4]
        w = torch.rand(bsz, 4, generator=gen)
        proj = rand_state(bsz, n, gen)
        # loss mixes expvals and a direct linear functional of the state (exercises complex grads)
        loss = (ev * w).sum() + (states * proj.conj()).real.sum()
        grads = torch.autograd.grad(loss, leaves + [init])
        # MeasureAll-style per-wire expval with mixed observables (mutates the device)
        obs_letters = [str(rng.choice(list("XYZ"))) for _ in range(n)]
        obs_objs = [{"X": tq.PauliX, "Y": tq.PauliY, "Z": tq.PauliZ}[c]() for c in obs_letters]
        qdev2 = run_ref(tq, tqf, n, init.detach(), [dict(o, params=o["params"].detach())
                                                    if "params" in o else o for o in ops])
        ev_wires = expval(qdev2, list(range(n)), obs_objs)
        pre = f"c{ci}_"
        arrays[pre + "init"] = init.detach().numpy()
        arrays[pre + "states"] = states.detach().numpy()
        arrays[pre + "ev"] = ev.detach().numpy()
        arrays[pre + "w"] = w.numpy()
        arrays[pre + "proj"] = proj.numpy()
        arrays[pre + "loss"] = np.array(loss.item())
        arrays[pre + "ev_wires"] = ev_wires.detach().numpy()
        arrays[pre + "states_after_expval"] = qdev2.get_states_1d().detach().numpy()
        for gi, g in enumerate(grads[:-1]):
            arrays[pre + f"g{gi}"] = g.numpy()
        arrays[pre + "ginit"] = grads[-1].numpy()
        circuits.append({"n": n, "bsz": bsz, "ops": jsonable_ops(ops), "obs": obs,
                         "obs_wires": obs_letters, "prefix": pre})
    arrays["manifest"] = np.array(json.dumps(circuits))
    np.savez_compressed(os.path.join(GOLDEN, "circuits.npz"), **arrays)
    print("circuits:", len(circuits))


def gen_measure(tq, tqf):
    from torchquantum.measurement import measure

    gen = torch.Generator().manual_seed(7)
    arrays, cases = {}, []
    for ci, (n, bsz, shots) in enumerate([(1, 2, 64), (3, 4, 512), (5, 3, 2048), (8, 2, 1024)]):
        states = rand_state(bsz, n, gen)
        if ci == 1:  # a basis state and an exactly-zero-tail state: CDF edge cases
            states[0] = 0
            states[0, 5] = 1
            states[1, 2 ** n // 2:] = 0
            states[1] /= states[1].abs().pow(2).sum().sqrt()
        qdev = tq.QuantumDevice(n_wires=n, bsz=bsz)
        qdev.set_states(states.clone())
        random.seed(1000 + ci)
        res = measure(qdev, n_shots=shots)
        after = random.random()  # position of the shared MT19937 stream after the call
        counts = np.array([[d["{:0{}b}".format(k, n)] for k in range(2 ** n)] for d in res])
        arrays[f"m{ci}_states"] = states.numpy()
        arrays[f"m{ci}_counts"] = counts
        cases.append({"n": n, "bsz": bsz, "shots": shots, "seed": 1000 + ci, "next_random": after})
    arrays["manifest"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(GOLDEN, "measure.npz"), **arrays)
    print("measure:", len(cases))


def gen_mnist_cfg1(tq, tqf):
    """BASELINE config 1: the QFCModel of examples/mnist/mnist.py:44-134 minus the dataset
    (SURVEY §8d row 1), built from the reference's own classes; bsz reduced to 32 for the
    fixture (the op list, parameters and per-op gradients are stored)."""
    import torch.nn.functional as F

    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    n, bsz = 4, 32

    class QLayer(tq.QuantumModule):
        def __init__(self):
            super().__init__()
            self.n_wires = n
            self.random_layer = tq.RandomLayer(n_ops=50, wires=list(range(n)))
            self.rx0 = tq.RX(has_params=True, trainable=True)
            self.ry0 = tq.RY(has_params=True, trainable=True)
            self.rz0 = tq.RZ(has_params=True, trainable=True)
            self.crx0 = tq.CRX(has_params=True, trainable=True)

        def forward(self, qdev):
            self.random_layer(qdev)
            self.rx0(qdev, wires=0)
            self.ry0(qdev, wires=1)
            self.rz0(qdev, wires=3)
            self.crx0(qdev, wires=[0, 2])
            qdev.h(wires=3)
            qdev.sx(wires=2)
            qdev.cnot(wires=[3, 0])
            qdev.rx(wires=1, params=torch.tensor([0.1]))

    encoder = tq.GeneralEncoder(tq.encoder_op_list_name_dict["4x4_u3_h_rx"])
    qlayer = QLayer()
    x = torch.rand(bsz, 16).requires_grad_(True)
    target = torch.randint(0, 2, (bsz,))
    qdev = tq.QuantumDevice(n_wires=n, bsz=bsz, record_op=True)
    encoder(qdev, x)
    n_enc = len(qdev.op_history)
    qlayer(qdev)
    out = tq.MeasureAll(tq.PauliZ)(qdev)
    logits = F.log_softmax(out.reshape(bsz, 2, 2).sum(-1), dim=1)
    loss = F.nll_loss(logits, target)
    loss.backward()

    enc_table = tq.encoder_op_list_name_dict["4x4_u3_h_rx"]
    ops, trainable = [], []
    for k, h in enumerate(qdev.op_history):
        op = {"name": h["name"], "wires": h["wires"] if isinstance(h["wires"], list) else [h["wires"]],
              "inverse": h["inverse"], "params": h["params"], "trainable": h["trainable"]}
        if k < n_enc and "input_idx" in enc_table[k]:
            op["input_idx"] = enc_table[k]["input_idx"]
        ops.append(op)
    # gradients of the module parameters, in op order (random layer ops then the 4 named ones)
    grads = []
    mods = list(qlayer.random_layer.op_list) + [qlayer.rx0, qlayer.ry0, qlayer.rz0, qlayer.crx0]
    for m in mods:
        if m.params is not None and m.params.requires_grad:
            grads.append({"name": m.name.lower(), "wires": [int(w) for w in m.wires],
                          "value": m.params.detach().reshape(-1).tolist(),
                          "grad": m.params.grad.reshape(-1).tolist()})
    arrays = {
        "x": x.detach().numpy(), "target": target.numpy(), "out": out.detach().numpy(),
        "loss": np.array(loss.item()), "grad_x": x.grad.numpy(),
        "states": qdev.get_states_1d().detach().numpy(),
        "manifest": np.array(json.dumps({"n": n, "bsz": bsz, "n_enc": n_enc, "ops": ops,
                                         "param_grads": grads})),
    }
    np.savez_compressed(os.path.join(GOLDEN, "mnist_cfg1.npz"), **arrays)
    print("mnist cfg1: ops", len(ops), "trainable", len(grads), "loss", loss.item())


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("reference tree not mounted; golden fixtures are generated offline")
    tq, tqf = ref
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(4)
    gen_gates(tq, tqf)
    gen_circuits(tq, tqf)
    gen_measure(tq, tqf)
    gen_mnist_cfg1(tq, tqf)


if __name__ == "__main__":
    main()

"""API shell (operators, modules, layers, encoders) on the CPU-emulated kernels: the BASELINE
config-1 model built from OUR classes with the reference's seeds must reproduce the reference's
golden loss / outputs / gradients (tests/golden/mnist_cfg1.npz)."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import assert_close, load
from tests.test_cpu_emulated import emulated  # noqa: F401  (autouse fixture: host emulation build)


def build_cfg1(tq):
    class QLayer(tq.QuantumModule):
        def __init__(self):
            super().__init__()
            self.n_wires = 4
            self.random_layer = tq.RandomLayer(n_ops=50, wires=list(range(4)))
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

    return tq.GeneralEncoder(tq.encoder_op_list_name_dict["4x4_u3_h_rx"]), QLayer()


def test_cfg1_model_from_classes_matches_reference_golden():
    import torchquantum_b200 as tq

    z, man = load("mnist_cfg1.npz")
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    encoder, qlayer = build_cfg1(tq)
    bsz = man["bsz"]
    x = torch.rand(bsz, 16).requires_grad_(True)
    assert_close(x, z["x"], what="same RNG stream as the reference model construction")
    target = torch.tensor(z["target"])
    qdev = tq.QuantumDevice(n_wires=4, bsz=bsz, record_op=True)
    encoder(qdev, x)
    qlayer(qdev)
    out = tq.MeasureAll(tq.PauliZ)(qdev)
    assert_close(out, z["out"], what="MeasureAll")
    loss = F.nll_loss(F.log_softmax(out.reshape(bsz, 2, 2).sum(-1), dim=1), target)
    assert_close(loss, z["loss"], what="loss")
    loss.backward()
    assert_close(x.grad, z["grad_x"], rtol=1e-4, what="grad x")
    mods = list(qlayer.random_layer.op_list) + [qlayer.rx0, qlayer.ry0, qlayer.rz0, qlayer.crx0]
    got = [m for m in mods if m.params is not None and m.params.requires_grad]
    assert len(got) == len(man["param_grads"]) == 45
    for m, w in zip(got, man["param_grads"]):
        assert m.name.lower() == w["name"]
        assert_close(m.params.detach().reshape(-1), torch.tensor(w["value"]), what="param init")
        assert_close(m.params.grad.reshape(-1), torch.tensor(w["grad"]), rtol=1e-4, what="param grad")
    # op_history wire format (functional/gate_wrapper.py:380-391)
    hist = qdev.op_history
    assert len(hist) == len(man["ops"]) == 70
    for h, w in zip(hist, man["ops"]):
        assert h["name"] == w["name"] and h["inverse"] == w["inverse"] and h["trainable"] == w["trainable"]
        hw = h["wires"] if isinstance(h["wires"], list) else [h["wires"]]
        assert hw == w["wires"]
    # four expval_joint_analytical strings equal MeasureAll(PauliZ)
    qdev2 = tq.QuantumDevice(n_wires=4, bsz=bsz)
    encoder(qdev2, x.detach())
    qlayer(qdev2)
    ev = torch.stack([tq.expval_joint_analytical(qdev2, s) for s in ("ZIII", "IZII", "IIZI", "IIIZ")], dim=-1)
    assert_close(ev, z["out"], what="joint analytical == MeasureAll")


def test_from_op_history_round_trip_and_get_unitary():
    import torchquantum_b200 as tq

    qdev = tq.QuantumDevice(n_wires=3, bsz=2, record_op=True)
    qdev.h(wires=0)
    qdev.crx(wires=[0, 2], params=torch.tensor([0.4]))
    qdev.u3(wires=1, params=torch.tensor([[0.1, 0.2, 0.3]]))
    qdev.cnot(wires=[2, 1], inverse=True)
    want = qdev.get_states_1d().clone()
    mod = tq.QuantumModule.from_op_history(qdev.op_history)
    q2 = tq.QuantumDevice(n_wires=3, bsz=2)
    mod(q2)
    assert_close(q2.get_states_1d(), want, what="from_op_history")
    mod.n_wires = 3
    u = mod.get_unitary(device="cpu")
    assert_close(u[:, 0], want[0], what="get_unitary column 0")
    assert_close(u.conj().T @ u, torch.eye(8, dtype=torch.complex64), what="unitarity")


def test_static_mode_is_equivalent_to_dynamic():
    import torchquantum_b200 as tq

    np.random.seed(3)
    torch.manual_seed(3)
    layer = tq.RandomLayerAllTypes(n_ops=60, wires=list(range(5)))
    q1 = tq.QuantumDevice(n_wires=5, bsz=3)
    layer(q1)
    dyn = q1.get_states_1d().clone()
    layer.static_on(wires_per_block=3)
    q2 = tq.QuantumDevice(n_wires=5, bsz=3)
    layer(q2)
    assert_close(q2.get_states_1d(), dyn, what="static == dynamic")  # states 1e-5 (test/static_mode_test.py)
    layer.static_off()


def test_layers_shapes_and_param_counts():
    import torchquantum_b200 as tq

    es = tq.EfficientSU2(n_wires=10, reps=20, entanglement_layer="linear")
    assert sum(p.numel() for p in es.parameters()) == 420  # BASELINE config 2
    n_gates = sum(len(l.ops_all) for l in es.layers_all)
    assert n_gates == 600
    q = tq.QuantumDevice(n_wires=6)
    q.x(wires=2)
    tq.QFTLayer(n_wires=6)(q)
    assert_close(q.get_states_1d().abs(), torch.full((1, 64), 1 / 8.0), what="QFT|x> is flat")


def test_encoder_tables_match_reference_if_mounted():
    import torchquantum_b200 as tq
    from oracle.refshim import load_reference

    ref = load_reference()
    if ref is None:
        pytest.skip("reference tree not mounted")
    ref_tq, _ = ref
    assert set(tq.encoder_op_list_name_dict) == set(ref_tq.encoder_op_list_name_dict)
    for name, table in ref_tq.encoder_op_list_name_dict.items():
        mine = tq.encoder_op_list_name_dict[name]
        assert len(mine) == len(table), name
        for a, b in zip(mine, table):
            assert a["func"] == b["func"] and list(a["wires"]) == list(b["wires"]), name
            assert list(a.get("input_idx") or []) == list(b.get("input_idx") or []), name
    assert set(tq.op_name_dict) == set(ref_tq.op_name_dict)
    for k, cls in ref_tq.op_name_dict.items():
        mine = tq.op_name_dict[k]
        assert int(mine.num_params) == int(cls.num_params) and int(mine.num_wires) == int(cls.num_wires), k

"""API shell (operators, modules, layers, encoders) on the CPU-emulated kernels: the BASELINE
config-1 model built from OUR classes with the reference's seeds must reproduce the reference's
golden loss / outputs / gradients (tests/golden/mnist_cfg1.npz)."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import GRAD_RTOL, assert_close, load
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
    assert_close(x.grad, z["grad_x"], rtol=GRAD_RTOL, what="grad x")
    mods = list(qlayer.random_layer.op_list) + [qlayer.rx0, qlayer.ry0, qlayer.rz0, qlayer.crx0]
    got = [m for m in mods if m.params is not None and m.params.requires_grad]
    assert len(got) == len(man["param_grads"]) == 45
    for m, w in zip(got, man["param_grads"]):
        assert m.name.lower() == w["name"]
        assert_close(m.params.detach().reshape(-1), torch.tensor(w["value"]), what="param init")
        assert_close(m.params.grad.reshape(-1), torch.tensor(w["grad"]), rtol=GRAD_RTOL, what="param grad")
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


@pytest.mark.parametrize("check", ["check_record_cache_follows_the_op_list", "check_grad_mode_is_sampled_when_gates_are_issued",
                                   "check_parameter_values_are_read_at_the_gate_call", "check_checkpoint_round_trip",
                                   "check_from_op_history_and_get_unitary", "check_static_mode_equals_dynamic",
                                   "check_layer_classes", "check_einsum_broadcast",
                                   "check_dropped_device_is_freed_by_refcount", "check_reference_test_vectors",
                                   "check_batch_collapse"])
def test_host_mirror_semantics(check):
    from tests import cases_api

    getattr(cases_api, check)("cpu")


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

"""Device-agnostic checks of the host mirror's semantics (lazy queue vs the reference's eager calls, record caches,
checkpoints, static mode, op_history).  Run on the CPU emulation by tests/test_api_layer.py and on the B200 by
tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import torchquantum_b200 as tq
import torchquantum_b200.functional as tqf
from oracle import tq_oracle as O
from tests.helpers import GRAD_RTOL, assert_close


def _oracle_state(n, bsz, ops):
    return O.run_ops(n, bsz, ops)


def check_record_cache_follows_the_op_list(device):
    """a layer's cached queue records must follow the live op list."""
    np.random.seed(11)
    torch.manual_seed(11)
    n = 4
    L = tq.RandomLayer(n_ops=12, wires=list(range(n))).to(device)
    other = tq.RandomLayer(n_ops=9, wires=list(range(n)), seed=5).to(device)

    def run(layer):
        q = tq.QuantumDevice(n_wires=n, bsz=2, device=device)
        q.h(wires=0)
        layer(q)
        return q.get_states_1d().clone()

    before = run(L)
    L.rebuild_random_layer_from_op_list(other.n_ops, other.wires, other.op_list)
    L.to(device)
    with torch.no_grad():
        for a, b in zip(L.op_list, other.op_list):
            if a.params is not None:
                a.params.copy_(b.params)
    after = run(L)
    assert_close(after, run(other), what="rebuilt RandomLayer runs the new op list")
    assert (after - before).abs().max().item() > 1e-3

    layer = tq.Op1QAllLayer(op=tq.RX, n_wires=n, has_params=True, trainable=True).to(device)
    s0 = run(layer)
    layer.ops_all[1] = tq.RY(has_params=True, trainable=True).to(device)
    s1 = run(layer)
    q = tq.QuantumDevice(n_wires=n, bsz=2, device=device)
    q.h(wires=0)
    for k, o in enumerate(layer.ops_all):
        o(q, wires=k)
    assert_close(s1, q.get_states_1d(), what="replaced entry of ops_all is used")
    assert (s1 - s0).abs().max().item() > 1e-3

    # the template-level chunk cache (LayerTemplate0) inherits the fix
    es = tq.EfficientSU2(n_wires=n, reps=2, entanglement_layer="linear").to(device)
    run(es)
    es.layers_all[0].ops_all[0] = tq.RX(has_params=True, trainable=True).to(device)
    q = tq.QuantumDevice(n_wires=n, bsz=2, device=device)
    q.h(wires=0)
    for layer_k in es.layers_all:
        for o, w in layer_k._layer_ops_wires():
            o(q, wires=w)
    assert_close(run(es), q.get_states_1d(), what="EfficientSU2 after replacing an operator")

    # ... and its epoch fast path (no walk over the layers while nothing was mutated) notices every kind of change
    def direct(model):
        q = tq.QuantumDevice(n_wires=n, bsz=2, device=device)
        q.h(wires=0)
        for layer_k in model.layers_all:
            for o, w in layer_k._layer_ops_wires():
                o(q, wires=w)
        return q.get_states_1d().clone()

    es = tq.EfficientSU2(n_wires=n, reps=2, entanglement_layer="linear").to(device)
    s0 = run(es)
    assert_close(run(es), s0, what="cached chunk, unchanged model")
    assert es.__dict__["_chunk_cache"][2] == tq.module.MUTATION_EPOCH[0]          # the fast path is what served it
    es.layers_all[0].ops_all[2].inverse = True                                     # attribute write on an operator
    s1 = run(es)
    assert_close(s1, direct(es), what="epoch: operator attribute")
    assert (s1 - s0).abs().max().item() > 1e-3
    es.layers_all[1].ops_all[0].params = torch.nn.Parameter(torch.tensor([[0.77]], device=device))   # new Parameter object
    s2 = run(es)
    assert_close(s2, direct(es), what="epoch: replaced Parameter")
    assert (s2 - s1).abs().max().item() > 1e-3
    es.layers_all[2].jump = 2                                                      # attribute write on a template layer
    assert_close(run(es), direct(es), what="epoch: layer attribute")
    es.layers_all[3] = tq.Op1QAllLayer(op=tq.RX, n_wires=n, has_params=True, trainable=True).to(device)   # list mutation
    s3 = run(es)
    assert_close(s3, direct(es), what="epoch: replaced layer")
    with torch.no_grad():                                                          # in-place edits need no invalidation
        es.layers_all[0].ops_all[0].params.add_(0.3)
    assert_close(run(es), direct(es), what="in-place parameter edit is read live")
    calls = []
    handle = es.layers_all[0].register_forward_hook(lambda *a: calls.append(1))   # a hook forces the per-layer path
    run(es)
    assert calls, "forward hook of a layer inside a template must run"
    handle.remove()
    qrec = tq.QuantumDevice(n_wires=n, bsz=2, device=device, record_op=True)       # op recording: no cached records
    es(qrec)
    assert len(qrec.op_history) == sum(len(layer_k._layer_ops_wires()) for layer_k in es.layers_all)


def check_grad_mode_is_sampled_when_gates_are_issued(device):
    """reading the state under no_grad mid-circuit must not drop the graph of the gates before it, and
    gates issued under no_grad must not be tracked when flushed outside it."""
    th = torch.tensor([[0.7]], device=device, requires_grad=True)
    ph = torch.tensor([[0.4]], device=device, requires_grad=True)

    def circuit(peek):
        q = tq.QuantumDevice(n_wires=2, bsz=1, device=device)
        tqf.rx(q, wires=0, params=th)
        if peek:
            with torch.no_grad():
                q.get_states_1d()
        tqf.ry(q, wires=0, params=ph)
        return tq.expval_joint_analytical(q, "ZI").sum()

    g_plain = torch.autograd.grad(circuit(False), [th, ph])
    g_peek = torch.autograd.grad(circuit(True), [th, ph])
    assert_close(torch.cat(g_peek).reshape(-1), torch.cat(g_plain).reshape(-1), rtol=GRAD_RTOL, what="grad after a no_grad peek")
    thc, phc = th.detach().cpu().requires_grad_(True), ph.detach().cpu().requires_grad_(True)
    s = O.run_ops(2, 1, [{"name": "rx", "wires": [0], "params": thc}, {"name": "ry", "wires": [0], "params": phc}])
    want = torch.autograd.grad(O.expval_joint_analytical(s, "ZI").sum(), [thc, phc])
    assert_close(torch.cat(g_peek).reshape(-1), torch.cat(want).reshape(-1), rtol=GRAD_RTOL, what="grad vs oracle")
    # the reverse: queued under no_grad, flushed outside -> constant w.r.t. th
    q = tq.QuantumDevice(n_wires=2, bsz=1, device=device)
    with torch.no_grad():
        tqf.rx(q, wires=0, params=th)
    tqf.ry(q, wires=0, params=ph)
    e = tq.expval_joint_analytical(q, "ZI").sum()
    g = torch.autograd.grad(e, [th, ph], allow_unused=True)
    assert g[0] is None and g[1] is not None


def check_parameter_values_are_read_at_the_gate_call(device):
    """the reference consumes parameter values when the gate is called."""
    q = tq.QuantumDevice(n_wires=2, bsz=1, device=device)
    p = torch.tensor([[0.3]], device=device)
    tqf.rx(q, wires=0, params=p)
    p.fill_(1.1)
    tqf.rx(q, wires=1, params=p)
    q.ry(wires=0, params=p)
    p.fill_(-2.0)
    want = _oracle_state(2, 1, [{"name": "rx", "wires": [0], "params": torch.tensor([[0.3]])},
                                {"name": "rx", "wires": [1], "params": torch.tensor([[1.1]])},
                                {"name": "ry", "wires": [0], "params": torch.tensor([[1.1]])}])
    assert_close(q.get_states_1d(), want, what="values at call time")
    # operators with non-trainable parameters go through the same snapshot
    op = tq.RX(has_params=True, trainable=False, init_params=0.5).to(device)
    q = tq.QuantumDevice(n_wires=1, bsz=1, device=device)
    op(q, wires=0)
    with torch.no_grad():
        op.params.fill_(0.9)
    op(q, wires=0)
    want = _oracle_state(1, 1, [{"name": "rx", "wires": [0], "params": torch.tensor([[0.5]])},
                                {"name": "rx", "wires": [0], "params": torch.tensor([[0.9]])}])
    assert_close(q.get_states_1d(), want, what="operator buffer values at call time")
    # a tensor that is part of the autograd graph is kept by reference; editing it in place is an error, not a
    # silently different circuit
    w = torch.tensor([[0.2]], device=device, requires_grad=True)
    v = w * 2
    q = tq.QuantumDevice(n_wires=1, bsz=1, device=device)
    tqf.rx(q, wires=0, params=v)
    with torch.no_grad():
        v.add_(1.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        q.get_states_1d()
    # layer templates queue their Parameters by reference (cached records): an optimizer step or any other in-place edit
    # between the layer call and the state read is an error too, for single layers and whole templates
    for layer in (tq.Op1QAllLayer(op=tq.RX, n_wires=3, has_params=True, trainable=True).to(device),
                  tq.EfficientSU2(n_wires=3, reps=1).to(device)):
        for _ in range(2):                       # second round: the records come from the layer's cache
            q = tq.QuantumDevice(n_wires=3, bsz=2, device=device)
            layer(q)
            q.get_states_1d()                    # untouched parameters: fine
            q = tq.QuantumDevice(n_wires=3, bsz=2, device=device)
            layer(q)
            with torch.no_grad():
                next(layer.parameters()).add_(0.25)
            with pytest.raises(RuntimeError, match="modified in place"):
                q.get_states_1d()


def check_checkpoint_round_trip(device):
    """`<prefix>q_device.states` is always a key, and a reference-format checkpoint loads strictly."""

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_device = tq.QuantumDevice(n_wires=3, bsz=2)
            self.rx = tq.RX(has_params=True, trainable=True)

        def forward(self):
            self.q_device.reset_states(2)
            self.rx(self.q_device, wires=1)
            return tq.expval_joint_analytical(self.q_device, "IZI")

    m = Model().to(device)
    sd0 = m.state_dict()
    assert "q_device.states" in sd0 and "q_device.state" in sd0
    assert tuple(sd0["q_device.states"].shape) == (2, 2, 2, 2)
    m()
    sd1 = m.state_dict()
    assert set(sd1) == set(sd0)
    # a checkpoint written by the reference: same keys, states [bsz, 2, 2, 2] complex64
    ref_sd = {"q_device.state": torch.zeros(2, 2, 2, dtype=torch.complex64),
              "q_device.states": torch.zeros(2, 2, 2, 2, dtype=torch.complex64),
              "rx.params": torch.tensor([[0.25]])}
    ref_sd["q_device.state"][0, 0, 0] = 1
    ref_sd["q_device.states"][:, 1, 0, 1] = 1
    fresh = Model().to(device)
    fresh.load_state_dict(ref_sd, strict=True)
    assert_close(fresh.q_device.states, ref_sd["q_device.states"], what="states restored")
    assert_close(fresh.rx.params.detach(), ref_sd["rx.params"], what="params restored")
    fresh2 = Model().to(device)
    fresh2.load_state_dict(sd1, strict=True)
    assert_close(fresh2.q_device.get_states_1d(), m.q_device.get_states_1d(), what="own checkpoint")
    with pytest.raises(RuntimeError, match="states"):
        bad = dict(ref_sd)
        del bad["q_device.states"]
        Model().to(device).load_state_dict(bad, strict=True)


def check_from_op_history_and_get_unitary(device):
    """op_history wire format -> module -> same state; get_unitary column check."""
    qdev = tq.QuantumDevice(n_wires=3, bsz=2, record_op=True, device=device)
    qdev.h(wires=0)
    qdev.crx(wires=[0, 2], params=torch.tensor([0.4]))
    qdev.u3(wires=1, params=torch.tensor([[0.1, 0.2, 0.3]]))
    qdev.cnot(wires=[2, 1], inverse=True)
    want = qdev.get_states_1d().clone()
    mod = tq.QuantumModule.from_op_history(qdev.op_history).to(device)
    q2 = tq.QuantumDevice(n_wires=3, bsz=2, device=device)
    mod(q2)
    assert_close(q2.get_states_1d(), want, what="from_op_history")
    mod.n_wires = 3
    u = mod.get_unitary(device=device)
    assert_close(u[:, 0], want[0], what="get_unitary column 0")
    assert_close((u.conj().T @ u).cpu(), torch.eye(8, dtype=torch.complex64), what="unitarity")
    ref = _oracle_state(3, 2, [{"name": "hadamard", "wires": [0]},
                               {"name": "crx", "wires": [0, 2], "params": torch.tensor([[0.4]])},
                               {"name": "u3", "wires": [1], "params": torch.tensor([[0.1, 0.2, 0.3]])},
                               {"name": "cnot", "wires": [2, 1], "inverse": True}])
    assert_close(want, ref, what="recorded circuit vs oracle")


def check_static_mode_equals_dynamic(device):
    """SURVEY §8 a9 / f2 (test/static_mode_test.py: static == dynamic, states 1e-5, gradients)."""
    np.random.seed(3)
    torch.manual_seed(3)
    layer = tq.RandomLayerAllTypes(n_ops=60, wires=list(range(5))).to(device)
    params = [p for p in layer.parameters()]

    def run():
        q = tq.QuantumDevice(n_wires=5, bsz=3, device=device)
        layer(q)
        s = q.get_states_1d()
        loss = tq.expval_joint_analytical(q, "ZXIYZ").sum()
        return s.detach().clone(), torch.autograd.grad(loss, params, allow_unused=True)

    dyn, gdyn = run()
    for wpb in (2, 3, 4):
        layer.static_on(wires_per_block=wpb)
        st, gst = run()
        assert_close(st, dyn, what=f"static({wpb}) == dynamic")
        for a, b in zip(gst, gdyn):
            assert (a is None) == (b is None)
            if a is not None:
                assert_close(a, b, rtol=GRAD_RTOL, what=f"static({wpb}) gradients")
        layer.static_off()
    assert not layer.static_mode


def check_layer_classes(device):
    """a15: RandomLayer / QFTLayer / MeasureAll classes (and Op2QAllLayer circular) against the oracle."""
    n = 6
    q = tq.QuantumDevice(n_wires=n, bsz=1, device=device)
    q.x(wires=2)
    tq.QFTLayer(n_wires=n)(q)
    assert_close(q.get_states_1d().abs().cpu(), torch.full((1, 64), 1 / 8.0), what="QFT|x> is flat")
    from tests import cases
    want = _oracle_state(n, 1, [{"name": "paulix", "wires": [2]}] + cases.qft_ops(n))
    assert_close(q.get_states_1d(), want, what="QFTLayer vs oracle QFT ops")
    z = tq.MeasureAll(tq.PauliZ)(q)
    zw, _ = O.expval(want.reshape([1] + [2] * n), list(range(n)), ["Z"] * n)
    assert_close(z, zw, what="MeasureAll(PauliZ) after QFT")
    np.random.seed(7)
    rl = tq.RandomLayer(n_ops=30, wires=list(range(n)), seed=7).to(device)
    q = tq.QuantumDevice(n_wires=n, bsz=2, device=device, record_op=True)
    rl(q)
    got = q.get_states_1d()
    ops = []
    for h in q.op_history:
        p = h["params"]
        ops.append({"name": h["name"], "wires": h["wires"] if isinstance(h["wires"], list) else [h["wires"]],
                    "params": None if p is None else torch.tensor(p, dtype=torch.float32).reshape(1, -1),
                    "inverse": h["inverse"]})
    assert_close(got, _oracle_state(n, 2, ops), what="RandomLayer vs oracle replay of its op_history")


def check_einsum_broadcast(device):
    """a3/a4: the pure functions; apply_unitary_einsum broadcasts a bsz-1 state over a batched matrix
    (functional/gate_wrapper.py:18-93), apply_unitary_bmm refuses a batch mismatch."""
    torch.manual_seed(1)
    n = 4
    st = torch.randn(1, 2 ** n, dtype=torch.complex64)
    st = (st / st.norm()).reshape([1] + [2] * n)
    mats = torch.linalg.qr(torch.randn(3, 4, 4, dtype=torch.complex64))[0]
    got = tqf.apply_unitary_einsum(st.to(device), mats.to(device), [2, 0])
    assert tuple(got.shape) == (3, 2, 2, 2, 2)
    for b in range(3):
        want = O.apply_unitary_bmm(st, mats[b], [2, 0])
        assert_close(got[b:b + 1], want, what=f"einsum broadcast sample {b}")
    st3 = torch.randn(3, 2 ** n, dtype=torch.complex64).reshape([3] + [2] * n)
    got = tqf.apply_unitary_bmm(st3.to(device), mats.to(device), [1, 3])
    want = O.apply_unitary_bmm(st3, mats, [1, 3])
    assert_close(got, want, what="bmm, batched matrix")
    got = tqf.apply_unitary_bmm(st3.to(device), mats[0].to(device), [0, 1])
    assert_close(got, O.apply_unitary_bmm(st3, mats[0], [0, 1]), what="bmm, shared matrix")
    with pytest.raises(RuntimeError):
        tqf.apply_unitary_bmm(st.to(device), mats.to(device), [2, 0])


def check_dropped_device_is_freed_by_refcount(device):
    """A device that goes out of scope must be freed at once (no reference cycle): its state can be GiBs, and a stale
    autograd graph kept alive behind it breaks CUDA-graph capture of the next step (AccumulateGrad stream mismatch)."""
    import gc
    import weakref

    gc.disable()
    try:
        q = tq.QuantumDevice(n_wires=3, bsz=2, device=device)
        p = torch.tensor([[0.3]], device=device, requires_grad=True)
        tqf.rx(q, wires=0, params=p)
        q.get_states_1d()
        q.ry(wires=1, params=p)
        ref = weakref.ref(q)
        del q
        assert ref() is None, "QuantumDevice is kept alive by a reference cycle"
    finally:
        gc.enable()


def check_reference_test_vectors(device, tmp_path=None):
    """tests/golden/reftests.npz (oracle/gen_reftests.py: outputs of the REAL reference, incl. the inputs of its own
    literal-vector tests): controlled_unitary / from_controlled_operation, matrix_exp, Hamiltonian matrices and file
    format, device construction, observable grouping, sampling estimators and parameter-shift gradients under the
    shared `random` stream (bit-exact), Hadamard-test gradients."""
    import json
    import os
    import random
    import tempfile

    from tests.helpers import load

    z, man = load("reftests.npz")
    # --- controlled unitaries (f3)
    for name, n, prep, (gname, gparam), cw, tw in man["controlled"]:
        for via in ("cu", "cop"):
            q = tq.QuantumDevice(n_wires=n, device=device)
            for g, w in prep:
                getattr(q, g)(wires=w)
            cls = tq.op_name_dict[gname]
            op = cls(has_params=True, init_params=gparam) if gparam is not None else cls()
            if via == "cu":
                q.controlled_unitary(params=op.matrix, c_wires=cw, t_wires=tw)
            else:
                tq.QubitUnitaryFast.from_controlled_operation(op=op, c_wires=cw, t_wires=tw, trainable=False).to(device)(q)
            assert_close(q.get_states_1d(), z[f"{via}_{name}"], what=f"controlled unitary {name} ({via})")
    # --- matrix_exp (f3): the reference's literals are non-unitary matrices (entries up to 750), same bar
    for k, (n, bsz, wires, mat) in enumerate(man["mat_exp"]):
        m = torch.tensor([[complex(re, im) for re, im in row] for row in mat], dtype=torch.complex64)
        q = tq.QuantumDevice(n_wires=n, device=device)
        q.reset_states(bsz=bsz)
        q.matrix_exp(wires=wires, params=m)
        assert_close(q.get_states_1d(), z[f"matexp_{k}"], what=f"matrix_exp case {k}")
    # --- Hamiltonian (f1)
    for k, (coeffs, paulis) in enumerate(man["hamiltonians"]):
        assert_close(tq.Hamiltonian(coeffs, paulis).get_matrix(), z[f"hamil_{k}"], what=f"Hamiltonian {paulis}")
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(man["h2_file"])
    try:
        assert_close(tq.Hamiltonian.from_file(f.name).get_matrix(), z["hamil_h2"], what="Hamiltonian.from_file(h2)")
    finally:
        os.unlink(f.name)
    # --- device construction
    got = tq.QuantumDevice(n_wires=2, bsz=2, device=device).get_states_1d()
    assert np.array_equal(got.cpu().numpy(), z["init_2_2"])
    # --- observable grouping
    assert tq.find_observable_groups(list(man["groups_in"])) == man["groups_out"]
    # --- sampling estimators: same `random` stream => same histograms => identical estimates (a14)
    ops = [{"name": o["name"], "wires": o["wires"], "inverse": o["inverse"],
            "params": None if o["params"] is None else torch.tensor(o["params"], dtype=torch.float32).reshape(1, -1)}
           for o in man["sampling_ops"]]
    from tests import cases
    q = tq.QuantumDevice(n_wires=4, bsz=2, device=device)
    cases.apply_ops(q, ops, device)
    obs = man["sampling_obs"]
    ana = torch.stack([tq.expval_joint_analytical(q, o) for o in obs])
    assert_close(ana, z["joint_analytical"], what="analytical expvals of the sampled state")
    random.seed(5)
    got = np.stack([tq.expval_joint_sampling(q, o, n_shots=2000).numpy() for o in obs])
    assert np.array_equal(got, z["joint_sampling"]), "expval_joint_sampling is not bit-exact with the reference"
    assert random.random() == float(z["rng_after_joint"][0]), "RNG stream position after expval_joint_sampling"
    random.seed(6)
    res = tq.expval_joint_sampling_grouping(q, obs, n_shots_per_group=3000)
    assert np.array_equal(np.stack([res[o].numpy() for o in obs]), z["grouping"]), "grouped sampling is not bit-exact"
    assert random.random() == float(z["rng_after_grouping"][0])
    assert np.abs(got - z["joint_analytical"]).max() < 0.08   # and they estimate the analytical values

    # --- parameter-shift gradient from shots (f4)
    class PSModel(tq.QuantumModule):
        def __init__(self):
            super().__init__()
            self.n_wires = 2
            self.q_device = tq.QuantumDevice(n_wires=2)
            self.rx = tq.RX(has_params=True, trainable=True, init_params=0.4)
            self.ry = tq.RY(has_params=True, trainable=True, init_params=-0.9)
            self.crz = tq.CRZ(has_params=True, trainable=True, init_params=1.3)

        def forward(self, x):
            self.q_device.reset_states(1)
            self.q_device.h(wires=1)
            self.rx(self.q_device, wires=0)
            self.ry(self.q_device, wires=1)
            self.crz(self.q_device, wires=[0, 1])
            self.q_device.rx(wires=1, params=x)

    model = PSModel().to(device)
    assert list(model.state_dict().keys()) == man["ps_state_dict_keys"]
    before = [p.detach().clone() for p in model.parameters()]
    random.seed(21)
    g = tq.parameter_shift_gradient(model, torch.tensor([[0.3]], device=device), "ZZ", shots=4000)
    assert np.array_equal(g.numpy(), z["param_shift"]), (g, z["param_shift"])
    assert random.random() == float(z["rng_after_param_shift"][0])
    for p, b in zip(model.parameters(), before):
        assert torch.equal(p.detach(), b), "parameters must be restored after the shifts"
    # the estimate agrees with the adjoint gradient of the analytical expectation value
    model.forward(torch.tensor([[0.3]], device=device))
    ev = tq.expval_joint_analytical(model.q_device, "ZZ").sum()
    exact = torch.autograd.grad(ev, list(model.parameters()))
    assert (torch.stack([e.reshape(()) for e in exact]).cpu() - g).abs().max().item() < 0.06

    # --- Hadamard-test gradients (f4): one batched (n+1)-qubit circuit for all (gate, Pauli term) pairs
    for k, gates in enumerate(man["pauli_exp_circuits"]):
        mods = [tq.OpPauliExp(coeffs=c, paulis=p, theta=th, trainable=True).to(device) for th, c, p in gates]
        q = tq.QuantumDevice(n_wires=4, bsz=1, record_op=True, device=device)
        q.h(wires=1)
        q.cnot(wires=[1, 0])
        q.ry(wires=2, params=torch.tensor([[0.37]]))
        for m in mods:
            m(q, wires=[0, 1, 2, 3])
            q.cnot(wires=[3, 2])
        ev = tq.expval_joint_analytical(q, "ZZZZ")
        assert_close(ev, z[f"hgrad_expval_{k}"], what=f"Pauli-exp circuit {k} expectation")
        assert len(q.op_history) == 3 + 2 * len(mods), "one op_history entry per gate"
        back = torch.autograd.grad(ev.sum(), [m.theta for m in mods])
        assert_close(torch.stack(back), z[f"hgrad_backprop_{k}"], rtol=GRAD_RTOL, what=f"adjoint gradient, circuit {k}")
        hg = [x for x in tq.hadamard_grad(q.op_history, 4, "ZZZZ", device=device) if x is not None]
        assert_close(torch.stack(hg), z[f"hgrad_{k}"], rtol=GRAD_RTOL, what=f"hadamard_grad vs reference, circuit {k}")
        # two different float32 circuits for the same derivative (the reference's own values differ by 4e-6 here; its
        # test allows 1e-3, test/hadamard_grad/test_hadamard_grad.py:33)
        assert (torch.stack(hg).cpu() - torch.stack(back).cpu()).abs().max().item() < 1e-4, f"hadamard_grad vs adjoint, circuit {k}"


def check_batch_collapse(device, n=7, bsz=300):
    """engine batch collapse: a long run of shared-parameter gates on a batch larger than the Hilbert space goes through
    ONE unitary (built on the identity batch) and the 3xTF32 tensor-core GEMM; states, expectation values and all gradients
    (shared parameters through the GEMM's backward + the adjoint sweep over the unitary's columns, per-sample encoder angles
    through the GEMM's other backward) must equal the gate-by-gate engine and the oracle."""
    from tests import cases
    from torchquantum_b200 import engine

    assert engine.COLLAPSE and bsz >= 2 * 2 ** n
    launches0 = engine.STATS["launches"]
    cases.check_cfg2_vqe(device, n=n, reps=4, bsz=bsz)          # vs the oracle, through the collapse (bsz not a multiple of 128)
    torch.manual_seed(3)
    ops = cases.brickwork_ops(n, 8, seed=2)                      # shared angles only: the whole circuit is one unitary
    x = torch.randn(bsz, 2 ** n, dtype=torch.complex64)
    x = (x / x.norm(dim=1, keepdim=True)).to(device)

    def run(collapse):
        saved = engine.COLLAPSE
        engine.COLLAPSE = collapse
        try:
            st = x.clone().requires_grad_(True)
            q = tq.QuantumDevice(n_wires=n, bsz=bsz, device=device)
            q.set_states(st)
            ps = [o["params"].clone().to(device).requires_grad_(True) if o.get("params") is not None else None for o in ops]
            cases.apply_ops(q, [dict(o, params=p) if p is not None else o for o, p in zip(ops, ps)], device)
            out = q.get_states_1d()
            w = torch.linspace(0.5, 1.5, 2 ** n, device=device)
            loss = (out.abs() ** 2 * w).sum() + tq.expval_joint_analytical(q, "ZX" + "I" * (n - 2)).sum()
            grads = torch.autograd.grad(loss, [st] + [p for p in ps if p is not None])
            return out.detach(), grads
        finally:
            engine.COLLAPSE = saved

    assert len(ops) >= engine.COLLAPSE_MIN_GATES
    s1, g1 = run(True)
    s0, g0 = run(False)
    assert_close(s1, s0, what="collapsed states vs gate-by-gate")
    assert_close(g1[0], g0[0], rtol=GRAD_RTOL, what="dL/dpsi_in through the collapsed run")
    assert_close(torch.cat([g.reshape(-1) for g in g1[1:]]), torch.cat([g.reshape(-1) for g in g0[1:]]), rtol=GRAD_RTOL,
                 what="shared-parameter gradients through the collapsed run")
    assert engine.STATS["launches"] > launches0
    if bsz == 300:
        # tile-aligned shapes take their GEMM operands from the fused preparation kernel (tqb_gemm_prep: conjugate /
        # transpose / real embedding / TF32 split in one pass), forward and backward
        check_batch_collapse(device, n=n, bsz=256)
        g = torch.Generator().manual_seed(9)
        a = torch.randn(256, 64, dtype=torch.complex64, generator=g).to(device)
        b = torch.randn(64, 128, dtype=torch.complex64, generator=g).to(device)
        assert_close(engine.complex_gemm(a, b), (a.cpu().to(torch.complex128) @ b.cpu().to(torch.complex128)).to(torch.complex64),
                     what="complex_gemm, fused operand preparation")
        x = torch.randn(64, 128, dtype=torch.complex64, generator=g).to(device)
        y = torch.randn(64, 256, dtype=torch.complex64, generator=g).to(device)
        want = (x.cpu().to(torch.complex128).conj().T @ y.cpu().to(torch.complex128)).to(torch.complex64)
        assert_close(engine.complex_gemm(x, y, a_h=True), want, what="complex_gemm(a^H b), fused operand preparation")

"""Differential probing: same API calls on the live reference and on this package (emulation)."""
import sys, random, traceback
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from oracle.refshim import load_reference
ref, _reftqf = load_reference()
from tests.test_cpu_emulated import build_emu
from torchquantum_b200 import _native
_native.use_library_for_testing(build_emu())
import torchquantum_b200 as tq

def both(label, fn):
    outs = []
    for mod in (ref, tq):
        try:
            random.seed(3); np.random.seed(3); torch.manual_seed(3)
            o = fn(mod)
            outs.append(("ok", o))
        except Exception as e:
            outs.append(("err", type(e).__name__ + ": " + str(e)[:70]))
    (ka, a), (kb, b) = outs
    if ka != kb:
        print(f"[DIFF] {label}: ref={ka}:{a if ka=='err' else ''} ours={kb}:{b if kb=='err' else ''}")
        return
    if ka == "err":
        same = a.split(":")[0] == b.split(":")[0]
        print(f"[{'same-err' if same else 'ERRTYPE'}] {label}: ref={a} | ours={b}")
        return
    def cmp(x, y):
        if isinstance(x, torch.Tensor):
            return x.shape == y.shape and torch.allclose(x.detach().cpu().to(torch.complex64), y.detach().cpu().to(torch.complex64), atol=2e-6)
        if isinstance(x, (list, tuple)):
            return len(x) == len(y) and all(cmp(p, q) for p, q in zip(x, y))
        if isinstance(x, dict):
            return list(x.keys()) == list(y.keys()) and all(cmp(x[k], y[k]) for k in x)
        if isinstance(x, float):
            return abs(x - y) < 2e-6
        return x == y
    ok = cmp(a, b)
    print(f"[{'ok' if ok else 'VALUE-DIFF'}] {label}" + ("" if ok else f"\n   ref={a}\n   ours={b}"))

def dev(m, n=3, bsz=2, **kw):
    return m.QuantumDevice(n_wires=n, bsz=bsz, **kw)

both("reset_identity_states", lambda m: (lambda q: (q.reset_identity_states(), q.get_states_1d())[1])(dev(m)))
both("reset_all_eq_states", lambda m: (lambda q: (q.reset_all_eq_states(4), q.get_states_1d())[1])(dev(m)))
both("states shape", lambda m: tuple(dev(m).states.shape))
both("state shape", lambda m: tuple(dev(m).state.shape))
both("set_states wrong bsz then gate", lambda m: (lambda q: (q.set_states(torch.ones(5, 8, dtype=torch.complex64) / 8 ** 0.5), q.h(wires=0), q.get_states_1d())[2])(dev(m)))
both("clone_states", lambda m: (lambda q: (q.clone_states(torch.eye(8, dtype=torch.complex64)[:3].reshape(3, 2, 2, 2)), q.x(wires=2), q.get_states_1d())[2])(dev(m)))
both("rx 1-D params", lambda m: (lambda q: (q.rx(wires=0, params=torch.tensor([0.3, 0.6])), q.get_states_1d())[1])(dev(m)))
both("rx scalar param", lambda m: (lambda q: (q.rx(wires=1, params=0.4), q.get_states_1d())[1])(dev(m)))
both("u3 list params", lambda m: (lambda q: (q.u3(wires=1, params=[0.4, 0.1, 0.2]), q.get_states_1d())[1])(dev(m)))
both("u3 [bsz,3]", lambda m: (lambda q: (q.u3(wires=1, params=torch.tensor([[0.4, 0.1, 0.2], [0.3, 0.2, 0.9]])), q.get_states_1d())[1])(dev(m)))
both("multirz n_wires 3", lambda m: (lambda q: (q.h(wires=0), q.multirz(wires=[0, 1, 2], params=0.7, n_wires=3), q.get_states_1d())[2])(dev(m)))
both("multicnot 3", lambda m: (lambda q: (q.x(wires=0), q.x(wires=1), q.multicnot(wires=[0, 1, 2], n_wires=3), q.get_states_1d())[3])(dev(m)))
both("multixcnot 3", lambda m: (lambda q: (q.multixcnot(wires=[0, 1, 2], n_wires=3), q.get_states_1d())[1])(dev(m)))
both("qft 3", lambda m: (lambda q: (q.x(wires=1), q.qft(wires=[0, 1, 2], n_wires=3), q.get_states_1d())[2])(dev(m)))
both("qft inverse", lambda m: (lambda q: (q.x(wires=1), q.qft(wires=[2, 0], n_wires=2, inverse=True), q.get_states_1d())[2])(dev(m)))
both("globalphase", lambda m: (lambda q: (q.globalphase(wires=0, params=0.7), q.get_states_1d())[1])(dev(m)))
both("inverse flag u2", lambda m: (lambda q: (q.u2(wires=2, params=[0.3, 0.8], inverse=True), q.get_states_1d())[1])(dev(m)))
both("cu2", lambda m: (lambda q: (q.x(wires=0), q.cu2(wires=[0, 2], params=[0.3, 0.8]), q.get_states_1d())[2])(dev(m)))
both("expval int wire", lambda m: (lambda q: (q.rx(wires=1, params=0.4), m.measurement.expval(q, 1, m.PauliZ()))[1])(dev(m)))
both("expval list X,Y", lambda m: (lambda q: (q.rx(wires=1, params=0.4), q.ry(wires=0, params=0.9), m.measurement.expval(q, [0, 1], [m.PauliX(), m.PauliY()]), q.get_states_1d())[2:])(dev(m)))
both("expval_joint lower", lambda m: (lambda q: (q.rx(wires=1, params=0.4), m.measurement.expval_joint_analytical(q, "zxi"))[1])(dev(m)))
both("expval_joint bad len", lambda m: m.measurement.expval_joint_analytical(dev(m), "ZZ"))
both("expval_joint bad char", lambda m: m.measurement.expval_joint_analytical(dev(m), "ZQZ"))
both("measure draw none", lambda m: (lambda q: (q.h(wires=0), [dict(d) for d in m.measurement.measure(q, n_shots=50)])[1])(dev(m)))
both("MeasureAll", lambda m: (lambda q: (q.rx(wires=1, params=0.4), m.MeasureAll(m.PauliZ)(q))[1])(dev(m)))
both("MeasureAll X", lambda m: (lambda q: (q.rx(wires=1, params=0.4), m.MeasureAll(m.PauliX)(q), q.get_states_1d())[1:])(dev(m)))
both("MeasureMultipleTimes", lambda m: (lambda q: (q.rx(wires=1, params=0.4), m.MeasureMultipleTimes([{"wires": [0, 1], "observables": ["x", "y"]}, {"wires": [2], "observables": ["z"]}])(q))[1])(dev(m)))
both("MeasureMultiPauliSum", lambda m: (lambda q: (q.rx(wires=1, params=0.4), m.MeasureMultiPauliSum([{"wires": [0, 1], "observables": ["x", "y"]}, {"wires": [2], "observables": ["z"]}])(q))[1])(dev(m)))
both("op_history", lambda m: (lambda q: (q.rx(wires=1, params=0.4), q.cnot(wires=[0, 2]), q.u3(wires=0, params=[0.1, 0.2, 0.3]), q.op_history)[3])(dev(m, record_op=True)))
both("Operator params overwrite", lambda m: (lambda q, o: (o(q, wires=0, params=torch.tensor([[0.25]])), o.params.detach().clone(), q.get_states_1d())[1:])(dev(m), m.RX(has_params=True, trainable=True)))
both("Operator init_params", lambda m: (lambda q, o: (o(q, wires=[0, 1]), q.get_states_1d())[1])(dev(m), m.CRX(has_params=True, init_params=0.5)))
both("Operator inverse", lambda m: (lambda q, o: (o(q, wires=0), q.get_states_1d())[1])(dev(m), m.U3(has_params=True, init_params=[0.5, 0.1, 0.2], inverse=True)))
both("QubitUnitary op", lambda m: (lambda q, o: (o(q, wires=[1, 0]), q.get_states_1d())[1])(dev(m), m.QubitUnitary(has_params=True, init_params=[[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1j], [0, 0, -1j, 0]])))
both("TrainableUnitary", lambda m: (lambda q, o: (o(q, wires=[1, 2]), q.get_states_1d().abs().sum() > 0)[1])(dev(m), m.TrainableUnitary(has_params=True, trainable=True, n_wires=2)))
both("RandomLayer seed", lambda m: (lambda q, L: (L(q), q.get_states_1d())[1])(dev(m), m.RandomLayer(n_ops=15, wires=[0, 1, 2], seed=7)))
both("RandomLayer qiskit_compatible", lambda m: (lambda q, L: (L(q), q.get_states_1d())[1])(dev(m), m.RandomLayer(n_ops=15, wires=[0, 1, 2], seed=7, qiskit_compatible=True)))
both("RandomLayerAllTypes", lambda m: (lambda q, L: (L(q), q.get_states_1d())[1])(dev(m, n=4), m.RandomLayerAllTypes(n_ops=20, wires=[0, 1, 2, 3], seed=9)))
both("QFTLayer", lambda m: (lambda q, L: (q.x(wires=2), L(q), q.get_states_1d())[2])(dev(m), m.layer.QFTLayer(n_wires=3)))
both("QFTLayer inverse no swaps", lambda m: (lambda q, L: (q.x(wires=2), L(q), q.get_states_1d())[2])(dev(m), m.layer.QFTLayer(n_wires=3, do_swaps=False, inverse=True)))
both("Op2QAllLayer circular", lambda m: (lambda q, L: (q.h(wires=0), L(q), q.get_states_1d())[2])(dev(m), m.Op2QAllLayer(op=m.CNOT, n_wires=3, circular=True)))
both("encoder 4x4_ryzxy", lambda m: (lambda q, E: (E(q, torch.arange(32.).reshape(2, 16) / 10), q.get_states_1d())[1])(dev(m, n=4), m.GeneralEncoder(m.encoder_op_list_name_dict["4x4_ryzxy"])))
both("PhaseEncoder", lambda m: (lambda q, E: (E(q, torch.arange(6.).reshape(2, 3) / 10), q.get_states_1d())[1])(dev(m), m.PhaseEncoder(m.functional.rx)))
both("StateEncoder", lambda m: (lambda q, E: (E(q, torch.arange(10.).reshape(2, 5) / 10), q.get_states_1d())[1])(dev(m), m.StateEncoder()))
both("MultiPhaseEncoder", lambda m: (lambda q, E: (E(q, torch.arange(12.).reshape(2, 6) / 10), q.get_states_1d())[1])(dev(m), m.MultiPhaseEncoder(["u3", "rx", "ry", "rz"][:2], wires=[0, 2])))

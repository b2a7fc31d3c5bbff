"""Layer templates with their constructor options: live reference vs this package (parameters copied by position)."""
import sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from oracle.refshim import load_reference
ref, reftqf = load_reference()
from tests.test_cpu_emulated import build_emu
from torchquantum_b200 import _native
_native.use_library_for_testing(build_emu())
import torchquantum_b200 as tq

n, bsz = 4, 2
rng = np.random.default_rng(8)
init = rng.normal(size=(bsz, 2 ** n)) + 1j * rng.normal(size=(bsz, 2 ** n))
init = torch.tensor((init / np.linalg.norm(init, axis=1, keepdims=True)).astype(np.complex64))

def run(mod, build, seed=3):
    torch.manual_seed(seed); np.random.seed(seed)
    layer = build(mod)
    q = mod.QuantumDevice(n_wires=n, bsz=bsz)
    q.set_states(init.clone().reshape([bsz] + [2] * n))
    layer(q)
    return layer, q.get_states_1d().detach().clone()

CASES = {
    "TwoLocal linear": lambda m: m.layer.TwoLocal(n_wires=n, rotation_ops=[m.RY, m.RZ], entanglement_ops=[m.CNOT], entanglement_layer="linear", reps=2),
    "TwoLocal reverse_linear": lambda m: m.layer.TwoLocal(n_wires=n, rotation_ops=[m.RY], entanglement_ops=[m.CNOT], entanglement_layer="reverse_linear", reps=2),
    "TwoLocal circular": lambda m: m.layer.TwoLocal(n_wires=n, rotation_ops=[m.RX], entanglement_ops=[m.CZ], entanglement_layer="circular", reps=1),
    "TwoLocal full": lambda m: m.layer.TwoLocal(n_wires=n, rotation_ops=[m.RY], entanglement_ops=[m.CNOT], entanglement_layer="full", reps=1),
    "TwoLocal skip_final": lambda m: m.layer.TwoLocal(n_wires=n, rotation_ops=[m.RY], entanglement_ops=[m.CRX], entanglement_layer="linear", reps=2, skip_final_rotation_layer=True, entanglement_layer_params={"has_params": True, "trainable": True}),
    "EfficientSU2 r3": lambda m: m.layer.EfficientSU2(n_wires=n, reps=3, entanglement_layer="reverse_linear"),
    "RealAmplitudes": lambda m: m.layer.RealAmplitudes(n_wires=n, reps=2, entanglement_layer="circular"),
    "Op2QDenseLayer": lambda m: m.layer.Op2QDenseLayer(op=m.CNOT, n_wires=n),
    "Op2QAllLayer jump2 reverse": lambda m: m.layer.Op2QAllLayer(op=m.CRY, n_wires=n, has_params=True, trainable=True, jump=2, wire_reverse=True),
    "Op1QAllLayer U3": lambda m: m.layer.Op1QAllLayer(op=m.U3, n_wires=n, has_params=True, trainable=True),
    "RandomLayer op_ratios": lambda m: m.RandomLayer(n_ops=25, wires=[0, 1, 2, 3], op_ratios=[0.2, 0.5, 0.3], op_types=[m.RX, m.CNOT, m.U3], seed=11),
    "RandomLayer n_params": lambda m: m.RandomLayer(n_ops=10, n_params=6, wires=[3, 1, 0], seed=2),
    "QFTLayer wires": lambda m: m.layer.QFTLayer(n_wires=3, wires=[2, 0, 3]),
}
ok = bad = 0
for name, build in CASES.items():
    try:
        lr, sr = run(ref, build)
    except Exception as e:
        print(f"[ref-err] {name}: {type(e).__name__}: {str(e)[:80]}"); continue
    try:
        lo, so = run(tq, build)
        pr, po = list(lr.parameters()), list(lo.parameters())
        same_init = len(pr) == len(po) and all(torch.allclose(a, b) for a, b in zip(pr, po))
        d = (sr - so).abs().max().item()
        if d < 2e-6:
            ok += 1; print(f"[ok] {name} ({len(po)} params, same init: {same_init})")
        else:
            bad += 1; print(f"[VALUE-DIFF] {name}: {d:.2e}  params {len(pr)} vs {len(po)} same_init={same_init}")
    except Exception as e:
        bad += 1; print(f"[ours-err] {name}: {type(e).__name__}: {str(e)[:100]}")
print("ok", ok, "bad", bad)

"""Every class of op_name_dict on random wires: live reference vs this package (emulation)."""
import sys, random, traceback
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from oracle.refshim import load_reference
ref, reftqf = load_reference()
from tests.test_cpu_emulated import build_emu
from torchquantum_b200 import _native
_native.use_library_for_testing(build_emu())
import torchquantum_b200 as tq

n, bsz = 5, 2
rng = np.random.default_rng(5)
init = rng.normal(size=(bsz, 2 ** n)) + 1j * rng.normal(size=(bsz, 2 ** n))
init = torch.tensor((init / np.linalg.norm(init, axis=1, keepdims=True)).astype(np.complex64))
names = sorted(set(ref.op_name_dict) | set(tq.op_name_dict))
print(len(ref.op_name_dict), "reference names,", len(tq.op_name_dict), "ours; missing here:", sorted(set(ref.op_name_dict) - set(tq.op_name_dict)), "extra here:", sorted(set(tq.op_name_dict) - set(ref.op_name_dict)))
ok = bad = skipped = 0
for name in names:
    if name not in ref.op_name_dict or name not in tq.op_name_dict:
        continue
    outs = []
    for mod in (ref, tq):
        try:
            cls = mod.op_name_dict[name]
            nw = cls.num_wires if type(cls.num_wires) is int and cls.num_wires > 0 else 3
            npar = cls.num_params if type(cls.num_params) is int and cls.num_params > 0 else 0
            if name == "multirz":
                npar = 1
            if name in ("globalphase", "reset"):
                raise KeyError("not a wires gate")
            r = np.random.default_rng(abs(hash(name)) % 1000)
            wires = [int(w) for w in r.choice(n, nw, replace=False)]
            kw = {}
            if npar > 0:
                kw = dict(has_params=True, trainable=False, init_params=[float(v) for v in r.uniform(-3, 3, npar)])
            if name in ("qubitunitary", "qubitunitaryfast", "qubitunitarystrict", "trainableunitary", "trainableunitarystrict"):
                raise KeyError("explicit-matrix class: covered by unitary_gates.npz")
            if not (type(cls.num_wires) is int and cls.num_wires > 0):
                kw["n_wires"] = nw
            op = cls(**kw)
            q = mod.QuantumDevice(n_wires=n, bsz=bsz)
            q.set_states(init.clone().reshape([bsz] + [2] * n))
            op(q, wires=wires if nw > 1 else wires[0])
            inv = cls(**kw, inverse=True) if "inverse" in cls.__init__.__code__.co_varnames else None
            outs.append(("ok", q.get_states_1d().detach().clone()))
        except Exception as e:
            outs.append(("err", f"{type(e).__name__}: {str(e)[:60]}"))
    (ka, a), (kb, b) = outs
    if ka == "ok" and kb == "ok":
        d = (a - b).abs().max().item()
        if d < 2e-6: ok += 1
        else:
            bad += 1; print(f"[VALUE-DIFF] {name}: {d:.2e}")
    elif ka == "err" and kb == "err":
        skipped += 1; print(f"[both-err] {name}: ref={a} | ours={b}")
    else:
        bad += 1; print(f"[DIFF] {name}: ref={ka} {a if ka=='err' else ''} | ours={kb} {b if kb=='err' else ''}")
print("ok", ok, "bad", bad, "both-err", skipped)

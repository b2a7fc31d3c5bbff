"""Fuzz the forward-only planner (blocks + quadphase + register-tile) against the oracle on the host emulation:
MODE=auto|rotate|inplace python tools/fuzz_blocks.py N SEED0 SEED1 GATES."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
import torchquantum_b200 as tq
from torchquantum_b200 import _native, gates as G, engine
from tests.test_cpu_emulated import build_emu
from tests import cases
from oracle import tq_oracle as O
_native.use_library_for_testing(build_emu())

NAMES = [k for k, s in G.SPECS.items() if s.kind in (G.U1Q, G.DIAG, G.U2Q) and k != "globalphase"]
def rand_ops(n, count, rng, names=NAMES, p_inv=0.3, local=0.6):
    ops = []
    for _ in range(count):
        name = names[rng.integers(len(names))]
        s = G.SPECS[name]
        k = s.n_wires if s.n_wires is not None else int(rng.integers(2, 5))
        if rng.random() < local:   # wires near each other (blocks absorb more)
            base = int(rng.integers(0, n - 6)); wires = [int(w) for w in rng.choice(np.arange(base, base + 6), k, replace=False)]
        else:
            wires = [int(w) for w in rng.choice(n, k, replace=False)]
        op = {"name": name, "wires": wires, "inverse": bool(rng.random() < p_inv)}
        if s.n_params:
            op["params"] = torch.tensor(rng.uniform(-np.pi, np.pi, (1, s.n_params)), dtype=torch.float32)
        ops.append(op)
    return ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seeds = range(int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else range(10)
count = int(sys.argv[4]) if len(sys.argv) > 4 else 120
import os
from torchquantum_b200 import plan as P, blockplan as B
P.BLOCK_MIN_RUN = int(os.environ.get("MIN_RUN", 4)); P.BLOCK_MIN_AVG = float(os.environ.get("MIN_AVG", 1.0))
B.MODE = os.environ.get("MODE", "auto")
bad = 0
for seed in seeds:
    rng = np.random.default_rng(seed)
    ops = rand_ops(n, count, rng)
    engine._PLAN_CACHE.clear()
    t = time.time()
    try:
        with torch.no_grad():
            q = tq.QuantumDevice(n_wires=n, bsz=1, device="cpu")
            cases.apply_ops(q, ops, "cpu")
            got = q.get_states_1d()
        kinds = [s.kind for s in next(reversed(engine._PLAN_CACHE.values())).stages]
        want = O.run_ops(n, 1, ops)
        rel = (torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want)).item()
        flag = "" if rel < 1e-5 else "  <<<<<< BAD"
        bad += rel >= 1e-5
        print(f"seed {seed}: stages {kinds} rel {rel:.2e} {time.time()-t:.1f}s{flag}", flush=True)
    except Exception as e:
        bad += 1
        import traceback; traceback.print_exc()
        print(f"seed {seed}: EXCEPTION {type(e).__name__}: {e}", flush=True)
print("bad:", bad)

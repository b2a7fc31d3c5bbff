"""Fuzz grad-mode programs (random catalogue gates, shared and per-sample parameters, optional forced small tiles:
python tools/fuzz_grad.py SEED0 SEED1 [T_SINGLE,T_MULTI,L_MIN]) against the oracle's autograd on the host emulation."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from tests.test_cpu_emulated import build_emu
from torchquantum_b200 import _native, engine, plan as P
_native.use_library_for_testing(build_emu())
from tests import cases
bad = 0
tiles = None
if len(sys.argv) > 3:
    tiles = tuple(int(x) for x in sys.argv[3].split(','))
    P.T_SINGLE, P.T_MULTI, P.L_MIN = tiles
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2, 10)); bsz = int(rng.integers(1, 5))
    ops = cases.random_catalogue_ops(n, int(rng.integers(5, 60)), seed, bsz=bsz, p_per_sample=0.4, max_wires=n)
    try:
        engine._PLAN_CACHE.clear()
        cases.check_against_oracle("cpu", n, bsz, ops, grad=True)
    except Exception as e:
        bad += 1
        print(f"seed {seed} n {n} bsz {bsz}: {type(e).__name__}: {str(e)[:300]}", flush=True)
print("bad", bad)

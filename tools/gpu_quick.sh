python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/cfg2_kernels.py 2>&1 | grep "cfg2"
python tools/pass_sweep.py --n 28 --mix rot --ks 1 24 64 2>&1 | tail -3
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
import torchquantum_b200 as tq, torchquantum_b200.functional as tqf
from dist_sharded import qft_ops, apply
dev = torch.device('cuda', 0)
n = 28
q = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)
ops = qft_ops(n)
def run():
    with torch.no_grad():
        q.reset_states(1); apply(q, ops, dev); return q.get_states_1d()
s = run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(3): s = run()
e1.record(); torch.cuda.synchronize()
amp = s[0, :4].abs().cpu().numpy() * 2 ** (n / 2)
print(f"QFT n={n}: {len(ops)} gates, {e0.elapsed_time(e1)/3:.2f} ms (incl. host queueing), |amp|*2^(n/2) = {amp}")
PY

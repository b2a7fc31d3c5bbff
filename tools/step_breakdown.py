"""Host/GPU time breakdown of one bench step (config 2) — where the wall clock goes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchquantum_b200 as tq
import bench

dev = torch.device("cuda", 0)
torch.manual_seed(0); np.random.seed(0)
model = bench.VQEModel(tq).to(dev)
params = [p for p in model.parameters() if p.requires_grad]
x = (torch.rand(bench.BSZ, bench.N_WIRES) * np.pi).to(dev)

def sync():
    torch.cuda.synchronize(); return time.perf_counter()

def step(breakdown=None):
    t = [sync()]
    for p in params: p.grad = None
    qdev = tq.QuantumDevice(n_wires=model.n, bsz=x.shape[0], device=x.device)
    model.encoder(qdev, x); model.ansatz(qdev)
    t.append(time.perf_counter())          # queue built (host only)
    qdev.get_states_1d()
    t.append(time.perf_counter())          # flush launched (host)
    t.append(sync())                       # forward kernels done
    energy = 0
    for term in model.terms:
        energy = energy + tq.expval_joint_analytical(qdev, term)
    loss = energy.mean()
    t.append(time.perf_counter())
    t.append(sync())
    loss.backward()
    t.append(time.perf_counter())
    t.append(sync())
    if breakdown is not None:
        breakdown.append([b - a for a, b in zip(t[:-1], t[1:])])

for _ in range(3): step()
bd = []
for _ in range(10): step(bd)
names = ["queue gates (host)", "flush launch (host)", "fwd kernels (wait)", "expval x19 launch (host)", "expval (wait)",
         "backward launch (host)", "backward (wait)"]
m = np.array(bd).mean(0) * 1e3
for n_, v in zip(names, m): print(f"{n_:28s} {v:8.3f} ms")
print(f"{'total':28s} {m.sum():8.3f} ms")

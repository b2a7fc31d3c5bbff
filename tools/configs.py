"""The five BASELINE configs as runnable workloads on the public API, with the size-independent
checks used where the reference / oracle cannot reach the full size.

  python tools/configs.py [--cfg 1 2 3 4] [--out gpurun_out/configs.json]

cfg 5 (34 qubits over 8 GPUs) is tools/dist_sharded.py."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchquantum_b200 as tq  # noqa: E402
import torchquantum_b200.functional as tqf  # noqa: E402
from torchquantum_b200 import engine  # noqa: E402

# networkx.random_regular_graph(3, 20, seed=0) (networkx 3.6.1), committed literally
QAOA_EDGES_20 = [(0, 1), (0, 9), (0, 15), (1, 4), (1, 18), (2, 4), (2, 6), (2, 17), (3, 6), (3, 7), (3, 11), (4, 14),
                 (5, 7), (5, 10), (5, 11), (6, 11), (7, 10), (8, 16), (8, 18), (8, 19), (9, 14), (9, 16), (10, 19),
                 (12, 13), (12, 16), (12, 17), (13, 17), (13, 19), (14, 15), (15, 18)]


def ev_ms(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


# ---------------------------------------------------------------------------------------
def qaoa_cost(betas, gammas, n, edges, dev):
    """examples/qaoa/max_cut_backprop.py:59-109 with per-sample betas / gammas [bsz, p]."""
    bsz, p = betas.shape
    q = tq.QuantumDevice(n_wires=n, bsz=bsz, device=dev)
    for w in range(n):
        tqf.hadamard(q, wires=w)
    for layer in range(p):
        b, g = betas[:, layer:layer + 1], gammas[:, layer:layer + 1]
        for w in range(n):
            tqf.rx(q, wires=w, params=b)
        for (i, j) in edges:
            tqf.cnot(q, wires=[i, j])
            tqf.rz(q, wires=j, params=g)
            tqf.cnot(q, wires=[i, j])
    terms = ["".join("Z" if w in e else "I" for w in range(n)) for e in edges]
    ev = tq.expval_joint_analytical_multi(q, terms)
    return 0.5 * ev.sum(1), q


def run_cfg3(dev, n=20, p=8, bsz=64, edges=QAOA_EDGES_20, timing=True):
    """fwd + adjoint bwd at full size; check: directional derivative of the batch cost along a random direction
    vs a central finite difference of two forward runs (a property that needs no reference at n=20)."""
    g = torch.Generator().manual_seed(0)
    betas = (0.4 * torch.rand(bsz, p, generator=g)).to(dev).requires_grad_(True)
    gammas = (0.4 * torch.rand(bsz, p, generator=g)).to(dev).requires_grad_(True)
    cost, q = qaoa_cost(betas, gammas, n, edges, dev)
    norm = float(q.get_states_1d().detach().abs().pow(2).sum(1).mean())
    gb, gg = torch.autograd.grad(cost.sum(), [betas, gammas])
    db, dg = torch.randn(bsz, p, generator=g).to(dev), torch.randn(bsz, p, generator=g).to(dev)
    analytic = (gb * db).sum(1) + (gg * dg).sum(1)
    h = 2e-2
    with torch.no_grad():
        cp, _ = qaoa_cost(betas + h * db, gammas + h * dg, n, edges, dev)
        cm, _ = qaoa_cost(betas - h * db, gammas - h * dg, n, edges, dev)
        cp2, _ = qaoa_cost(betas + 2 * h * db, gammas + 2 * h * dg, n, edges, dev)
        cm2, _ = qaoa_cost(betas - 2 * h * db, gammas - 2 * h * dg, n, edges, dev)
    fd = (8 * (cp - cm) - (cp2 - cm2)) / (12 * h)   # 4th-order central difference
    rel = float((analytic - fd).norm() / fd.norm())
    res = {"workload": f"cfg3: {n}-qubit QAOA MaxCut p={p}, bsz={bsz}, {len(edges)} edges, {n + p * (n + 3 * len(edges))} gates",
           "state_norm": norm, "grad_vs_finite_difference_rel": rel, "ok": bool(abs(norm - 1) < 1e-4 and rel < 2e-3)}
    if timing:
        def step():
            betas.grad = gammas.grad = None
            c, _ = qaoa_cost(betas, gammas, n, edges, dev)
            c.sum().backward()

        def fwd():
            with torch.no_grad():
                qaoa_cost(betas, gammas, n, edges, dev)

        l0 = engine.STATS["launches"]
        step()
        res["launches_per_step"] = engine.STATS["launches"] - l0
        res["fwd_ms"], res["fwd_bwd_ms"] = ev_ms(fwd), ev_ms(step)
        res["circuits_per_s_fwd_bwd"] = bsz / (res["fwd_bwd_ms"] * 1e-3)
        ngates = n + p * (n + 3 * len(edges))
        res["gate_amp_updates_per_s_fwd"] = ngates * bsz * 2.0 ** n / (res["fwd_ms"] * 1e-3)
    return res


# ---------------------------------------------------------------------------------------
def brickwork(n, layers, seed=0):
    rng = np.random.default_rng(seed)
    ops = []
    for layer in range(layers):
        for w in range(n):
            ops.append((["rx", "ry", "rz"][rng.integers(3)], [w], float(rng.uniform(-np.pi, np.pi))))
        for w in range(layer % 2, n - 1, 2):
            ops.append(("cnot", [w, w + 1], None))
    return ops


def apply_ops(q, ops, dev, inverse=False):
    seq = reversed(ops) if inverse else ops
    for name, wires, p in seq:
        if p is None:
            getattr(tqf, name)(q, wires=wires, inverse=inverse)
        else:
            getattr(tqf, name)(q, wires=wires, params=torch.tensor([[p]], dtype=torch.float32, device=dev), inverse=inverse)


def run_cfg4(dev, n=28, layers=100, timing=True):
    """Forward only.  Checks: norm, all <Z_w> in [-1, 1], and circuit followed by its inverse returns |0...0>."""
    ops = brickwork(n, layers)
    q = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)
    with torch.no_grad():
        apply_ops(q, ops, dev)
        s = q.get_states_1d()
        norm = float(s.abs().pow(2).sum())
        z = tq.expval_joint_analytical_multi(q, ["I" * w + "Z" + "I" * (n - 1 - w) for w in range(0, n, 5)])
        apply_ops(q, ops, dev, inverse=True)
        s0 = q.get_states_1d()
        back = float(s0[0, 0].abs())
        leak = float(s0[0, 1:].abs().max())
    res = {"workload": f"cfg4: {n}-qubit brickwork rx/ry/rz + cnot, depth {layers}, {len(ops)} gates, single statevector",
           "state_norm": norm, "z_max_abs": float(z.abs().max()), "amp0_after_inverse": back, "max_leak_after_inverse": leak,
           "ok": bool(abs(norm - 1) < 2e-3 and abs(back - 1) < 2e-3 and leak < 1e-3 and float(z.abs().max()) <= 1.001)}
    if timing:
        def run():
            with torch.no_grad():
                q.reset_states(1)
                apply_ops(q, ops, dev)
                q.get_states_1d()

        l0 = engine.STATS["launches"]
        run()
        res["passes"] = engine.STATS["launches"] - l0 - 1
        res["fwd_ms"] = ev_ms(run, warm=1, iters=3)
        res["gate_amp_updates_per_s"] = len(ops) * 2.0 ** n / (res["fwd_ms"] * 1e-3)
        res["algorithmic_GBps"] = res["passes"] * 16 * 2.0 ** n / (res["fwd_ms"] * 1e-3) / 1e9
        res["gates_per_pass"] = len(ops) / max(res["passes"], 1)
    return res


# ---------------------------------------------------------------------------------------
def run_cfg1(dev, bsz=256):
    """4-wire MNIST QNN step (examples/mnist/mnist.py:44-134 minus the dataset); parity is pinned by the golden
    fixture (tests/golden/mnist_cfg1.npz), this only times the step."""
    import random
    import torch.nn.functional as F

    random.seed(0); np.random.seed(0); torch.manual_seed(0)

    class QLayer(tq.QuantumModule):
        def __init__(self):
            super().__init__()
            self.random_layer = tq.RandomLayer(n_ops=50, wires=list(range(4)))
            self.rx0, self.ry0 = tq.RX(has_params=True, trainable=True), tq.RY(has_params=True, trainable=True)
            self.rz0, self.crx0 = tq.RZ(has_params=True, trainable=True), tq.CRX(has_params=True, trainable=True)

        def forward(self, qdev):
            self.random_layer(qdev)
            self.rx0(qdev, wires=0); self.ry0(qdev, wires=1); self.rz0(qdev, wires=3); self.crx0(qdev, wires=[0, 2])
            qdev.h(wires=3); qdev.sx(wires=2); qdev.cnot(wires=[3, 0])

    enc = tq.GeneralEncoder(tq.encoder_op_list_name_dict["4x4_u3_h_rx"])
    layer = QLayer().to(dev)
    measure = tq.MeasureAll(tq.PauliZ)
    x = torch.rand(bsz, 16, device=dev)
    target = torch.randint(0, 2, (bsz,), device=dev)
    params = list(layer.parameters())

    def loss_fn(xx):
        q = tq.QuantumDevice(n_wires=4, bsz=xx.shape[0], device=dev)
        enc(q, xx)
        layer(q)
        out = measure(q)
        return F.nll_loss(F.log_softmax(out.reshape(xx.shape[0], 2, 2).sum(-1), dim=1), target)

    def step():
        for p in params:
            p.grad = None
        loss_fn(x).backward()

    eager = ev_ms(step, warm=3, iters=20)
    g = tq.GraphedStep(loss_fn, [x], params)
    graphed = ev_ms(lambda: g(x), warm=3, iters=20)
    return {"workload": f"cfg1: 4-wire MNIST QNN, bsz={bsz}, fwd+bwd", "eager_ms": eager, "graphed_ms": graphed,
            "circuits_per_s_graphed": bsz / (graphed * 1e-3), "ok": True}


def run_cfg3_batch_sharded(out_path):
    """cfg 3 batch-sharded over the ranks of a torchrun launch (no data-path collective, per-sample angles and
    gradients stay on their rank): 64 circuits in total, 64 / world per GPU; time = max over ranks."""
    import torch.distributed as dist

    local, world = int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = run_cfg3(dev, bsz=64 // world)
    t = torch.tensor([res["fwd_ms"], res["fwd_bwd_ms"], 0.0 if res["ok"] else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if dist.get_rank() == 0:
        out = {"workload": res["workload"] + f" per rank, batch-sharded over {world} B200 (64 circuits in total)",
               "world": world, "fwd_ms": t[0].item(), "fwd_bwd_ms": t[1].item(),
               "circuits_per_s_fwd_bwd": 64 / (t[1].item() * 1e-3), "scaling": "strong", "ok": t[2].item() == 0.0,
               "grad_vs_finite_difference_rel": res["grad_vs_finite_difference_rel"]}
        print(json.dumps({"cfg3_batch_sharded": out}))
        if out_path:
            json.dump(out, open(out_path, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, nargs="*", default=[1, 3, 4])
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_cfg3_batch_sharded(a.out)
    dev = torch.device("cuda", 0)
    out = {"gpu": torch.cuda.get_device_name(0)}
    for c in a.cfg:
        t0 = time.perf_counter()
        out[f"cfg{c}"] = {1: run_cfg1, 3: run_cfg3, 4: run_cfg4}[c](dev)
        out[f"cfg{c}"]["wall_s"] = round(time.perf_counter() - t0, 2)
        print(json.dumps({f"cfg{c}": out[f"cfg{c}"]}))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)
    sys.exit(0 if all(v.get("ok", True) for v in out.values() if isinstance(v, dict)) else 1)


if __name__ == "__main__":
    main()

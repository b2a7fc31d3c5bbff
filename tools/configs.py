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


def qaoa_ops(betas, gammas, n, edges):
    """the same circuit as an op list for the oracle (tests/cases.apply_ops format)"""
    p = betas.shape[1]
    ops = [{"name": "hadamard", "wires": [w]} for w in range(n)]
    for layer in range(p):
        for w in range(n):
            ops.append({"name": "rx", "wires": [w], "params": betas[:, layer:layer + 1]})
        for (i, j) in edges:
            ops.append({"name": "cnot", "wires": [i, j]})
            ops.append({"name": "rz", "wires": [j], "params": gammas[:, layer:layer + 1]})
            ops.append({"name": "cnot", "wires": [i, j]})
    return ops


def cfg3_oracle_parity(dev, betas, gammas, states, cost, gb, gg, n, edges, samples):
    """Full-size parity of chosen samples against the float64 twin of the oracle run with torch's CUDA ops
    (oracle/tq_oracle64.py; test infrastructure): states, cost and adjoint gradients, rel-L2 each."""
    from oracle.tq_oracle64 import run_ops64

    idx = torch.tensor(samples, device=dev)
    b64 = betas.detach()[idx].double().cpu().requires_grad_(True)
    g64 = gammas.detach()[idx].double().cpu().requires_grad_(True)
    st = run_ops64(n, len(samples), qaoa_ops(b64, g64, n, edges), device=dev, keep_graph=True)
    probs = st.real ** 2 + st.imag ** 2
    k = torch.arange(2 ** n, device=dev)
    c64 = 0
    for (i, j) in edges:   # <Z_i Z_j> = sum_k p_k (-1)^(k_i + k_j), wire w = index bit n-1-w (device/devices.py:57-71)
        sign = 1.0 - 2.0 * (((k >> (n - 1 - i)) ^ (k >> (n - 1 - j))) & 1).double()
        c64 = c64 + 0.5 * (probs * sign).sum(1)
    gb64, gg64 = torch.autograd.grad(c64.sum(), [b64, g64])

    def rel(a, b):
        a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
        return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))

    got_states = states.detach()[idx].to(torch.complex128)
    return {"samples": list(samples),
            "states_rel_l2": float(torch.linalg.vector_norm(got_states - st.detach()) / torch.linalg.vector_norm(st.detach())),
            "cost_rel": rel(cost[idx], c64), "grad_beta_rel_l2": rel(gb[idx], gb64), "grad_gamma_rel_l2": rel(gg[idx], gg64)}


def run_cfg3(dev, n=20, p=8, bsz=64, edges=QAOA_EDGES_20, timing=True, parity_samples=(0, -1)):
    """fwd + adjoint bwd at full size.  Checks at the FULL size (n=20, p=8): unit norm (float64 accumulation) and, for
    `parity_samples` of the batch, states / cost / gradients against the float64 oracle twin (bars 1e-5)."""
    g = torch.Generator().manual_seed(0)
    betas = (0.4 * torch.rand(bsz, p, generator=g)).to(dev).requires_grad_(True)
    gammas = (0.4 * torch.rand(bsz, p, generator=g)).to(dev).requires_grad_(True)
    cost, q = qaoa_cost(betas, gammas, n, edges, dev)
    states = q.get_states_1d()
    norms = torch.linalg.vector_norm(torch.view_as_real(states.detach()).reshape(bsz, -1), dim=1, dtype=torch.float64)
    norm_err = float((norms - 1).abs().max())
    gb, gg = torch.autograd.grad(cost.sum(), [betas, gammas])
    samples = sorted({s % bsz for s in parity_samples})
    par = cfg3_oracle_parity(dev, betas, gammas, states, cost, gb, gg, n, edges, samples)
    worst = max(par["states_rel_l2"], par["cost_rel"], par["grad_beta_rel_l2"], par["grad_gamma_rel_l2"])
    res = {"workload": f"cfg3: {n}-qubit QAOA MaxCut p={p}, bsz={bsz}, {len(edges)} edges, {n + p * (n + 3 * len(edges))} gates",
           "max_norm_err": norm_err, "oracle64_parity": par, "ok": bool(norm_err < 1e-5 and worst < 1e-5)}
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


def apply_ops(q, ops, dev, inverse=False, thetas=None):
    """thetas: the circuit's angles as [1, 1] views of ONE device tensor made by `device_angles` (no per-gate upload);
    without it every gate uploads its own angle, as a user script written against the reference would."""
    seq = list(enumerate(ops))
    for i, (name, wires, p) in (reversed(seq) if inverse else seq):
        if p is None:
            getattr(tqf, name)(q, wires=wires, inverse=inverse)
        elif thetas is not None:
            getattr(tqf, name)(q, wires=wires, params=thetas[i], inverse=inverse)
        else:
            getattr(tqf, name)(q, wires=wires, params=torch.tensor([[p]], dtype=torch.float32, device=dev), inverse=inverse)


def device_angles(ops, dev):
    """All angles of the circuit in one device tensor; [1, 1] views of it that take part in autograd (the circuit's
    trainable parameters) are handed to the gate calls by reference."""
    flat = torch.tensor([0.0 if p is None else p for _, _, p in ops], dtype=torch.float32, device=dev).requires_grad_(True)
    return [None if p is None else flat[i:i + 1].view(1, 1) for i, (_, _, p) in enumerate(ops)]


def norm64(s):
    return float(torch.linalg.vector_norm(torch.view_as_real(s).reshape(-1), dtype=torch.float64))


def run_cfg4(dev, n=28, layers=100, timing=True):
    """Forward only.  Checks (bars 1e-5, norms accumulated in float64): norm, all <Z_w> in [-1, 1], and circuit followed
    by its inverse returns |0...0>.  Full-depth parity against the fp64 oracle: tools/depth_parity.py."""
    ops = brickwork(n, layers)
    thetas = device_angles(ops, dev)
    q = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)
    with torch.no_grad():
        apply_ops(q, ops, dev, thetas=thetas)
        s = q.get_states_1d()
        norm = norm64(s)
        z = tq.expval_joint_analytical_multi(q, ["I" * w + "Z" + "I" * (n - 1 - w) for w in range(0, n, 5)])
        apply_ops(q, ops, dev, inverse=True, thetas=thetas)
        s0 = q.get_states_1d()
        back = float(s0[0, 0].abs())
        s0[0, 0] = 0
        leak = norm64(s0)
    res = {"workload": f"cfg4: {n}-qubit brickwork rx/ry/rz + cnot, depth {layers}, {len(ops)} gates, single statevector",
           "state_norm": norm, "z_max_abs": float(z.abs().max()), "amp0_after_inverse": back,
           "leak_norm_after_inverse": leak,
           "ok": bool(abs(norm - 1) < 1e-5 and abs(back - 1) < 2e-5 and leak < 2e-5 and float(z.abs().max()) <= 1.00001)}
    if timing:
        def queue():
            q.reset_states(1)
            apply_ops(q, ops, dev, thetas=thetas)

        def run():
            with torch.no_grad():
                queue()
                q.get_states_1d()

        l0 = engine.STATS["launches"]
        run()
        res["passes"] = engine.STATS["launches"] - l0 - 1
        res["fwd_ms_incl_python_queueing"] = ev_ms(run, warm=1, iters=3)
        # the gates are queued by Python first (host only); the timed region is the flush = every kernel of the circuit
        times = []
        for _ in range(4):
            with torch.no_grad():
                queue()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                q.get_states_1d()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        res["fwd_ms"] = float(np.median(times[1:]))
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


# ---------------------------------------------------------------------------------------
# cfg 5: QFT + brickwork layers on a statevector sharded over the ranks of an initialised process group
def qft_gate_list(n):
    ops = []
    for top in range(n):
        ops.append(("hadamard", [top], None))
        for w in range(top + 1, n):
            ops.append(("cu1", [w, top], float(np.pi / 2 ** (w - top))))
    ops += [("swap", [w, n - 1 - w], None) for w in range(n // 2)]
    return ops


def run_cfg5(dev, n, layers=4, exchange="auto", reps=2):
    """QFT on a random basis state |x>, then `layers` brickwork layers, on a ShardedQuantumDevice over the default group.
    Checks: the QFT amplitudes at sampled indices against the analytic value exp(2 pi i x k / 2^n) / 2^(n/2)
    and the norm after the brickwork layers accumulated in float64 (bars 1e-5).  Exchange time is reported against the
    NVLink peak (900 GB/s per direction per GPU)."""
    import torch.distributed as dist

    world = dist.get_world_size()
    rng = np.random.default_rng(7)
    x = int(rng.integers(1 << 62)) & ((1 << n) - 1)
    prep = [("paulix", [w], None) for w in range(n) if (x >> (n - 1 - w)) & 1]
    circuit = prep + qft_gate_list(n)
    bricks = brickwork(n, layers, seed=2)
    th_c, th_b = device_angles(circuit, dev), device_angles(bricks, dev)
    q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device=dev, exchange=exchange)
    q.time_exchanges = True
    with torch.no_grad():
        for rep in range(reps):   # rep 0 warms up (plan build, NCCL channels / symmetric-memory rendezvous)
            q.reset_states(1)
            q.stats.update(exchanges=0, exchange_bytes_sent=0, exchange_ms=[], local_permute_passes=0, local_runs=[],
                           exchange_victim_bits=[])
            dist.barrier()
            torch.cuda.synchronize()
            l0 = engine.STATS["launches"]
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            apply_ops(q, circuit, dev, thetas=th_c)
            q.flush()
            e1.record()
            ks = [0, 1, (1 << n) - 1] + [int(v) & ((1 << n) - 1) for v in rng.integers(1 << 62, size=13)]
            amps = q.read_amplitudes(ks)[0].cpu().numpy() if rep == reps - 1 else None
            apply_ops(q, bricks, dev, thetas=th_b)
            q.flush()
            e2.record()
            torch.cuda.synchronize()
        norm2 = torch.linalg.vector_norm(torch.view_as_real(q._flat).reshape(-1), dtype=torch.float64) ** 2
        dist.all_reduce(norm2)
    qft_ms, brick_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    ex_ms = q.exchange_times_ms()
    t = torch.tensor([qft_ms, brick_ms, sum(ex_ms)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    want = np.array([np.exp(2j * np.pi * ((x * k) % (1 << n)) / (1 << n)) / 2 ** (n / 2) for k in ks])
    amp_err = float(np.abs(amps - want).max() * 2 ** (n / 2))   # relative to the amplitude magnitude
    norm = float(norm2.sqrt())
    sent = q.stats["exchange_bytes_sent"]
    n_gates = len(circuit) + len(bricks)
    total_ms = t[0].item() + t[1].item()
    return {
        "workload": f"cfg5: {n}-qubit QFT on |x> + {layers} brickwork layers ({n_gates} gates), statevector sharded over "
                    f"{world} B200",
        "n": n, "world": world, "gates": n_gates, "exchange_backend": q.exchange_backend,
        "total_ms": round(total_ms, 3), "qft_ms": round(t[0].item(), 3), "brickwork_ms": round(t[1].item(), 3),
        "exchanges": q.stats["exchanges"], "local_permute_passes": q.stats["local_permute_passes"],
        "exchange_ms_total": round(t[2].item(), 3), "exchange_ms_each": [round(v, 3) for v in ex_ms],
        "exchange_bytes_sent_per_rank": sent, "exchange_victim_bits": q.stats.get("exchange_victim_bits"),
        "exchange_GBps_per_rank_per_direction": round(sent / max(t[2].item(), 1e-9) / 1e6, 1),
        "nvlink_peak_GBps_per_direction": 900.0,
        "exchange_frac_of_nvlink": round(sent / max(t[2].item(), 1e-9) / 1e6 / 900.0, 4),
        "kernel_launches_per_rank": engine.STATS["launches"] - l0,
        "local_runs_gates_launches_ms": q.local_run_times(),
        "gate_amp_updates_per_s": round(n_gates * 2.0 ** n / (total_ms * 1e-3), 1),
        "qft_amp_rel_err_max": amp_err, "norm_after": norm,
        "ok": bool(amp_err < 1e-5 and abs(norm - 1) < 1e-5),
    }


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
               "oracle64_parity": res["oracle64_parity"]}
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

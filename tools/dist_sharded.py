"""Sharded-statevector runs under torchrun (one rank per GPU, NCCL).

  torchrun --nproc-per-node N tools/dist_sharded.py --check 22     parity of the cfg-5 family vs one GPU
  torchrun --nproc-per-node N tools/dist_sharded.py --bench 34     cfg 5: QFT on |x> + 4 brickwork layers

--bench verifies QFT|x> analytically (amplitude at flat index k = exp(2 pi i x k / 2^n) / 2^(n/2),
SURVEY §8c) at sampled indices before the brickwork layers, and the norm after them, and reports the
exchange (all_to_all) time against the NVLink peak separately from the gate passes."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchquantum_b200 as tq  # noqa: E402
import torchquantum_b200.functional as tqf  # noqa: E402
from torchquantum_b200 import engine  # noqa: E402


def qft_ops(n):
    ops = []
    for top in range(n):
        ops.append(("hadamard", [top], None))
        for w in range(top + 1, n):
            ops.append(("cu1", [w, top], np.pi / 2 ** (w - top)))
    ops += [("swap", [w, n - 1 - w], None) for w in range(n // 2)]
    return ops


def brick_ops(n, layers, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for layer in range(layers):
        for w in range(n):
            ops.append((["rx", "ry", "rz"][rng.integers(3)], [w], float(rng.uniform(-np.pi, np.pi))))
        for w in range(layer % 2, n - 1, 2):
            ops.append(("cnot", [w, w + 1], None))
    return ops


def apply(q, ops, dev):
    for name, wires, p in ops:
        if p is None:
            getattr(tqf, name)(q, wires=wires)
        else:
            getattr(tqf, name)(q, wires=wires, params=torch.tensor([[p]], dtype=torch.float32, device=dev))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--check-bsz", type=int, default=1)
    ap.add_argument("--grad", type=int, default=0, help="n: adjoint gradients through the sharded device vs one GPU")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"], help="gloo: several ranks may share one GPU")
    ap.add_argument("--bench", type=int, default=0)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "collective", "both"])
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if a.backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    results = {}

    if a.check:
        n = a.check
        rng = np.random.default_rng(5)
        xbits = [w for w in range(n) if rng.integers(2)]
        ops = [("paulix", [w], None) for w in xbits] + qft_ops(n) + brick_ops(n, 3, seed=1)
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=a.check_bsz, device=dev,
                                    exchange="auto" if a.exchange == "both" else a.exchange)
        apply(q, ops, dev)
        obs = ["".join("ZXYI"[(i + s) % 4] for i in range(n)) for s in range(2)] + ["Z" * n]
        ev = tq.expval_joint_analytical_multi(q, obs)
        slab = q.get_states_1d()
        ref = tq.QuantumDevice(n_wires=n, bsz=a.check_bsz, device=dev)   # every rank: the same circuit on one GPU
        apply(ref, ops, dev)
        want = ref.get_states_1d()
        ev_want = tq.expval_joint_analytical_multi(ref, obs)
        nl = q.nl
        mine = want[:, rank << nl:(rank + 1) << nl]
        d2 = (slab - mine).abs().double().pow(2).sum().item()
        w2 = mine.abs().double().pow(2).sum().item()
        ev_err = (ev - ev_want).abs().max().item()
        cpu = a.backend == "gloo"
        t = torch.tensor([d2, w2], dtype=torch.float64, device="cpu" if cpu else dev)
        dist.all_reduce(t)
        m = torch.tensor([ev_err], dtype=torch.float64, device="cpu" if cpu else dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        rel = float((t[0] / t[1]).sqrt())
        results["check"] = {"n": n, "world": world, "exchange_backend": q.exchange_backend, "bsz": a.check_bsz,
                            "state_rel_l2_vs_one_gpu": rel, "max_abs_expval_err": m[0].item(),
                            "exchanges": q.stats["exchanges"], "ok": bool(rel < 1e-5 and m[0].item() < 1e-5)}
        assert results["check"]["ok"], results["check"]

    if a.grad:
        n = a.grad
        torch.manual_seed(4)
        thetas = [torch.randn(1, 1, device=dev).requires_grad_(True) for _ in range(3 * n)]
        obs = ["Z" * n, "ZZ" + "I" * (n - 2), "X" + "I" * (n - 2) + "Y", "I" * (n - 1) + "Z"]
        wts = torch.tensor([0.7, -1.3, 0.4, 2.0], device=dev)

        def circuit(q):
            k = 0
            for w in range(n):
                tqf.hadamard(q, wires=w)
            for layer in range(3):
                for w in range(n):   # rotations on EVERY wire, incl. the rank-index ones
                    (tqf.rx, tqf.ry, tqf.rz)[(layer + w) % 3](q, wires=w, params=thetas[k])
                    k += 1
                for w in range(layer % 2, n - 1, 2):
                    tqf.cnot(q, wires=[w, w + 1])
                tqf.crx(q, wires=[n - 1, 0], params=thetas[(5 * layer) % len(thetas)])
            return (tq.expval_joint_analytical_multi(q, obs) * wts).sum()

        ref_loss = circuit(tq.QuantumDevice(n_wires=n, bsz=2, device=dev))
        want = torch.cat([g.reshape(-1) for g in torch.autograd.grad(ref_loss, thetas)])
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=2, device=dev, exchange="auto" if a.exchange == "both" else a.exchange)
        loss = circuit(q)
        loss.backward()
        q.allreduce_gradients(thetas)
        got = torch.cat([t.grad.reshape(-1) for t in thetas])
        rel = float((got - want).norm() / want.norm())
        results["grad"] = {"n": n, "world": world, "exchanges": q.stats["exchanges"], "loss_abs_err": abs(float(loss - ref_loss)),
                           "grad_rel_l2_vs_one_gpu": rel, "ok": bool(rel < 1e-5 and abs(float(loss - ref_loss)) < 1e-5)}
        assert results["grad"]["ok"], results["grad"]

    if a.bench:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import configs

        for backend in ([a.exchange] if a.exchange != "both" else ["p2p", "collective"]):
            results["bench_" + backend] = configs.run_cfg5(dev, a.bench, layers=a.layers, exchange=backend)
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(results))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(results, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not all(v.get("ok", True) for v in results.values()):
        sys.exit(1)


if __name__ == "__main__":
    main()

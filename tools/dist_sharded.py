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
    ap.add_argument("--bench", type=int, default=0)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    results = {}

    if a.check:
        n = a.check
        rng = np.random.default_rng(5)
        xbits = [w for w in range(n) if rng.integers(2)]
        ops = [("paulix", [w], None) for w in xbits] + qft_ops(n) + brick_ops(n, 3, seed=1)
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device=dev)
        apply(q, ops, dev)
        obs = ["".join("ZXYI"[(i + s) % 4] for i in range(n)) for s in range(2)] + ["Z" * n]
        ev = tq.expval_joint_analytical_multi(q, obs)
        slab = q.get_states_1d()
        ref = tq.QuantumDevice(n_wires=n, bsz=1, device=dev)   # every rank: the same circuit on one GPU
        apply(ref, ops, dev)
        want = ref.get_states_1d()
        ev_want = tq.expval_joint_analytical_multi(ref, obs)
        nl = q.nl
        err = (slab - want[:, rank << nl:(rank + 1) << nl]).abs().max().item()
        ev_err = (ev - ev_want).abs().max().item()
        t = torch.tensor([err, ev_err], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results["check"] = {"n": n, "world": world, "max_abs_state_err": t[0].item(), "max_abs_expval_err": t[1].item(),
                            "exchanges": q.stats["exchanges"], "ok": bool(t[0].item() < 2e-6 and t[1].item() < 2e-5)}
        assert results["check"]["ok"], results["check"]

    if a.bench:
        n = a.bench
        rng = np.random.default_rng(7)
        x = int(rng.integers(1 << 62)) & ((1 << n) - 1)
        prep = [("paulix", [w], None) for w in range(n) if (x >> (n - 1 - w)) & 1]
        circuit = prep + qft_ops(n)
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device=dev)
        q.time_exchanges = True
        for rep in range(2):   # rep 0 warms up (plan build, NCCL channels)
            q.reset_states(1)
            q.stats.update(exchanges=0, exchange_bytes_sent=0, exchange_ms=[], local_permute_passes=0, local_runs=[])
            dist.barrier(); torch.cuda.synchronize()
            l0 = engine.STATS["launches"]
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            apply(q, circuit, dev)
            q.flush()
            e1.record()
            ks = [0, 1, (1 << n) - 1] + [int(v) & ((1 << n) - 1) for v in rng.integers(1 << 62, size=13)]
            amps = q.read_amplitudes(ks)[0].cpu().numpy() if rep == 1 else None
            apply(q, brick_ops(n, a.layers, seed=2), dev)
            q.flush()
            e2.record()
            torch.cuda.synchronize()
            z = q.expval_z_all()
        qft_ms, brick_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
        ex_ms = q.exchange_times_ms()
        t = torch.tensor([qft_ms, brick_ms, sum(ex_ms)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        want = np.array([np.exp(2j * np.pi * ((x * k) % (1 << n)) / (1 << n)) / 2 ** (n / 2) for k in ks])
        amp_err = float(np.abs(amps - want).max() * 2 ** (n / 2))   # relative to the amplitude magnitude
        norm = z[0, n].item()
        sent = q.stats["exchange_bytes_sent"]
        n_gates = len(circuit) + len(brick_ops(n, a.layers, seed=2))
        results["bench"] = {
            "workload": f"cfg5: {n}-qubit QFT on |x> + {a.layers} brickwork layers, statevector sharded over {world} B200",
            "n": n, "world": world, "gates": n_gates, "qft_ms": round(t[0].item(), 3), "brickwork_ms": round(t[1].item(), 3),
            "exchanges": q.stats["exchanges"], "local_permute_passes": q.stats["local_permute_passes"],
            "exchange_ms_total": round(t[2].item(), 3), "exchange_ms_each": [round(v, 3) for v in ex_ms],
            "exchange_bytes_sent_per_rank": sent,
            "exchange_GBps_per_rank_per_direction": round(sent / max(t[2].item(), 1e-9) / 1e6, 1),
            "nvlink_peak_GBps_per_direction": 900.0,
            "exchange_frac_of_nvlink": round(sent / max(t[2].item(), 1e-9) / 1e6 / 900.0, 4),
            "kernel_launches": engine.STATS["launches"] - l0,
            "local_runs_gates_launches_ms": q.local_run_times(),
            "gate_amp_updates_per_s": round(n_gates * 2.0 ** n / ((t[0].item() + t[1].item()) * 1e-3), 1),
            "qft_amp_rel_err_max": amp_err, "norm_after": norm,
            "ok": bool(amp_err < 1e-4 and abs(norm - 1) < 1e-3),
        }
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

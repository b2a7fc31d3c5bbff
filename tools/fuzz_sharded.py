"""Fuzz ShardedQuantumDevice (gloo, host emulation) with random catalogue gates against the oracle:
python tools/fuzz_sharded.py WORLD N SEED0 SEED1 GATES [BSZ] [grad]."""
import os, sys, socket
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist, torch.multiprocessing as mp

def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); return s.getsockname()[1]

def worker(rank, world, port, n, seeds, count, emu, bsz, grad):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torchquantum_b200 as tq
    from torchquantum_b200 import _native, gates as G
    from torchquantum_b200.measurement import expval_joint_analytical_multi
    from tests import cases
    from oracle import tq_oracle as O
    _native.use_library_for_testing(emu)
    names = [k for k, s in G.SPECS.items() if s.kind in (G.U1Q, G.DIAG, G.U2Q) and k != "globalphase"]
    g = int(np.log2(world))
    bad = 0
    for seed in seeds:
        rng = np.random.default_rng(seed)
        ops = []
        for _ in range(count):
            name = names[rng.integers(len(names))]
            s = G.SPECS[name]
            k = s.n_wires if s.n_wires is not None else int(rng.integers(2, 4))
            if k > n - 2 * g: continue
            wires = [int(w) for w in rng.choice(n, k, replace=False)]
            op = {"name": name, "wires": wires, "inverse": bool(rng.random() < 0.3)}
            if s.n_params:
                shape = (bsz if rng.random() < 0.3 else 1, s.n_params)
                op["params"] = torch.tensor(rng.uniform(-np.pi, np.pi, shape), dtype=torch.float32)
            ops.append(op)
        try:
            if grad:
                leaves = [o["params"].requires_grad_(True) for o in ops if o.get("params") is not None]
                if not leaves:
                    continue
                obs = ["".join("ZXYI"[(i + s_) % 4] for i in range(n)) for s_ in range(2)]
                wts = torch.tensor([0.7, -1.3])
                q = tq.ShardedQuantumDevice(n_wires=n, bsz=bsz, device="cpu")
                cases.apply_ops(q, ops, "cpu")
                loss = (expval_joint_analytical_multi(q, obs) * wts).sum()
                for t in leaves:
                    t.grad = None
                if loss.requires_grad:
                    loss.backward()
                    q.allreduce_gradients(leaves)
                got = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).reshape(-1) for t in leaves])
                ref = tq.QuantumDevice(n_wires=n, bsz=bsz, device="cpu")
                cases.apply_ops(ref, ops, "cpu")
                ref_loss = (expval_joint_analytical_multi(ref, obs) * wts).sum()
                want = torch.autograd.grad(ref_loss, leaves, allow_unused=True)
                want = torch.cat([(g if g is not None else torch.zeros_like(t)).reshape(-1) for g, t in zip(want, leaves)])
                scale = max(want.abs().max().item(), 1e-3)
                err = (got - want).abs().max().item() / scale
                lerr = abs(loss.item() - ref_loss.item())
                ok = err < 3e-5 and lerr < 2e-5
                bad += not ok
                if rank == 0:
                    print(f"world {world} n {n} seed {seed}: grad err {err:.2e} loss err {lerr:.2e} exchanges {q.stats['exchanges']}" + ("" if ok else "  <<<<< BAD"), flush=True)
                continue
            with torch.no_grad():
                q = tq.ShardedQuantumDevice(n_wires=n, bsz=bsz, device="cpu")
                cases.apply_ops(q, ops, "cpu")
                obs = ["".join("ZXYI"[(i + s_) % 4] for i in range(n)) for s_ in range(2)]
                ev = expval_joint_analytical_multi(q, obs)
                full = q.gather_states_1d()
            want = O.run_ops(n, bsz, ops)
            ev_want = torch.stack([O.expval_joint_rotated(want, o) for o in obs], dim=-1)
            rel = (torch.linalg.vector_norm(full - want) / torch.linalg.vector_norm(want)).item()
            eerr = (ev - ev_want).abs().max().item()
            ok = rel < 1e-5 and eerr < 2e-5
            bad += not ok
            if rank == 0:
                print(f"world {world} n {n} seed {seed}: rel {rel:.2e} ev_err {eerr:.2e} exchanges {q.stats['exchanges']}" + ("" if ok else "  <<<<< BAD"), flush=True)
        except Exception as e:
            bad += 1
            import traceback; traceback.print_exc()
            print(f"rank {rank} seed {seed}: EXCEPTION {type(e).__name__}: {e}", flush=True)
            break
    dist.destroy_process_group()
    assert bad == 0, f"rank {rank}: {bad} mismatching circuit(s)"

if __name__ == "__main__":
    world, n, s0, s1, count = [int(a) for a in sys.argv[1:6]]
    bsz = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    grad = len(sys.argv) > 7 and sys.argv[7] == "grad"
    from tests.test_cpu_emulated import build_emu
    emu = build_emu()
    mp.spawn(worker, args=(world, free_port(), n, list(range(s0, s1)), count, emu, bsz, grad), nprocs=world, join=True)

"""Sharded statevector (amplitude axis) on CPU: world_size-2 and -4 `gloo` process groups
driving the host-emulation build of the kernel source; every rank checks its slab against the
single-process CPU oracle.  (The GPU suite repeats this over NCCL on real B200s.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ops_qft_brick(n, seed):
    """config-5 family: X-prepared basis state, QFT (H + CU1 ladder + swaps), brickwork layers."""
    from tests import cases

    rng = np.random.default_rng(seed)
    ops = [{"name": "paulix", "wires": [w]} for w in range(n) if rng.integers(2)]
    for top in range(n):
        ops.append({"name": "hadamard", "wires": [top]})
        for d, w in enumerate(range(top + 1, n)):
            ops.append({"name": "cu1", "wires": [w, top], "params": torch.tensor([[np.pi / 2 ** (d + 1)]], dtype=torch.float32)})
    for w in range(n // 2):
        ops.append({"name": "swap", "wires": [w, n - 1 - w]})
    return ops + cases.brickwork_ops(n, 3, seed=seed + 1)


def _worker(rank, world, port, n, seed, emu_path, tiles):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchquantum_b200 as tq
        import torchquantum_b200.functional as tqf
        from oracle import tq_oracle as O
        from tests import cases
        from tests.helpers import GRAD_RTOL, RTOL, assert_close
        from torchquantum_b200 import _native
        from torchquantum_b200.measurement import expval, expval_joint_analytical_multi

        _native.use_library_for_testing(emu_path)
        if tiles is not None:  # force several tiles per slab: the rank bits sit above the tile index
            from torchquantum_b200 import plan as P
            P.T_SINGLE, P.T_MULTI, P.L_MIN = tiles
        ops = _ops_qft_brick(n, seed)
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device="cpu")
        cases.apply_ops(q, ops, "cpu")
        obs = ["".join("ZXYI"[(i + s) % 4] for i in range(n)) for s in range(3)] + ["Z" * n]
        ev = expval_joint_analytical_multi(q, obs)
        want = O.run_ops(n, 1, ops)
        ev_want = torch.stack([O.expval_joint_rotated(want, o) for o in obs], dim=-1)
        assert_close(ev, ev_want, rtol=GRAD_RTOL, what=f"rank {rank} expvals")
        nl = q.nl
        assert q.stats["exchanges"] >= 1
        if nl >= 18:   # rotating block stages left their layouts to the device instead of sorting the qubits home
            assert q.stats.get("layouts_kept", 0) >= 1, q.stats
        slab = q.get_states_1d()
        assert_close(slab, want[:, rank << nl:(rank + 1) << nl], rtol=RTOL, what=f"rank {rank} slab")
        full = q.gather_states_1d()
        assert_close(full, want, rtol=RTOL, what="gathered state")
        # per-wire <Z> after a few more gates that leave qubits in the rank index
        tqf.rx(q, wires=0, params=torch.tensor([[0.3]]))
        tqf.cnot(q, wires=[n - 1, 0])
        z = expval(q, list(range(n)), [cases._Obs("Z")] * n)
        want2 = O.run_ops(n, 1, [{"name": "rx", "wires": [0], "params": torch.tensor([[0.3]])},
                                 {"name": "cnot", "wires": [n - 1, 0]}], want.reshape([1] + [2] * n))
        zw, _ = O.expval(want2.reshape([1] + [2] * n), list(range(n)), ["Z"] * n)
        # (n = 19: the complex64 oracle's own float32 marginals over 2^19 probabilities are only good to a few 1e-6)
        assert_close(z, zw, rtol=GRAD_RTOL if n < 16 else 5e-5, what="per-wire Z")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,tiles", [(2, 8, None), (4, 9, None), (2, 9, (4, 5, 2)), (4, 10, (5, 6, 3)),
                                           (2, 19, None)])   # 18 local qubits: rotating tensor-core blocks, quadratic-phase passes, layouts kept across exchanges
def test_sharded_qft_brickwork_matches_oracle(world, n, tiles):
    emu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libtqb200_emu.so")
    if not os.path.exists(emu):
        pytest.skip("host emulation library not built (run __graft_entry__.build())")
    mp.spawn(_worker, args=(world, _free_port(), n, 3, emu, tiles), nprocs=world, join=True)


def _worker_unrunnable(rank, world, port, emu_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchquantum_b200 as tq
        import torchquantum_b200.functional as tqf
        from torchquantum_b200 import _native

        _native.use_library_for_testing(emu_path)
        # a dense gate that needs more qubits on one GPU than can stay local during an exchange: must raise, not spin
        q = tq.ShardedQuantumDevice(n_wires=5, bsz=1, device="cpu")
        tqf.qft(q, wires=[0, 1, 2, 3, 4])
        with pytest.raises(NotImplementedError, match="decompose"):
            q.flush()
        with pytest.raises(ValueError, match="too few qubits"):
            tq.ShardedQuantumDevice(n_wires=2, bsz=1, device="cpu")
        # a 2-qubit gate whose qubits are both needed on chip while the rank qubit comes in: victims avoid them
        q = tq.ShardedQuantumDevice(n_wires=4, bsz=1, device="cpu")
        q.h(wires=3)
        tqf.rxx(q, wires=[0, 3], params=torch.tensor([[0.6]]))
        tqf.rxx(q, wires=[0, 1], params=torch.tensor([[0.2]]))
        full = q.gather_states_1d()
        from oracle import tq_oracle as O
        from tests.helpers import assert_close
        want = O.run_ops(4, 1, [{"name": "hadamard", "wires": [3]},
                                {"name": "rxx", "wires": [0, 3], "params": torch.tensor([[0.6]])},
                                {"name": "rxx", "wires": [0, 1], "params": torch.tensor([[0.2]])}])
        assert_close(full, want, what="rxx across the rank qubit")
    finally:
        dist.destroy_process_group()


def test_sharded_flush_makes_progress_or_raises():
    """no endless exchange loop for gates that can never be made local."""
    emu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libtqb200_emu.so")
    if not os.path.exists(emu):
        pytest.skip("host emulation library not built (run __graft_entry__.build())")
    mp.spawn(_worker_unrunnable, args=(2, _free_port(), emu), nprocs=2, join=True)


def _worker_autograd(rank, world, port, n, emu_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchquantum_b200 as tq
        import torchquantum_b200.functional as tqf
        from tests.helpers import GRAD_RTOL, assert_close
        from torchquantum_b200 import _native
        from torchquantum_b200.measurement import expval_joint_analytical_multi

        _native.use_library_for_testing(emu_path)
        torch.manual_seed(4)
        thetas = [torch.randn(1, 1).requires_grad_(True) for _ in range(3 * n)]
        obs = ["Z" * n, "ZZ" + "I" * (n - 2), "X" + "I" * (n - 2) + "Y", "I" * (n - 1) + "Z"]
        wts = torch.tensor([0.7, -1.3, 0.4, 2.0])

        def circuit(q):
            k = 0
            for w in range(n):
                tqf.hadamard(q, wires=w)
            for layer in range(3):
                for w in range(n):                      # rotations on EVERY wire, incl. the rank-index ones
                    (tqf.rx, tqf.ry, tqf.rz)[(layer + w) % 3](q, wires=w, params=thetas[k])
                    k += 1
                for w in range(layer % 2, n - 1, 2):
                    tqf.cnot(q, wires=[w, w + 1])
                tqf.crx(q, wires=[n - 1, 0], params=thetas[(5 * layer) % len(thetas)])
            return (expval_joint_analytical_multi(q, obs) * wts).sum()

        ref = tq.QuantumDevice(n_wires=n, bsz=2, device="cpu")
        ref_loss = circuit(ref)
        want = torch.autograd.grad(ref_loss, thetas)
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=2, device="cpu")
        loss = circuit(q)
        assert q.stats["exchanges"] >= 1
        assert_close(loss.detach().reshape(1), ref_loss.detach().reshape(1), what="sharded loss")
        for t in thetas:
            t.grad = None
        loss.backward()
        q.allreduce_gradients(thetas)
        assert_close(torch.cat([t.grad.reshape(-1) for t in thetas]), torch.cat([g.reshape(-1) for g in want]),
                     rtol=GRAD_RTOL, what=f"rank {rank}: gradients through exchanges")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 7), (4, 8)])
def test_sharded_adjoint_gradients_match_one_device(world, n):
    """north_star (2) + (4): adjoint differentiation through the sharded device (exchange = involution in the backward,
    expectation values all-reduced with a pass-through gradient, shared-parameter partials summed over the ranks)."""
    emu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libtqb200_emu.so")
    if not os.path.exists(emu):
        pytest.skip("host emulation library not built (run __graft_entry__.build())")
    mp.spawn(_worker_autograd, args=(world, _free_port(), n, emu), nprocs=world, join=True)


def _worker_semantics(rank, world, port, emu_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchquantum_b200 as tq
        import torchquantum_b200.functional as tqf
        from tests.helpers import GRAD_RTOL, assert_close
        from torchquantum_b200 import _native
        from torchquantum_b200.measurement import expval, expval_joint_analytical_multi

        _native.use_library_for_testing(emu_path)
        n = 5

        def circuit(q, p, peek):
            tqf.hadamard(q, wires=1)
            tqf.rx(q, wires=0, params=p)              # wire 0 lives in the rank index: exchange
            tqf.cnot(q, wires=[0, 3])
            if peek:
                with torch.no_grad():                 # logging / debugging read in the middle of the circuit
                    q.get_states_1d()
            tqf.ry(q, wires=1, params=p)
            tqf.crx(q, wires=[4, 0], params=p)
            return (expval_joint_analytical_multi(q, ["ZZIII", "XIIZY"]) * torch.tensor([0.6, -1.1])).sum()

        p0 = torch.tensor([[0.45]], requires_grad=True)
        want_loss = circuit(tq.QuantumDevice(n_wires=n, bsz=2, device="cpu"), p0, False)
        want, = torch.autograd.grad(want_loss, p0)
        for peek in (False, True):
            p = torch.tensor([[0.45]], requires_grad=True)
            q = tq.ShardedQuantumDevice(n_wires=n, bsz=2, device="cpu")
            loss = circuit(q, p, peek)
            assert_close(loss.detach().reshape(1), want_loss.detach().reshape(1), what=f"loss (peek={peek})")
            loss.backward()
            q.allreduce_gradients([p])
            assert_close(p.grad, want, rtol=GRAD_RTOL, what=f"rank {rank}: gradient with a no_grad peek={peek}")

        # gates issued under no_grad are not tracked even when the flush happens outside it, and they do not overwrite the
        # slab a pending backward still needs
        p = torch.tensor([[0.45]], requires_grad=True)
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=2, device="cpu")
        tqf.rx(q, wires=2, params=p)
        first = expval_joint_analytical_multi(q, ["IIZII"]).sum()
        with torch.no_grad():
            tqf.rx(q, wires=2, params=p)
            tqf.rx(q, wires=0, params=p)
        second = expval_joint_analytical_multi(q, ["IIZII"])
        assert not second.requires_grad
        first.backward()
        q.allreduce_gradients([p])
        assert_close(p.grad.reshape(1), -2 * torch.sin(torch.tensor([0.45])), rtol=GRAD_RTOL, what="first read-out's gradient")
        assert_close(second, torch.cos(torch.tensor([0.9])).expand(2, 1), what="second read-out")

        # parameter values are those at the gate call: plain tensors are snapshotted, an in-place edit of a tensor that
        # takes part in autograd raises at the flush
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device="cpu")
        v = torch.tensor([[0.3]])
        tqf.rx(q, wires=2, params=v)
        v.fill_(1.1)
        assert_close(expval_joint_analytical_multi(q, ["IIZII"]), torch.cos(torch.tensor([[0.3]])), what="snapshot")
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device="cpu")
        w = torch.tensor([[0.3]], requires_grad=True)
        tqf.rx(q, wires=2, params=w)
        with torch.no_grad():
            w.fill_(1.1)
        with pytest.raises(RuntimeError, match="modified in place"):
            q.flush()

        # per-wire read-out has no backward on the sharded device: loud, not silently detached
        q = tq.ShardedQuantumDevice(n_wires=n, bsz=1, device="cpu")
        tqf.rx(q, wires=2, params=torch.tensor([[0.3]], requires_grad=True))
        from tests import cases
        with pytest.raises(NotImplementedError, match="not differentiable"):
            expval(q, [2], [cases._Obs("Z")])
    finally:
        dist.destroy_process_group()


def test_sharded_device_call_time_semantics():
    """ADVICE r1 for the sharded device too: the autograd mode is the one the gates were issued under (a no_grad peek in
    the middle of a circuit neither drops the graph nor corrupts it), parameter values are those at the call."""
    emu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libtqb200_emu.so")
    if not os.path.exists(emu):
        pytest.skip("host emulation library not built (run __graft_entry__.build())")
    mp.spawn(_worker_semantics, args=(2, _free_port(), emu), nprocs=2, join=True)


def _worker_batch_axis(rank, world, port, emu_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        import torchquantum_b200 as tq
        from tests.helpers import GRAD_RTOL, assert_close
        from torchquantum_b200 import _native

        _native.use_library_for_testing(emu_path)
        n, reps, total = 5, 3, 12
        torch.manual_seed(0)
        np.random.seed(0)
        model = bench.VQEModel(tq, n=n, reps=reps)          # the bench's model class, same seed on every rank
        model.encoder = tq.GeneralEncoder([{"input_idx": [k], "func": "ry", "wires": [k]} for k in range(n)])
        params = [p for p in model.parameters() if p.requires_grad]
        x_all = torch.rand(total, n, generator=torch.Generator().manual_seed(7)) * np.pi
        per = total // world
        x = x_all[rank * per:(rank + 1) * per]               # contiguous batch shard
        loss = model(x)
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in params])   # what bench.py's reduce_grads() does
        dist.all_reduce(flat)
        flat /= world                                            # mean over equal shards = mean over the whole batch
        losses = torch.tensor([loss.item()])
        dist.all_reduce(losses)
        for p in params:
            p.grad = None
        whole = model(x_all)
        whole.backward()
        want = torch.cat([p.grad.reshape(-1) for p in params])
        assert_close(losses / world, whole.detach().reshape(1), what="mean energy over the shards")
        assert_close(flat, want, rtol=GRAD_RTOL, what=f"rank {rank}: all-reduced shared-parameter gradients")
    finally:
        dist.destroy_process_group()


def test_batch_sharded_step_matches_one_process():
    """SURVEY §8e, batch axis (configs 1-3): every rank runs its contiguous shard of the batch, no data-path collective;
    the one exchange is the all-reduce of the shared-parameter gradients, which must equal the whole-batch gradient."""
    emu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libtqb200_emu.so")
    if not os.path.exists(emu):
        pytest.skip("host emulation library not built (run __graft_entry__.build())")
    mp.spawn(_worker_batch_axis, args=(2, _free_port(), emu), nprocs=2, join=True)


@pytest.mark.parametrize("grad", [False, True])
def test_sharded_random_catalogue_gates(grad):
    """Random gates from the whole in-tile catalogue (controls / parity bits in the rank index, 2-qubit dense gates across it,
    per-sample parameters, inverses) on a sharded device vs the oracle (states, expectation values) and, with `grad`, the
    sharded adjoint vs one device: a few seeds of tools/fuzz_sharded.py."""
    emu = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libtqb200_emu.so")
    if not os.path.exists(emu):
        pytest.skip("host emulation library not built (run __graft_entry__.build())")
    from tools.fuzz_sharded import worker

    mp.spawn(worker, args=(2, _free_port(), 7, [0, 1, 2], 40, emu, 2, grad), nprocs=2, join=True)

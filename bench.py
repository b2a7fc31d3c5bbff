"""Benchmark of the gate-application hot path (BASELINE.json metric: batched circuit fwd+bwd
circuits/s; fused gate-pass HBM GB/s vs peak).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE config 2): 10-qubit EfficientSU2 ansatz, reps=20 (600 gates, 420 shared
parameters) behind a per-sample `10_ry` encoder, bsz=4096 per GPU, 19-term energy
(sum Z_i Z_{i+1} + sum X_i), forward + adjoint backward through the public drop-in API.
  value    : circuits/s with inputs resident in HBM (device-timed, max over ranks)
  e2e      : the same step fed from pinned HOST memory, loss read back every step
  roofline : every gate-pass launch of a config-4 style circuit as the planner schedules it (tensor-core block
             passes + register-tile passes), timed live on an n=28 state (2 GiB, larger than L2) against the
             measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference : the reference's OWN code (the unmodified package in baseline/_ref: gate_wrapper +
             bmm + torch autograd on the host cores, all threads) running the same model class on the same workload in
             bsz-1024 chunks; the oracle port (oracle/tq_oracle.py) only if that package cannot be imported (`kind` says which).
  eager    : the same step through the plain eager API (no tq.GraphedStep), device-timed and end to end.
N > 1 (torchrun): the batch is sharded, 4096 circuits per rank (weak scaling); the only exchange is
the all-reduce of the 420 shared-parameter gradients; the line also carries `sharded_configs`: config 3 strong-scaled
over the ranks and the config-5 family (QFT + brickwork on a statevector sharded over the ranks, n = 30 + log2 N).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WIRES, REPS, BSZ = 10, 20, 4096
WORKLOAD = ("cfg2: 10-qubit EfficientSU2 reps=20 (600 gates, 420 shared params) + 10_ry encoder, "
            "bsz=4096 per GPU, 19-term energy, fwd + adjoint bwd")


def energy_terms(n):
    return ["I" * i + "ZZ" + "I" * (n - 2 - i) for i in range(n - 1)] + ["I" * i + "X" + "I" * (n - 1 - i) for i in range(n)]


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
class VQEModel(torch.nn.Module):
    def __init__(self, tq, n=N_WIRES, reps=REPS):
        super().__init__()
        self.tq, self.n = tq, n
        self.encoder = tq.GeneralEncoder(tq.encoder_op_list_name_dict["10_ry"])
        self.ansatz = tq.EfficientSU2(n_wires=n, reps=reps, entanglement_layer="linear")
        self.terms = energy_terms(n)

    def forward(self, x):
        tq = self.tq
        qdev = tq.QuantumDevice(n_wires=self.n, bsz=x.shape[0], device=x.device)
        self.encoder(qdev, x)
        self.ansatz(qdev)
        energy = 0
        for t in self.terms:  # the reference API: one call per Pauli string (algorithm/vqe.py:80-113)
            energy = energy + tq.expval_joint_analytical(qdev, t)
        return energy.mean()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [s.strip() for s in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _cfg4_family(n, layers, seed=0):
    """plan key + angles of a config-4 style circuit: rx/ry/rz on every wire + brickwork CNOTs per layer"""
    rng = np.random.default_rng(seed)
    key, thetas = [], []
    for layer in range(layers):
        for w in range(n):
            key.append((["rx", "ry", "rz"][rng.integers(3)], (w,), False, "sh", False, 0))
            thetas.append(rng.uniform(-np.pi, np.pi))
        for w in range(layer % 2, n - 1, 2):
            key.append(("cnot", (w, w + 1), False, None, False, 0))
            thetas.append(None)
    return key, thetas


def _time_planned_circuit(n, key, thetas, dev, reps=3):
    """Runs the forward-only plan of `key` on a 2^n state, one ABI call per kernel launch, CUDA events around each.
    -> {"block": [ms...], "pass": [ms...]}, launches per repetition, gates per kind, norm of the final state."""
    from torchquantum_b200 import _native, engine
    from torchquantum_b200.engine import _build_mats, _desc, _dev

    plan = engine.get_plan(n, tuple(key), False)
    state = torch.zeros(1, 2 ** n, dtype=torch.complex64, device=dev)
    state[0, 0] = 1
    other = torch.empty_like(state) if any(getattr(s, "n_rotating", 0) for s in plan.stages) else None
    evs = {"block": [], "pass": []}
    per_rep = {"block": 0, "pass": 0}
    pps = torch.zeros(1, 1, dtype=torch.float32, device=dev)
    for rep in range(reps + 1):
        for stage in plan.stages:
            psh = torch.tensor([thetas[r] for r in stage.sh_recs] or [0.0], dtype=torch.float32, device=dev)
            if stage.kind == "blocks":
                t = []
                out = engine.run_blocks(state, stage, n, 0, psh, True, timing=t, scratch=other)
                if out is not state:      # rotating blocks move the state between the two buffers
                    state, other = out, state
                if rep > 0:
                    evs["block"] += t
                elif rep == 0:
                    per_rep["block"] += len(t)
            elif stage.kind == "fused":
                msh, mps = _build_mats(stage, 1, psh, pps, [], dev)
                for pp in stage.passes:
                    prog = _dev(pp, "prog", dev)
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    _native.call("tqb_pass_forward", state, state, n, 1, prog, _desc(pp), msh,
                                 mps, stage.M_ps, 0)
                    e.record()
                    if rep > 0:
                        evs["pass"].append((s, e))
                    else:
                        per_rep["pass"] += 1
            else:
                raise RuntimeError(f"unexpected stage kind {stage.kind} in the roofline circuit")
    torch.cuda.synchronize()
    norm = float(torch.linalg.vector_norm(torch.view_as_real(state).reshape(-1), dtype=torch.float64).item())
    return {k: [s.elapsed_time(e) for s, e in v] for k, v in evs.items()}, per_rep, norm


def _sustained(ms_sus, dom, per_rep, alg_bytes, peak):
    """the last third of a ~1.5 s run of the same launches (power cap reached)"""
    v = ms_sus[dom]
    tail = v[2 * len(v) // 3:]
    avg = float(np.mean(tail))
    return {"avg_launch_ms": round(avg, 4), "GBps": round(alg_bytes / avg / 1e6, 1), "frac": round(alg_bytes / avg / 1e6 / peak, 4),
            "launches_timed": len(tail), "run_ms": round(float(sum(np.sum(x) for x in ms_sus.values())), 1),
            "note": "same launches back to back until the board is power capped (~1000 W, SM clock ~1.3 GHz): the block "
                    "kernel's per-tile pipeline (TF32 split, 24 tcgen05.mma, epilogue: ~2 500 cycles per tile) then "
                    "takes longer than the tile's HBM traffic; `frac` above is the un-capped figure"}


def gate_pass_roofline(dev, n=28, layers=16):
    """Time every launch of the gate-pass kernels on a config-4 style circuit (rx/ry/rz on every wire + brickwork
    CNOTs) over a 2^n state (2 GiB at n=28: far larger than the 126 MB L2), exactly as the planner schedules it:
    tensor-core block passes (tqb_block6_kernel: one dense 6-qubit block unitary per pass, whatever number of gates it
    absorbed) and register-tile passes (tqb_pass_kernel) for what is left.  The roofline object describes the kernel
    with the larger share of the time; a sweep of single register-tile passes holding k gates is kept beside it."""
    key, thetas = _cfg4_family(n, layers)
    n_gates = len(key)
    ms, per_rep, norm = _time_planned_circuit(n, key, thetas, dev)
    # the same launches back to back for ~1.5 s: after ~0.1 s of this load the board sits at its 1000 W power cap and the
    # SM clock drops from 1.97 to ~1.3 GHz (a plain device copy also reaches the cap, at 1.75 GHz, without losing bandwidth)
    n_sus = max(int(1500.0 / max(sum(float(np.sum(v)) for v in ms.values()) / 3.0, 1.0)), 3)
    ms_sus, _, _ = _time_planned_circuit(n, key, thetas, dev, reps=n_sus)
    alg_bytes = 16.0 * 2 ** n
    peak, peak_src = measured_peak()
    tot = {k: float(np.sum(v)) for k, v in ms.items()}
    dom = "block" if tot["block"] >= tot["pass"] else "pass"
    avg_ms = float(np.mean(ms[dom]))
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    reps = max(len(ms[dom]) // max(per_rep[dom], 1), 1)
    circuit_ms = (tot["block"] + tot["pass"]) / reps
    launches = per_rep["block"] + per_rep["pass"]
    sweep = []
    for k in (1, 4, 8, 16):  # k gates on the qubits of one register tile -> exactly one tqb_pass_kernel launch
        r2 = np.random.default_rng(k)
        bits = list(range(5)) + list(range(n - 7, n))
        key_k, th_k = [], []
        for i in range(k):
            b = bits[r2.integers(len(bits))]
            key_k.append((["rx", "ry", "rz"][r2.integers(3)], (n - 1 - b,), False, "sh", False, 0))
            th_k.append(r2.uniform(-np.pi, np.pi))
        ms_k, rep_k, _ = _time_planned_circuit(n, key_k, th_k, dev)
        if rep_k["pass"] == 1 and rep_k["block"] == 0:
            t = float(np.mean(ms_k["pass"]))
            sweep.append({"gates_in_pass": k, "ms": round(t, 4), "GBps": round(alg_bytes / t / 1e6, 1),
                          "frac": round(alg_bytes / t / 1e6 / peak, 4)})
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        ent = tj.get("tqb_block6_kernel" if dom == "block" else "tqb_pass_kernel", {})
        traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
    kernel = "tqb_block6_kernel (tcgen05 3xTF32, TMA tile IO)" if dom == "block" else "tqb_pass_kernel<R=4,fwd>"
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": kernel, "algorithmic_bytes_per_launch": alg_bytes,
            "avg_launch_ms": round(avg_ms, 4), "launches_timed": len(ms[dom]),
            "share_of_circuit_time": round(tot[dom] / max(tot["block"] + tot["pass"], 1e-9), 4),
            "workload": f"cfg4 family: n={n} single statevector, {layers} layers = {n_gates} gates in "
                        f"{per_rep['block']} tensor-core block passes + {per_rep['pass']} register-tile passes "
                        f"({n_gates / max(launches, 1):.1f} gates/pass)",
            "state_norm_after": round(norm, 7),
            "circuit": {"ms": round(circuit_ms, 3), "launches": launches,
                        "algorithmic_GBps": round(launches * alg_bytes / (circuit_ms * 1e-3) / 1e9, 1),
                        "frac_of_peak": round(launches * alg_bytes / (circuit_ms * 1e-3) / 1e9 / peak, 4),
                        "gate_amp_updates_per_s": round(n_gates * 2.0 ** n / (circuit_ms * 1e-3), 1)},
            "per_kind_avg_ms": {k: (round(float(np.mean(v)), 4) if v else None) for k, v in ms.items()},
            "sustained": _sustained(ms_sus, dom, per_rep, alg_bytes, peak),
            "register_tile_single_pass_sweep": sweep,
            "note": "a block pass costs the same whatever it holds (one 64x64 complex unitary per 2^12-amplitude tile, "
                    "24 tcgen05.mma per tile); the register-tile kernel is HBM-bound up to ~4 gates per pass and "
                    "issue-bound beyond (sweep)"}


def sharded_configs(dev, world, rank):
    """N > 1 only, after the headline measurement: the two multi-GPU configs BASELINE.json names, each with its own
    correctness check, so that the driver's scaling run carries them (tools/configs.py holds the workloads).
      cfg 3: 20-qubit QAOA p=8, 64 circuits in total, batch-sharded (strong scaling, no data-path collective);
      cfg 5: QFT + 4 brickwork layers on a statevector of 31 + log2(N) qubits sharded over the N GPUs (32 qubits on 2,
             33 on 4, 34 on 8: 16 GiB per GPU), global <-> local qubit exchanges timed against the NVLink peak."""
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import configs

    out = {}
    try:
        r3 = configs.run_cfg3(dev, bsz=max(64 // world, 1))
        t = torch.tensor([r3["fwd_ms"], r3["fwd_bwd_ms"], 0.0 if r3["ok"] else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["cfg3"] = {"workload": r3["workload"] + f" per rank: 64 circuits in total, batch-sharded over {world} B200",
                       "scaling": "strong", "fwd_ms": round(t[0].item(), 3), "fwd_bwd_ms": round(t[1].item(), 3),
                       "circuits_per_s_fwd_bwd": round(64 / (t[1].item() * 1e-3), 1), "ok": t[2].item() == 0.0,
                       "oracle64_parity_rank0": r3["oracle64_parity"], "max_norm_err_rank0": r3["max_norm_err"]}
    except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
        out["cfg3"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    try:
        g = int(np.log2(world))
        out["cfg5"] = configs.run_cfg5(dev, 31 + g)
    except Exception as e:  # noqa: BLE001
        out["cfg5"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    return out


def load_real_reference():
    """The UNMODIFIED reference package installed into baseline/_ref (pip --no-deps from /root/reference; git-ignored, travels
    to the GPU box).  Its __init__ imports qiskit / torchpack / matplotlib ... which are not installed and never touched on
    the statevector path: those top-level names resolve to inert stand-ins (the same shim as oracle/refshim.py, without its
    `mat_dict` repair: config 2 does not reach the functions that need it).  -> the module, or None."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "torchquantum")):
        return None
    try:
        import warnings

        from oracle import refshim

        warnings.filterwarnings("ignore")
        if not any(isinstance(f, refshim._FakeFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, refshim._FakeFinder())
        if ref_root not in sys.path:
            sys.path.insert(0, ref_root)
        import torchquantum as ref_tq

        assert os.path.realpath(ref_tq.__file__).startswith(os.path.realpath(ref_root)), ref_tq.__file__
        return ref_tq
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: reference package in baseline/_ref is not importable ({e!r}); using the oracle port", file=sys.stderr)
        return None


def cpu_reference_step(bsz, threads):
    """One fwd+bwd of the config-2 workload on the host cores -> (step function, kind, description).
    kind "reference": the reference's own code (baseline/_ref) running the SAME model class as the GPU arm (VQEModel:
    GeneralEncoder('10_ry') + EfficientSU2(10, reps=20, linear) + 19 expval_joint_analytical calls + backward through
    torch autograd); kind "port": the oracle restatement (oracle/tq_oracle.py) when the package cannot be imported."""
    torch.set_num_threads(threads)
    n = N_WIRES
    g = torch.Generator().manual_seed(0)
    ref_tq = load_real_reference()
    if ref_tq is not None:
        torch.manual_seed(0)
        np.random.seed(0)
        model = VQEModel(ref_tq)
        params = [p for p in model.parameters() if p.requires_grad]
        assert sum(p.numel() for p in params) == 420, "reference EfficientSU2(10, reps=20) should hold 420 parameters"
        x = torch.rand(bsz, n, generator=g) * np.pi

        def step():
            for p in params:
                p.grad = None
            loss = model(x)
            loss.backward()
            return float(loss.detach())

        return step, "reference", ("mit-han-lab/torchquantum itself (baseline/_ref, unmodified; gate_wrapper + bmm + torch "
                                   "autograd on the host cores)")
    from oracle import tq_oracle as O

    x = (torch.rand(bsz, n, generator=g) * np.pi).requires_grad_(True)
    thetas = [(torch.rand(1, 1, generator=g) * 6 - 3).requires_grad_(True) for _ in range(2 * n * (REPS + 1))]
    terms = energy_terms(n)

    def step():
        ops = [{"name": "ry", "wires": [w], "params": x[:, [w]]} for w in range(n)]
        k = 0
        for r in range(REPS + 1):
            for name in ("ry", "rz"):
                for w in range(n):
                    ops.append({"name": name, "wires": [w], "params": thetas[k]})
                    k += 1
            if r < REPS:
                ops += [{"name": "cnot", "wires": [w, w + 1]} for w in range(n - 1)]
        states = O.run_ops(n, bsz, ops)
        energy = sum(O.expval_joint_analytical(states, t) for t in terms).mean()
        energy.backward()
        return float(energy.detach())

    return step, "port", "oracle port of the reference path (torch CPU restatement of gate_wrapper/bmm + autograd)"


def run_reference_arm(args):
    """The reference's own CPU implementation of the headline workload: each step = the 4096-circuit batch of config 2 in
    4 chunks of 1024 (10.7 GB RSS per chunk), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    chunk, chunks = 1024, BSZ // 1024
    step, kind, what = cpu_reference_step(chunk, threads)
    t0 = time.perf_counter()
    step()                                   # warm-up chunk (also sizes the run)
    t_chunk = time.perf_counter() - t0
    # a step is the whole batch (4 chunks) when K steps of it fit in ~3 minutes on this host, else a bounded sample of it
    chunks = int(max(1, min(chunks, 180.0 / max(args.steps * t_chunk, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(chunks):
            step()
    dt = (time.perf_counter() - t0) / args.steps
    value = chunks * chunk / dt
    line = {"metric": "batched circuit fwd+bwd circuits/s", "value": round(value, 2), "unit": "circuits/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": 1,
            "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BSZ,
                       "sample": f"every step = {chunks * chunk} circuits of the {BSZ}-circuit batch, run as {chunks} chunk(s) of "
                                 f"{chunk} on the host (one warm-up chunk; 4 chunks = the whole batch)"},
            "cpu_baseline": {"value": round(value, 2), "unit": "circuits/s", "cores": threads, "kind": kind,
                             "sample": f"{what}; {chunks} x bsz={chunk} chunks per step = {chunks * chunk} circuits"},
            "e2e": {"value": round(value, 2), "unit": "circuits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints "NCCL version ..." on stdout when its communicator is created (NCCL_DEBUG >= VERSION, which WARN
        # includes): send fd 1 to stderr while that happens so that rank 0's stdout is exactly one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    import torchquantum_b200 as tq
    from torchquantum_b200 import engine

    torch.manual_seed(0)
    np.random.seed(0)
    model = VQEModel(tq).to(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = (torch.rand(BSZ, N_WIRES, generator=gen) * np.pi).pin_memory()
    x_dev = x_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def reduce_grads():
        if world > 1:   # eager data parallelism: one all-reduce of the flattened gradients, written back into .grad
            grads = [p.grad for p in params]
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            torch._foreach_copy_([g.view(-1) for g in grads], list(flat.split([g.numel() for g in grads])))

    def eager_step(x, read_loss):
        for p in params:
            p.grad = None
        loss = model(x)
        loss.backward()
        reduce_grads()
        return loss.item() if read_loss else loss

    graphed = None

    def graphed_step(x, read_loss):
        loss = graphed(x)  # copies x into the captured input buffer, replays fwd + bwd
        if world > 1:
            graphed.allreduce_gradients()   # one all-reduce of the replay's flat gradient + a captured scatter into .grad
        return loss.item() if read_loss else loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, k, from_host):
        evs = []
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            flush.zero_()  # L2 flush between steps, outside the per-step events
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if from_host:
                step(x_host if step is graphed_step else x_host.to(dev, non_blocking=True), True)
            else:
                step(x_dev, False)
            e.record()
            evs.append((s, e))
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item() / k, t[1].item() / k

    for _ in range(max(args.warmup, 3)):
        eager_step(x_dev, False)
    launches0 = engine.STATS["launches"]
    eager_step(x_dev, False)
    launches_per_step = engine.STATS["launches"] - launches0   # kernels of libtqb200 per step (the graph replays the same)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eager_ms, eager_wall_ms = timed(eager_step, args.steps, False)
    eager_e2e_ms, _ = timed(eager_step, args.steps, True)
    # the same step captured once into a CUDA graph (tq.GraphedStep) and replayed: the public API for
    # launch-bound batched workloads; the eager numbers stay in the line as `eager`
    graphed = tq.GraphedStep(lambda x: model(x), [x_dev], params, warmup=3)
    for _ in range(max(args.warmup, 3)):
        graphed_step(x_dev, False)
    step_ms, wall_ms = timed(graphed_step, args.steps, False)
    e2e_step_ms, e2e_wall_ms = timed(graphed_step, args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    sharded = sharded_configs(dev, world, rank) if world > 1 and not args.no_sharded_configs else None
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    roof = gate_pass_roofline(dev)
    line = {
        "metric": "batched circuit fwd+bwd circuits/s", "value": round(BSZ * world / (step_ms * 1e-3), 1),
        "unit": "circuits/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BSZ * world, "parallelism": f"batch-sharded x{world}",
                   "l2": "flushed by a 256 MiB device write between timed steps",
                   "timing": "per-step CUDA events on the current stream, max over ranks",
                   "gate_amplitude_updates_per_step": 610 * BSZ * world * 2 ** N_WIRES},
        "e2e": {"value": round(BSZ * world / (e2e_step_ms * 1e-3), 1), "unit": "circuits/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4) * world, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": round(e2e_step_ms, 4),
                "wall_ms_per_step_incl_l2_flush": round(e2e_wall_ms, 4),
                "note": "x copied from pinned host memory into the step's input buffer every step, loss.item() read "
                        "back every step"},
        "step_api": "tq.GraphedStep (fwd + adjoint bwd captured once as a CUDA graph, replayed per step)",
        "eager": {"value": round(BSZ * world / (eager_ms * 1e-3), 1), "ms_per_step": round(eager_ms, 4),
                  "wall_ms_per_step": round(eager_wall_ms, 4),
                  "e2e_value": round(BSZ * world / (eager_e2e_ms * 1e-3), 1), "e2e_ms_per_step": round(eager_e2e_ms, 4),
                  "note": "same step through the plain eager API (Python queues 610 gates and walks autograd every step)"},
        "gpu_launches": launches, "gpu_launches_per_step": launches_per_step,
        "device_ms_per_step": round(step_ms, 4), "wall_ms_per_step": round(wall_ms, 4),
        "clocks": clocks, "roofline": roof,
    }
    if sharded is not None:
        line["sharded_configs"] = sharded
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample_bsz = 1024
        cstep, kind, what = cpu_reference_step(sample_bsz, threads)
        cstep()
        t0 = time.perf_counter()
        reps = 3   # ~2 s per 1024-circuit chunk on 16 cores: ~8 s of CPU work incl. the warm-up chunk
        for _ in range(reps):
            cstep()
        dt = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = {"value": round(sample_bsz / dt, 2), "unit": "circuits/s", "cores": threads, "kind": kind,
                                "sample": f"{what}; one bsz={sample_bsz} chunk of the same workload per step (a quarter of "
                                          f"the batch), {reps} steps after 1 warm-up"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded-configs", action="store_true", help="N > 1: skip the cfg 3 / cfg 5 runs after the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""GPU parity suite: libtqb200.so (sm_100a) through the public API / C ABI vs golden vectors of the
real reference and vs the CPU oracle.  Tolerances (tests/helpers.py): 1e-5 relative on states, expectation values AND
gradients (north_star); histograms bit-exact.  Depth-100 circuits are compared with the float64 twin of the oracle."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def real_library():
    from torchquantum_b200 import _native, engine

    _native.use_library_for_testing(None)
    engine._PLAN_CACHE.clear()
    yield
    torch.cuda.synchronize()


def test_native_library_loaded():
    from torchquantum_b200 import _native

    assert not _native.is_emulated()
    assert _native.abi_version() == _native.ABI_VERSION


def test_every_gate():
    cases.check_all_gates(DEV)


def test_random_circuits_fwd_bwd():
    cases.check_random_circuits(DEV)


@pytest.mark.parametrize("tiles", [(4, 5, 2), (5, 6, 3), (6, 7, 5)])
def test_random_circuits_multi_tile(tiles):
    from torchquantum_b200 import plan as P

    saved = (P.T_SINGLE, P.T_MULTI, P.L_MIN)
    try:
        P.T_SINGLE, P.T_MULTI, P.L_MIN = tiles
        cases.check_random_circuits(DEV, min_n=tiles[0] + 1)
    finally:
        P.T_SINGLE, P.T_MULTI, P.L_MIN = saved


def test_expval_per_wire_mutates_device():
    cases.check_expval_per_wire(DEV)


def test_measure_bit_exact():
    cases.check_measure_bit_exact(DEV)


def test_mnist_cfg1():
    cases.check_mnist_cfg1(DEV)


def test_unitary_gates():
    """qubitunitary / qubitunitarystrict vs the reference's outputs (56 cases: wire orders, inverse, per-sample, SVD projection)"""
    cases.check_unitary_gates(DEV)


def test_large_batch_readout():
    cases.check_large_batch_readout(DEV)


def test_quadphase_kernel():
    """K12: fused quadratic-phase diagonal pass vs its formula in float64 (local bits, rank bits, batch)"""
    cases.check_quadphase_kernel(DEV)
    cases.check_quadphase_kernel(DEV, nl=16, g=3, rank=5, bsz=1)


def test_qft_blocks_and_quadphase():
    """QFT-like circuits on streamed states: 6-qubit QFT blocks + quadratic-phase twiddle passes + relabelled SWAPs"""
    cases.check_qft_blocks_and_quadphase(DEV, n=18)
    cases.check_qft_blocks_and_quadphase(DEV, n=21, bsz=2)


def test_random_gates_through_blocks():
    """random gates from the whole in-tile catalogue (multi-controlled, parity diagonals, 2-qubit dense, inverses) fused into
    tensor-core blocks between register-tile stages, in place and rotating, vs the oracle"""
    cases.check_random_gates_through_blocks(DEV, seeds=(1, 2, 4, 5))


def test_block_write_back_modes():
    """tensor-core block stages, in place and rotating (contiguous tile write-back into a second buffer), vs the oracle"""
    cases.check_block_write_back_modes(DEV, n=20, depth=24)


@pytest.mark.parametrize("n,bsz,depth", [(10, 8, 20), (13, 2, 12), (14, 2, 10), (16, 1, 10), (18, 1, 8)])
def test_brickwork_vs_oracle(n, bsz, depth):
    """config-4 family at sizes the oracle finishes in seconds; n >= 14 streams tiles through HBM."""
    cases.check_against_oracle(DEV, n, bsz, cases.brickwork_ops(n, depth, seed=n), grad=(n <= 16))


def test_cfg2_vqe_class_api():
    cases.check_cfg2_vqe(DEV, n=6, reps=4, bsz=40)


def test_cfg2_vqe_cuda_graph_step():
    """tq.GraphedStep (whole fwd+bwd step replayed as a CUDA graph) gives the same energy and gradients."""
    cases.check_cfg2_vqe(DEV, n=6, reps=4, bsz=40, graphed=True)


def test_cfg2_full_shape_readout_kernels():
    """Warp-per-sample Pauli read-out at the config-2 shape (n=10, bsz=4096) vs the multi-term launch and the
    per-wire marginals; shared backward of 19 single-string read-outs vs one multi-string read-out."""
    import torchquantum_b200 as tq
    from torchquantum_b200.measurement import expval_joint_analytical, expval_joint_analytical_multi

    n, bsz = 10, 4096
    torch.manual_seed(0)
    terms = ["I" * i + "ZZ" + "I" * (n - 2 - i) for i in range(n - 1)] + ["I" * i + "X" + "I" * (n - 1 - i) for i in range(n)] \
        + ["XYZIIZYXII", "YYIIIIIIZZ"]
    st = torch.randn(bsz, 2 ** n, dtype=torch.complex64, device=DEV)
    st = (st / st.norm(dim=1, keepdim=True)).requires_grad_(True)
    q = tq.QuantumDevice(n_wires=n, bsz=bsz, device=DEV)
    q.set_states(st)
    multi = expval_joint_analytical_multi(q, terms)
    singles = torch.stack([expval_joint_analytical(q, t) for t in terms], dim=1)
    cases.assert_close(singles, multi, what="single-string vs multi-string read-out")
    # dense check of a few samples against the oracle's Kronecker form
    from oracle import tq_oracle as O
    want = torch.stack([O.expval_joint_analytical(st[:4].detach().cpu(), t) for t in terms], dim=1)
    cases.assert_close(multi[:4], want, what="read-out vs oracle")
    w = torch.randn(bsz, len(terms), device=DEV)
    (g1,) = torch.autograd.grad((singles * w).sum(), st, retain_graph=True)
    (g2,) = torch.autograd.grad((multi * w).sum(), st)
    cases.assert_close(g1, g2, rtol=cases.GRAD_RTOL, what="shared backward of single-string read-outs")


def test_cfg3_qaoa_per_sample_params():
    cases.check_cfg3_qaoa(DEV)


def test_cfg3_qaoa_n14_streamed():
    """Same family on a state that streams through HBM in tiles (n=14), 3-regular ring-with-chords graph."""
    n = 14
    edges = [(i, (i + 1) % n) for i in range(n)] + [(i, i + n // 2) for i in range(n // 2)]
    cases.check_cfg3_qaoa(DEV, n=n, p=2, bsz=2, edges=edges)


def test_forward_only_forms():
    """Scaled rx/ry/h/one-bit diagonals + folded CNOTs (programs without a backward) vs the oracle: one tile per
    sample, forced small tiles, and a really streamed state (n=15, T=12)."""
    from torchquantum_b200 import plan as P

    cases.check_forward_only_forms(DEV)
    cases.check_forward_only_forms(DEV, n=15, bsz=2)
    saved = (P.T_SINGLE, P.T_MULTI, P.L_MIN)
    try:
        P.T_SINGLE, P.T_MULTI, P.L_MIN = 5, 6, 3
        cases.check_forward_only_forms(DEV, n=9, bsz=2)
    finally:
        P.T_SINGLE, P.T_MULTI, P.L_MIN = saved


def test_qft_phase_ladders():
    """QFT / inverse QFT / constant rungs through the folded controlled-phase ladders vs the oracle, on-chip (n=8, 12)
    and streamed (n=15: ladder bits in the outer index)."""
    cases.check_qft_phase_ladders(DEV)
    cases.check_qft_phase_ladders(DEV, n=12, bsz=1)
    cases.check_qft_phase_ladders(DEV, n=15, bsz=1)


def test_dense_blocks_tensor_core():
    """K4: qubitunitaryfast / qft on 4-7 wires through the 3xTF32 tensor-core kernel (n >= 12) vs the oracle, and the
    same circuit through the CUDA-core dense kernel (n = 11)."""
    cases.check_dense_blocks(DEV, n=13)
    cases.check_dense_blocks(DEV, n=11)


def test_speculative_readout():
    cases.check_speculative_readout(DEV)


def test_op_hamil_exp():
    cases.check_op_hamil_exp(DEV)


def test_vqe_hamiltonian(tmp_path):
    cases.check_vqe_hamiltonian(DEV, tmp_path)


def test_per_sample_params_large_batch():
    ops = cases.brickwork_ops(6, 6, seed=1, per_sample_bsz=300)
    cases.check_against_oracle(DEV, 6, 300, ops, grad=True)


def test_full_size_properties_n24():
    """Size-independent properties at a size the oracle cannot reach quickly: unitarity (norm),
    circuit followed by its inverse is the identity, <Z> bounds."""
    import torchquantum_b200 as tq
    from torchquantum_b200.measurement import expval_joint_analytical

    n = 24
    ops = cases.brickwork_ops(n, 6, seed=7)
    q = tq.QuantumDevice(n_wires=n, bsz=1, device=DEV)
    cases.apply_ops(q, [dict(o, params=o["params"].to(DEV)) if "params" in o else o for o in ops], DEV)
    s = q.get_states_1d()
    norm = torch.linalg.vector_norm(torch.view_as_real(s).reshape(-1), dtype=torch.float64).item()
    assert abs(norm - 1.0) < 1e-5, norm
    ez = expval_joint_analytical(q, "Z" + "I" * (n - 1)).item()
    assert -1.00001 <= ez <= 1.00001
    inv = [dict(o, inverse=True, params=o["params"].to(DEV)) if "params" in o else dict(o, inverse=True)
           for o in reversed(ops)]
    cases.apply_ops(q, inv, DEV)
    s0 = q.get_states_1d()
    assert abs(s0[0, 0].abs().item() - 1.0) < 1e-5
    assert s0[0, 1:].abs().max().item() < 1e-5


@pytest.mark.parametrize("check", ["check_record_cache_follows_the_op_list", "check_grad_mode_is_sampled_when_gates_are_issued",
                                   "check_parameter_values_are_read_at_the_gate_call", "check_checkpoint_round_trip",
                                   "check_from_op_history_and_get_unitary", "check_static_mode_equals_dynamic",
                                   "check_layer_classes", "check_einsum_broadcast",
                                   "check_dropped_device_is_freed_by_refcount", "check_reference_test_vectors",
                                   "check_batch_collapse"])
def test_host_mirror_semantics(check):
    """lazy-queue semantics, record caches, checkpoints, static mode, op_history, layer classes, the pure
    apply_unitary_* functions on the real kernels"""
    from tests import cases_api

    getattr(cases_api, check)(DEV)


@pytest.mark.parametrize("n,mode,blocks", [(20, "fwd", True), (20, "fwd", False), (20, "grad", None),
                                           (24, "fwd", True), (24, "fwd", False)])
def test_depth100_parity_vs_fp64_oracle(n, mode, blocks):
    """the config-4 family at FULL depth (100 layers) against the float64 oracle twin: rel-L2 and
    |norm - 1| <= 1e-5 for the tensor-core block program, the register-tile forward-only program and the program
    autograd builds."""
    cases.check_depth_parity(DEV, n, 100, mode, blocks=blocks, expect_blocks=blocks)


def test_cfg4_full_size_parity_vs_fp64_oracle():
    """Config 4 itself (n=28, depth 100, 4150 gates) against the float64 oracle twin run with torch's CUDA ops
    (a 4 GiB complex128 state): every amplitude, rel-L2 and norm within 1e-5."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs 40 GiB of device memory for the complex128 oracle state")
    cases.check_depth_parity(DEV, 28, 100, "fwd", blocks=True, expect_blocks=True)


def test_cfg3_n16_p8_vs_oracle():
    """Config 3's own depth (p=8) on the 3-regular graph family at a size the CPU oracle finishes quickly."""
    n = 16
    edges = [(i, (i + 1) % n) for i in range(n)] + [(i, i + n // 2) for i in range(n // 2)]
    cases.check_cfg3_qaoa(DEV, n=n, p=8, bsz=2, edges=edges)


def test_cfg3_full_size_properties():
    """Config 3 at its full size (n=20, p=8, bsz=64, 900 gates): unit norm for all 64 samples and, for the first and
    last sample, states / cost / adjoint gradients against the float64 oracle twin (1e-5)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import configs

    r = configs.run_cfg3(torch.device(DEV), timing=False)
    assert r["ok"], r


def test_cfg4_full_size_properties():
    """Config 4 at its full size (n=28, depth 100, ~4150 gates): unit norm, |<Z>| <= 1, circuit . circuit^-1 = 1."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import configs

    r = configs.run_cfg4(torch.device(DEV), timing=False)
    assert r["ok"], r


def _torchrun(nproc, args, port, timeout=900):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "dist_sharded.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world,n,bsz", [(2, 20, 1), (4, 18, 2)])
def test_sharded_statevector_ranks_sharing_one_gpu(world, n, bsz):
    """SURVEY §8e, amplitude axis, on a ONE-GPU box: `world` processes share cuda:0 (gloo process group, exchanges staged
    through the host), every rank runs the real sm_100a kernels on its slab with its rank bits, and the slabs are compared
    with the same circuit (QFT + brickwork, config-5 family) on one device.  Covers the rank-index controls / diagonals
    (K11), the pack pass and the layout bookkeeping on the GPU; the NVLink paths need >= 2 GPUs (next test)."""
    out = _torchrun(world, ["--check", str(n), "--check-bsz", str(bsz), "--backend", "gloo", "--exchange", "collective"],
                    29541 + world)
    assert '"ok": true' in out, out[-1500:]


def test_sharded_adjoint_gradients_ranks_sharing_one_gpu():
    """north_star (2) + (4): adjoint gradients through the sharded device (rank-index rotations, controls, exchanges) equal
    the one-GPU adjoint; 2 processes on cuda:0 over gloo."""
    out = _torchrun(2, ["--grad", "16", "--backend", "gloo", "--exchange", "collective"], 29551)
    assert '"ok": true' in out, out[-1500:]


def test_sharded_statevector_two_gpus():
    """Amplitude-sharded device over NCCL / NVLink peer stores vs the same circuit on one GPU (needs >= 2 visible GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for backend in ("p2p", "collective"):
        out = _torchrun(2, ["--check", "22", "--exchange", backend], 29533)
        assert '"ok": true' in out and f'"exchange_backend": "{backend}"' in out, out[-1500:]

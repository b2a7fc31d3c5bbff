"""CPU suite for the host side (planner, program wire format, autograd wiring, lazy device):
the kernel SOURCE (csrc/tqb_core.cuh) is compiled for the host by tests/emu and injected behind
the same C ABI.  This is test infrastructure; the product has no CPU path (see test_abi.py)."""
import os
import subprocess

import pytest
import torch

from torchquantum_b200 import _native, engine, plan as P
from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_SRC = os.path.join(HERE, "emu", "tqb_emu.cpp")
EMU_LIB = os.path.join(HERE, "emu", "libtqb200_emu.so")


def build_emu():
    deps = [EMU_SRC, os.path.join(HERE, "..", "torchquantum_b200", "csrc", "tqb_core.cuh"),
            os.path.join(HERE, "..", "torchquantum_b200", "csrc", "tqb_program.h"),
            os.path.join(HERE, "..", "include", "tqb200.h")]
    if os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(d) for d in deps):
        return EMU_LIB
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-Wno-unknown-pragmas", "-o", EMU_LIB, EMU_SRC])
    return EMU_LIB


@pytest.fixture(autouse=True)
def emulated():
    _native.use_library_for_testing(build_emu())
    saved = (P.T_SINGLE, P.T_MULTI, P.L_MIN)
    engine._PLAN_CACHE.clear()
    yield
    P.T_SINGLE, P.T_MULTI, P.L_MIN = saved
    engine._PLAN_CACHE.clear()
    _native.use_library_for_testing(None)


def test_every_gate():
    cases.check_all_gates("cpu")


def test_unitary_gates():
    cases.check_unitary_gates("cpu")


def test_random_circuits_fwd_bwd():
    cases.check_random_circuits("cpu")


@pytest.mark.parametrize("tiles", [(4, 5, 2), (3, 6, 1), (5, 5, 3), (6, 7, 5)])
def test_random_circuits_multi_tile(tiles):
    """Force tiny tiles so that the multi-pass planner, outer-bit controls/parities and the atomic
    gradient accumulation are exercised at sizes the emulation finishes in seconds."""
    P.T_SINGLE, P.T_MULTI, P.L_MIN = tiles
    cases.check_random_circuits("cpu", min_n=tiles[0] + 1)


def test_expval_per_wire_mutates_device():
    cases.check_expval_per_wire("cpu")


def test_measure_bit_exact():
    cases.check_measure_bit_exact("cpu")


def test_mnist_cfg1():
    cases.check_mnist_cfg1("cpu")


def test_brickwork_vs_oracle_multi_tile():
    P.T_SINGLE, P.T_MULTI, P.L_MIN = 5, 6, 3
    cases.check_against_oracle("cpu", 9, 2, cases.brickwork_ops(9, 6, seed=3), grad=True)


def test_plan_cache_and_lazy_queue():
    import torchquantum_b200 as tq

    q = tq.QuantumDevice(n_wires=3, bsz=2)
    before = engine.STATS["plans_built"]
    for _ in range(3):
        q.reset_states(2)
        q.h(wires=0)
        q.cnot(wires=[0, 1])
        q.rx(wires=2, params=torch.tensor([0.3]))
        assert len(q._queue) == 3  # nothing ran yet
        s = q.get_states_1d()
        assert len(q._queue) == 0 and s.shape == (2, 8)
    assert engine.STATS["plans_built"] == before + 1
    assert q.states.shape == (2, 2, 2, 2) and q.states.is_contiguous()


def test_errors_match_reference_behaviour():
    import torchquantum_b200 as tq
    import torchquantum_b200.functional as tqf

    q = tq.QuantumDevice(n_wires=3, bsz=2)
    with pytest.raises(AssertionError):
        tqf.cnot(q, wires=[0])  # functional/gate_wrapper.py:433: matrix size vs wires
    with pytest.raises(AssertionError):
        tqf.qubitunitary(q, wires=[0], params=torch.tensor([[1.0, 1.0], [0.0, 1.0]]))  # not unitary
    tqf.rx(q, wires=0, params=torch.rand(5))  # batch mismatch: bmm raises RuntimeError in the reference
    with pytest.raises(RuntimeError):
        q.get_states_1d()
    # wire indices fail at the gate call, as the reference's permute does (gate_wrapper.py:96-140), not at the lazy flush
    q = tq.QuantumDevice(n_wires=3, bsz=2)
    with pytest.raises(IndexError, match="out of range"):
        tqf.rx(q, wires=3, params=torch.tensor([[0.1]]))
    with pytest.raises(IndexError, match="out of range"):
        q.cnot(wires=[0, -1])
    with pytest.raises(RuntimeError, match="repeated wire"):
        q.cnot(wires=[1, 1])
    with pytest.raises(IndexError, match="out of range"):
        tq.RX(has_params=True)(q, wires=7)
    with pytest.raises(IndexError, match="out of range"):
        tq.Op1QAllLayer(op=tq.RX, n_wires=4, has_params=True)(q)      # a 4-wire template on a 3-wire device
    q.h(wires=2)
    assert q.get_states_1d().shape == (2, 8)                         # nothing half-queued by the failed calls
    for template in (tq.Op1QAllLayer(op=tq.RX, n_wires=4, has_params=True), tq.EfficientSU2(n_wires=4, reps=1)):
        template(tq.QuantumDevice(n_wires=4, bsz=2))                 # records cached for a 4-wire device ...
        with pytest.raises(IndexError, match="out of range"):
            template(tq.QuantumDevice(n_wires=3, bsz=2))             # ... are not replayed on a narrower one


def test_cfg2_vqe_class_api_record_cache():
    cases.check_cfg2_vqe("cpu")


def test_cfg3_qaoa_per_sample_params():
    cases.check_cfg3_qaoa("cpu", n=6, p=2, bsz=3, edges=[(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (0, 3)])


def test_forward_only_forms():
    cases.check_forward_only_forms("cpu")


@pytest.mark.parametrize("tiles", [(4, 5, 2), (5, 6, 3)])
def test_forward_only_forms_multi_tile(tiles):
    P.T_SINGLE, P.T_MULTI, P.L_MIN = tiles
    cases.check_forward_only_forms("cpu", n=8, bsz=2)


def test_qft_phase_ladders():
    cases.check_qft_phase_ladders("cpu")


def test_qft_phase_ladders_multi_tile():
    P.T_SINGLE, P.T_MULTI, P.L_MIN = 4, 5, 2
    cases.check_qft_phase_ladders("cpu", n=8, bsz=1)


def test_vqe_hamiltonian(tmp_path):
    cases.check_vqe_hamiltonian("cpu", tmp_path)


def test_speculative_readout():
    cases.check_speculative_readout("cpu")


def test_dense_blocks():
    cases.check_dense_blocks("cpu", n=8)


def test_block_write_back_modes():
    cases.check_block_write_back_modes("cpu", n=18, depth=8)


def test_random_gates_through_blocks():
    cases.check_random_gates_through_blocks("cpu", seeds=(1,))


def test_random_catalogue_gradients():
    cases.check_random_catalogue_gradients("cpu")


@pytest.mark.parametrize("tiles", [(4, 5, 2)])
def test_random_catalogue_gradients_multi_tile(tiles):
    P.T_SINGLE, P.T_MULTI, P.L_MIN = tiles      # (restored by the autouse fixture)
    cases.check_random_catalogue_gradients("cpu", seeds=range(20, 26))


def test_large_batch_readout():
    cases.check_large_batch_readout("cpu", bsz=66000)


def test_quadphase_kernel():
    cases.check_quadphase_kernel("cpu")


def test_qft_blocks_and_quadphase():
    cases.check_qft_blocks_and_quadphase("cpu", n=18)


def test_op_hamil_exp():
    cases.check_op_hamil_exp("cpu")

"""Lowering pass vs the oracle, through the NumPy interpreter of the binary program format
(``tests/_emulator.py``): fusion, pass grouping, control handling and table fields are all
validated on the CPU; the GPU tests then only have to prove the kernels match the same format."""
import dataclasses

import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200.circuit import QuantumCircuit, flatten
from qls_b200.hhl import HHL
from qls_b200.library import QFT
from qls_b200.lower import LoweringOptions, lower_circuit
from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample, VolterraProblem

from _emulator import run_program
from _util import circuit_to_stream, random_circuit, random_unitary


def _check(qc, opts, fused=True, tol=1e-12):
    want = numpy_sv.run(qc.num_qubits, circuit_to_stream(qc))
    low = lower_circuit(qc, opts, fused=fused)
    got = run_program(low)
    assert np.max(np.abs(got - want)) < tol
    return low


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("fused", [False, True])
def test_random_circuits(seed, fused):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3, 11))
    qc = random_circuit(n, 40, rng, max_k=4)
    opts = LoweringOptions(tile_bits=int(rng.integers(3, 8)), max_high=int(rng.integers(1, 4)),
                           max_fuse_qubits=int(rng.integers(1, 5)))
    _check(qc, opts, fused)


def _wake_up_circuit(n, rng):
    """Qubits leave |0> one by one, with controlled ops on sleeping and awake qubits in between."""
    from qls_b200.circuit import Gate

    qc = QuantumCircuit(n)
    order = [int(q) for q in rng.permutation(n)]
    for step, q in enumerate(order):
        qc.h(q) if rng.integers(2) else qc.ry(float(rng.uniform(-3, 3)), q)
        for _ in range(3):
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            kind = rng.integers(4)
            if kind == 0:
                qc.cx(a, b)
            elif kind == 1:
                qc.cp(float(rng.uniform(-3, 3)), a, b)
            elif kind == 2:
                qc.unitary(random_unitary(1, rng), [order[int(rng.integers(step + 1))]])
            else:
                qc.append(Gate("unitary", 1, [], random_unitary(1, rng)).control(1, int(rng.integers(2))), [a, b])
    return qc


@pytest.mark.parametrize("seed", range(16))
@pytest.mark.parametrize("form", ["sweep", "rs"])
def test_zero_qubit_tracking(seed, form):
    """zero_start: passes only enumerate the tiles in which the qubits that are still |0> are 0, and
    controls on such qubits are resolved at lowering time.  Sparse circuits (few qubits touched early)
    so that the fixed bits really occur; both pass forms."""
    rng = np.random.default_rng(400 + seed)
    n = 16 if form == "rs" else int(rng.integers(6, 12))
    qc = _wake_up_circuit(n, rng)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    opts = LoweringOptions(zero_start=True) if form == "rs" else LoweringOptions(tile_bits=int(rng.integers(3, 6)), max_high=2,
                                                                                 zero_start=True)
    low = lower_circuit(qc, opts)
    assert any(p.fixed_bits > 0 for p in low.pass_info)
    if form == "rs":
        assert any(p.rs_stages > 0 for p in low.pass_info)
    got = run_program(low, form=form)
    assert np.max(np.abs(got - want)) < 1e-12
    # same amplitudes as without the tracking
    ref = run_program(lower_circuit(qc, dataclasses.replace(opts, zero_start=False)), form=form)
    assert np.max(np.abs(got - ref)) < 1e-13


def test_fusion_invariance_and_fewer_passes():
    rng = np.random.default_rng(7)
    qc = random_circuit(9, 80, rng, max_k=2)
    opts = LoweringOptions(tile_bits=6, max_high=3)
    a = _check(qc, opts, fused=False)
    b = _check(qc, opts, fused=True)
    assert sum(p.n_ops for p in b.pass_info) < sum(p.n_ops for p in a.pass_info)
    assert np.max(np.abs(run_program(a) - run_program(b))) < 1e-13


def test_does_not_mutate_circuit():
    rng = np.random.default_rng(3)
    qc = random_circuit(5, 20, rng)
    before = [(i.operation.name, i.qubits) for i in qc.data]
    lower_circuit(qc, LoweringOptions(tile_bits=4, max_high=2))
    assert before == [(i.operation.name, i.qubits) for i in qc.data]


@pytest.mark.parametrize("m", [3, 6, 9])
def test_qft_is_few_diag_runs(m):
    qc = QuantumCircuit(m)
    qc.h(0)
    qc.compose(QFT(m, inverse=True, do_swaps=False).reverse_bits(), inplace=True)
    low = _check(qc, LoweringOptions(tile_bits=5, max_high=2))
    # m H gates, m(m-1)/2 controlled phases -> at most ~2m ops after merging the ladders
    assert sum(p.n_ops for p in low.pass_info) <= 2 * m + 2


@pytest.mark.parametrize("model_cls", [Qiskit4QubitExample, ClassiqDemoExample, lambda: VolterraProblem(2)])
@pytest.mark.parametrize("mode", ["repeat", "eig"])
def test_hhl_circuits(model_cls, mode):
    model = model_cls()
    if mode == "repeat" and model.matrix_a.shape[0] > 4:
        pytest.skip("slow in the CPU oracle")
    qc = HHL(power_mode=mode).construct_circuit(model.matrix_a, model.vector_b)
    low = _check(qc, LoweringOptions(tile_bits=6, max_high=3))
    # ExactReciprocal (UCRY + 2 controlled UCRY) must collapse to a single multiplexed-RY block
    from qls_b200 import _abi
    kinds = [low.ops[i].kind for i in range(sum(p.n_ops for p in low.pass_info))]
    assert kinds.count(_abi.QLS_OP_MUXRY) == 1


def test_fp32_program():
    rng = np.random.default_rng(11)
    qc = random_circuit(8, 40, rng)
    want = numpy_sv.run(8, circuit_to_stream(qc))
    low = lower_circuit(qc, LoweringOptions(dtype="fp32", tile_bits=6, max_high=3))
    got = run_program(low)
    assert got.dtype == np.complex64
    assert np.max(np.abs(got - want)) < 1e-5


def test_pass_level_control_skips_tiles():
    # controlled blocks whose (shared) control is outside the tile touch only half the tiles
    rng = np.random.default_rng(5)
    from _util import random_unitary
    from qls_b200.circuit import Gate
    n = 10
    qc = QuantumCircuit(n)
    qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1), [9, 0, 1])
    qc.append(Gate("unitary", 1, [], random_unitary(1, rng)).control(1), [9, 3])
    psi0 = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi0 /= np.linalg.norm(psi0)
    want = numpy_sv.run(n, circuit_to_stream(qc), psi0=psi0)
    low = lower_circuit(qc, LoweringOptions(tile_bits=5, max_high=2))
    assert np.max(np.abs(run_program(low, psi0) - want)) < 1e-12
    assert low.n_passes == 1 and low.pass_info[0].fixed_bits == 1
    assert low.pass_bytes(0) == 2 * 16 * 2**9
    assert low.amp_updates() == 2 * 2**9


def test_sharded_index_offset_semantics():
    # rank bits enter only through index_offset: run the two halves of an 8-qubit state as two
    # 7-qubit shards for a circuit whose top qubit is only used as control / in diagonals
    rng = np.random.default_rng(9)
    from _util import random_unitary
    from qls_b200.circuit import Gate
    n = 8
    qc = QuantumCircuit(n)
    for q in range(n - 1):
        qc.h(q)
        qc.cp(0.3 * (q + 1), n - 1, q)
    qc.append(Gate("unitary", 2, [], random_unitary(2, rng)).control(1), [n - 1, 2, 5])
    qc.rz(0.7, n - 1)
    psi0 = rng.standard_normal(2**n) + 1j * rng.standard_normal(2**n)
    psi0 /= np.linalg.norm(psi0)
    want = numpy_sv.run(n, circuit_to_stream(qc), psi0=psi0)
    low = lower_circuit(qc, LoweringOptions(tile_bits=5, max_high=2, n_local=n - 1))
    half = 2 ** (n - 1)
    got = np.concatenate([run_program(low, psi0[:half], 0), run_program(low, psi0[half:], half)])
    assert np.max(np.abs(got - want)) < 1e-12


# ---- register-stage form (csrc/rs_pass.cu) ---------------------------------------------------


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("dtype", ["fp64", "fp32"])
def test_register_stage_form_matches_oracle(seed, dtype):
    rng = np.random.default_rng(900 + seed)
    T = 12 if dtype == "fp64" else 13
    n = T + int(rng.integers(0, 3))
    qc = random_circuit(n, 40, rng, max_k=2)
    want = numpy_sv.run(n, circuit_to_stream(qc))
    low = lower_circuit(qc, LoweringOptions(dtype=dtype, max_fuse_qubits=2, max_high=int(rng.integers(2, 6))))
    assert any(p.rs_stages > 0 for p in low.pass_info)
    got = run_program(low, form="rs")
    tol = 1e-12 if dtype == "fp64" else 1e-5
    assert np.max(np.abs(got - want)) < tol
    assert np.max(np.abs(run_program(low, form="sweep") - want)) < tol


def test_register_stage_form_of_hhl_shaped_circuit():
    from qls_b200.synthetic import hhl_shaped_circuit

    qc = hhl_shaped_circuit(14)
    want = numpy_sv.run(14, circuit_to_stream(qc))
    low = lower_circuit(qc, LoweringOptions(max_fuse_qubits=2))
    assert any(p.rs_stages > 0 for p in low.pass_info)
    assert np.max(np.abs(run_program(low, form="rs") - want)) < 1e-12
    # at full size every pass of the benchmark circuit has the register-stage form
    big = lower_circuit(hhl_shaped_circuit(33), LoweringOptions(max_fuse_qubits=2))
    assert all(p.rs_stages > 0 for p in big.pass_info)


def test_wide_dense_blocks_keep_the_sweep_form():
    rng = np.random.default_rng(4)
    from _util import random_unitary
    qc = QuantumCircuit(13)
    qc.h(3)
    qc.unitary(random_unitary(3, rng), [1, 5, 12])
    low = lower_circuit(qc, LoweringOptions())
    assert all(p.rs_stages == 0 for p in low.pass_info)


# ---- edge cases ------------------------------------------------------------------------------------


def _edge_circuits():
    from qls_b200.circuit import Gate
    out = {}
    qc = QuantumCircuit(3, name="empty")
    out["empty"] = qc
    qc = QuantumCircuit(1, name="one_qubit")
    qc.h(0)
    qc.t(0)
    qc.h(0)
    out["one_qubit"] = qc
    qc = QuantumCircuit(4, name="diag_only", global_phase=0.3)
    for q in range(4):
        qc.rz(0.1 * (q + 1), q)
    qc.cp(0.7, 0, 3)
    qc.cz(1, 2)
    out["diag_only"] = qc
    qc = QuantumCircuit(4, name="cancelling")
    for q in range(4):
        qc.h(q)
        qc.h(q)
    qc.cx(0, 1)
    qc.cx(0, 1)
    out["cancelling"] = qc
    qc = QuantumCircuit(2, name="phase_only", global_phase=1.234)
    out["phase_only"] = qc
    qc = QuantumCircuit(5, name="negative_controls")
    for q in range(5):
        qc.h(q)
    qc.append(Gate("x", 1).control(2, ctrl_state=0b01), [0, 3, 4])
    qc.append(Gate("ry", 1, [0.9]).control(1, ctrl_state=0), [2, 1])
    out["negative_controls"] = qc
    qc = QuantumCircuit(6, name="prep_then_gates")
    amps = np.arange(1, 9, dtype=complex)
    amps[3] *= 1j
    amps /= np.linalg.norm(amps)
    qc.prepare_state(amps, [0, 1, 2])
    qc.h(4)
    qc.cx(4, 0)
    out["prep_then_gates"] = qc
    return out


@pytest.mark.parametrize("name", sorted(_edge_circuits()))
def test_edge_case_circuits(name):
    qc = _edge_circuits()[name]
    n = qc.num_qubits
    want = numpy_sv.run(n, circuit_to_stream(qc))
    for fused in (True, False):
        low = lower_circuit(qc, LoweringOptions(tile_bits=3, max_high=1), fused=fused)
        got = run_program(low)
        assert np.max(np.abs(got - want)) < 1e-14, (name, fused)
    assert abs(np.linalg.norm(want) - 1) < 1e-14

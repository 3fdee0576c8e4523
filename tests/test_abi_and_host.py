"""CPU-side checks: the C-ABI library loads and exports every symbol include/qls_abi.h declares
(no compute calls without a GPU), struct layouts match, golden fixtures match the oracle, and the
host-side mirror of the reference interface behaves like the reference's own unit tests
(tests/test_utils.py, tests/test_toymodels.py, tests/test_quantum_linear_solver.py,
tests/test_vqls_implementations.py:39-58 in the reference tree)."""
import os
import re

import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200 import _abi

from _util import circuit_to_stream, random_circuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    if not os.path.exists(_abi.lib_path()):
        import __graft_entry__ as g

        g.build()


def test_library_exports_every_declared_symbol():
    _built()
    header = open(os.path.join(ROOT, "include", "qls_abi.h")).read()
    declared = set(re.findall(r"\b(qls_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    lib = _abi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.qls_abi_version() == 1


def test_struct_sizes_match_header():
    import ctypes as C

    assert C.sizeof(_abi.QlsOp) == 128 and C.sizeof(_abi.QlsPass) == 128 and C.sizeof(_abi.QlsSeg) == 4
    assert C.sizeof(_abi.QlsRsOp) == 192


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qls_b200.circuit import QuantumCircuit
    from qls_b200.statevector import Statevector

    qc = QuantumCircuit(2)
    qc.h(0)
    with pytest.raises(_abi.QlsError):
        Statevector(qc)


def test_golden_random10_matches_oracle():
    qc = random_circuit(10, 80, np.random.default_rng(424242), max_k=4)
    want = np.load(os.path.join(ROOT, "tests", "golden", "random10_state.npy"))
    assert np.max(np.abs(numpy_sv.run(10, circuit_to_stream(qc)) - want)) < 1e-14


def test_golden_kat_is_appendix_b():
    v = np.load(os.path.join(ROOT, "tests", "golden", "hhl_2x2_state.npy"))
    want = np.zeros(32, dtype=complex)
    want[0], want[1], want[16], want[17] = np.sqrt(3) / 4, -np.sqrt(3) / 4, 0.75, 0.25
    assert np.max(np.abs(v - want)) < 1e-14


# ---- host mirror of the reference interface ---------------------------------------------------


def test_utils_hermitian_embedding_and_expansion():
    from qls_b200.utils import (expand_b_vector, extract_hhl_solution_vector_from_state_vector,
                                extract_x_from_expanded, make_matrix_hermitian)

    vector = np.array([[7], [8]])
    for mat in (np.array([[1, 2, 3], [4, 5, 6]]), np.array([[1, 2], [4, 5]])):
        herm = make_matrix_hermitian(mat)
        assert np.array_equal(herm, herm.conj().T)
        ev = expand_b_vector(vector, non_square_matrix=mat)
        assert ev.shape == (mat.shape[0] + mat.shape[1],)
        assert np.array_equal(ev, np.array([7, 8] + [0] * mat.shape[1]))
    assert np.array_equal(extract_x_from_expanded(expand_b_vector(vector)), np.zeros(2))
    assert np.array_equal(extract_x_from_expanded(np.array([0, 0, 1, 1])), np.array([1, 1]))
    herm = make_matrix_hermitian(np.array([[1, 2, 3], [4, 5, 6]]))
    sol = extract_hhl_solution_vector_from_state_vector(herm, np.array([0, 0, 1, 0, 0, 0, 1, 0]))
    assert sol.shape == (4,)  # reference tests/test_utils.py:52-59 pins the slice start 2**(n-1)


def test_toymodel_sizes():
    from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample, VolterraProblem

    assert Qiskit4QubitExample().num_qubits == 1
    assert VolterraProblem(2).num_qubits == 3
    assert ClassiqDemoExample().num_qubits == 2
    a = VolterraProblem.volterra_a_matrix(4, 0.125)
    want = np.array([[1, 0, 0, 0], [0.125, 1.125, 0, 0], [0.125, 0.25, 1.125, 0], [0.125, 0.25, 0.25, 1.125]])
    assert np.array_equal(a, want)


def test_facade_checks():
    from qls_b200.quantum_linear_solver import QuantumLinearSolver as Q

    assert Q.check_matrix_condition_number(np.eye(3)) == pytest.approx(1.0)
    assert Q.check_matrix_condition_number(np.diag([2.0, 3.0, 5.0])) == pytest.approx(2.5)
    assert Q.check_matrix_sparsity(np.zeros((3, 3))) == 1.0
    assert Q.check_matrix_sparsity(np.diag([1.0, 2.0, 3.0])) == pytest.approx(6 / 9)
    with pytest.raises(ValueError):
        Q.check_matrix_square_hermitian(np.ones((2, 3)))
    with pytest.raises(ValueError):
        Q.check_matrix_square_hermitian(np.array([[1, 2], [3, 4]]))
    with pytest.raises(NotImplementedError):
        Q().solve(np.eye(2), np.ones(2), method="nope", file_basename="x")


def test_qasm3_round_trip():
    from qls_b200 import qasm3
    from qls_b200.hhl import HHL
    from qls_b200.toymodels import ClassiqDemoExample

    m = ClassiqDemoExample()
    qc = HHL().construct_circuit(m.matrix_a, m.vector_b)
    for circ in (qc, qc.decompose(reps=10), random_circuit(6, 40, np.random.default_rng(1))):
        text = qasm3.dumps(circ)
        back = qasm3.loads(text)
        assert qasm3.dumps(back) == text
        a = numpy_sv.run(circ.num_qubits, circuit_to_stream(circ))
        b = numpy_sv.run(back.num_qubits, circuit_to_stream(back))
        assert np.max(np.abs(a - b)) < 1e-14

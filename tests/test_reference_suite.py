"""The reference's own unit tests (``/root/reference/tests``), restated one to one against this package: same test
names, same inputs, same expectations.  The two tests that simulate circuits (``test_hhl_implementations.py:23-29``
and ``test_vqls_implementations.py:26-37``) need the GPU and live in ``tests/test_gpu_parity.py``
(``test_kat_hhl_2x2_through_the_facade``, ``test_vqls_2x2_through_the_facade``); the Classiq test needs an online
account (out of scope)."""
import numpy as np

from qls_b200.implementations.vqls_qiskit_implementation import postprocess_solution
from qls_b200.quantum_linear_solver import QuantumLinearSolver
from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample, ScalingTestModel, VolterraProblem
from qls_b200.utils import (expand_b_vector, extract_hhl_solution_vector_from_state_vector, extract_x_from_expanded,
                            make_matrix_hermitian)

MATRIX = np.array([[1, 2, 3], [4, 5, 6]])
SQUARE = np.array([[1, 2], [4, 5]])
VECTOR = np.array([[7], [8]])


# ---- tests/test_utils.py ------------------------------------------------------------------------
def test_make_matrix_hermitian():  # test_utils.py:23-27
    for mat in (MATRIX, SQUARE):
        h = make_matrix_hermitian(mat)
        assert np.array_equal(h, h.conj().T)


def test_expand_b_vector():  # test_utils.py:29-39
    for mat in (MATRIX, SQUARE):
        v = expand_b_vector(VECTOR, non_square_matrix=mat)
        assert v.shape == (mat.shape[0] + mat.shape[1],)
        assert np.array_equal(v, np.array([7, 8] + mat.shape[1] * [0]))


def test_extract_x_from_expanded():  # test_utils.py:41-50
    assert np.array_equal(extract_x_from_expanded(expand_b_vector(VECTOR)), np.zeros(len(VECTOR)))
    assert np.array_equal(extract_x_from_expanded(np.array([0, 0, 1, 1])), np.array([1, 1]))


def test_extract_hhl_solution_vector_from_state_vector():  # test_utils.py:52-59
    h = make_matrix_hermitian(MATRIX)
    x = extract_hhl_solution_vector_from_state_vector(h, np.array([0, 0, 1, 0, 0, 0, 0, 0]))
    assert x.shape == (4,)


# ---- tests/test_quantum_linear_solver.py ---------------------------------------------------------
def test_check_matrix_condition_number():  # test_quantum_linear_solver.py:9-19
    qls = QuantumLinearSolver()
    assert qls.check_matrix_condition_number(np.eye(3)) == 1.0
    assert qls.check_matrix_condition_number(np.diag([2, 3, 5])) == 2.5


def test_check_matrix_sparsity():  # test_quantum_linear_solver.py:21-31
    qls = QuantumLinearSolver()
    assert qls.check_matrix_sparsity(np.zeros((3, 3))) == 1.0
    assert qls.check_matrix_sparsity(np.eye(3)) == 6 / 9


# ---- tests/test_toymodels.py ---------------------------------------------------------------------
def test_default_values():  # test_toymodels.py:14-26
    assert Qiskit4QubitExample().num_qubits == 1
    assert VolterraProblem(2).num_qubits == 2 + 1  # made hermitian: one more qubit
    assert ClassiqDemoExample().num_qubits == 2
    assert ScalingTestModel().num_qubits == 3


# ---- tests/test_vqls_implementations.py ----------------------------------------------------------
def test_postprocessing():  # test_vqls_implementations.py:39-58
    rng = np.random.default_rng(0)
    a, b = rng.random((2, 2)), rng.random(2)
    true = np.linalg.solve(a, b)
    unit = true / np.linalg.norm(true)
    assert not np.allclose(true, unit)
    assert np.allclose(true, postprocess_solution(a, b, unit))
    assert not np.allclose(true, -unit)
    assert np.allclose(true, postprocess_solution(a, b, -unit))

"""Pin the oracle and the host-side HHL builder (CPU only).

The reference's only exact known-answer test is ``tests/test_hhl_implementations.py:23-29`` in the
reference tree: ``Qiskit4QubitExample`` -> circuit width 5 and x = [1.125, 0.375] at
``np.allclose`` defaults.  SURVEY.md Appendix B derives the full final state for it."""
import math

import numpy as np
import pytest

from oracle import hhl_analytic, kron_unitary, numpy_sv
from qls_b200.circuit import flatten
from qls_b200.hhl import HHL
from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample, VolterraProblem
from qls_b200.utils import extract_hhl_solution_vector_from_state_vector

from _util import circuit_to_stream, random_circuit


def _oracle_state(qc):
    return numpy_sv.run(qc.num_qubits, circuit_to_stream(qc))


def test_kat_2x2_width_and_solution():
    model = Qiskit4QubitExample()
    hhl = HHL(power_mode="repeat")
    qc = hhl.construct_circuit(model.matrix_a, model.vector_b)
    assert qc.width() == 5  # tests/test_hhl_implementations.py:28
    assert (hhl.nb, hhl.nl) == (1, 3)
    assert hhl.delta == 0.25
    assert math.isclose(hhl.evolution_time, 3 * math.pi / 8, rel_tol=1e-15)
    psi = _oracle_state(qc)
    # exactly four non-zero amplitudes
    want = np.zeros(32, dtype=complex)
    want[0], want[1], want[16], want[17] = math.sqrt(3) / 4, -math.sqrt(3) / 4, 0.75, 0.25
    assert np.max(np.abs(psi - want)) < 1e-14
    norm = numpy_sv.hhl_norm(psi, 2, hhl.scaling)
    assert math.isclose(norm, 1.1858541225631416, rel_tol=1e-14)
    x = norm * extract_hhl_solution_vector_from_state_vector(model.matrix_a, psi)
    assert np.allclose(model.classical_solution, x)  # tests/test_hhl_implementations.py:29
    assert np.max(np.abs(x - np.array([1.125, 0.375]))) < 1e-13


@pytest.mark.parametrize("mode", ["repeat", "eig"])
@pytest.mark.parametrize("model_cls", [Qiskit4QubitExample, ClassiqDemoExample, lambda: VolterraProblem(2)])
def test_gate_stream_matches_closed_form(model_cls, mode):
    model = model_cls()
    if mode == "repeat" and model.matrix_a.shape[0] > 4:
        pytest.skip("literal repetition is slow in the CPU oracle")
    hhl = HHL(power_mode=mode)
    qc = hhl.construct_circuit(model.matrix_a, model.vector_b)
    psi = _oracle_state(qc)
    ref = hhl_analytic.hhl_state(model.matrix_a, model.vector_b)
    assert psi.shape == ref.shape
    assert np.max(np.abs(psi - ref)) < 1e-12
    nb, nl, delta, t, lmin = hhl_analytic.hhl_parameters(model.matrix_a)
    assert (hhl.nb, hhl.nl) == (nb, nl)
    assert hhl.delta == delta and math.isclose(hhl.evolution_time, t, rel_tol=1e-15)


def test_toy_parameters_match_survey():
    # ClassiqDemo nb=2, nl=4, delta=0.125; Volterra(2) nb=3, nl=5, delta=0.625
    h = HHL()
    m = ClassiqDemoExample()
    qc = h.construct_circuit(m.matrix_a, m.vector_b)
    assert (h.nb, h.nl, h.delta, qc.width()) == (2, 4, 0.125, 7)
    m = VolterraProblem(2)
    qc = h.construct_circuit(m.matrix_a, m.vector_b)
    assert (h.nb, h.nl, h.delta, qc.width()) == (3, 5, 0.625, 9)


def test_solution_quality_within_reference_gate():
    # plotting.py:136-143 rejects > 20 % relative distance; SURVEY quotes ~1.1 % / ~5.8 %
    for model, bound in ((ClassiqDemoExample(), 3.0), (VolterraProblem(2), 8.0)):
        x = hhl_analytic.hhl_solution(model.matrix_a, model.vector_b)
        if model.name.startswith("Volterra"):
            x = x[len(x) // 2:]
        cs = model.classical_solution
        xn = x / np.linalg.norm(x) * np.linalg.norm(cs)
        rel = np.linalg.norm(cs - xn) / np.linalg.norm(cs) * 100
        assert rel < bound, (model.name, rel)


@pytest.mark.parametrize("matrix, vector", [
    ([[0.0, 1.0], [1.0, 0.0]], [1.0, 0.0]),   # eigenvalues +-1: x = [0, 1]
    ([[1.0, 2.0], [2.0, 1.0]], [1.0, 0.0]),   # eigenvalues 3, -1: x = [-1/3, 2/3]
    ([[1.0, 2.0], [2.0, 1.0]], [0.6, -0.8]),
])
def test_negative_eigenvalues_solve_the_system_passed_in(matrix, vector):
    """``ExactReciprocal(neg_vals=True)``: with the sign qubit set, QPE holds the eigenvalue in two's
    complement and the second controlled UCRY rotates by ``scaling / -(1 - i/nl)``.  Both eigenvalues of
    these matrices are exactly representable on the clock register, so HHL is exact: compare with
    ``np.linalg.solve`` of the very system that was passed in (no model-side rescaling)."""
    a, b = np.asarray(matrix), np.asarray(vector)
    hhl = HHL(power_mode="repeat")
    qc = hhl.construct_circuit(a, b)
    psi = _oracle_state(qc)
    assert np.max(np.abs(psi - hhl_analytic.hhl_state(a, b))) < 1e-12
    x = numpy_sv.hhl_norm(psi, 2, hhl.scaling) * extract_hhl_solution_vector_from_state_vector(a, psi)
    assert np.max(np.abs(x - np.linalg.solve(a, b / np.linalg.norm(b)))) < 1e-12


@pytest.mark.parametrize("seed", range(6))
def test_evolve_operator_matches_dense_unitary(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 8))
    qc = random_circuit(n, 25, rng)
    stream = circuit_to_stream(qc)
    a = numpy_sv.run(n, stream)
    b = kron_unitary.run(n, stream)
    assert np.max(np.abs(a - b)) < 1e-12
    assert abs(np.linalg.norm(a) - 1) < 1e-12


def test_little_endian_convention():
    # X on qubit 0 of |00> gives index 1; CX(control 0 -> target 1) then gives index 3
    psi = numpy_sv.run(2, [("u", np.array([[0, 1], [1, 0]]), [0])])
    assert abs(psi[1] - 1) < 1e-15
    cx = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]])
    psi = numpy_sv.run(2, [("u", np.array([[0, 1], [1, 0]]), [0]), ("u", cx, [0, 1])])
    assert abs(psi[3] - 1) < 1e-15
    assert numpy_sv.expectation_z(psi, 0) == -1.0

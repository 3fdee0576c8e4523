"""Host-side VQLS logic on the CPU: the Estimator is replaced by a stand-in that evaluates the
Hadamard-test circuits with the ORACLE, so the restated algorithm (decomposition, Hadamard tests,
cost assembly, optimiser, post-processing) is checked without a GPU; the batched program format is
checked through the emulator."""
import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200.circuit import QuantumCircuit
from qls_b200.library import RealAmplitudes
from qls_b200.lower import LoweringOptions, lower_circuit
from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample
from qls_b200.vqls import COBYLA, VQLS, HadamardTest, symmetric_decomposition

from _emulator import run_program
from _util import circuit_to_stream, random_unitary


class OracleEstimator:
    """BaseEstimator-shaped stand-in: bind, simulate with the oracle, <Z_q>."""

    def run(self, circuits, observables, parameter_values=None, **_):
        vals = []
        for i, (c, ob) in enumerate(zip(circuits, observables)):
            bound = c.assign_parameters(parameter_values[i]) if c.num_parameters else c
            psi = numpy_sv.run(bound.num_qubits, circuit_to_stream(bound))
            vals.append(numpy_sv.expectation_z(psi, ob[1]))

        class _R:
            values = np.array(vals)

        class _J:
            @staticmethod
            def result():
                return _R

        return _J


def test_symmetric_decomposition_reconstructs_a():
    a = ClassiqDemoExample().matrix_a
    terms = symmetric_decomposition(a)
    assert len(terms) == 2
    rec = sum(t.coeff * t.matrix for t in terms)
    assert np.allclose(rec, a, atol=1e-12)
    for t in terms:
        assert np.allclose(t.matrix.conj().T @ t.matrix, np.eye(4), atol=1e-10)


def test_hadamard_test_measures_real_and_imag():
    rng = np.random.default_rng(0)
    u = random_unitary(2, rng)
    v = random_unitary(2, rng)
    state = QuantumCircuit(2)
    state.unitary(v, [0, 1])
    op = QuantumCircuit(2)
    op.unitary(u, [0, 1])
    val = HadamardTest([op], apply_initial_state=state).get_value(OracleEstimator(), [])
    psi = v[:, 0]
    assert abs(val - np.vdot(psi, u @ psi)) < 1e-12


def test_cost_is_zero_at_the_solution_and_matches_closed_form():
    model = Qiskit4QubitExample()
    ansatz = RealAmplitudes(1, reps=3)
    vq = VQLS(OracleEstimator(), ansatz, COBYLA(maxiter=1))
    norm_t, over_t = vq.construct_circuit(model.matrix_a, model.vector_b)
    cm = vq.get_coefficient_matrix(np.array([t.coeff for t in vq.matrix_terms]))
    cost = vq.get_cost_evaluation_function(norm_t, over_t, cm)
    x = np.linalg.solve(model.matrix_a, model.vector_b)
    x /= np.linalg.norm(x)
    theta_sol = [2 * np.arctan2(x[1], x[0]), 0, 0, 0]  # RY(t)|0> = (cos t/2, sin t/2)
    assert abs(cost(theta_sol)) < 1e-12
    for th in ([0.3, -1.0, 0.7, 0.2], [2.0, 0.1, -0.4, 1.1]):
        psi = numpy_sv.run(1, circuit_to_stream(ansatz.assign_parameters(th)))
        ax = model.matrix_a @ psi
        want = 1 - abs(np.vdot(model.vector_b / np.linalg.norm(model.vector_b), ax)) ** 2 / np.vdot(ax, ax).real
        assert abs(cost(th) - want) < 1e-12


def test_vqls_solves_2x2_like_the_reference_test():
    """Reference tests/test_vqls_implementations.py:26-37: 1-qubit RealAmplitudes(reps=3), COBYLA
    <= 250 iterations, width 1, atol 1e-3."""
    from qls_b200.implementations.vqls_qiskit_implementation import postprocess_solution

    model = Qiskit4QubitExample()
    ansatz = RealAmplitudes(num_qubits=1, entanglement="full", reps=3)
    res = VQLS(OracleEstimator(), ansatz, COBYLA(maxiter=250), seed=3).solve(model.matrix_a, model.vector_b)
    assert res.state.width() == 1
    x = np.real(numpy_sv.run(1, circuit_to_stream(res.state)))
    sol = postprocess_solution(model.matrix_a, model.vector_b, x)
    assert np.allclose(model.classical_solution, sol, atol=1e-3)


def test_postprocessing_scale_and_sign():
    """Reference tests/test_vqls_implementations.py:39-58 (seeded here)."""
    from qls_b200.implementations.vqls_qiskit_implementation import postprocess_solution

    rng = np.random.default_rng(12)
    a, b = rng.random((2, 2)), rng.random(2)
    true = np.linalg.solve(a, b)
    sol = true / np.linalg.norm(true)
    assert not np.allclose(true, sol)
    assert np.allclose(true, postprocess_solution(a, b, sol))
    assert np.allclose(true, postprocess_solution(a, b, -sol))


@pytest.mark.parametrize("nq", [1, 2, 3])
def test_batched_program_format_matches_oracle(nq):
    """Hadamard-test circuits lowered in batch mode (PARAM rotations, controlled ansatz, MATVEC or
    dense operator blocks) and run through the emulator vs bind + oracle."""
    rng = np.random.default_rng(nq)
    dim = 2**nq
    a = rng.standard_normal((dim, dim))
    a = a + a.T
    b = rng.standard_normal(dim)
    ansatz = RealAmplitudes(nq, reps=2)
    vq = VQLS(None, ansatz, None)
    norm_t, over_t = vq.construct_circuit(a, b)
    theta = rng.uniform(-np.pi, np.pi, ansatz.num_parameters)
    for test in norm_t + over_t:
        for circ in test.circuits:
            low = lower_circuit(circ, LoweringOptions(batch=True))
            assert low.n_passes == 1
            got = run_program(low, params=theta)
            want = numpy_sv.run(circ.num_qubits, circuit_to_stream(circ.assign_parameters(theta)))
            assert np.max(np.abs(got - want)) < 1e-12


def test_batched_program_with_wide_operator_uses_matvec():
    from qls_b200 import _abi

    rng = np.random.default_rng(5)
    nq = 7
    ansatz = RealAmplitudes(nq, reps=1)
    op = QuantumCircuit(nq)
    op.unitary(random_unitary(nq, rng), list(range(nq)))
    test = HadamardTest([op], apply_initial_state=ansatz)
    theta = rng.uniform(-np.pi, np.pi, ansatz.num_parameters)
    low = lower_circuit(test.circuits[1], LoweringOptions(batch=True))
    kinds = [low.ops[i].kind for i in range(low.pass_info[0].n_ops)]
    assert _abi.QLS_OP_MATVEC in kinds
    got = run_program(low, params=theta)
    want = numpy_sv.run(nq + 1, circuit_to_stream(test.circuits[1].assign_parameters(theta)))
    assert np.max(np.abs(got - want)) < 1e-12


def test_batched_lowering_uses_still_zero_qubits():
    """Batched circuits start from |0...0>: a block that does not use a qubit nothing has acted on yet skips the amplitudes
    with that bit set (a control on value 0), and a block controlled on such a qubit is the identity and is dropped.  Same
    state as bind + oracle."""
    rng = np.random.default_rng(9)
    qc = QuantumCircuit(5)
    qc.ry(0.3, 1)
    qc.cx(4, 2)            # control qubit 4 is still |0>: dropped
    qc.cx(1, 2)
    qc.unitary(random_unitary(2, rng), [1, 3])
    qc.h(0)                # the ancilla wakes up last
    qc.cx(0, 4)
    low = lower_circuit(qc, LoweringOptions(batch=True, max_fuse_qubits=1))
    ops = [low.ops[i] for i in range(low.pass_info[0].n_ops)]
    assert len(ops) == 5                                          # cx(4, 2) is gone
    assert ops[0].lc_mask == 0b11101 and ops[0].lc_val == 0        # ry on qubit 1: every other qubit is still |0>
    assert ops[1].lc_mask & 0b10001 == 0b10001 and (ops[1].lc_val & 0b10001) == 0  # cx(1, 2): qubits 0 and 4 still |0>
    assert ops[-1].lc_mask == 0b00001 and ops[-1].lc_val == 1      # cx(0, 4): only its real control is left
    got = run_program(low)
    want = numpy_sv.run(5, circuit_to_stream(qc))
    assert np.max(np.abs(got - want)) < 1e-14

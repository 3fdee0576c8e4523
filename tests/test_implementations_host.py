"""Host logic of ``solve_hhl_qiskit`` / ``solve_vqls_qiskit`` on the CPU: the engine's ``Statevector`` (and the
Estimator) are replaced by stand-ins that evolve the circuit with the ORACLE, so signature, return tuple, side
effects, prints and error behaviour of the two entry points are checked without a GPU.  The same calls with the
real engine are in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from oracle import numpy_sv
from qls_b200.circuit import QuantumCircuit
from qls_b200.library import RealAmplitudes
from qls_b200.toymodels import ClassiqDemoExample, Qiskit4QubitExample, VolterraProblem

from _util import circuit_to_stream
from test_vqls_host import OracleEstimator


class OracleStatevector:
    """The members of qls_b200.statevector.Statevector that the two entry points and HHL._calculate_norm use."""

    def __init__(self, data, precision="fp64", device=None, **_):
        assert isinstance(data, QuantumCircuit)
        self.precision, self.num_qubits = precision, data.num_qubits
        self._psi = numpy_sv.run(data.num_qubits, circuit_to_stream(data))

    dim = property(lambda self: 2**self.num_qubits)
    data = property(lambda self: self._psi)

    def read_slice(self, offset, count):
        return self._psi[offset:offset + count]

    def norm_sq_range(self, offset, count):
        return float(np.sum(np.abs(self._psi[offset:offset + count]) ** 2))


@pytest.fixture
def oracle_engine(monkeypatch, tmp_path):
    import qls_b200.statevector as sv_mod

    monkeypatch.setattr(sv_mod, "Statevector", OracleStatevector)
    import qls_b200.implementations.hhl_qiskit_implementation as hhl_impl

    monkeypatch.setattr(hhl_impl, "Statevector", OracleStatevector)
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_solve_hhl_qiskit_kat(oracle_engine, capsys):
    """Reference tests/test_hhl_implementations.py:23-29."""
    from qls_b200.implementations.hhl_qiskit_implementation import solve_hhl_qiskit

    model = Qiskit4QubitExample()
    a0, b0 = model.matrix_a.copy(), model.vector_b.copy()
    x, qasm, depth, width, seconds = solve_hhl_qiskit(model.matrix_a, model.vector_b, show_circuit=False)
    assert width == 5 and depth > 0 and seconds > 0
    assert np.allclose(model.classical_solution, x) and np.max(np.abs(x - [1.125, 0.375])) < 1e-12
    assert qasm.startswith("OPENQASM 3") and open("hhl_qiskit_circuit.qasm3").read() == qasm
    assert np.array_equal(a0, model.matrix_a) and np.array_equal(b0, model.vector_b)  # inputs are not mutated
    out = capsys.readouterr().out
    assert "Size of solution register is 1 , QPE registers is 3." in out and "Comparing depths original" in out


@pytest.mark.parametrize("model", [ClassiqDemoExample(), VolterraProblem(2)], ids=["classiq_demo", "volterra2"])
def test_solve_hhl_qiskit_toy_models(oracle_engine, model):
    """Solution quality inside the reference's 20 % gate (plotting.py:136-143); Volterra is Hermitian-embedded, so
    the zero half of the expanded solution must be cut off."""
    from qls_b200.implementations.hhl_qiskit_implementation import solve_hhl_qiskit
    from qls_b200.utils import relative_distance_quantum_classical_solution

    x, _, _, width, _ = solve_hhl_qiskit(model.matrix_a, model.vector_b)
    assert x.shape == np.asarray(model.classical_solution).reshape(-1).shape
    # the reference's gate normalises both vectors first (plotting.py:123-124, 136-143); ToyModel.normalize_model
    # also rescales classical_solution although A and b are divided by the same factor (toymodels.py:59-64)
    cs = np.asarray(model.classical_solution).reshape(-1)
    assert relative_distance_quantum_classical_solution(x / np.linalg.norm(x), cs / np.linalg.norm(cs)) < 20.0
    assert width >= 2 * model.num_qubits + 2


def test_solve_vqls_qiskit_kat(oracle_engine):
    """Reference tests/test_vqls_implementations.py:26-37."""
    from qls_b200.implementations.vqls_qiskit_implementation import solve_vqls_qiskit

    model = Qiskit4QubitExample()
    ansatz = RealAmplitudes(num_qubits=1, entanglement="full", reps=3, insert_barriers=False)
    x, qasm, depth, width, seconds = solve_vqls_qiskit(model.matrix_a, model.vector_b.reshape(-1, 1), ansatz=ansatz,
                                                       estimator=OracleEstimator(), show_circuit=False, seed=3)
    assert width == 1 and depth > 0 and seconds > 0
    assert np.allclose(model.classical_solution, x, atol=1e-3)
    assert os.path.exists("vqls_qiskit_circuit.qasm3") and qasm.startswith("OPENQASM 3")


def test_solve_vqls_qiskit_rejects_unknown_optimizer(oracle_engine):
    from qls_b200.implementations.vqls_qiskit_implementation import solve_vqls_qiskit

    model = Qiskit4QubitExample()
    with pytest.raises(ValueError, match="Invalid optimizer_name"):
        solve_vqls_qiskit(model.matrix_a, model.vector_b, estimator=OracleEstimator(), optimizer_name="adam")

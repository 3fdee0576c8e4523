"""Runs only where qiskit is importable (it is not in the build image: every test here is skipped there).  On a machine
that has the reference's stack installed this pins the restated pieces against qiskit itself:

* ``qiskit.quantum_info.Statevector(circuit).data`` (the call being replaced, ``hhl_qiskit_implementation.py:49``) against
  the oracle's restatement of ``_evolve_operator`` fed through the adapter's walk of the same ``QuantumCircuit``;
* the text ``qiskit.qasm3.dumps`` writes (``utils.py:10-14``) read by ``qasm3.loads_stdgates``;
* with a GPU as well: the adapter's ``Statevector`` (evolved on the B200) against qiskit's.
"""
import numpy as np
import pytest

qiskit = pytest.importorskip("qiskit")

from oracle import numpy_sv  # noqa: E402
from qls_b200 import qasm3  # noqa: E402
from qls_b200.qiskit_adapter import from_qiskit  # noqa: E402

from _util import circuit_to_stream  # noqa: E402


def _circuit():
    from qiskit import QuantumCircuit
    from qiskit.circuit.library import QFT

    qc = QuantumCircuit(5, global_phase=0.3)
    qc.h(range(5))
    qc.cp(0.7, 0, 3)
    qc.cry(1.1, 2, 4)
    qc.append(QFT(3, inverse=True, do_swaps=False).to_gate(), [1, 2, 3])
    qc.ccx(0, 1, 4)
    qc.rz(-0.4, 2)
    qc.u(0.1, 0.2, 0.3, 0)
    return qc


def test_oracle_matches_qiskit_statevector():
    from qiskit.quantum_info import Statevector

    qc = _circuit()
    want = np.asarray(Statevector(qc).data)
    got = numpy_sv.run(qc.num_qubits, circuit_to_stream(from_qiskit(qc)))
    assert np.max(np.abs(got - want)) <= 1e-13


def test_reader_takes_qiskit_export():
    from qiskit import qasm3 as qk_qasm3
    from qiskit.quantum_info import Statevector

    qc = _circuit().decompose(reps=3)
    text = qk_qasm3.dumps(qc)
    back = qasm3.loads_stdgates(text)
    want = np.asarray(Statevector(qc).data)
    got = numpy_sv.run(back.num_qubits, circuit_to_stream(back))
    assert np.max(np.abs(got - want)) <= 1e-12


def test_hhl_reference_kat_through_qiskit_if_linear_solvers_present():
    ls = pytest.importorskip("linear_solvers")
    from qiskit.quantum_info import Statevector

    from oracle import hhl_analytic

    a = np.array([[1, -1 / 3], [-1 / 3, 1]])
    b = np.array([1.0, 0.0])
    res = ls.HHL().solve(a, b)
    want = np.asarray(Statevector(res.state).data)
    got = hhl_analytic.hhl_state(a, b)
    assert np.max(np.abs(got - want)) <= 1e-10


@pytest.mark.gpu
def test_adapter_statevector_on_the_device(cuda_device):
    from qiskit.quantum_info import Statevector as QkStatevector

    from qls_b200.qiskit_adapter import Statevector

    qc = _circuit()
    assert np.max(np.abs(Statevector(qc).data - np.asarray(QkStatevector(qc).data))) <= 1e-12

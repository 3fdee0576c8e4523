"""qiskit adapter exercised with duck-typed stand-ins for qiskit's QuantumCircuit / Instruction objects
(qiskit itself is not installable in this image): same walk as Statevector._evolve_instruction --
operation with a matrix -> leaf, otherwise recurse into .definition and pick up its global phase."""
import math

import numpy as np

from oracle import numpy_sv
from qls_b200.circuit import QuantumCircuit
from qls_b200.qiskit_adapter import from_qiskit

from _util import circuit_to_stream


class _Qubit:
    pass


class _Op:
    def __init__(self, name, matrix=None, definition=None):
        self.name, self._m, self.definition = name, matrix, definition

    def to_matrix(self):
        if self._m is None:
            raise RuntimeError("to_matrix not defined for this instruction")  # qiskit raises CircuitError here
        return self._m


class _Ins:
    def __init__(self, operation, qubits):
        self.operation, self.qubits = operation, qubits


class _Circuit:
    def __init__(self, n, global_phase=0.0, name="fake"):
        self.qubits = [_Qubit() for _ in range(n)]
        self.num_qubits, self.global_phase, self.name, self.data = n, global_phase, name, []

    def add(self, op, idx):
        self.data.append(_Ins(op, [self.qubits[i] for i in idx]))
        return self


H = np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
CX = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)  # qargs[0] = control = LSB


def _rz(t):
    return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])


def test_from_qiskit_walks_definitions_and_phases():
    inner = _Circuit(2, global_phase=0.3).add(_Op("cx", CX), [1, 0]).add(_Op("rz", _rz(0.7)), [0])
    mid = _Circuit(3, global_phase=-0.1).add(_Op("inner", definition=inner), [2, 0]).add(_Op("barrier"), [0, 1, 2]) \
        .add(_Op("h", H), [1])
    top = _Circuit(4, global_phase=0.25).add(_Op("h", H), [3]).add(_Op("mid", definition=mid), [3, 1, 0]) \
        .add(_Op("mid", definition=mid), [0, 2, 3])
    got = from_qiskit(top)

    ref = QuantumCircuit(4, global_phase=0.25 + 2 * (-0.1 + 0.3))
    ref.h(3)
    for a, b, c in ((3, 1, 0), (0, 2, 3)):  # mid's qubits 0, 1, 2
        ref.cx(a, c)  # inner on (mid 2, mid 0): cx with control = inner qubit 1 = mid 0, target = inner 0 = mid 2
        ref.rz(0.7, c)
        ref.h(b)
    a = numpy_sv.run(4, circuit_to_stream(got))
    b = numpy_sv.run(4, circuit_to_stream(ref))
    assert abs(got.global_phase - ref.global_phase) < 1e-15
    assert np.max(np.abs(a - b)) < 1e-15

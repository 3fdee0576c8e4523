"""Circuit library used by the HHL / VQLS builders.

Restates the constructions of the third-party circuits the reference's qiskit path relies on
(none of their sources are in the reference tree; semantics per SURVEY.md section 8a rows a4-a6,
a12 and Appendix A):

* ``QFT``                -> ``qiskit.circuit.library.QFT``
* ``PhaseEstimation``    -> ``qiskit.circuit.library.PhaseEstimation``
* ``ExactReciprocal``    -> ``qiskit.circuit.library.ExactReciprocal``
* ``NumPyMatrix``        -> ``linear_solvers.matrices.NumPyMatrix``
* ``RealAmplitudes``     -> ``qiskit.circuit.library.RealAmplitudes`` (``vqls_qiskit_implementation.py:38-44``)
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from .circuit import Gate, ParameterVector, QuantumCircuit, QuantumRegister, UCRYGate


def QFT(num_qubits: int, inverse: bool = False, do_swaps: bool = True, name: Optional[str] = None) -> QuantumCircuit:
    """Quantum Fourier transform in qiskit's gate order: for ``j = m-1 .. 0``: ``H(j)`` then
    ``CP(pi * 2**(k-j))`` with control ``j`` and target ``k`` for ``k = j-1 .. 0``; optional swaps;
    ``inverse=True`` inverts the whole circuit."""
    m = int(num_qubits)
    qc = QuantumCircuit(m, name=name or ("IQFT" if inverse else "QFT"))
    for j in reversed(range(m)):
        qc.h(j)
        for k in reversed(range(j)):
            qc.cp(math.pi * (2.0 ** (k - j)), j, k)
    if do_swaps:
        for i in range(m // 2):
            qc.swap(i, m - 1 - i)
    if inverse:
        inv = qc.inverse()
        inv.name = qc.name
        return inv
    return qc


class NumPyMatrix:
    """Hamiltonian-evolution block ``U = expm(i * A * t)`` on ``log2(N)`` qubits.

    ``power(p)`` follows ``linear_solvers.NumPyMatrix.power`` when ``mode="repeat"`` (the unitary
    gate repeated ``p`` times -- the reference's literal gate stream) and collapses the repetition
    to one gate ``V diag(exp(i*lambda*t*p)) V^dagger`` when ``mode="eig"`` (what the engine uses
    for wide data registers)."""

    def __init__(self, matrix: np.ndarray, evolution_time: float = 2 * math.pi, mode: str = "repeat"):
        self.matrix = np.asarray(matrix)
        self.evolution_time = float(evolution_time)
        self.num_state_qubits = int(round(math.log2(self.matrix.shape[0])))
        self.num_qubits = self.num_state_qubits
        self.num_ancillas = 0
        self.mode = mode
        self._eig = None

    def _is_hermitian(self) -> bool:
        return bool(np.allclose(self.matrix, self.matrix.conj().T))

    def condition_bounds(self):
        # Hermitian input (HHL requires it): singular values = |eigenvalues|, and the eigen-decomposition is needed for the
        # evolution anyway -- np.linalg.cond (an SVD) and np.linalg.eigvals were 2.1 of the 3.1 s of construct_circuit at N = 2^10
        if self._is_hermitian():
            ev = np.abs(self._eigh()[0])
            kappa = float(np.max(ev) / np.min(ev))
        else:
            kappa = float(np.linalg.cond(self.matrix))
        return kappa, kappa

    def eigs_bounds(self):
        ev = np.abs(self._eigh()[0]) if self._is_hermitian() else np.abs(np.linalg.eigvals(self.matrix))
        return float(np.min(ev)), float(np.max(ev))

    def _eigh(self):
        if self._eig is None:
            self._eig = np.linalg.eigh(self.matrix)
        return self._eig

    def evolved(self, power: int = 1) -> np.ndarray:
        """``expm(i A t)**power`` from the eigen-decomposition.  The angle ``lambda t power`` reaches thousands of radians
        for the top clock qubits (``power = 2**(nl-1)``); it is reduced modulo 2 pi in extended precision so that the
        phases stay good to ~1e-16 (a double would lose ``nl - 1`` bits: 1e-12 at nl = 12)."""
        w, v = self._eigh()
        ld = np.longdouble
        two_pi = 2 * np.arccos(ld(-1))
        turns = w.astype(ld) * ld(self.evolution_time) / two_pi * ld(int(power))
        turns -= np.floor(turns)
        ang = (two_pi * turns).astype(np.float64)
        return (v * (np.cos(ang) + 1j * np.sin(ang))) @ v.conj().T

    def power(self, power: int) -> QuantumCircuit:
        nq = self.num_state_qubits
        if self.mode == "eig":
            qc = QuantumCircuit(nq, name=f"U**{power}")
            qc.unitary(self.evolved(power), range(nq), label=f"exp(iAt)^{power}")
            return qc
        base = QuantumCircuit(nq, name="circuit")
        base.unitary(self.evolved(1), range(nq), label="exp(iAt)")
        return base.power(power)


def PhaseEstimation(num_evaluation_qubits: int, unitary, iqft: Optional[QuantumCircuit] = None, name: str = "QPE") -> QuantumCircuit:
    """QPE with the evaluation register first and the state register second (qiskit layout):
    ``H`` on every evaluation qubit, ``controlled-U**(2**j)`` from evaluation qubit ``j``, then
    ``QFT(inverse=True, do_swaps=False).reverse_bits()``."""
    ne = int(num_evaluation_qubits)
    ns = unitary.num_qubits
    qc = QuantumCircuit(QuantumRegister(ne, "eval"), QuantumRegister(ns, "q"), name=name)
    if iqft is None:
        iqft = QFT(ne, inverse=True, do_swaps=False).reverse_bits()
    state = list(range(ne, ne + ns))
    for j in range(ne):
        qc.h(j)
    for j in range(ne):
        qc.append(unitary.power(2**j).control(1), [j] + state)
    qc.compose(iqft, qubits=list(range(ne)), inplace=True)
    return qc


def exact_reciprocal_angles(num_state_qubits: int, scaling: float, neg_vals: bool):
    nl = 2 ** (num_state_qubits - 1) if neg_vals else 2**num_state_qubits
    angles = [0.0]
    for i in range(1, nl):
        r = scaling * nl / i
        if math.isclose(r, 1, abs_tol=1e-5):
            angles.append(math.pi)
        elif r < 1:
            angles.append(2 * math.asin(r))
        else:
            angles.append(0.0)
    # sign qubit set: QPE holds a negative eigenvalue in two's complement, i.e. the value bits i encode
    # lambda ~ -(1 - i/nl); the rotation is by scaling / lambda  [qiskit ExactReciprocal, neg_vals branch]
    angles_neg = [0.0]
    for i in range(1, nl):
        r = scaling * (-1) / (1 - i / nl)
        if math.isclose(r, -1, abs_tol=1e-5):
            angles_neg.append(-math.pi)
        elif abs(r) < 1:
            angles_neg.append(2 * math.asin(r))
        else:
            angles_neg.append(0.0)
    return np.array(angles), np.array(angles_neg)


def ExactReciprocal(num_state_qubits: int, scaling: float, neg_vals: bool = False, name: str = "1/x") -> QuantumCircuit:
    """Eigenvalue inversion: ``UCRY(angles)`` on the flag selected by the value bits; with
    ``neg_vals`` two more UCRYs controlled on the sign qubit (the last state qubit)."""
    ns = int(num_state_qubits)
    qc = QuantumCircuit(QuantumRegister(ns, "state"), QuantumRegister(1, "flag"), name=name)
    angles, angles_neg = exact_reciprocal_angles(ns, scaling, neg_vals)
    state = list(range(ns))
    flag = ns
    qc.append(UCRYGate(angles), [flag] + state[: ns - int(neg_vals)])
    if neg_vals:
        qc.append(UCRYGate(-angles).control(1), [state[-1], flag] + state[:-1])
        qc.append(UCRYGate(angles_neg).control(1), [state[-1], flag] + state[:-1])
    return qc


def RealAmplitudes(num_qubits: int, entanglement: str = "full", reps: int = 3, insert_barriers: bool = False,
                   parameter_prefix: str = "θ") -> QuantumCircuit:
    """``(reps + 1)`` RY layers with CX entanglement between them; ``num_qubits * (reps + 1)``
    parameters ordered layer by layer."""
    n = int(num_qubits)
    if entanglement == "full":
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    elif entanglement == "linear":
        pairs = [(i, i + 1) for i in range(n - 1)]
    elif entanglement == "reverse_linear":
        pairs = [(i, i + 1) for i in reversed(range(n - 1))]
    elif entanglement == "circular":
        pairs = ([(n - 1, 0)] if n > 2 else []) + [(i, i + 1) for i in range(n - 1)]
    else:
        raise ValueError(f"unsupported entanglement {entanglement!r}")
    theta = ParameterVector(parameter_prefix, n * (reps + 1))
    qc = QuantumCircuit(n, name="RealAmplitudes")
    k = 0
    for layer in range(reps + 1):
        for q in range(n):
            qc.ry(theta[k], q)
            k += 1
        if layer < reps:
            for a, b in pairs:
                qc.cx(a, b)
    qc.ordered_parameters = list(theta)  # type: ignore[attr-defined]
    return qc

"""HHL algorithm front-end with the ``linear_solvers.HHL`` surface the reference calls.

Reference call sites: ``HHL(quantum_instance=backend)``, ``.solve(matrix=, vector=)``,
``result.state`` and ``result.euclidean_norm``
(``quantum_linear_systems/implementations/hhl_qiskit_implementation.py:35-40,54``).  The class
they come from is the third-party package quantum-linear-solvers 0.1.1 (not in the reference
tree, not installable here); its construction is restated per SURVEY.md section 8a rows a3-a8.

Differences that are deliberate:

* ``_calculate_norm`` does not run a second simulation: the state is evolved once on the GPU and
  the norm is one warp-shuffle reduction over ``psi[2**(n-1) : 2**(n-1)+N]``; the evolved device
  state is cached on the result so that ``Statevector(result.state)`` does not evolve again.
* ``power_mode="eig"`` collapses the ``2**j``-fold repetition of ``exp(iAt)`` into one block per
  clock qubit (needed for wide data registers); ``"repeat"`` keeps the reference's literal stream.
"""
from __future__ import annotations

import math
from typing import Any, Optional, Sequence, Union

import numpy as np

from .circuit import QuantumCircuit, QuantumRegister
from .library import ExactReciprocal, NumPyMatrix, PhaseEstimation


class LinearSolverResult:
    """Same attribute bag as ``linear_solvers.LinearSolverResult``."""

    def __init__(self) -> None:
        self.state: Optional[QuantumCircuit] = None
        self.euclidean_norm: Optional[float] = None
        self.observable: Any = None
        self.circuit_results: Any = None


class HHL:
    def __init__(self, epsilon: float = 1e-2, expectation: Any = None, quantum_instance: Any = None,
                 power_mode: str = "auto", precision: str = "fp64", device: Optional[int] = None):
        self._epsilon = epsilon
        self._expectation = expectation
        self.quantum_instance = quantum_instance  # accepted and ignored, like the reference's use
        self.scaling = 1.0
        self._exact_reciprocal = True
        self.power_mode = power_mode
        self.precision = precision
        self.device = device
        # parameters of the last constructed circuit (handy for tests and benches)
        self.nb = self.nl = 0
        self.delta = self.evolution_time = 0.0

    @staticmethod
    def _get_delta(n_l: int, lambda_min: float, lambda_max: float) -> float:
        """``lambda_min`` expressed on ``n_l`` bits relative to ``lambda_max`` (truncated), read
        back as a binary fraction -- so that the smallest eigenvalue is represented exactly."""
        scaled = abs(lambda_min * (2**n_l - 1) / lambda_max)
        if abs(scaled - 1) < 1e-7:  # guard against 0.99999...
            scaled = 1
        integer = int(scaled)
        rep = 0.0
        for pos in range(n_l):  # MSB first
            if (integer >> (n_l - 1 - pos)) & 1:
                rep += 1.0 / 2 ** (pos + 1)
        return rep

    def construct_circuit(self, matrix: Union[np.ndarray, Sequence], vector: Union[np.ndarray, Sequence, QuantumCircuit],
                          neg_vals: bool = True) -> QuantumCircuit:
        if isinstance(vector, QuantumCircuit):
            nb = vector.num_qubits
            vector_circuit = vector
        else:
            vec = np.asarray(vector).reshape(-1)
            nb = int(np.log2(len(vec)))
            if 2**nb != len(vec):
                raise ValueError("Input vector dimension must be 2^n!")
            vector_circuit = QuantumCircuit(nb, name="isometry")
            vector_circuit.prepare_state(vec / np.linalg.norm(vec), list(range(nb)))
        nf = 1

        mat = np.asarray(matrix)
        if mat.shape[0] != mat.shape[1]:
            raise ValueError("Input matrix must be square!")
        if np.log2(mat.shape[0]) % 1 != 0:
            raise ValueError("Input matrix dimension must be 2^n!")
        if not np.allclose(mat, mat.conj().T):
            raise ValueError("Input matrix must be hermitian!")
        if mat.shape[0] != 2**nb:
            raise ValueError(f"Input vector dimension does not match input matrix dimension! Vector dimension: {2**nb}. "
                             f"Matrix dimension: {mat.shape[0]}")
        matrix_circuit = NumPyMatrix(mat, evolution_time=2 * np.pi)

        kappa = matrix_circuit.condition_bounds()[1]
        nl = max(nb + 1, int(np.ceil(np.log2(kappa + 1)))) + int(neg_vals)
        lambda_min, lambda_max = matrix_circuit.eigs_bounds()
        delta = self._get_delta(nl - int(neg_vals), lambda_min, lambda_max)
        matrix_circuit.evolution_time = 2 * np.pi * delta / lambda_min / (2 ** int(neg_vals))
        self.scaling = lambda_min

        mode = self.power_mode
        if mode == "auto":
            # the literal stream repeats U (2**nl - 1) times per QPE; collapse it once that is large
            mode = "repeat" if (nb <= 3 and nl <= 6) else "eig"
        matrix_circuit.mode = mode

        reciprocal_circuit = ExactReciprocal(nl, delta, neg_vals=neg_vals)

        qb = QuantumRegister(nb, "q_b")
        ql = QuantumRegister(nl, "q_l")
        qf = QuantumRegister(nf, "q_f")
        qc = QuantumCircuit(qb, ql, qf, name="hhl")
        qb_q, ql_q, qf_q = qc.register_qubits(0), qc.register_qubits(1), qc.register_qubits(2)

        qc.append(vector_circuit.to_gate(), qb_q)
        phase_estimation = PhaseEstimation(nl, matrix_circuit)
        qc.append(phase_estimation.to_gate(), ql_q + qb_q)
        qc.append(reciprocal_circuit.to_gate(), ql_q[::-1] + [qf_q[0]])
        qc.append(phase_estimation.inverse().to_gate(), ql_q + qb_q)

        self.nb, self.nl, self.delta = nb, nl, delta
        self.evolution_time = matrix_circuit.evolution_time
        return qc

    def _calculate_norm(self, qc: QuantumCircuit) -> float:
        """sqrt(P(flag=1, clock=0)) / scaling with the engine's reduction kernel."""
        from .statevector import Statevector

        sv = Statevector(qc, precision=self.precision, device=self.device)
        nb = qc.qregs[0].size
        half = 2 ** (qc.num_qubits - 1)
        norm2 = sv.norm_sq_range(half, 2**nb)
        qc._sv_cache = ((len(qc.data), self.precision, getattr(sv, "device", self.device)), sv)  # type: ignore[attr-defined]
        return float(np.real(math.sqrt(norm2) / self.scaling))

    def solve(self, matrix: Union[np.ndarray, Sequence], vector: Union[np.ndarray, Sequence, QuantumCircuit],
              observable: Any = None, observable_circuit: Any = None, post_processing: Any = None) -> LinearSolverResult:
        if observable is not None or observable_circuit is not None or post_processing is not None:
            raise NotImplementedError("observables are not used by the reference's path and are not supported")
        solution = LinearSolverResult()
        solution.state = self.construct_circuit(matrix, vector)
        solution.euclidean_norm = self._calculate_norm(solution.state)
        return solution

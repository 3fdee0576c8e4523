"""``QuantumLinearSolver`` facade with the reference's surface
(``quantum_linear_systems/quantum_linear_solver.py:19-145``): ``solve(matrix_a, vector_b, method,
file_basename, **kwargs)``, the five result attributes, ``circuit_data``, ``save_qasm`` and the
static checks.  ``hhl_classiq`` needs Classiq's authenticated cloud service and is out of scope."""
from __future__ import annotations

from typing import Any, Tuple

import numpy as np


class QuantumLinearSolver:
    """Solve ``Ax = b`` with "hhl_qiskit" or "vqls_qiskit" on the B200 statevector engine."""

    def __init__(self) -> None:
        self.matrix_a: np.ndarray
        self.vector_b: np.ndarray
        self.name: str
        self.method: str
        self.solution: np.ndarray
        self.qasm_circuit: str
        self.circuit_width: int
        self.circuit_depth: int
        self.run_time: float

    def circuit_data(self) -> Tuple[str, int, int]:
        return self.qasm_circuit, self.circuit_depth, self.circuit_width

    def normalize_model(self) -> None:
        norm_b = np.linalg.norm(self.vector_b)
        self.matrix_a = self.matrix_a / norm_b
        self.vector_b = self.vector_b / norm_b

    def solve(self, matrix_a: np.ndarray, vector_b: np.ndarray, method: str, file_basename: str, **kwargs: Any) -> np.ndarray:
        self.matrix_a, self.vector_b, self.name, self.method = matrix_a, vector_b, file_basename, method
        self.check_matrix_square_hermitian(self.matrix_a)
        if method == "hhl_qiskit":
            from .implementations.hhl_qiskit_implementation import solve_hhl_qiskit as impl
        elif method == "vqls_qiskit":
            from .implementations.vqls_qiskit_implementation import solve_vqls_qiskit as impl
        elif method == "hhl_classiq":
            raise NotImplementedError("hhl_classiq needs Classiq's authenticated cloud service (out of scope)")
        else:
            raise NotImplementedError(f"Unsupported method: {method}")
        (self.solution, self.qasm_circuit, self.circuit_depth, self.circuit_width, self.run_time) = impl(
            matrix_a=matrix_a, vector_b=vector_b, **kwargs
        )
        return self.solution

    def save_qasm(self) -> None:
        with open(f"{self.name}_{self.method}.qasm", "w", encoding="utf-8") as qasm_file:
            qasm_file.write(self.qasm_circuit)

    @staticmethod
    def check_matrix_square_hermitian(matrix_a: np.ndarray) -> None:
        rows, cols = matrix_a.shape[0], matrix_a.shape[1]
        if rows != cols:
            raise ValueError(f"Input matrix A needs to be square, not {rows}x{cols}.")
        if not np.allclose(matrix_a, np.conj(matrix_a.T)):
            raise ValueError("Input matrix A is not hermitian!")

    @staticmethod
    def check_matrix_condition_number(matrix_a: np.ndarray) -> float:
        mags = np.abs(np.linalg.eigvals(matrix_a))
        return float(np.max(mags) / np.min(mags[np.nonzero(mags)]))

    @staticmethod
    def check_matrix_sparsity(matrix_a: np.ndarray) -> float:
        return float(np.sum(matrix_a == 0) / (matrix_a.shape[0] * matrix_a.shape[1]))

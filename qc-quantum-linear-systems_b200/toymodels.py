"""Input generators for the parity cases (BASELINE.json configs 1-2): the reference's toy models
that do not need third-party packages (``quantum_linear_systems/toymodels.py:18-185``), plus the
synthetic tridiagonal family of SURVEY.md section 8d.  ``HEPTrackReconstruction`` needs the
uninstallable ``trackhhl`` and is out of scope."""
from __future__ import annotations

import numpy as np

from .utils import expand_b_vector, make_matrix_hermitian


class ToyModel:
    """A linear system ``A x = b`` normalised by ``|b|`` (``toymodels.py:59-64``)."""

    def __init__(self, name: str, matrix: np.ndarray, vector: np.ndarray, csol: np.ndarray):
        if not isinstance(name, str):
            raise TypeError(f"Name of ToyModel must be of type str not {type(name)}.")
        for arr in (matrix, vector, csol):
            if not isinstance(arr, np.ndarray):
                raise TypeError(f"Matrix, vector and csol of ToyModel must be of type np.ndarray not {type(arr)}.")
        self.name = name
        self.matrix_a = matrix
        self.vector_b = vector
        self.classical_solution = csol.flatten()
        self.normalize_model()

    @property
    def num_qubits(self) -> int:
        return int(np.log2(self.matrix_a.shape[0]))

    def normalize_model(self) -> None:
        scale = np.linalg.norm(self.vector_b)
        self.matrix_a = self.matrix_a / scale
        self.classical_solution = self.classical_solution / scale
        self.vector_b = self.vector_b / scale

    @staticmethod
    def classically_solve(mat: np.ndarray, vec: np.ndarray) -> np.ndarray:
        return np.linalg.solve(mat, vec)


class Qiskit4QubitExample(ToyModel):
    """A = [[1, -1/3], [-1/3, 1]], b = (1, 0), x = (1.125, 0.375)."""

    def __init__(self) -> None:
        super().__init__("Qiskit4QubitExample", np.array([[1, -1 / 3], [-1 / 3, 1]]), np.array([1, 0]),
                         np.array([[1.125], [0.375]]))


class VolterraProblem(ToyModel):
    """Discretised Volterra equation x(t) = 1 - int_0^t x(s) ds on ``2**num_qubits`` points,
    embedded into a Hermitian system of twice the size."""

    def __init__(self, num_qubits: int) -> None:
        total = 2**num_qubits
        alpha = (1 / total) / 2
        mat = self.volterra_a_matrix(total, alpha)
        vec = np.ones((total, 1))
        super().__init__(f"VolterraProblem(n={num_qubits})", make_matrix_hermitian(mat), expand_b_vector(vec),
                         self.classically_solve(mat, vec))

    @staticmethod
    def volterra_a_matrix(size: int, alpha: float) -> np.ndarray:
        a = np.tril(np.full((size, size), 2 * alpha), k=-1)
        a[:, 0] = alpha
        np.fill_diagonal(a, 1 + alpha)
        a[0, 0] = 1
        return a


class ClassiqDemoExample(ToyModel):
    def __init__(self) -> None:
        a = np.array([[0.28, -0.01, 0.02, -0.1], [-0.01, 0.5, -0.22, -0.07],
                      [0.02, -0.22, 0.43, -0.05], [-0.1, -0.07, -0.05, 0.42]])
        b = np.array([1, 2, 4, 3])
        b = b / np.linalg.norm(b)
        super().__init__("ClassiqDemoExample", a, b, self.classically_solve(a, b))


class TridiagonalToeplitz(ToyModel):
    """Synthetic Hermitian ``tridiag(off, diag, off)`` of size ``2**num_qubits`` (BASELINE.json
    config 2: diag=1, off=-1/3, N=2**10, b ~ N(0,1) seeded) -- the 2x2 KAT matrix extended."""

    def __init__(self, num_qubits: int, diag: float = 1.0, off: float = -1.0 / 3.0, seed: int = 1234) -> None:
        n = 2**num_qubits
        a = diag * np.eye(n) + off * (np.eye(n, k=1) + np.eye(n, k=-1))
        b = np.random.default_rng(seed).standard_normal(n)
        b = b / np.linalg.norm(b)  # unit right-hand side: ToyModel.normalize_model is then a no-op
        super().__init__(f"TridiagonalToeplitz(n={num_qubits})", a, b, self.classically_solve(a, b))

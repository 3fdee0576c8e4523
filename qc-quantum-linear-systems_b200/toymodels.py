"""Input generators for the parity cases (BASELINE.json configs 1-2): the reference's toy models
that do not need third-party packages (``quantum_linear_systems/toymodels.py:18-185``), plus the
synthetic tridiagonal family of SURVEY.md section 8d.  ``HEPTrackReconstruction`` needs the
uninstallable ``trackhhl`` and is out of scope."""
from __future__ import annotations

import numpy as np

from .utils import (expand_b_vector, extract_x_from_expanded, generate_random_vector, generate_s_sparse_matrix,
                    is_matrix_well_conditioned, make_matrix_hermitian, vector_uniformity_entropy)


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


class RandomNQubitProblem(ToyModel):
    """Random symmetric ``2**num_qubits`` system: ``A = R + R^T`` with ``R``, ``b`` ~ U(0, 1) from NumPy's
    global generator (``toymodels.py:188-208``)."""

    def __init__(self, num_qubits: int) -> None:
        dim = 2**num_qubits
        r = np.random.rand(dim, dim)
        a = r + r.T
        b = np.random.rand(dim)
        super().__init__(f"RandomNQubitProblem(N={num_qubits})", a, b, self.classically_solve(a, b))


def integro_differential_a_matrix(a_matrix: np.ndarray, time_discretization_steps: int) -> np.ndarray:
    """Block matrix of the explicit-Euler integro-differential toy model (``toymodels.py:211-247``):
    identity on the block diagonal, ``-(1 + dt*A)`` on the block sub-diagonal, ``dt = 1/steps``."""
    steps = int(time_discretization_steps)
    m = a_matrix.shape[0]
    out = np.zeros((steps * m, steps * m), dtype=np.result_type(a_matrix, float))
    sub = -np.identity(m) - a_matrix / steps
    for i in range(steps):
        out[i * m:(i + 1) * m, i * m:(i + 1) * m] = np.identity(m)
        if i:
            out[i * m:(i + 1) * m, (i - 1) * m:i * m] = sub
    return out


class ScalingTestModel(ToyModel):
    """Scaling study input (``toymodels.py:250-330``): an ``s``-sparse random matrix made Hermitian by
    embedding, drawn again until its condition number is (``matrix_well_conditioned``) at most
    ``10 * matrix_size`` / (otherwise) above 1000, and a random ``b`` of the given uniformity."""

    def __init__(self, matrix_size: int = 4, matrix_s: int = 2, matrix_well_conditioned: bool = True,
                 vector_uniformity: float = 1.0, max_num_iterations: int = 100):
        b = generate_random_vector(size=matrix_size, uniformity_level=vector_uniformity)
        entropy = vector_uniformity_entropy(b)
        b = expand_b_vector(b)
        assert np.isclose(entropy, vector_uniformity_entropy(b))
        threshold = 10 * matrix_size if matrix_well_conditioned else 1000
        a = make_matrix_hermitian(generate_s_sparse_matrix(matrix_size=matrix_size, s_non_zero_entries=matrix_s))
        tries = 1
        while matrix_well_conditioned != is_matrix_well_conditioned(a, threshold):
            a = make_matrix_hermitian(generate_s_sparse_matrix(matrix_size=matrix_size, s_non_zero_entries=matrix_s))
            tries += 1
            if tries == max_num_iterations:
                raise ValueError(f"Could not generate matrix where matrix_well_conditioned={matrix_well_conditioned} "
                                 f"within {tries} iterations.")
        csol = extract_x_from_expanded(self.classically_solve(a, b))
        name = f"TestNxN_n={matrix_size}_s={matrix_s}_c={np.linalg.cond(a):.1f}_e={vector_uniformity_entropy(b):.1f}"
        super().__init__(name=name, matrix=a, vector=b, csol=csol)


class TridiagonalToeplitz(ToyModel):
    """Synthetic Hermitian ``tridiag(off, diag, off)`` of size ``2**num_qubits`` (BASELINE.json
    config 2: diag=1, off=-1/3, N=2**10, b ~ N(0,1) seeded) -- the 2x2 KAT matrix extended."""

    def __init__(self, num_qubits: int, diag: float = 1.0, off: float = -1.0 / 3.0, seed: int = 1234) -> None:
        n = 2**num_qubits
        a = diag * np.eye(n) + off * (np.eye(n, k=1) + np.eye(n, k=-1))
        b = np.random.default_rng(seed).standard_normal(n)
        b = b / np.linalg.norm(b)  # unit right-hand side: ToyModel.normalize_model is then a no-op
        super().__init__(f"TridiagonalToeplitz(n={num_qubits})", a, b, self.classically_solve(a, b))

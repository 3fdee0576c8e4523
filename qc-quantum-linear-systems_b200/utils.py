"""Host-side helpers on the hot path's output side (drop-in for the reference's
``quantum_linear_systems/utils.py``; only the functions the qiskit path uses: ``:10-14``,
``:17-39``, ``:42-94``, ``:97-119``).  Same names, arguments, return values and error behaviour;
written from the behaviour, not the text."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np


def circuit_to_qasm3(circuit: Any, filename: str) -> str:
    """OpenQASM 3 text of ``circuit``; also written to ``filename`` (reference side effect,
    ``utils.py:10-14``)."""
    from .qasm3 import dumps

    text = dumps(circuit)
    with open(filename, "w", encoding="utf-8") as fh:
        fh.write(text)
    return text


def make_matrix_hermitian(matrix: np.ndarray) -> np.ndarray:
    """``[[0, A], [A^dagger, 0]]`` for an ``N x M`` matrix ``A``."""
    rows, cols = matrix.shape
    out = np.zeros((rows + cols, rows + cols), dtype=np.result_type(matrix.dtype, float))
    out[:rows, rows:] = matrix
    out[rows:, :rows] = matrix.conj().T
    assert np.array_equal(out, out.conj().T)
    return out


def expand_b_vector(unexpanded_vector: np.ndarray, non_square_matrix: Optional[np.ndarray] = None) -> np.ndarray:
    """``b -> (b, 0)`` to match the Hermitian embedding of ``A``."""
    pad = non_square_matrix.shape[1] if non_square_matrix is not None else unexpanded_vector.shape[0]
    flat = np.asarray(unexpanded_vector).reshape(-1)
    return np.concatenate([flat, np.zeros(pad, dtype=np.result_type(flat.dtype, float))])


def is_expanded(matrix_a: np.ndarray, vector_b: np.ndarray) -> bool:
    """True when ``b``'s second half is zero and ``A`` has zero top-left / bottom-right quarters."""
    if matrix_a.shape[0] != matrix_a.shape[1]:
        raise ValueError("Matrix not square")
    q = matrix_a.shape[0] // 2
    corners_zero = not (np.any(matrix_a[:q, :q] != 0) or np.any(matrix_a[q:, q:] != 0))
    tail_zero = all(e == 0 for e in vector_b[len(vector_b) // 2:])
    return bool(tail_zero and corners_zero)


def extract_x_from_expanded(expanded_solution_vector: np.ndarray) -> np.ndarray:
    """The expanded system's solution is ``y = (0, x)``; return ``x``."""
    y = np.asarray(expanded_solution_vector)
    return y[len(y) // 2:].flatten()


def extract_hhl_solution_vector_from_state_vector(hermitian_matrix: np.ndarray, state_vector: np.ndarray) -> np.ndarray:
    """Real part of the amplitudes with flag = 1 and every other non-solution qubit 0 (the slice
    starting at ``2**(n-1)``, ``A.shape[1]`` long), L2-normalised.  Accepts a host array or an
    engine state (only the slice is copied back from the GPU)."""
    size = hermitian_matrix.shape[1]
    if hasattr(state_vector, "read_slice"):
        start = state_vector.dim // 2
        part = state_vector.read_slice(start, min(size, state_vector.dim - start))
    else:
        sv = np.asarray(state_vector)
        start = 1 << (int(np.log2(len(sv))) - 1)
        part = sv[start:start + size]
    vec = np.real(part)
    return vec / np.linalg.norm(vec)


def normalize_quantum_by_classical_solution(quantum_solution: np.ndarray, classical_solution: np.ndarray) -> np.ndarray:
    return quantum_solution / np.linalg.norm(quantum_solution) * np.linalg.norm(classical_solution)


def relative_distance_quantum_classical_solution(quantum_solution: np.ndarray, classical_solution: np.ndarray) -> float:
    """Relative distance in percent (``utils.py:106-119``)."""
    if quantum_solution.shape != classical_solution.shape:
        raise ValueError(
            f"Can't compute relative distance. Shape of quantum solution {quantum_solution.shape} "
            f"different from classical {classical_solution.shape}."
        )
    return float(np.linalg.norm(classical_solution - quantum_solution) / np.linalg.norm(classical_solution) * 100)


# ---- input generators of the scaling studies (reference ``utils.py:122-262``; same names and
# semantics, NumPy's global generator as in the reference so that ``np.random.seed`` pins them) ----


def generate_random_vector(size: int, uniformity_level: float) -> np.ndarray:
    """Normalised random column vector ``(size, 1)``: a fraction ``uniformity_level`` of its entries is
    drawn from U(0, 1), the rest from N(0, 1), shuffled (0: all normal, 1: all uniform)."""
    if not 0 <= uniformity_level <= 1:
        raise ValueError("uniformity_level should be between 0 and 1")
    if uniformity_level == 0:
        vec = np.random.randn(size, 1)
    elif uniformity_level == 1:
        vec = np.random.rand(size, 1)
    else:
        n_uniform = int(size * uniformity_level)
        vec = np.concatenate((np.random.rand(n_uniform, 1), np.random.randn(size - n_uniform, 1)))
        np.random.shuffle(vec)
    return vec / np.linalg.norm(vec)


def vector_uniformity_entropy(vector: np.ndarray) -> float:
    """Shannon entropy (bits) of the vector normalised to unit sum; epsilon 1e-10 inside the log."""
    v = np.array(vector)
    p = v / np.sum(v)
    return float(-np.sum(p * np.log2(p + 1e-10)))


def generate_s_sparse_matrix(matrix_size: int, s_non_zero_entries: int) -> np.ndarray:
    """Square matrix with exactly ``s`` U(0, 1) entries per row at random distinct columns."""
    if s_non_zero_entries <= 0:
        raise ValueError("s must be a positive integer")
    if matrix_size <= 0:
        raise ValueError("matrix_size must be a positive integer")
    if s_non_zero_entries > matrix_size:
        raise ValueError("s cannot be greater than matrix_size")
    out = np.zeros((matrix_size, matrix_size), dtype=float)
    for row in range(matrix_size):
        cols = np.random.choice(matrix_size, s_non_zero_entries, replace=False)
        out[row, cols] = np.random.rand(s_non_zero_entries)
    return out


def is_matrix_well_conditioned(matrix: np.ndarray, threshold: float = 10.0) -> bool:
    """2-norm condition number at most ``threshold``."""
    return bool(np.linalg.cond(matrix) <= threshold)

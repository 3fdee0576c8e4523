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

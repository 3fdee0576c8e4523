"""``solve_hhl_qiskit`` on the B200 engine: same signature, return tuple, side effects and error
behaviour as the reference's ``quantum_linear_systems/implementations/hhl_qiskit_implementation.py:20-79``.

Where the reference evolves the circuit twice on the CPU (``HHL.solve`` -> ``_calculate_norm``,
then ``Statevector(hhl_circuit).data`` at ``:49``), this evolves it once on the GPU, reduces the
norm there, and copies back only the ``N`` solution amplitudes."""
from __future__ import annotations

import time
from typing import Any, Tuple

import numpy as np

from ..hhl import HHL, LinearSolverResult
from ..statevector import Statevector
from ..utils import circuit_to_qasm3, extract_hhl_solution_vector_from_state_vector, extract_x_from_expanded, is_expanded


def solve_hhl_qiskit(matrix_a: np.ndarray, vector_b: np.ndarray, show_circuit: bool = False, backend: Any = None,
                     precision: str = "fp64", device: Any = None, power_mode: str = "auto",
                     ) -> Tuple[np.ndarray, str, int, int, float]:
    np.set_printoptions(precision=3, suppress=True)
    start_time = time.time()

    hhl_implementation = HHL(quantum_instance=backend, precision=precision, device=device, power_mode=power_mode)
    naive_hhl_solution: LinearSolverResult = hhl_implementation.solve(matrix=matrix_a, vector=vector_b)
    hhl_circuit = naive_hhl_solution.state
    print(f"Size of solution register is {hhl_circuit.qregs[0].size} , QPE registers is {hhl_circuit.qregs[1].size}.")

    naive_state_vec = Statevector(hhl_circuit, precision=precision, device=device)  # reuses the evolved device state
    hhl_solution_vector = extract_hhl_solution_vector_from_state_vector(
        hermitian_matrix=matrix_a, state_vector=naive_state_vec
    )
    hhl_solution_vector = naive_hhl_solution.euclidean_norm * hhl_solution_vector

    if show_circuit:
        print(hhl_circuit.draw())

    print("x quantum vs classical solution")
    if is_expanded(matrix_a, vector_b):
        hhl_solution_vector = extract_x_from_expanded(hhl_solution_vector)

    qc_basis = hhl_circuit.decompose(reps=10)
    print(f"Comparing depths original {hhl_circuit.depth()} vs. decomposed {qc_basis.depth()}")
    qasm_content = circuit_to_qasm3(circuit=qc_basis, filename="hhl_qiskit_circuit.qasm3")

    return (hhl_solution_vector, qasm_content, qc_basis.depth(), hhl_circuit.width(), time.time() - start_time)

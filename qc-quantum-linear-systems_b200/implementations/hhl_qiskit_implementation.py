"""``solve_hhl_qiskit`` on the B200 engine.

Interface of the reference's ``quantum_linear_systems/implementations/hhl_qiskit_implementation.py:20-79``:
``(matrix_a, vector_b, show_circuit=False, backend=None) -> (x, qasm3 text, depth, width, seconds)``, the
``hhl_qiskit_circuit.qasm3`` file written into the working directory, the global NumPy print options and the
progress lines on stdout.  Extra keyword arguments select the engine (``precision``, ``device``) and how the
controlled powers of ``exp(iAt)`` are built (``power_mode``, see ``qls_b200.hhl``).

The reference evolves the circuit twice on the CPU (once inside ``HHL.solve`` for the norm, once through
``Statevector(circuit).data`` at ``:49``).  Here it is evolved once, on the GPU: the norm is a device reduction
over the flag = 1 / clock = 0 slice and only those ``N`` amplitudes are copied back to the host."""
from __future__ import annotations

import time
from typing import Any, Tuple

import numpy as np

from .. import utils
from ..hhl import HHL
from ..statevector import Statevector

QASM_FILE = "hhl_qiskit_circuit.qasm3"  # utils.py:12-13 of the reference writes the same name


def _scaled_solution(matrix_a: np.ndarray, vector_b: np.ndarray, state: Statevector, norm: float) -> np.ndarray:
    """Unit solution slice of the final state, scaled to ``HHL._calculate_norm``'s value, zero padding removed."""
    x = norm * utils.extract_hhl_solution_vector_from_state_vector(hermitian_matrix=matrix_a, state_vector=state)
    return utils.extract_x_from_expanded(x) if utils.is_expanded(matrix_a, vector_b) else x


def solve_hhl_qiskit(matrix_a: np.ndarray, vector_b: np.ndarray, show_circuit: bool = False, backend: Any = None,
                     precision: str = "fp64", device: Any = None, power_mode: str = "auto",
                     ) -> Tuple[np.ndarray, str, int, int, float]:
    t0 = time.time()
    np.set_printoptions(precision=3, suppress=True)  # the reference sets this process-wide as well (:31)

    solver = HHL(quantum_instance=backend, precision=precision, device=device, power_mode=power_mode)
    result = solver.solve(matrix=matrix_a, vector=vector_b)
    circuit = result.state
    data_reg, clock_reg = circuit.qregs[0], circuit.qregs[1]
    print(f"Size of solution register is {data_reg.size} , QPE registers is {clock_reg.size}.")

    evolved = Statevector(circuit, precision=precision, device=device)  # picks up the state HHL.solve left in HBM
    x = _scaled_solution(matrix_a, vector_b, evolved, result.euclidean_norm)
    if show_circuit:
        print(circuit.draw())
    print("x quantum vs classical solution")

    # QASM and depth come from the un-fused circuit expanded to leaf gates; width from the original (:64-77)
    leaves = circuit.decompose(reps=10)
    print(f"Comparing depths original {circuit.depth()} vs. decomposed {leaves.depth()}")
    qasm_text = utils.circuit_to_qasm3(circuit=leaves, filename=QASM_FILE)
    return x, qasm_text, leaves.depth(), circuit.width(), time.time() - t0

"""``solve_vqls_qiskit`` on the B200 engine: same signature, return tuple, side effects and error
behaviour as the reference's
``quantum_linear_systems/implementations/vqls_qiskit_implementation.py:24-126``.  The default
``estimator`` is a ``B200Estimator`` (created on first use, then shared across calls like the
reference's import-time ``Estimator()`` default, ``:28``)."""
from __future__ import annotations

import time
from typing import Any, Optional, Tuple

import numpy as np

from ..circuit import QuantumCircuit
from ..library import RealAmplitudes
from ..utils import circuit_to_qasm3, extract_x_from_expanded, is_expanded
from ..vqls import COBYLA, SLSQP, VQLS

_DEFAULT_ESTIMATOR = None


def _default_estimator():
    global _DEFAULT_ESTIMATOR
    if _DEFAULT_ESTIMATOR is None:
        from ..estimator import B200Estimator

        _DEFAULT_ESTIMATOR = B200Estimator()
    return _DEFAULT_ESTIMATOR


def solve_vqls_qiskit(matrix_a: np.ndarray, vector_b: np.ndarray, ansatz: Optional[QuantumCircuit] = None,
                      estimator: Any = None, optimizer_name: str = "cobyla", optimizer_max_iter: int = 250,
                      show_circuit: bool = False, seed: Optional[int] = None) -> Tuple[np.ndarray, str, int, int, float]:
    start_time = time.time()
    np.set_printoptions(precision=3, suppress=True)

    if ansatz is None:
        ansatz = RealAmplitudes(num_qubits=int(np.log2(matrix_a.shape[0])), entanglement="full", reps=3,
                                insert_barriers=False)
    if optimizer_name.lower() == "cobyla":
        optimizer = COBYLA(maxiter=optimizer_max_iter, disp=True)
    elif optimizer_name.lower() == "slsqp":
        optimizer = SLSQP(maxiter=optimizer_max_iter, disp=True)
    else:
        raise ValueError(f"Invalid optimizer_name: {optimizer_name}")
    if estimator is None:
        estimator = _default_estimator()

    if vector_b.ndim == 2:
        vector_b = vector_b.flatten()

    vqls = VQLS(estimator=estimator, ansatz=ansatz, optimizer=optimizer, sampler=None,
                options={"use_overlap_test": False, "use_local_cost_function": False}, seed=seed)
    res = vqls.solve(matrix_a, vector_b)

    vqls_circuit = res.state
    from ..statevector import Statevector

    vqls_solution_vector = np.real(Statevector(res.state).data)
    quantum_solution = postprocess_solution(matrix_a=matrix_a, vector_b=vector_b, solution_x=vqls_solution_vector)

    qc_basis = vqls_circuit.decompose(reps=10)
    if show_circuit:
        print(qc_basis)
    qasm_content = circuit_to_qasm3(circuit=vqls_circuit, filename="vqls_qiskit_circuit.qasm3")
    print(f"Comparing depths original {vqls_circuit.depth()} vs. decomposed {qc_basis.depth()}")

    return (quantum_solution, qasm_content, qc_basis.depth(), vqls_circuit.width(), time.time() - start_time)


def postprocess_solution(matrix_a: np.ndarray, vector_b: np.ndarray, solution_x: np.ndarray) -> np.ndarray:
    """Rescale the normalised VQLS state to a solution of ``Ax = b`` (norm of ``Ax`` vs ``b``), fix
    the global sign (first component where ``Ax`` and ``b`` disagree in sign decides), and drop the
    zero half of a Hermitian-embedded system (reference ``:94-126``)."""
    lhs = np.matmul(matrix_a, solution_x)
    scaled = solution_x * (np.linalg.norm(vector_b) / np.linalg.norm(lhs))
    same_sign = True
    for i in range(len(vector_b)):
        li = lhs[i] if lhs.ndim == 1 else lhs[0, i]
        same_sign = abs(vector_b[i]) + abs(li) == abs(vector_b[i] + li)
        if not same_sign:
            break
    if not same_sign:
        scaled = -scaled
    if is_expanded(matrix_a, vector_b):
        scaled = extract_x_from_expanded(scaled)
    return scaled

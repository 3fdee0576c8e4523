"""``solve_vqls_qiskit`` on the B200 engine.

Interface of the reference's ``quantum_linear_systems/implementations/vqls_qiskit_implementation.py:24-126``:
``(matrix_a, vector_b, ansatz=None, estimator=..., optimizer_name="cobyla", optimizer_max_iter=250,
show_circuit=False) -> (x, qasm3 text, depth, width, seconds)``, ``ValueError`` for an unknown optimizer, the
``vqls_qiskit_circuit.qasm3`` file in the working directory, and ``postprocess_solution``.

``estimator`` is the reference's plug-in hole (it swaps in a Braket estimator at ``aws_execution.py:174-182``);
the default here is one process-wide ``B200Estimator`` -- created on first use instead of at import time
(``:28`` of the reference), so importing this module does not need a GPU.  ``seed`` fixes the optimizer's random
start point, which the reference leaves to ``vqls_prototype``."""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np

from .. import utils
from ..circuit import QuantumCircuit
from ..library import RealAmplitudes
from ..vqls import COBYLA, SLSQP, VQLS

QASM_FILE = "vqls_qiskit_circuit.qasm3"
_OPTIMIZERS: Dict[str, Callable[..., Any]] = {"cobyla": COBYLA, "slsqp": SLSQP}
_shared_estimator: Any = None


def _default_estimator() -> Any:
    global _shared_estimator
    if _shared_estimator is None:
        from ..estimator import B200Estimator

        _shared_estimator = B200Estimator()
    return _shared_estimator


def _make_optimizer(name: str, max_iter: int) -> Any:
    try:
        factory = _OPTIMIZERS[name.lower()]
    except KeyError:
        raise ValueError(f"Invalid optimizer_name: {name}") from None
    return factory(maxiter=max_iter, disp=True)


def solve_vqls_qiskit(matrix_a: np.ndarray, vector_b: np.ndarray, ansatz: Optional[QuantumCircuit] = None,
                      estimator: Any = None, optimizer_name: str = "cobyla", optimizer_max_iter: int = 250,
                      show_circuit: bool = False, seed: Optional[int] = None) -> Tuple[np.ndarray, str, int, int, float]:
    from ..statevector import Statevector

    t0 = time.time()
    np.set_printoptions(precision=3, suppress=True)  # process-wide, like the reference (:36)
    n_data = int(np.log2(matrix_a.shape[0]))
    rhs = np.ravel(vector_b) if np.ndim(vector_b) == 2 else vector_b

    solver = VQLS(
        estimator=estimator if estimator is not None else _default_estimator(),
        ansatz=ansatz if ansatz is not None else RealAmplitudes(num_qubits=n_data, entanglement="full", reps=3, insert_barriers=False),
        optimizer=_make_optimizer(optimizer_name, optimizer_max_iter),
        sampler=None,
        options={"use_overlap_test": False, "use_local_cost_function": False},  # global cost through Hadamard tests
        seed=seed,
    )
    outcome = solver.solve(matrix_a, rhs)
    circuit = outcome.state  # the ansatz bound to the optimal parameters

    amplitudes = np.real(Statevector(circuit).data)  # RealAmplitudes: the imaginary parts are rounding noise
    x = postprocess_solution(matrix_a=matrix_a, vector_b=rhs, solution_x=amplitudes)

    leaves = circuit.decompose(reps=10)
    if show_circuit:
        print(leaves)
    qasm_text = utils.circuit_to_qasm3(circuit=circuit, filename=QASM_FILE)  # the reference exports the UN-decomposed circuit (:76-79)
    print(f"Comparing depths original {circuit.depth()} vs. decomposed {leaves.depth()}")
    return x, qasm_text, leaves.depth(), circuit.width(), time.time() - t0


def postprocess_solution(matrix_a: np.ndarray, vector_b: np.ndarray, solution_x: np.ndarray) -> np.ndarray:
    """From the unit vector VQLS returns to a solution of ``Ax = b`` (reference ``:94-126``):

    * length: ``|b| / |A x|``;
    * sign: walk the components of ``A x`` and ``b`` until one pair disagrees in sign
      (``|b_i| + |(Ax)_i| != |b_i + (Ax)_i|``); if the LAST pair looked at disagrees, flip ``x``;
    * a Hermitian-embedded system ``[[0, A], [A^H, 0]]`` keeps only the second half."""
    image = np.asarray(np.matmul(matrix_a, solution_x))
    row = image if image.ndim == 1 else image[0]
    x = solution_x * (np.linalg.norm(vector_b) / np.linalg.norm(image))
    agrees = True
    for b_i, ax_i in zip(vector_b, row):
        agrees = abs(b_i) + abs(ax_i) == abs(b_i + ax_i)
        if not agrees:
            break
    if not agrees:
        x = -x
    return utils.extract_x_from_expanded(x) if utils.is_expanded(matrix_a, vector_b) else x

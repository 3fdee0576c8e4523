"""Variational quantum linear solver with the ``vqls_prototype.VQLS`` surface the reference calls
(``VQLS(estimator=, ansatz=, optimizer=, sampler=, options=)`` and ``.solve(A, b)`` returning an
object with ``.state``; ``vqls_qiskit_implementation.py:55-65``).  ``vqls-prototype 0.2.0`` is a
third-party dependency that is neither in the reference tree nor installable here; its algorithm
is restated per SURVEY.md section 8a rows a10-a12:

* ``A = sum_i c_i A_i`` by the default "symmetric" decomposition: ``A_hat = A/|A|_F``,
  ``U_pm = A_hat +- i*sqrtm(I - A_hat^2)``, ``c = |A|_F / 2`` (an analogous pair with coefficient
  ``i*|A|_F/2`` when ``Im A != 0``);
* Hadamard tests with the ancilla on qubit 0 (``H``, optional ``Sdg``, controlled operators,
  ``H``; value = ``<Z_0>``; the real and the imaginary variant form one complex number);
* global cost ``1 - |sum_i c_i beta_i|^2 / sum_ij conj(c_i) c_j gamma_ij`` with
  ``beta_i = <0|Ub^dagger A_i V(theta)|0>`` and ``gamma_ij = <0|V^dagger A_i^dagger A_j V|0>``;
* a classical optimiser (COBYLA / SLSQP through SciPy) minimises the cost from a random initial
  point in [-pi, pi] (``qiskit_algorithms.utils.validate_initial_point``).

Every expectation value goes through ``estimator.run`` -- that is the hot loop the B200 batched
kernel replaces."""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as spla
import scipy.optimize as spopt

from .circuit import QuantumCircuit, QuantumRegister

# --------------------------------------------------------------------------------------
# optimisers (qiskit_algorithms.optimizers.COBYLA / SLSQP call shapes)
# --------------------------------------------------------------------------------------


class OptimizerResult:
    def __init__(self) -> None:
        self.x: Optional[np.ndarray] = None
        self.fun: Optional[float] = None
        self.nfev: Optional[int] = None
        self.nit: Optional[int] = None


class _SciPyOptimizer:
    method = ""

    def __init__(self, maxiter: int = 1000, disp: bool = False, tol: Optional[float] = None, **options: Any):
        self.options = {"maxiter": maxiter, "disp": disp, **options}
        self.tol = tol

    def minimize(self, fun: Callable, x0: np.ndarray, jac: Any = None, bounds: Any = None) -> OptimizerResult:
        kwargs: Dict[str, Any] = {}
        if self.method == "SLSQP" and bounds is not None:
            kwargs["bounds"] = bounds
        raw = spopt.minimize(fun, x0=np.asarray(x0, dtype=float), method=self.method, jac=jac, tol=self.tol,
                             options=self.options, **kwargs)
        res = OptimizerResult()
        res.x, res.fun, res.nfev, res.nit = raw.x, float(raw.fun), int(raw.get("nfev", 0)), raw.get("nit", None)
        return res


class COBYLA(_SciPyOptimizer):
    method = "COBYLA"

    def __init__(self, maxiter: int = 1000, disp: bool = False, rhobeg: float = 1.0, tol: Optional[float] = None):
        super().__init__(maxiter=maxiter, disp=disp, tol=tol, rhobeg=rhobeg)


class SLSQP(_SciPyOptimizer):
    method = "SLSQP"

    def __init__(self, maxiter: int = 100, disp: bool = False, ftol: float = 1e-06, tol: Optional[float] = None,
                 eps: float = 1.4901161193847656e-08):
        super().__init__(maxiter=maxiter, disp=disp, tol=tol, ftol=ftol, eps=eps)


# --------------------------------------------------------------------------------------
# matrix decomposition
# --------------------------------------------------------------------------------------


class UnitaryTerm:
    def __init__(self, coeff: complex, matrix: np.ndarray, name: str):
        self.coeff = complex(coeff)
        self.matrix = matrix
        nq = int(round(math.log2(matrix.shape[0])))
        self.circuit = QuantumCircuit(nq, name=name)
        self.circuit.unitary(matrix, list(range(nq)))


def symmetric_decomposition(matrix: np.ndarray) -> List[UnitaryTerm]:
    norm = np.linalg.norm(matrix)
    mat = matrix / norm
    mat_real, mat_imag = np.real(mat), np.imag(mat)
    terms: List[UnitaryTerm] = []

    def aux(x: np.ndarray) -> np.ndarray:
        return spla.sqrtm(np.eye(len(x)) - x @ x)

    if not np.allclose(mat_real, 0.0):
        a = aux(mat_real)
        terms.append(UnitaryTerm(0.5 * norm, mat_real + 1.0j * a, "b+"))
        terms.append(UnitaryTerm(0.5 * norm, mat_real - 1.0j * a, "b-"))
    if not np.allclose(mat_imag, 0.0):
        a = aux(mat_imag)
        terms.append(UnitaryTerm(0.5j * norm, mat_imag + 1.0j * a, "c+"))
        terms.append(UnitaryTerm(0.5j * norm, mat_imag - 1.0j * a, "c-"))
    return terms


# --------------------------------------------------------------------------------------
# Hadamard test
# --------------------------------------------------------------------------------------


class HadamardTest:
    """Real and imaginary Hadamard-test circuits for ``<psi| O_k ... O_1 |psi>``: ancilla = qubit 0,
    operators controlled on it, observable ``Z`` on the ancilla."""

    def __init__(self, operators: Sequence[QuantumCircuit], apply_initial_state: Optional[QuantumCircuit] = None):
        nq = operators[0].num_qubits
        if apply_initial_state is not None and apply_initial_state.num_qubits != nq:
            raise ValueError("The operator and the initial state circuits have different numbers of qubits")
        self.num_qubits = nq + 1
        self.observable = ("Z", 0)
        self.circuits: List[QuantumCircuit] = []
        for imaginary in (False, True):
            qc = QuantumCircuit(QuantumRegister(1, "qctl"), QuantumRegister(nq, "qreg"), name="hdmr_imag" if imaginary else "hdmr_real")
            qreg = list(range(1, nq + 1))
            if apply_initial_state is not None:
                qc.append(apply_initial_state.to_gate(), qreg)
            qc.h(0)
            if imaginary:
                qc.sdg(0)
            for op in operators:
                qc.append(op.control(1), [0] + qreg)
            qc.h(0)
            self.circuits.append(qc)

    def get_value(self, estimator: Any, parameter_sets: Sequence[float]) -> complex:
        n = len(self.circuits)
        job = estimator.run(self.circuits, [self.observable] * n, [parameter_sets] * n)
        vals = job.result().values
        return complex(vals[0], vals[1])


# --------------------------------------------------------------------------------------
# VQLS
# --------------------------------------------------------------------------------------


class VariationalLinearSolverResult:
    def __init__(self) -> None:
        self.state: Optional[QuantumCircuit] = None
        self.optimal_point: Optional[np.ndarray] = None
        self.optimal_parameters: Optional[dict] = None
        self.optimal_value: Optional[float] = None
        self.cost_function_evals: Optional[int] = None


class VQLS:
    def __init__(self, estimator: Any, ansatz: QuantumCircuit, optimizer: Any, sampler: Any = None,
                 initial_point: Optional[np.ndarray] = None, options: Optional[dict] = None,
                 callback: Optional[Callable] = None, seed: Optional[int] = None):
        self.estimator = estimator
        self.ansatz = ansatz
        self.optimizer = optimizer
        self.sampler = sampler
        self.initial_point = initial_point
        self.callback = callback
        self.options = {"use_overlap_test": False, "use_local_cost_function": False,
                        "matrix_decomposition": "symmetric", "shots": None, "reuse_matrix": False, "verbose": False}
        self.options.update(options or {})
        if self.options["use_overlap_test"] or self.options["use_local_cost_function"]:
            raise NotImplementedError("only the global Hadamard-test cost (the reference's options) is supported")
        if self.options["matrix_decomposition"] != "symmetric":
            raise NotImplementedError("only the default symmetric matrix decomposition is supported")
        self._rng = np.random.default_rng(seed)
        self._eval_count = 0
        self.matrix_terms: List[UnitaryTerm] = []
        self.vector_circuit: Optional[QuantumCircuit] = None

    # -- circuits -------------------------------------------------------------------------
    def construct_circuit(self, matrix: np.ndarray, vector: np.ndarray) -> Tuple[List[HadamardTest], List[HadamardTest]]:
        vector = np.asarray(vector).astype("float64").reshape(-1)
        nb = int(np.log2(len(vector)))
        vec_norm = np.linalg.norm(vector)
        if vec_norm == 0:
            raise ValueError("Norm of b vector is null!")
        self.vector_circuit = QuantumCircuit(nb, name="Ub")
        self.vector_circuit.prepare_state(vector / vec_norm)
        matrix = np.asarray(matrix)
        if matrix.shape[0] != matrix.shape[1]:
            raise ValueError("Input matrix must be square!")
        if np.log2(matrix.shape[0]) % 1 != 0:
            raise ValueError("Input matrix dimension must be 2^N!")
        if not np.allclose(matrix, matrix.conj().T):
            raise ValueError("Input matrix must be hermitian!")
        if matrix.shape[0] != 2**nb:
            raise ValueError("Input vector dimension does not match input matrix dimension!")
        if self.ansatz.num_qubits != nb:
            raise ValueError("Ansatz and problem have different numbers of qubits")
        self.matrix_terms = symmetric_decomposition(matrix.astype(complex) if np.iscomplexobj(matrix) else matrix.astype("float64"))
        m = len(self.matrix_terms)
        norm_tests = []
        for i in range(m):
            for j in range(i + 1, m):
                norm_tests.append(HadamardTest(
                    [self.matrix_terms[i].circuit.inverse(), self.matrix_terms[j].circuit], apply_initial_state=self.ansatz))
        ub_dagger = self._inverse_state_preparation()
        overlap_tests = [HadamardTest([self.ansatz, t.circuit, ub_dagger]) for t in self.matrix_terms]
        return norm_tests, overlap_tests

    def _inverse_state_preparation(self) -> QuantumCircuit:
        """``Ub^dagger`` as one unitary block (the inverse of an isometry is not an initialisation)."""
        prep = self.vector_circuit.data[0].operation
        qc = QuantumCircuit(self.vector_circuit.num_qubits, name="Ub_dg")
        qc.unitary(prep.to_matrix().conj().T, list(range(qc.num_qubits)))
        return qc

    # -- cost -----------------------------------------------------------------------------
    @staticmethod
    def get_coefficient_matrix(coeffs: np.ndarray) -> np.ndarray:
        return coeffs[:, None].conj() @ coeffs[None, :]

    @staticmethod
    def _normalization_term(coeff_matrix: np.ndarray, hdmr_values: np.ndarray) -> complex:
        out = (hdmr_values * coeff_matrix[np.triu_indices_from(coeff_matrix, k=1)]).sum()
        out = out + np.conj(out)
        out = out + coeff_matrix.diagonal().sum()  # <0|V^+ A_i^+ A_i V|0> = 1
        return complex(out)

    @staticmethod
    def _global_terms(coeff_matrix: np.ndarray, hdmr_values: np.ndarray) -> complex:
        return complex((coeff_matrix * np.outer(hdmr_values.conj(), hdmr_values)).sum())

    def _assemble_cost(self, norm_values: np.ndarray, overlap_values: np.ndarray, coeff_matrix: np.ndarray) -> float:
        norm = self._normalization_term(coeff_matrix, norm_values)
        return float(1.0 - np.real(self._global_terms(coeff_matrix, overlap_values) / norm))

    def get_cost_evaluation_function(self, norm_tests: List[HadamardTest], overlap_tests: List[HadamardTest],
                                     coeff_matrix: np.ndarray) -> Callable[[np.ndarray], float]:
        def cost_evaluation(parameters: np.ndarray) -> float:
            theta = np.asarray(parameters, dtype=float).reshape(-1)
            norm_values = np.array([t.get_value(self.estimator, theta) for t in norm_tests], dtype=complex)
            overlap_values = np.array([t.get_value(self.estimator, theta) for t in overlap_tests], dtype=complex)
            cost = self._assemble_cost(norm_values, overlap_values, coeff_matrix)
            self._eval_count += 1
            if self.callback is not None:
                self.callback(self._eval_count, cost, theta)
            return cost

        return cost_evaluation

    # -- solve ----------------------------------------------------------------------------
    def solve(self, matrix: np.ndarray, vector: np.ndarray) -> VariationalLinearSolverResult:
        norm_tests, overlap_tests = self.construct_circuit(matrix, vector)
        coeff_matrix = self.get_coefficient_matrix(np.array([t.coeff for t in self.matrix_terms]))
        num_params = self.ansatz.num_parameters
        if self.initial_point is None:
            initial_point = self._rng.uniform(-np.pi, np.pi, num_params)
        else:
            initial_point = np.asarray(self.initial_point, dtype=float)
            if initial_point.size != num_params:
                raise ValueError("initial_point has the wrong length")
        bounds = [(-np.pi, np.pi)] * num_params
        self._eval_count = 0
        cost = self.get_cost_evaluation_function(norm_tests, overlap_tests, coeff_matrix)
        if callable(self.optimizer):
            opt = self.optimizer(fun=cost, x0=initial_point, jac=None, bounds=bounds)
        else:
            opt = self.optimizer.minimize(fun=cost, x0=initial_point, jac=None, bounds=bounds)
        sol = VariationalLinearSolverResult()
        sol.optimal_point = opt.x
        sol.optimal_parameters = dict(zip(self.ansatz.parameters, opt.x))
        sol.optimal_value = opt.fun
        sol.cost_function_evals = opt.nfev
        sol.state = self.ansatz.assign_parameters(sol.optimal_parameters)
        return sol

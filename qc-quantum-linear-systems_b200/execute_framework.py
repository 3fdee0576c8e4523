"""Command line with the flags of the reference's ``quantum_linear_systems/execute_framework.py``:

    python -m qls_b200.execute_framework -i hhl_qiskit [-m A.csv] [-v b.csv] [-iargs args.yaml] [-y]

``-m`` / ``-v`` default to the ClassiqDemoExample toy model, ``-iargs`` is a YAML mapping forwarded as keyword
arguments to the implementation.  Like the reference, the tool prints the matrix checks and asks on stdin before it
solves; ``-y`` skips the question (batch jobs).  ``hhl_classiq`` is a known name but needs Classiq's cloud service."""
from __future__ import annotations

import argparse
import sys
from typing import Any, Dict, List, Optional

import numpy as np

from .quantum_linear_solver import QuantumLinearSolver
from .toymodels import ClassiqDemoExample

# spelling accepted on the command line -> method name of QuantumLinearSolver.solve (the reference's parser only
# knows the misspelt "vqls_qiksit"; both are taken here)
METHODS: Dict[str, str] = {"hhl_qiskit": "hhl_qiskit", "vqls_qiskit": "vqls_qiskit", "vqls_qiksit": "vqls_qiskit",
                           "hhl_classiq": "hhl_classiq"}

_FLAGS = (
    (("-m", "--matrix_csv"), dict(type=str, help="comma-separated file holding A")),
    (("-v", "--vector_csv"), dict(type=str, help="comma-separated file holding b")),
    (("-i", "--implementation"), dict(type=str, required=True, help="one of: " + ", ".join(sorted(set(METHODS.values()))))),
    (("-iargs", "--implementation_args"), dict(type=str, help="YAML file with keyword arguments for the implementation")),
    (("-y", "--yes"), dict(action="store_true", help="solve without asking")),
)


def _csv_or_default(path: Optional[str], what: str, default: np.ndarray) -> np.ndarray:
    if path:
        return np.loadtxt(path, delimiter=",")
    print(f"No {what} provided. Falling back to ToyModel.")
    return default


def _yaml_kwargs(path: Optional[str]) -> Dict[str, Any]:
    if not path:
        return {}
    import yaml

    try:
        with open(path, encoding="utf-8") as fh:
            loaded = yaml.safe_load(fh)
    except yaml.YAMLError as exc:
        raise ValueError(f"Error loading YAML argument file: {exc}") from exc
    if loaded is not None and not isinstance(loaded, dict):
        raise ValueError("the YAML argument file must hold a mapping")
    return loaded or {}


def parse_arguments(argv: Optional[List[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description="Quantum Linear Systems Framework (B200 statevector engine)")
    for names, kw in _FLAGS:
        parser.add_argument(*names, **kw)
    ns = parser.parse_args(argv)
    if ns.implementation not in METHODS:
        raise ValueError(f"Unknown implementation {ns.implementation}.")
    ns.implementation = METHODS[ns.implementation]
    fallback = ClassiqDemoExample()
    ns.matrix_a = _csv_or_default(ns.matrix_csv, "matrix", fallback.matrix_a)
    ns.vector_b = _csv_or_default(ns.vector_csv, "vector", fallback.vector_b)
    ns.implementation_args = _yaml_kwargs(ns.implementation_args)
    return ns


def _report_checks(solver: QuantumLinearSolver, a: np.ndarray) -> None:
    solver.check_matrix_square_hermitian(a)
    sparsity, kappa = solver.check_matrix_sparsity(a), solver.check_matrix_condition_number(a)
    print(f"\n\n###\nSuccessfully checked matrix A.\nSparsity: {sparsity}\nCondition number: {kappa}\n"
          f"Number of matrix elements: {a.shape[0]}x{a.shape[1]}={a.size}")


def main(argv: Optional[List[str]] = None) -> Optional[QuantumLinearSolver]:
    ns = parse_arguments(argv)
    solver = QuantumLinearSolver()
    _report_checks(solver, ns.matrix_a)
    if not ns.yes:
        answer = input("\nDo you want to continue with solving the linear system? (y/n): ")
        if answer.strip().lower() == "n":
            return None
    solver.solve(ns.matrix_a, ns.vector_b, method=ns.implementation, file_basename="default_run", **ns.implementation_args)
    for label, value in (("Solution", solver.solution), ("Circuit depth", solver.circuit_depth),
                         ("Circuit width", solver.circuit_width), ("Runtime", solver.run_time)):
        print(("\n###\n" if label == "Solution" else "") + f"{label}: {value}")
    return solver


if __name__ == "__main__":
    sys.exit(0 if main() is not None else 1)

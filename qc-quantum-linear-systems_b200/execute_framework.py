"""Command line of the reference (``quantum_linear_systems/execute_framework.py``), same flags:

    python -m qls_b200.execute_framework -i hhl_qiskit [-m A.csv] [-v b.csv] [-iargs args.yaml] [-y]

Without ``-m`` / ``-v`` the ClassiqDemoExample toy model is used.  The reference asks for a
confirmation on stdin after printing the matrix checks; ``-y`` answers it (needed for batch jobs).
``hhl_classiq`` is accepted by the parser like in the reference but needs Classiq's cloud service."""
from __future__ import annotations

import argparse
import sys
from typing import Any, List, Optional

import numpy as np

from .quantum_linear_solver import QuantumLinearSolver
from .toymodels import ClassiqDemoExample

IMPLEMENTATIONS = ["hhl_qiskit", "vqls_qiskit", "vqls_qiksit", "hhl_classiq"]  # the reference's spelling is accepted too


def parse_arguments(argv: Optional[List[str]] = None) -> argparse.Namespace:
    ap = argparse.ArgumentParser(description="Quantum Linear Systems Framework (B200 statevector engine)")
    ap.add_argument("-m", "--matrix_csv", type=str, required=False, help="CSV file containing the matrix A.")
    ap.add_argument("-v", "--vector_csv", type=str, required=False, help="CSV file containing the vector b.")
    ap.add_argument("-i", "--implementation", type=str, required=True, help="Implementation to solve the problem Ax=b.")
    ap.add_argument("-iargs", "--implementation_args", type=str, required=False,
                    help="Path to a YAML file containing parameters to be passed to the implementation.")
    ap.add_argument("-y", "--yes", action="store_true", help="do not ask before solving")
    args = ap.parse_args(argv)
    toy = ClassiqDemoExample()
    if args.matrix_csv:
        args.matrix_a = np.loadtxt(args.matrix_csv, delimiter=",")
    else:
        print("No matrix provided. Falling back to ToyModel.")
        args.matrix_a = toy.matrix_a
    if args.vector_csv:
        args.vector_b = np.loadtxt(args.vector_csv, delimiter=",")
    else:
        print("No vector provided. Falling back to ToyModel.")
        args.vector_b = toy.vector_b
    if args.implementation not in IMPLEMENTATIONS:
        raise ValueError(f"Unknown implementation {args.implementation}.")
    if args.implementation == "vqls_qiksit":
        args.implementation = "vqls_qiskit"
    if args.implementation_args:
        import yaml

        with open(args.implementation_args, "r", encoding="utf-8") as fh:
            try:
                args.implementation_args = yaml.safe_load(fh) or {}
            except yaml.YAMLError as exc:
                raise ValueError(f"Error loading YAML argument file: {exc}") from exc
    else:
        args.implementation_args = {}
    return args


def main(argv: Optional[List[str]] = None) -> Any:
    args = parse_arguments(argv)
    qls = QuantumLinearSolver()
    qls.check_matrix_square_hermitian(args.matrix_a)
    s = qls.check_matrix_sparsity(args.matrix_a)
    k = qls.check_matrix_condition_number(args.matrix_a)
    rows, cols = args.matrix_a.shape
    print(f"\n\n###\nSuccessfully checked matrix A.\nSparsity: {s}\nCondition number: {k}\n"
          f"Number of matrix elements: {rows}x{cols}={rows * cols}")
    if not args.yes and input("\nDo you want to continue with solving the linear system? (y/n): ").lower() == "n":
        return None
    qls.solve(matrix_a=args.matrix_a, vector_b=args.vector_b, method=args.implementation, file_basename="default_run",
              **args.implementation_args)
    print(f"\n###\nSolution: {qls.solution}\nCircuit depth: {qls.circuit_depth}\nCircuit width: {qls.circuit_width}\n"
          f"Runtime: {qls.run_time}")
    return qls


if __name__ == "__main__":
    sys.exit(0 if main() is not None else 1)

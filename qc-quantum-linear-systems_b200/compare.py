"""Comparison harness on the B200 engine: the reference's ``solve_models`` loop and CSV schema
(``quantum_linear_systems/compare_classiq_qiskit.py:30-116``: one row per model with ``model_name,
solver, quantum_solution, classical_solution, circuit_depth, run_time, relative_distance``) so that
existing studies can be re-run unchanged.  Plotting and the Classiq solver are out of scope."""
from __future__ import annotations

import csv
from datetime import datetime
from typing import Any, List, Sequence, Tuple

import numpy as np

from .quantum_linear_solver import QuantumLinearSolver
from .utils import relative_distance_quantum_classical_solution

CSV_HEADERS = ["model_name", "solver", "quantum_solution", "classical_solution", "circuit_depth", "run_time",
               "relative_distance"]


def append_to_csv(filename: str, data: Sequence[Any]) -> None:
    with open(filename, mode="a", newline="", encoding="utf-8") as fh:
        csv.writer(fh).writerow(data)


def solve_models(models: Sequence[Any], method: str, save_file: str, **kwargs: Any) -> Tuple[List[np.ndarray], List[np.ndarray],
                                                                                             Tuple[List[int], List[float], List[float]]]:
    """Solve every model with ``method``; both solutions are normalised to unit length before the
    relative distance (percent) is taken, as in the reference."""
    qsols: List[np.ndarray] = []
    csols: List[np.ndarray] = []
    depths: List[int] = []
    times: List[float] = []
    dists: List[float] = []
    for model in models:
        qls = QuantumLinearSolver()
        q = qls.solve(matrix_a=model.matrix_a, vector_b=model.vector_b, method=method, file_basename=model.name, **kwargs)
        q = q / np.linalg.norm(q)
        c = model.classical_solution / np.linalg.norm(model.classical_solution)
        d = relative_distance_quantum_classical_solution(quantum_solution=q, classical_solution=c)
        qsols.append(q)
        csols.append(c)
        depths.append(qls.circuit_depth)
        times.append(qls.run_time)
        dists.append(d)
        append_to_csv(save_file, [model.name, method, q, c, qls.circuit_depth, qls.run_time, d])
    return qsols, csols, (depths, times, dists)


def compare_qls(models: Sequence[Any], methods: Sequence[str] = ("hhl_qiskit",), filebasename: str = "comparison") -> str:
    """Write ``<filebasename>_<timestamp>.csv`` with one row per (model, method); returns the file name."""
    csv_file = f"{filebasename}_{datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}.csv"
    with open(csv_file, mode="w", newline="", encoding="utf-8") as fh:
        csv.writer(fh).writerow(CSV_HEADERS)
    for method in methods:
        solve_models(models, method, csv_file)
    return csv_file

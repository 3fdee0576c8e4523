"""``B200Estimator``: the object that goes into ``solve_vqls_qiskit(estimator=...)``.

The reference's plug-in hole for the VQLS hot loop is the ``estimator=`` keyword
(``vqls_qiskit_implementation.py:28,56``; the AWS variant swaps in a ``BackendEstimator`` through
it, ``execution/aws_execution.py:174-182``).  This class has the ``BaseEstimator`` call shape the
reference relies on: ``run(circuits, observables, parameter_values, **opts)`` returns a job whose
``result().values`` is a float64 array in submission order, exact (``shots=None``).

Execution: entries that share a circuit object and an observable are evaluated by ONE launch of
the batched kernel (``qls_batch_expect_z``: one CTA per parameter vector, whole state in shared
memory) with a program lowered once per circuit and cached; circuits too wide for shared memory
go through the streaming engine one by one.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .circuit import QuantumCircuit
from .lower import LoweringOptions, lower_circuit


class EstimatorResult:
    def __init__(self, values: np.ndarray, metadata: List[dict]):
        self.values = values
        self.metadata = metadata


class _Job:
    def __init__(self, result: EstimatorResult):
        self._result = result

    def result(self) -> EstimatorResult:
        return self._result


def observable_z_qubit(observable: Any, num_qubits: int) -> int:
    """Qubit index of a single-qubit Z observable given as ``("Z", q)``, a Pauli label
    (rightmost character = qubit 0), or a qiskit ``SparsePauliOp``/``Pauli`` with one Z."""
    if isinstance(observable, tuple) and len(observable) == 2 and str(observable[0]).upper() == "Z":
        return int(observable[1])
    label = None
    if isinstance(observable, str):
        label = observable
    elif hasattr(observable, "paulis") and hasattr(observable, "coeffs"):
        if len(observable.paulis) != 1 or abs(complex(observable.coeffs[0]) - 1) > 1e-12:
            raise NotImplementedError("only a single Pauli-Z term with coefficient 1 is supported")
        label = observable.paulis[0].to_label()
    elif hasattr(observable, "to_label"):
        label = observable.to_label()
    if label is not None:
        label = label.upper()
        if len(label) != num_qubits or set(label) - {"I", "Z"} or label.count("Z") != 1:
            raise NotImplementedError(f"unsupported observable {label!r}: need one Z on a {num_qubits}-qubit register")
        return len(label) - 1 - label.index("Z")
    raise NotImplementedError(f"unsupported observable type {type(observable)!r}")


class BatchProgram:
    """A parameterised circuit lowered for the batched kernel and uploaded to one GPU."""

    def __init__(self, circuit: QuantumCircuit, precision: str = "fp64", device: Optional[int] = None):
        from .engine import DeviceProgram

        self.num_qubits = circuit.num_qubits
        self.num_parameters = circuit.num_parameters
        fuse_q = int(os.environ.get("QLS_BATCH_FUSE", "2"))
        self.lowered = lower_circuit(circuit, LoweringOptions(dtype=precision, batch=True, max_fuse_qubits=fuse_q))
        if self.lowered.n_passes != 1:
            raise _abi.QlsError("batch lowering must produce exactly one pass")
        self.program = DeviceProgram(self.lowered, device)
        self.device = self.program.device

    def expect_z(self, z_qubit: int, params) -> "Any":
        """``params``: float64 CUDA tensor ``[K, P]`` (or ``[K, 0]``); returns a float64 CUDA tensor ``[K]``."""
        import torch

        k = int(params.shape[0])
        out = torch.empty(k, dtype=torch.float64, device=params.device)
        with torch.cuda.device(self.device):
            _abi.check(self.program.lib.qls_batch_expect_z(
                self.program.handle, 0, self.num_qubits, int(z_qubit), params.data_ptr() if params.numel() else None,
                k, self.num_parameters, out.data_ptr(), int(torch.cuda.current_stream().cuda_stream)))
        return out


class B200Estimator:
    MAX_BATCH_QUBITS = {"fp64": 12, "fp32": 13}

    def __init__(self, precision: str = "fp64", device: Optional[int] = None, options: Optional[dict] = None):
        self.precision = precision
        self.device = device
        self.options = dict(options or {})
        self._cache: Dict[int, Tuple[QuantumCircuit, BatchProgram]] = {}
        self.launches = 0

    def _program(self, circuit: QuantumCircuit) -> BatchProgram:
        hit = self._cache.get(id(circuit))
        if hit is not None and hit[0] is circuit:
            return hit[1]
        prog = BatchProgram(circuit, self.precision, self.device)
        self._cache[id(circuit)] = (circuit, prog)
        return prog

    def run(self, circuits: Sequence[QuantumCircuit], observables: Sequence[Any],
            parameter_values: Optional[Sequence[Sequence[float]]] = None, **run_options: Any) -> _Job:
        import torch

        if isinstance(circuits, QuantumCircuit):
            circuits = [circuits]
        if not isinstance(observables, (list, tuple)):
            observables = [observables]
        n = len(circuits)
        if len(observables) != n:
            raise ValueError("circuits and observables must have the same length")
        if parameter_values is None:
            parameter_values = [[] for _ in range(n)]
        if len(parameter_values) != n:
            raise ValueError("circuits and parameter_values must have the same length")
        values = np.empty(n, dtype=np.float64)
        # group submissions by (circuit object, observable qubit), keeping submission order
        groups: Dict[Tuple[int, int], List[int]] = {}
        for i, (c, ob) in enumerate(zip(circuits, observables)):
            q = observable_z_qubit(ob, c.num_qubits)
            pv = np.asarray(parameter_values[i], dtype=np.float64).reshape(-1)
            if pv.size != c.num_parameters:
                raise ValueError(f"circuit {i} has {c.num_parameters} parameters, got {pv.size} values")
            groups.setdefault((id(c), q), []).append(i)
        for (_, q), idxs in groups.items():
            circ = circuits[idxs[0]]
            theta = np.stack([np.asarray(parameter_values[i], dtype=np.float64).reshape(-1) for i in idxs])
            if circ.num_qubits <= self.MAX_BATCH_QUBITS[self.precision]:
                prog = self._program(circ)
                dev = f"cuda:{prog.device}"
                t = torch.from_numpy(theta).to(dev) if theta.size else torch.empty((len(idxs), 0), dtype=torch.float64, device=dev)
                out = prog.expect_z(q, t)
                self.launches += 1
                values[idxs] = out.cpu().numpy()
            else:
                from .statevector import Statevector

                for i, th in zip(idxs, theta):
                    bound = circ.assign_parameters(th) if circ.num_parameters else circ
                    values[i] = Statevector(bound, precision=self.precision, device=self.device).expectation_z(q)
        return _Job(EstimatorResult(values, [{} for _ in range(n)]))

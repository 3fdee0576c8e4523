"""Adapter for real qiskit circuits (used only when qiskit is importable; it is not in the build
image, so this module never imports it and is tested with duck-typed stand-ins, tests/test_qiskit_adapter.py).

``from_qiskit`` walks ``QuantumCircuit.data`` the way ``Statevector._evolve_instruction`` does
(reference call sites ``hhl_qiskit_implementation.py:49``, ``vqls_qiskit_implementation.py:65``):
an operation with ``to_matrix()`` becomes a ``unitary`` leaf on its qubits, otherwise its
``definition`` is walked with remapped qubits and its ``global_phase`` accumulated.  ``Statevector``
here is a drop-in for ``qiskit.quantum_info.Statevector(circuit)``: ``.data`` is evolved on the B200.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .circuit import QuantumCircuit
from .statevector import Statevector as _EngineStatevector


def from_qiskit(circuit: Any) -> QuantumCircuit:
    out = QuantumCircuit(circuit.num_qubits, name=getattr(circuit, "name", "circuit"),
                         global_phase=float(circuit.global_phase))
    index = {q: i for i, q in enumerate(circuit.qubits)}

    def walk(circ: Any, qmap) -> None:
        local = {q: i for i, q in enumerate(circ.qubits)}
        for ins in circ.data:
            op = ins.operation
            qs = [qmap[local[q]] for q in ins.qubits]
            if op.name == "barrier":
                continue
            mat = None
            try:
                mat = op.to_matrix()
            except Exception:  # qiskit raises CircuitError / QiskitError when there is no matrix
                mat = None
            if mat is not None:
                out.unitary(np.asarray(mat, dtype=complex), qs, label=op.name)
                continue
            if op.definition is None:
                raise ValueError(f"instruction {op.name!r} has neither a matrix nor a definition")
            out.global_phase += float(op.definition.global_phase)
            walk(op.definition, qs)

    walk(circuit, [index[q] for q in circuit.qubits])
    return out


class Statevector(_EngineStatevector):
    def __init__(self, data: Any, **kwargs: Any):
        if hasattr(data, "qubits") and hasattr(data, "data") and not isinstance(data, QuantumCircuit):
            data = from_qiskit(data)
        super().__init__(data, **kwargs)

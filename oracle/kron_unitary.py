"""Dense-unitary oracle (n <= ~11): builds the full ``2**n x 2**n`` matrix of each gate
application by explicit index arithmetic and multiplies it onto the state.

TEST INFRASTRUCTURE ONLY (see ``oracle/numpy_sv.py``).  Shares no code with
``numpy_sv.evolve_operator`` so that the two restatements of qiskit's little-endian gate
convention check each other.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def embed(mat: np.ndarray, qargs: Sequence[int], num_qubits: int) -> np.ndarray:
    """Full-register matrix of ``mat`` acting on ``qargs`` (``qargs[0]`` = LSB of ``mat``'s index)."""
    n = num_qubits
    k = len(qargs)
    dim = 2**n
    full = np.zeros((dim, dim), dtype=complex)
    mask = 0
    for q in qargs:
        mask |= 1 << q
    for col in range(dim):
        rest = col & ~mask
        sub_c = 0
        for j, q in enumerate(qargs):
            sub_c |= ((col >> q) & 1) << j
        for sub_r in range(2**k):
            row = rest
            for j, q in enumerate(qargs):
                row |= ((sub_r >> j) & 1) << q
            full[row, col] = mat[sub_r, sub_c]
    return full


def run(num_qubits: int, stream, psi0=None) -> np.ndarray:
    psi = np.zeros(2**num_qubits, dtype=complex)
    psi[0] = 1.0
    if psi0 is not None:
        psi = np.array(psi0, dtype=complex)
    for item in stream:
        if item[0] == "u":
            psi = embed(np.asarray(item[1]), item[2], num_qubits) @ psi
        elif item[0] == "phase":
            psi = psi * np.exp(1j * float(item[1]))
        elif item[0] == "init":
            amps, qargs = np.asarray(item[1], dtype=complex), list(item[2])
            new = np.zeros_like(psi)
            mask = sum(1 << q for q in qargs)
            for i in range(psi.size):
                if i & mask:
                    if abs(psi[i]) > 1e-12:
                        raise ValueError("init on a register that is not |0>")
                    continue
                for s, a in enumerate(amps):
                    j = i
                    for b, q in enumerate(qargs):
                        j |= ((s >> b) & 1) << q
                    new[j] = a * psi[i]
            psi = new
        else:
            raise ValueError(item[0])
    return psi

"""Synthetic HHL-shaped circuits (BASELINE.json configs 4-5).

Same structure as ``HHL.construct_circuit`` -- state preparation, QPE (H layer, one controlled
block per clock qubit, inverse QFT without swaps, bit-reversed), ``ExactReciprocal``, inverse QPE
-- but the controlled ``e^{iAt 2^j}`` of clock qubit ``j`` is a seeded random 5-qubit unitary on
data qubits followed by a random diagonal phase run, because a dense ``2^nb x 2^nb`` matrix stops
being representable long before the state stops fitting in HBM."""
from __future__ import annotations

import numpy as np

from .circuit import QuantumCircuit, QuantumRegister
from .library import ExactReciprocal, PhaseEstimation


def haar_unitary(k: int, rng: np.random.Generator) -> np.ndarray:
    dim = 2**k
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


class SyntheticEvolution:
    """Stand-in for ``NumPyMatrix``: ``power(2**j)`` is block ``j``."""

    def __init__(self, num_qubits: int, num_blocks: int, seed: int = 2024, dense_qubits: int = 5, diag_qubits: int = 6):
        self.num_qubits = self.num_state_qubits = num_qubits
        self.num_ancillas = 0
        self._blocks = []
        for j in range(num_blocks):
            rng = np.random.default_rng(seed + j)
            kd = min(dense_qubits, num_qubits)
            kg = min(diag_qubits, num_qubits)
            dq = sorted(int(q) for q in rng.choice(num_qubits, kd, replace=False))
            gq = sorted(int(q) for q in rng.choice(num_qubits, kg, replace=False))
            self._blocks.append((dq, haar_unitary(kd, rng), gq, np.exp(1j * rng.uniform(-np.pi, np.pi, 2**kg))))

    def power(self, power: int) -> QuantumCircuit:
        j = int(round(np.log2(power)))
        dq, u, gq, d = self._blocks[j]
        qc = QuantumCircuit(self.num_qubits, name=f"U**{power}")
        qc.unitary(u, dq)
        qc.diagonal(d, gq)
        return qc


def hhl_shaped_circuit(num_qubits: int, seed: int = 2024, b_seed: int = 99, nb: int = 0) -> QuantumCircuit:
    """``num_qubits = nb + nl + 1``; default split ``nb = (n-1)//2`` (n = 33 -> nb = nl = 16)."""
    n = int(num_qubits)
    nb = nb or (n - 1) // 2
    nl = n - 1 - nb
    if nb < 1 or nl < 2:
        raise ValueError("need at least 1 data qubit and 2 clock qubits")
    rng = np.random.default_rng(b_seed)
    b = rng.standard_normal(2**nb)
    b /= np.linalg.norm(b)
    qb, ql, qf = QuantumRegister(nb, "q_b"), QuantumRegister(nl, "q_l"), QuantumRegister(1, "q_f")
    qc = QuantumCircuit(qb, ql, qf, name="hhl_shaped")
    qb_q, ql_q, qf_q = qc.register_qubits(0), qc.register_qubits(1), qc.register_qubits(2)
    prep = QuantumCircuit(nb, name="isometry")
    prep.prepare_state(b, list(range(nb)))
    qc.append(prep.to_gate(), qb_q)
    qpe = PhaseEstimation(nl, SyntheticEvolution(nb, nl, seed))
    qc.append(qpe.to_gate(), ql_q + qb_q)
    delta = 2.0 ** (-(nl // 2))
    qc.append(ExactReciprocal(nl, delta, neg_vals=True).to_gate(), ql_q[::-1] + [qf_q[0]])
    qc.append(qpe.inverse().to_gate(), ql_q + qb_q)
    return qc

"""Synthetic HHL-shaped circuits (BASELINE.json configs 4-5).

Same structure as ``HHL.construct_circuit`` -- state preparation, QPE (H layer, one controlled
block per clock qubit, inverse QFT without swaps, bit-reversed), ``ExactReciprocal``, inverse QPE
-- but the controlled ``e^{iAt 2^j}`` of clock qubit ``j`` is a seeded Trotter-like step: a chain
of Haar-random two-qubit unitaries over five data qubits followed by a random diagonal phase run
on six, because a dense ``2^nb x 2^nb`` matrix stops being representable long before the state
stops fitting in HBM.  (``block="dense5"`` replaces the chain by one Haar-random five-qubit
unitary: 256 flop per 32 bytes, beyond the FP64 ridge of a B200 -- kept as a stress case.)"""
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

    def __init__(self, num_qubits: int, num_blocks: int, seed: int = 2024, dense_qubits: int = 5, diag_qubits: int = 6,
                 block: str = "chain2"):
        self.num_qubits = self.num_state_qubits = num_qubits
        self.num_ancillas = 0
        self._blocks = []
        for j in range(num_blocks):
            rng = np.random.default_rng(seed + j)
            kd = min(dense_qubits, num_qubits)
            kg = min(diag_qubits, num_qubits)
            dq = sorted(int(q) for q in rng.choice(num_qubits, kd, replace=False))
            gq = sorted(int(q) for q in rng.choice(num_qubits, kg, replace=False))
            if block == "dense5" or kd < 2:
                gates = [(dq, haar_unitary(kd, rng))]
            elif block == "chain2":
                gates = [([dq[i], dq[i + 1]], haar_unitary(2, rng)) for i in range(kd - 1)]
            else:
                raise ValueError(block)
            self._blocks.append((gates, gq, np.exp(1j * rng.uniform(-np.pi, np.pi, 2**kg))))

    def power(self, power: int) -> QuantumCircuit:
        j = int(round(np.log2(power)))
        gates, gq, d = self._blocks[j]
        qc = QuantumCircuit(self.num_qubits, name=f"U**{power}")
        for qs, u in gates:
            qc.unitary(u, qs)
        qc.diagonal(d, gq)
        return qc


def hhl_shaped_circuit(num_qubits: int, seed: int = 2024, b_seed: int = 99, nb: int = 0,
                       block: str = "chain2") -> QuantumCircuit:
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
    qpe = PhaseEstimation(nl, SyntheticEvolution(nb, nl, seed, block=block))
    qc.append(qpe.to_gate(), ql_q + qb_q)
    delta = 2.0 ** (-(nl // 2))
    qc.append(ExactReciprocal(nl, delta, neg_vals=True).to_gate(), ql_q[::-1] + [qf_q[0]])
    qc.append(qpe.inverse().to_gate(), ql_q + qb_q)
    return qc

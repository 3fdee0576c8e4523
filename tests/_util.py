"""Test helpers: turn the product's flattened leaf stream into the plain ``(tag, matrix, qargs)``
stream the oracle consumes, using explicit full matrices (controls expanded), so that the oracle
never sees product data structures."""
import math

import numpy as np


from oracle.stream import controlled_matrix, leaves_to_stream, ucry_matrix  # noqa: F401,E402


def circuit_to_stream(circuit):
    from qls_b200.circuit import flatten

    leaves, phase = flatten(circuit)
    return leaves_to_stream(leaves, phase)


def random_unitary(k, rng):
    dim = 2**k
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_circuit(n, n_gates, rng, max_k=3, with_controls=True):
    """Random mix of every leaf kind the engine supports."""
    from qls_b200.circuit import Gate, QuantumCircuit, UCRYGate

    qc = QuantumCircuit(n, name="rand", global_phase=float(rng.uniform(-1, 1)))
    names1 = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx"]
    for _ in range(n_gates):
        kind = rng.integers(0, 10)
        if kind == 0:
            qc._g(names1[rng.integers(len(names1))], [int(rng.integers(n))])
        elif kind == 1:
            nm = ["rx", "ry", "rz", "p"][rng.integers(4)]
            qc._g(nm, [int(rng.integers(n))], [float(rng.uniform(-3, 3))])
        elif kind == 2 and n >= 2:
            a, b = rng.choice(n, 2, replace=False)
            qc.cx(int(a), int(b))
        elif kind == 3 and n >= 2:
            a, b = rng.choice(n, 2, replace=False)
            qc.cp(float(rng.uniform(-3, 3)), int(a), int(b))
        elif kind == 4:
            k = int(rng.integers(1, min(max_k, n) + 1))
            qs = [int(q) for q in rng.choice(n, k, replace=False)]
            qc.unitary(random_unitary(k, rng), qs)
        elif kind == 5 and n >= 3 and with_controls:
            k = int(rng.integers(1, min(max_k, n - 1) + 1))
            nc = int(rng.integers(1, min(2, n - k) + 1))
            qs = [int(q) for q in rng.choice(n, k + nc, replace=False)]
            g = Gate("unitary", k, [], random_unitary(k, rng)).control(nc, int(rng.integers(2**nc)))
            qc.append(g, qs)
        elif kind == 6 and n >= 2:
            nc = int(rng.integers(1, min(4, n - 1) + 1))
            qs = [int(q) for q in rng.choice(n, nc + 1, replace=False)]
            qc.append(UCRYGate(rng.uniform(-3, 3, 2**nc)), qs)
        elif kind == 7 and n >= 2:
            k = int(rng.integers(1, min(4, n) + 1))
            qs = [int(q) for q in rng.choice(n, k, replace=False)]
            qc.diagonal(np.exp(1j * rng.uniform(-3, 3, 2**k)), qs)
        elif kind == 8 and n >= 3 and with_controls:
            qs = [int(q) for q in rng.choice(n, 3, replace=False)]
            qc.append(UCRYGate(rng.uniform(-3, 3, 2)).control(1), qs)
        else:
            a = int(rng.integers(n))
            qc.u(float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)), a)
    return qc

"""CPU oracle for the hot path: gate-stream application to a complex statevector.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and only as the checker / the reported CPU baseline.

What it restates
----------------
The reference simulates its circuits with ``qiskit.quantum_info.Statevector`` (call sites
``quantum_linear_systems/implementations/hhl_qiskit_implementation.py:49`` and
``quantum_linear_systems/implementations/vqls_qiskit_implementation.py:65`` in the reference
tree).  That class lives in the third-party dependency **qiskit-terra 0.46.3**
(``poetry.lock:3144-3145``), which is neither vendored under ``/root/reference`` nor installed in
this image, so its published algorithm is restated here:

* ``Statevector.from_instruction``: start from ``|0...0>`` (``psi[0] = 1``), walk the circuit;
* ``Statevector._evolve_operator(statevec, oper, qargs)``: view the state as an ``n``-axis
  tensor (axis ``n-1-q`` is qubit ``q``), move the axes ``[n-1-q for q in reversed(qargs)]`` to
  the front, flatten to ``(2**k, 2**(n-k))``, left-multiply by the ``2**k x 2**k`` gate matrix
  with ``np.dot`` (``qargs[0]`` = LSB of the gate index), undo the transpose, flatten.

PARITY STATUS: **parity unpinned at amplitude level** -- the reference holds no per-amplitude
golden statevector and qiskit cannot be imported here.  The oracle is pinned (a) end to end by
the reference's only exact known-answer test (``tests/test_hhl_implementations.py:23-29``:
width 5, x = [1.125, 0.375]) via ``tests/test_oracle_kat.py``, (b) against an independent
dense-unitary (Kronecker) construction in ``oracle/kron_unitary.py``, and (c) against the
closed-form HHL state in ``oracle/hhl_analytic.py``.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple, Union

import numpy as np

GateApp = Tuple  # ("u", matrix, qargs) | ("init", amplitudes, qargs) | ("phase", angle)


def zero_state(num_qubits: int, dtype=np.complex128) -> np.ndarray:
    """``Statevector.from_instruction`` initial state: ``|0...0>``."""
    psi = np.zeros(2**num_qubits, dtype=dtype)
    psi[0] = 1.0
    return psi


def evolve_operator(psi: np.ndarray, mat: np.ndarray, qargs: Sequence[int]) -> np.ndarray:
    """One ``Statevector._evolve_operator`` step (reshape -> transpose -> ``np.dot`` -> inverse
    transpose).  Returns a new array, as the reference path does."""
    n = int(round(np.log2(psi.size)))
    k = len(qargs)
    mat = np.asarray(mat, dtype=psi.dtype)
    assert mat.shape == (2**k, 2**k)
    indices = [n - 1 - q for q in reversed(list(qargs))]
    axes = indices + [i for i in range(n) if i not in indices]
    axes_inv = np.argsort(axes).tolist()
    tensor = np.transpose(np.reshape(psi, (2,) * n), axes)
    tshape = tensor.shape
    tensor = np.reshape(np.dot(mat, np.reshape(tensor, (2**k, 2 ** (n - k)))), tshape)
    return np.reshape(np.transpose(tensor, axes_inv), 2**n)


def initialize_subsystem(psi: np.ndarray, amps: np.ndarray, qargs: Sequence[int]) -> np.ndarray:
    """State preparation ``|0..0> -> sum_i amps[i]|i>`` on ``qargs`` (must still be in |0..0>).

    qiskit's ``isometry`` circuit maps ``|0>`` to ``amps`` exactly including phase (SURVEY.md
    section 8a row a7); on a register known to be ``|0>`` that equals this direct assignment."""
    n = int(round(np.log2(psi.size)))
    qargs = list(qargs)
    k = len(qargs)
    idx = np.arange(psi.size)
    sub = np.zeros(psi.size, dtype=np.int64)
    for j, q in enumerate(qargs):
        sub |= ((idx >> q) & 1) << j
    if np.max(np.abs(psi[sub != 0])) > 1e-12:
        raise ValueError("initialize_subsystem: target qubits are not in |0...0>")
    mask = 0
    for q in qargs:
        mask |= 1 << q
    base = idx & ~mask
    return (np.asarray(amps, dtype=psi.dtype)[sub] * psi[base]).astype(psi.dtype)


def run(num_qubits: int, stream: Iterable[GateApp], dtype=np.complex128, psi0: np.ndarray = None) -> np.ndarray:
    """Apply a gate stream the way ``Statevector(circuit)`` would.

    ``stream`` items: ``("u", matrix, qargs)`` -> ``evolve_operator``; ``("init", amps, qargs)``
    -> state preparation; ``("phase", angle)`` -> ``psi *= exp(1j*angle)`` (a definition's
    ``global_phase``)."""
    psi = zero_state(num_qubits, dtype) if psi0 is None else np.array(psi0, dtype=dtype)
    for item in stream:
        tag = item[0]
        if tag == "u":
            psi = evolve_operator(psi, item[1], item[2])
        elif tag == "init":
            psi = initialize_subsystem(psi, item[1], item[2])
        elif tag == "phase":
            psi = psi * np.exp(1j * float(item[1]))
        else:
            raise ValueError(f"unknown stream item {tag!r}")
    return psi


# -- reductions the reference performs on the final state --------------------------------------


def prob_mask(psi: np.ndarray, mask: int, value: int) -> float:
    """sum |psi_i|^2 over indices with ``i & mask == value``."""
    idx = np.arange(psi.size)
    sel = (idx & mask) == value
    return float(np.sum(np.abs(psi[sel]) ** 2))


def expectation_z(psi: np.ndarray, qubit: int) -> float:
    """``Statevector.expectation_value(Z_qubit)`` = sum_i (-1)^{bit_q(i)} |psi_i|^2."""
    idx = np.arange(psi.size)
    sign = 1.0 - 2.0 * ((idx >> qubit) & 1)
    return float(np.sum(sign * np.abs(psi) ** 2))


def hhl_norm(psi: np.ndarray, size: int, scaling: float) -> float:
    """``HHL._calculate_norm`` (linear_solvers 0.1.1): sqrt(P(flag=1, clock=0)) / scaling, i.e.
    the norm of the slice ``[2**(n-1), 2**(n-1) + size)`` divided by ``lambda_min``."""
    half = psi.size // 2
    return float(np.linalg.norm(psi[half : half + size]) / scaling)


def extract_solution(psi: np.ndarray, size: int) -> np.ndarray:
    """``extract_hhl_solution_vector_from_state_vector`` (reference ``utils.py:77-94``): real part
    of the slice with flag = 1 and all other non-solution qubits 0, L2-normalised."""
    half = psi.size // 2
    v = np.real(psi[half : half + size])
    return v / np.linalg.norm(v)

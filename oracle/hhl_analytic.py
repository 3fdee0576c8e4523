"""Closed-form HHL final state (dense linear algebra in A's eigenbasis).

TEST INFRASTRUCTURE ONLY (see ``oracle/numpy_sv.py``).

Independent restatement of what ``linear_solvers.HHL.construct_circuit`` (quantum-linear-solvers
0.1.1, ``poetry.lock:3185-3186``; called from the reference at
``hhl_qiskit_implementation.py:35-38``) builds, without going through any gate stream:

* ``nb = log2 N``; ``nl = max(nb+1, ceil(log2(kappa+1))) + 1`` (``neg_vals=True``);
* ``delta = _get_delta(nl-1, lambda_min, lambda_max)``; ``t = 2*pi*delta/lambda_min/2``;
* QPE on the clock register = ``W_phi = R F^dagger diag(exp(2*pi*i*phi*x)) H^{(x)nl}`` for an
  eigenvector with eigenphase ``phi = lambda*t/(2*pi)`` (``R`` = bit reversal, ``F`` = DFT on the
  little-endian value), see SURVEY.md Appendix A;
* ``ExactReciprocal(nl, delta, neg_vals=True)`` rotates the flag by ``RY(theta(c))`` with ``c``
  the bit-reversed clock register value;
* final state ``sum_k beta_k |u_k> (x) (W_k^dagger (x) I_flag) Rot (W_k|0> (x) |0>_flag)``.

Used to check the gate-level builder + oracle (and through them the CUDA engine) on the toy
problems; the 2x2 case reproduces Appendix B of SURVEY.md.
"""
from __future__ import annotations

import math

import numpy as np


def get_delta(n_l: int, lambda_min: float, lambda_max: float) -> float:
    """``HHL._get_delta``: ``lambda_min`` scaled to ``n_l`` bits, truncated, as a binary fraction."""
    lam = abs(lambda_min * (2**n_l - 1) / lambda_max)
    if abs(lam - 1) < 1e-7:
        lam = 1
    bits = format(int(lam), "#0" + str(n_l + 2) + "b")[2:]
    return sum(int(ch) / 2 ** (i + 1) for i, ch in enumerate(bits))


def hhl_parameters(matrix: np.ndarray):
    n = matrix.shape[0]
    nb = int(np.log2(n))
    kappa = np.linalg.cond(matrix)
    nl = max(nb + 1, int(np.ceil(np.log2(kappa + 1)))) + 1
    ev = np.abs(np.linalg.eigvals(matrix))
    lmin, lmax = float(np.min(ev)), float(np.max(ev))
    delta = get_delta(nl - 1, lmin, lmax)
    t = 2 * np.pi * delta / lmin / 2
    return nb, nl, delta, t, lmin


def reciprocal_angle_by_value(nl: int, delta: float) -> np.ndarray:
    """Net flag rotation angle for every clock *value* ``c`` (sign bit = MSB of ``c``)."""
    m = 2 ** (nl - 1)
    out = np.zeros(2**nl)
    for c in range(2**nl):
        v, sign = c & (m - 1), c >> (nl - 1)
        if v == 0:
            continue
        # sign = 0: the value bits read lambda ~ v/m;  sign = 1 (two's complement): lambda ~ -(1 - v/m)
        lam = v / m if not sign else -(1.0 - v / m)
        r = delta / lam
        if math.isclose(abs(r), 1, abs_tol=1e-5):
            th = math.copysign(math.pi, r)
        elif abs(r) < 1:
            th = 2 * math.asin(r)
        else:
            th = 0.0
        out[c] = th
    return out


def _bitrev(x: int, nbits: int) -> int:
    return int(format(x, f"0{nbits}b")[::-1], 2) if nbits else 0


def hhl_state(matrix: np.ndarray, vector: np.ndarray) -> np.ndarray:
    """Final ``2**(nb+nl+1)`` amplitude vector of the HHL circuit (register order b, clock, flag)."""
    matrix = np.asarray(matrix, dtype=complex)
    nb, nl, delta, t, _ = hhl_parameters(matrix)
    dim_c = 2**nl
    w_eig, v_eig = np.linalg.eigh(matrix)
    b = np.asarray(vector, dtype=complex).reshape(-1)
    b = b / np.linalg.norm(b)
    beta = v_eig.conj().T @ b
    x = np.arange(dim_c)
    dft_dag = np.exp(-2j * np.pi * np.outer(x, x) / dim_c) / math.sqrt(dim_c)  # F^dagger
    rev = np.array([_bitrev(i, nl) for i in range(dim_c)])
    perm = np.zeros((dim_c, dim_c))
    perm[rev, x] = 1.0  # R |c> = |rev(c)>
    theta_by_value = reciprocal_angle_by_value(nl, delta)
    theta_by_reg = theta_by_value[rev]  # register index r holds value rev(r)
    psi = np.zeros((2, dim_c, 2**nb), dtype=complex)
    for k in range(len(w_eig)):
        phi = w_eig[k] * t / (2 * np.pi)
        full_w = perm @ dft_dag @ np.diag(np.exp(2j * np.pi * phi * x)) @ _hadamard(nl)
        w0 = full_w[:, 0]  # W_k |0...0>
        chi0 = np.cos(theta_by_reg / 2) * w0
        chi1 = np.sin(theta_by_reg / 2) * w0
        psi[0] += beta[k] * np.outer(full_w.conj().T @ chi0, v_eig[:, k])
        psi[1] += beta[k] * np.outer(full_w.conj().T @ chi1, v_eig[:, k])
    return psi.reshape(-1)


def _hadamard(m: int) -> np.ndarray:
    h = np.array([[1.0]])
    h1 = np.array([[1, 1], [1, -1]]) / math.sqrt(2)
    for _ in range(m):
        h = np.kron(h1, h)
    return h


def hhl_solution(matrix: np.ndarray, vector: np.ndarray) -> np.ndarray:
    """The reference's post-processing (``hhl_qiskit_implementation.py:49-54``) on the closed-form
    state: slice, real part, normalise, multiply by ``euclidean_norm``."""
    psi = hhl_state(matrix, vector)
    n = matrix.shape[0]
    _, _, _, _, lmin = hhl_parameters(np.asarray(matrix))
    half = psi.size // 2
    sl = psi[half : half + n]
    v = np.real(sl)
    return np.linalg.norm(sl) / lmin * v / np.linalg.norm(v)
